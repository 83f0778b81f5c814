"""`eval_mode`, `parse_config` and small tensor helpers with the names of the reference's modules/utils.py.
`parse_config` reads the same .gin files through hidvae_b200.gin_lite (gin-config itself is not required)."""
import argparse

import torch
from torch import Tensor

from hidvae_b200 import gin_lite


def eval_mode(fn):
    """Run a method with the module in eval() and restore the previous training flag afterwards."""
    def inner(self, *args, **kwargs):
        was_training = self.training
        self.eval()
        try:
            return fn(self, *args, **kwargs)
        finally:
            self.train(was_training)
    return inner


def select_columns_per_row(x: Tensor, indices: Tensor) -> Tensor:
    assert x.shape[0] == indices.shape[0]
    assert indices.shape[1] <= x.shape[1]
    return x.gather(1, indices)


def maybe_repeat_interleave(x, repeats, dim):
    return x.repeat_interleave(repeats, dim=dim) if isinstance(x, Tensor) else x


def parse_config(argv=None) -> str:
    parser = argparse.ArgumentParser()
    parser.add_argument("config_path", type=str, help="Path to gin config file.")
    args, _ = parser.parse_known_args(argv)
    gin_lite.parse_config_file(args.config_path)
    return args.config_path
