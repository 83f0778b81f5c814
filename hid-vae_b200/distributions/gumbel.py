"""Gumbel-softmax sampling and the temperature schedule (reference distributions/gumbel.py:8-41).
The GUMBEL_SOFTMAX training forward runs in hv_gumbel_forward (csrc/gumbel.cu) with the same formulas; these functions
remain for shapes without a fused instantiation and as the API of the reference module."""
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


def sample_gumbel(shape: Tuple, device: torch.device, eps: float = 1e-20) -> Tensor:
    uniform = torch.rand(shape, device=device)
    return -torch.log(eps - torch.log(uniform + eps))


def gumbel_softmax_sample(logits: Tensor, temperature: float, device: torch.device) -> Tensor:
    return F.softmax((logits + sample_gumbel(logits.shape, device)) / temperature, dim=-1)


class TemperatureScheduler:
    def __init__(self, t0: float, min_t: float, anneal_rate: float, step_size: int) -> None:
        self.t0, self.min_t, self.anneal_rate, self.step_size = t0, min_t, anneal_rate, step_size
        self.t = t0

    def update_t(self, iter):
        if iter % self.step_size == self.step_size - 1:
            self.t = np.maximum(self.t * np.exp(-self.anneal_rate * iter), self.min_t)

    def get_t(self, iter):
        self.update_t(iter)
        return self.t
