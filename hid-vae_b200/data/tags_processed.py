"""`RecDataset` and `ItemData` with the names the reference's trainer and gin files use (data/tags_processed.py:20-278).

The reference builds ItemData from downloaded review dumps, LLM-generated tags and sentence-T5 embeddings (network,
torch_geometric, polars: out of scope here and unavailable offline).  This ItemData serves the SAME batch schema
(`TaggedSeqBatch`: x [B, 768], tags_emb [B, L, 768], tags_indices [B, L]) from either
  * a processed tensor file `<root>/processed/items.pt` ({"x", "tags_emb", "tags_indices", optional "is_train"}), or
  * a seeded synthetic catalogue of the dataset's shape, resident on the device so that the
    training loop never waits on host memory.
"""
import logging
import os
from enum import Enum
from typing import Optional, Sequence

import torch
import torch.nn.functional as F
from torch.utils.data import Dataset

from data.schemas import SeqBatch, TaggedSeqBatch
from hidvae_b200.gin_lite import constants_from_enum


@constants_from_enum(module="data.tags_processed")
class RecDataset(Enum):
    AMAZON = 1
    ML_1M = 2
    ML_32M = 3
    KUAIRAND = 4


# catalogue sizes / tag vocabularies of the shipped configs (configs/h_rqvae_*.gin; Amazon-Beauty has 12,101 items;
# the reference does not state KuaiRand's item count -- 32,768 is this repo's stand-in)
SYNTHETIC_SHAPES = {
    RecDataset.AMAZON: dict(n_items=12101, tag_class_counts=(38, 168, 348)),
    RecDataset.KUAIRAND: dict(n_items=32768, tag_class_counts=(37, 168, 353)),
    RecDataset.ML_1M: dict(n_items=3883, tag_class_counts=(10, 100, 1000)),
    RecDataset.ML_32M: dict(n_items=87585, tag_class_counts=(10, 100, 1000)),
}


class ItemData(Dataset):
    """Item catalogue with the reference's batch schema (data/tags_processed.py:44-278).

    The processed catalogue `<root>/processed/items.pt` is loaded when it exists -- `force_process` never discards it
    (re-processing the raw downloads is outside this repo: it needs polars / torch_geometric / an LLM service).  A seeded
    SYNTHETIC catalogue of the same shape is built only when the caller opts in (`synthetic=True`, or an explicit
    `n_items`); otherwise a missing file raises, so that a shipped gin config can never silently train on noise."""

    def __init__(self, root: str, *args, force_process: bool = False, dataset: RecDataset = RecDataset.ML_1M,
                 train_test_split: str = "all", n_items: Optional[int] = None, input_dim: int = 768,
                 tag_embed_dim: int = 768, tag_class_counts: Optional[Sequence[int]] = None, seed: int = 0,
                 device: Optional[torch.device] = None, synthetic: Optional[bool] = None, **kwargs) -> None:
        shape = SYNTHETIC_SHAPES[dataset]
        path = os.path.join(root, "processed", "items.pt") if root else None
        have_file = bool(path) and os.path.isfile(path)
        if have_file and not synthetic:
            if force_process:
                logging.getLogger(__name__).warning(
                    "force_process=True: raw re-processing is not part of this repo; loading the processed catalogue %s", path)
            blob = torch.load(path, map_location="cpu")
            x, tags_emb, tags_indices = blob["x"][:, :input_dim], blob.get("tags_emb"), blob.get("tags_indices")
            is_train = blob.get("is_train")
            self.synthetic = False
        elif synthetic or (synthetic is None and n_items is not None):
            n = n_items or shape["n_items"]
            counts = list(tag_class_counts or shape["tag_class_counts"])
            g = torch.Generator().manual_seed(seed)
            x = F.normalize(torch.randn(n, input_dim, generator=g), dim=-1)
            tags_emb = torch.randn(n, len(counts), tag_embed_dim, generator=g)
            tags_indices = torch.stack([torch.randint(0, c, (n,), generator=g) for c in counts], dim=1)
            tags_indices[torch.rand(n, len(counts), generator=g) < 0.05] = -1   # 5 % missing tags
            is_train = torch.rand(n, generator=g) > 0.05                          # 95 / 5 split (tags_amazon.py:413)
            self.synthetic = True
        else:
            raise FileNotFoundError(
                f"processed catalogue {path!r} not found.  Pass synthetic=True (gin: train.synthetic_data = True) or an "
                "explicit n_items (train.synthetic_items) to build a seeded synthetic catalogue of the same shape.")
        if is_train is None:
            is_train = torch.rand(x.shape[0], generator=torch.Generator().manual_seed(42)) > 0.05
        keep = {"train": is_train, "eval": ~is_train, "all": torch.ones_like(is_train)}[train_test_split]
        self.item_data = x[keep]
        self.has_tags = tags_emb is not None and tags_indices is not None
        self.tags_emb = tags_emb[keep] if self.has_tags else None
        self.tags_indices = tags_indices[keep] if self.has_tags else None
        if device is not None:
            self.to(device)

    def to(self, device):
        self.item_data = self.item_data.to(device)
        if self.has_tags:
            self.tags_emb, self.tags_indices = self.tags_emb.to(device), self.tags_indices.to(device)
        return self

    def __len__(self) -> int:
        return self.item_data.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, int):
            idx = slice(idx, idx + 1)
        if isinstance(idx, (list, tuple)):
            idx = torch.as_tensor(idx)
        if isinstance(idx, torch.Tensor):
            idx = idx.to(self.item_data.device)
            ids = idx
        else:
            ids = torch.arange(len(self), device=self.item_data.device)[idx]
        filler = -torch.ones_like(ids)
        fields = dict(user_ids=filler, ids=ids, ids_fut=filler, x=self.item_data[idx], x_fut=filler,
                      seq_mask=torch.ones_like(ids, dtype=torch.bool))
        if self.has_tags:
            return TaggedSeqBatch(**fields, tags_emb=self.tags_emb[idx], tags_indices=self.tags_indices[idx])
        return SeqBatch(**fields)
