"""`cycle` and `batch_to` of the reference's data/utils.py."""
from data.schemas import SeqBatch, TaggedSeqBatch


def cycle(dataloader):
    while True:
        yield from dataloader


def batch_to(batch, device, non_blocking: bool = False):
    """Move every tensor field of a (Tagged)SeqBatch to `device`; None fields stay None."""
    if not isinstance(batch, (SeqBatch, TaggedSeqBatch)):
        raise TypeError(f"batch_to: unsupported batch type {type(batch).__name__}")
    moved = [None if f is None else f.to(device, non_blocking=non_blocking) for f in batch]
    return type(batch)(*moved)
