"""Batch and output containers with the field names of the reference's data/schemas.py, so code written
against the reference (`batch.x`, `out.sem_ids`, `losses.rqvae_loss`, ...) runs unchanged."""
from typing import NamedTuple, Optional

from torch import Tensor

FUT_SUFFIX = "_fut"


class SeqBatch(NamedTuple):
    user_ids: Tensor
    ids: Tensor
    ids_fut: Tensor
    x: Tensor
    x_fut: Tensor
    seq_mask: Tensor


class TaggedSeqBatch(NamedTuple):
    user_ids: Tensor
    ids: Tensor
    ids_fut: Tensor
    x: Tensor
    x_fut: Tensor
    seq_mask: Tensor
    tags_emb: Tensor
    tags_indices: Tensor


class TokenizedSeqBatch(NamedTuple):
    user_ids: Tensor
    sem_ids: Tensor
    sem_ids_fut: Tensor
    seq_mask: Tensor
    token_type_ids: Tensor
    token_type_ids_fut: Tensor


class TaggedTokenizedSeqBatch(NamedTuple):
    user_ids: Tensor
    sem_ids: Tensor
    sem_ids_fut: Tensor
    seq_mask: Tensor
    token_type_ids: Tensor
    token_type_ids_fut: Tensor
    tags_emb: Tensor
    tags_indices: Tensor


class HRqVaeComputedLosses(NamedTuple):
    loss: Tensor
    reconstruction_loss: Tensor
    rqvae_loss: Tensor
    tag_align_loss: Tensor
    tag_pred_loss: Tensor
    tag_pred_accuracy: Tensor
    embs_norm: Tensor
    p_unique_ids: Tensor
    tag_align_loss_by_layer: Optional[Tensor] = None
    tag_pred_loss_by_layer: Optional[Tensor] = None
    tag_pred_accuracy_by_layer: Optional[Tensor] = None
    sem_id_uniqueness_loss: Optional[Tensor] = None


class HRqVaeOutput:
    """Attribute bag returned by HRqVae.get_semantic_ids (a plain class in the reference too, data/schemas.py:74-97)."""

    _FIELDS = ("embeddings", "residuals", "sem_ids", "quantize_loss", "tag_align_loss", "tag_pred_loss",
               "tag_pred_accuracy", "tag_align_loss_by_layer", "tag_pred_loss_by_layer", "tag_pred_accuracy_by_layer")

    def __init__(self, embeddings, residuals, sem_ids, quantize_loss, tag_align_loss, tag_pred_loss, tag_pred_accuracy,
                 tag_align_loss_by_layer=None, tag_pred_loss_by_layer=None, tag_pred_accuracy_by_layer=None) -> None:
        values = (embeddings, residuals, sem_ids, quantize_loss, tag_align_loss, tag_pred_loss, tag_pred_accuracy,
                  tag_align_loss_by_layer, tag_pred_loss_by_layer, tag_pred_accuracy_by_layer)
        for name, value in zip(self._FIELDS, values):
            setattr(self, name, value)
