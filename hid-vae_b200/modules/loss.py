"""Loss modules with the class names and call signatures of the reference's modules/loss.py.

`QuantizeLoss` is on the hot path, and inside `Quantize`/`HRqVae` it is evaluated by the fused CUDA kernels; the
module here is the stand-alone operator with the same value/gradients (reference modules/loss.py:36-44).
Reconstruction and tag losses are HiD-VAE's supervision around the hot path: plain PyTorch GPU ops (out of
scope as kernels), kept so that the drop-in `HRqVae` is complete."""
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn


class ReconstructionLoss(nn.Module):
    """Per-row sum of squared errors (reference :7-12)."""

    def forward(self, x_hat: Tensor, x: Tensor) -> Tensor:
        return (x_hat - x).pow(2).sum(dim=-1)


class CategoricalReconstructionLoss(nn.Module):
    """SSE on the dense part + BCE-with-logits on the trailing `n_cat_feats` columns (reference :15-33)."""

    def __init__(self, n_cat_feats: int) -> None:
        super().__init__()
        self.reconstruction_loss = ReconstructionLoss()
        self.n_cat_feats = n_cat_feats

    def forward(self, x_hat: Tensor, x: Tensor) -> Tensor:
        c = self.n_cat_feats
        out = self.reconstruction_loss(x_hat[:, :-c], x[:, :-c])
        if c > 0:
            out = out + F.binary_cross_entropy_with_logits(x_hat[:, -c:], x[:, -c:], reduction="none").sum(dim=-1)
        return out


# the reference imports this class under a misspelt name (modules/h_rqvae.py:8); accept both
CategoricalReconstuctionLoss = CategoricalReconstructionLoss


class QuantizeLoss(nn.Module):
    """|sg(query) - value|^2 + commitment_weight * |query - sg(value)|^2 per row (reference :36-44)."""

    def __init__(self, commitment_weight: float = 1.0) -> None:
        super().__init__()
        self.commitment_weight = commitment_weight

    def forward(self, query: Tensor, value: Tensor) -> Tensor:
        codebook_term = (query.detach() - value).pow(2).sum(dim=-1)
        commit_term = (query - value.detach()).pow(2).sum(dim=-1)
        return codebook_term + self.commitment_weight * commit_term


class TagAlignmentLoss(nn.Module):
    """InfoNCE between the (concatenated) code embeddings and the projected tag embeddings of the same rows,
    scaled by alignment_weight / (1 + layer_idx / 2) (reference :48-84)."""

    def __init__(self, alignment_weight: float = 1.0, temperature: float = 0.1) -> None:
        super().__init__()
        self.alignment_weight = alignment_weight
        self.temperature = temperature

    def forward(self, codebook_emb: Tensor, tag_emb: Tensor, layer_idx: int) -> Tensor:
        a = F.normalize(codebook_emb, p=2, dim=-1)
        b = F.normalize(tag_emb, p=2, dim=-1)
        logits = a @ b.t() / self.temperature
        target = torch.arange(a.shape[0], device=a.device)
        return F.cross_entropy(logits, target) * self.alignment_weight / (1.0 + 0.5 * layer_idx)


class TagPredictionLoss(nn.Module):
    """Tag classification loss + accuracy over rows with a valid (>= 0) target (reference :88-321).

    Training-time behaviour kept from the reference: mixup of the logits with Beta(0.2, 0.2), label-smoothed
    cross entropy plus a small KL-to-uniform term, or (use_focal_loss) the smoothed focal variants with optional
    inverse-sqrt-frequency class weights."""

    def __init__(self, use_focal_loss: bool = False, focal_params: Optional[dict] = None,
                 class_counts: Optional[dict] = None) -> None:
        super().__init__()
        self.use_focal_loss = use_focal_loss
        self.focal_params = focal_params or {"gamma": 2.0, "alpha": 0.25}
        self.class_counts = class_counts
        self.use_label_smoothing = True
        self.label_smoothing_alpha = 0.1
        self.use_mixup = True
        self.mixup_alpha = 0.2
        self.weight_scheduler = None

    # -- helpers ---------------------------------------------------------------------------------------------
    def _soft_targets(self, logits: Tensor, targets: Tensor, gamma: float) -> Tensor:
        n_cls = logits.shape[-1]
        hot = torch.zeros_like(logits).scatter_(1, targets.unsqueeze(1), 1.0)   # (no host-side range check: capturable)
        if self.use_label_smoothing and logits.requires_grad:
            eps = min(0.25, self.label_smoothing_alpha + 0.015 * gamma + min(0.3, 0.05 * (n_cls / 100)))
            hot = hot * (1.0 - eps) + eps / n_cls
        return hot

    # the three loss forms return ONE VALUE PER ROW; `forward` averages them over the rows with a valid target
    def _focal_loss_with_smoothing(self, logits: Tensor, targets: Tensor, gamma: float = 2.0, alpha: float = 0.25) -> Tensor:
        soft = self._soft_targets(logits, targets, gamma)
        logp = F.log_softmax(logits, dim=-1)
        pt = (soft * logp.exp()).sum(dim=1)
        return alpha * (1.0 - pt).pow(gamma) * -(soft * logp).sum(dim=1)

    def _focal_loss_with_weights_and_smoothing(self, logits: Tensor, targets: Tensor, gamma: float = 2.0,
                                               class_weights: Optional[Tensor] = None) -> Tensor:
        n_cls = logits.shape[-1]
        soft = self._soft_targets(logits, targets, gamma)
        logp = F.log_softmax(logits, dim=-1)
        probs = logp.exp()
        pt = (soft * probs).sum(dim=1)
        row_w = class_weights[targets] if class_weights is not None else torch.ones_like(targets, dtype=torch.float)
        sharpened = gamma * (1.0 + 0.25 * min(1.0, n_cls / 250))
        loss = row_w * (1.0 - pt).pow(sharpened) * -(soft * logp).sum(dim=1)
        if n_cls > 100 and logits.requires_grad:
            kl = F.kl_div(torch.log(probs + 1e-8), torch.full_like(probs, 1.0 / n_cls), reduction="none").sum(dim=1)
            loss = loss + min(0.12, 0.015 * (n_cls / 100)) * kl   # (batchmean = the row sums averaged over the batch)
        return loss

    def _class_weights(self, layer_idx: int, device) -> Optional[Tensor]:
        if self.class_counts is None or layer_idx not in self.class_counts:
            return None
        counts = self.class_counts[layer_idx]
        if not isinstance(counts, Tensor) or counts.numel() == 0:
            return None
        freq = (counts.float() / counts.sum()).clamp(min=1e-6)
        w = freq.rsqrt()
        return (w / w.mean()).clamp(min=0.5, max=3.0).to(device)

    # -- forward ---------------------------------------------------------------------------------------------
    def forward(self, pred_logits: Tensor, target_indices: Tensor, layer_idx: int = 0):
        """Loss and accuracy over the rows whose target is >= 0 (reference :232-321).

        The reference compacts those rows with a boolean mask (`pred_logits[valid_mask]`), which costs a device -> host
        synchronisation per level and gives every step another shape.  Here all rows go through the same fixed-shape
        expressions and the averages are MASKED means -- same values (every term of the reference is a mean over the kept
        rows), no synchronisation, and the whole training step can be captured in a CUDA graph.  Mixup draws its partner
        among the kept rows, as the reference's randperm over the compacted batch does, and its Beta sample on the device."""
        keep = target_indices >= 0
        w = keep.to(pred_logits.dtype)
        denom = w.sum().clamp(min=1.0)                 # no kept row: every masked sum is 0, like the reference's early return
        mean = lambda per_row: (per_row * w).sum() / denom
        logits, targets = pred_logits, target_indices.clamp(min=0)
        accuracy = mean((logits.argmax(dim=-1) == targets).to(pred_logits.dtype))
        clean_probs = F.softmax(logits, dim=-1)

        mixed = self.use_mixup and logits.shape[0] > 1 and logits.requires_grad
        if mixed:
            n, dev = logits.shape[0], logits.device
            # a uniformly random permutation of the KEPT rows: kept rows in row order <-> kept rows in random order
            in_order = torch.argsort((~keep).to(torch.int8), stable=True)
            shuffled = torch.argsort(torch.rand(n, device=dev) + (~keep).to(torch.float32) * 2.0)
            perm = torch.empty_like(in_order).scatter_(0, in_order, shuffled)
            conc = torch.full((), float(self.mixup_alpha), device=dev)
            lam = torch.distributions.Beta(conc, conc, validate_args=False).sample()   # (argument validation reads back to the host)
            logits = lam * logits + (1 - lam) * logits[perm]
            target_sets = ((lam, targets), (1 - lam, targets[perm]))
        else:
            target_sets = ((1.0, targets),)

        if self.use_focal_loss:
            p = self.focal_params
            gamma = p.get(f"gamma_{layer_idx}", p.get("gamma", 2.0)) * (1 + 0.35 * layer_idx)
            alpha = max(0.08, p.get(f"alpha_{layer_idx}", p.get("alpha", 0.25)) - 0.06 * layer_idx)
            weights = self._class_weights(layer_idx, logits.device)
            if weights is not None:
                one = lambda t: self._focal_loss_with_weights_and_smoothing(logits, t, gamma, weights)
            else:
                one = lambda t: self._focal_loss_with_smoothing(logits, t, gamma, alpha)
            loss = sum(wt * mean(one(t)) for wt, t in target_sets)
        else:
            smoothing = min(0.25, 0.05 + 0.06 * layer_idx)
            ce = sum(wt * mean(F.cross_entropy(logits, t, reduction="none", label_smoothing=smoothing)) for wt, t in target_sets)
            uniform = torch.full_like(clean_probs, 1.0 / clean_probs.shape[-1])
            kl = 0.05 * mean(F.kl_div(torch.log(clean_probs + 1e-8), uniform, reduction="none").sum(dim=1))
            loss = ce + kl  # the reference's "L2 regulariser" iterates over the parameters of a Tensor: always 0
        return loss, accuracy
