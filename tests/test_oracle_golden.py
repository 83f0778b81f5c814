"""The oracle (oracle/*.py) against the fixtures recorded from the real reference (oracle/make_golden.py).

CPU only.  These pin the checker: forward values must match the reference to fp32 round-off (same ATen
operators in the same order), autograd through the oracle must reproduce the reference's gradients, and the
closed-form backward recursion the CUDA kernel implements must agree with autograd.
"""
import os

import numpy as np
import pytest
import torch

from oracle import rq as O
from oracle import kmeans as OK

MODES = {"ste": O.MODE_STE, "rot": O.MODE_ROTATION_TRICK}


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("mname", ["ste", "rot"])
@pytest.mark.parametrize("normalize", [0, 1])
@pytest.mark.parametrize("training", [0, 1])
def test_quantize_level_matches_reference(golden_dir, mname, normalize, training):
    g = _load(golden_dir, "quantize_levels.npz")
    tag = f"{mname}_norm{normalize}_train{training}"
    x = _t(g[f"{tag}/x"]).requires_grad_(True)
    w = _t(g[f"{tag}/weight"]).requires_grad_(True)
    beta = float(g[f"{tag}/beta"])
    cb = O.effective_codebook(w, bool(normalize))
    out = O.quantize_level(x, cb, MODES[mname], beta, bool(training))
    assert torch.equal(out.ids, _t(g[f"{tag}/ids"]))
    torch.testing.assert_close(out.embeddings, _t(g[f"{tag}/emb_out"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(out.loss, _t(g[f"{tag}/loss"]), rtol=1e-6, atol=1e-7)
    ((out.embeddings * _t(g[f"{tag}/g_emb"])).sum() + (out.loss * _t(g[f"{tag}/g_loss"])).sum()).backward()
    torch.testing.assert_close(x.grad, _t(g[f"{tag}/grad_x"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(w.grad, _t(g[f"{tag}/grad_weight"]), rtol=1e-5, atol=1e-6)


def test_gumbel_level_matches_reference(golden_dir):
    g = _load(golden_dir, "quantize_levels.npz")
    x, w, u = _t(g["gumbel/x"]), _t(g["gumbel/weight"]), _t(g["gumbel/uniform"])
    out = O.quantize_level(x, O.effective_codebook(w, False), O.MODE_GUMBEL_SOFTMAX, 0.25, True, 0.2, uniform=u)
    assert torch.equal(out.ids, _t(g["gumbel/ids"]))
    torch.testing.assert_close(out.embeddings, _t(g["gumbel/emb_out"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out.loss, _t(g["gumbel/loss"]), rtol=1e-5, atol=1e-6)


def _rq_case(g, tag):
    enc = _t(g[f"{tag}/enc"])
    weights = _t(g[f"{tag}/weights"])
    beta = float(g[f"{tag}/beta"])
    return enc, weights, beta


@pytest.mark.parametrize("mname", ["ste", "rot"])
@pytest.mark.parametrize("training", [0, 1])
def test_rq_forward_and_autograd_match_reference(golden_dir, mname, training):
    g = _load(golden_dir, "rq_c1.npz")
    tag = f"{mname}_train{training}"
    enc, weights, beta = _rq_case(g, tag)
    enc = enc.clone().requires_grad_(True)
    ws = [weights[l].clone().requires_grad_(True) for l in range(weights.shape[0])]
    cbs = [O.effective_codebook(w, normalize=(l == 0)) for l, w in enumerate(ws)]  # h_rqvae.py:295
    out = O.rq_forward(enc, cbs, MODES[mname], beta, bool(training))
    assert torch.equal(out.sem_ids, _t(g[f"{tag}/sem_ids"]))
    assert out.embeddings.shape == _t(g[f"{tag}/embeddings"]).shape
    torch.testing.assert_close(out.embeddings, _t(g[f"{tag}/embeddings"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(out.residuals, _t(g[f"{tag}/residuals"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(out.quantize_loss, _t(g[f"{tag}/quantize_loss"]), rtol=1e-6, atol=1e-7)
    ((out.embeddings * _t(g[f"{tag}/g_emb"])).sum() + (out.quantize_loss * _t(g[f"{tag}/g_loss"])).sum()).backward()
    torch.testing.assert_close(enc.grad, _t(g[f"{tag}/grad_enc"]), rtol=1e-5, atol=1e-6)
    for l, w in enumerate(ws):
        torch.testing.assert_close(w.grad, _t(g[f"{tag}/grad_weights"][l]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mname", ["ste", "rot"])
def test_closed_form_backward_matches_reference_autograd(golden_dir, mname):
    """SURVEY section 8a backward recursion == the reference's autograd (training mode)."""
    g = _load(golden_dir, "rq_c1.npz")
    tag = f"{mname}_train1"
    enc, weights, beta = _rq_case(g, tag)
    ws = [weights[l].clone().requires_grad_(True) for l in range(weights.shape[0])]
    cbs = [O.effective_codebook(w, normalize=(l == 0)) for l, w in enumerate(ws)]
    sem_ids = _t(g[f"{tag}/sem_ids"])
    g_enc, g_cbs = O.rq_backward_closed_form(enc, [c.detach() for c in cbs], sem_ids, _t(g[f"{tag}/g_emb"]),
                                             _t(g[f"{tag}/g_loss"]), MODES[mname], beta)
    torch.testing.assert_close(g_enc, _t(g[f"{tag}/grad_enc"]), rtol=2e-5, atol=2e-6)
    # chain the effective-codebook gradient through out_proj (row L2 norm on level 0) with autograd
    torch.autograd.backward(cbs, g_cbs)
    for l, w in enumerate(ws):
        torch.testing.assert_close(w.grad, _t(g[f"{tag}/grad_weights"][l]), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["dups", "margin", "nodup"])
def test_uniqueness_loss_matches_reference(golden_dir, name):
    g = _load(golden_dir, "uniqueness.npz")
    ids = _t(g[f"{name}/ids"])
    feats = _t(g[f"{name}/feats"]).requires_grad_(True)
    margin, weight = float(g[f"{name}/margin"]), float(g[f"{name}/weight"])
    loss = O.uniqueness_loss(ids, feats, margin, weight)
    torch.testing.assert_close(loss, _t(g[f"{name}/loss"]), rtol=1e-6, atol=1e-7)
    if loss.requires_grad:
        loss.backward()
        torch.testing.assert_close(feats.grad, _t(g[f"{name}/grad_feats"]), rtol=1e-5, atol=1e-7)
    # the call as wired in HRqVae.forward (transposed ids) -- SURVEY quirk 1
    wired = O.uniqueness_loss(ids.transpose(0, 1), feats.detach(), margin, weight)
    torch.testing.assert_close(wired, _t(g[f"{name}/loss_as_wired"]))
    torch.testing.assert_close(O.p_unique_ids(ids), _t(g[f"{name}/p_unique"]))


@pytest.mark.parametrize("name", ["blobs", "unit32"])
def test_kmeans_matches_reference(golden_dir, name):
    g = _load(golden_dir, "kmeans.npz")
    x = _t(g[f"{name}/x"])
    init_idx = g[f"{name}/init_idx"]
    k = init_idx.shape[0]
    torch.set_num_threads(1)
    res = OK.kmeans_run(x, k, init_idx=init_idx)
    assert torch.equal(res.assignment, _t(g[f"{name}/assignment"]))
    torch.testing.assert_close(res.centroids, _t(g[f"{name}/centroids"]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("normalize", [0, 1])
def test_encoder_oracle_matches_reference_mlp(golden_dir, normalize):
    """oracle/encoder.py against the reference's own MLP (modules/encoder.py) at the gin shape 768-512-256-128-32."""
    from oracle import encoder as OE
    g = _load(golden_dir, "encoder.npz")
    weights = OE.seeded_weights([int(v) for v in g["dims"]], int(g["seed"]))
    z = OE.mlp_forward(_t(g["x"]), weights, bool(normalize))
    torch.testing.assert_close(z, _t(g[f"z_norm{normalize}"]), rtol=1e-5, atol=1e-7)
