"""CPU tests of the host-side logic around the kernels: the gin-syntax reader, the batch schema / synthetic
catalogue, the id-table layout of the tokenizer, and the data-parallel plumbing (world_size-2 gloo processes:
flat-buffer gradient all-reduce, sharded k-means row exchange, item sharding)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hid-vae_b200")


# ------------------------------------------------------------------------------------------------------------------
# gin_lite
# ------------------------------------------------------------------------------------------------------------------
def test_gin_lite_reads_the_shipped_and_reference_style_configs():
    from hidvae_b200 import gin_lite
    from modules.quantize import QuantizeForwardMode
    from data.tags_processed import RecDataset
    gin_lite.clear_config()
    gin_lite.parse_config_file(os.path.join(PKG, "configs", "h_rqvae_amazon.gin"))
    b = gin_lite.bindings("train")
    assert b["vae_codebook_mode"] is QuantizeForwardMode.ROTATION_TRICK and b["dataset"] is RecDataset.AMAZON
    assert b["vae_hidden_dims"] == [512, 256, 128] and b["commitment_weight"] == 0.4 and b["lr_scheduler_eta_min"] == 7e-8
    assert b["lr_scheduler_type"] == "cosine" and b["use_kmeans_init"] is True and b["tag_class_counts"] == [38, 168, 348]
    gin_lite.clear_config()
    gin_lite.parse_config_file(os.path.join(PKG, "configs", "h_rqvae_kuairand.gin"))
    b = gin_lite.bindings("train")
    assert b["dataset"] is RecDataset.KUAIRAND and b["sem_id_uniqueness_margin"] == 0.5 and b["commitment_weight"] == 0.5
    # reference formatting: no spaces around '=', trailing spaces, comment lines, '#'-commented bindings
    gin_lite.clear_config()
    gin_lite.parse_config("import modules.quantize\ntrain.iterations=400000\n#train.tag_class_counts=[6, 130, 927]\n"
                          "train.vae_input_dim=768  \ntrain.dataset_split=\"sports\"  # trailing comment\n"
                          "train.vae_codebook_mode=%modules.quantize.QuantizeForwardMode.STE\n"
                          "train.vae_hidden_dims=[512,\n   256, 128]\n")
    b = gin_lite.bindings("train")
    assert b == dict(iterations=400000, vae_input_dim=768, dataset_split="sports",
                     vae_codebook_mode=QuantizeForwardMode.STE, vae_hidden_dims=[512, 256, 128])
    gin_lite.clear_config()


def test_gin_lite_configurable_and_errors():
    from hidvae_b200 import gin_lite
    gin_lite.clear_config()

    @gin_lite.configurable
    def job(a=1, b=2, c="x"):
        return a, b, c

    gin_lite.parse_config("job.a = 10\njob.c = 'y'\n")
    assert job() == (10, 2, "y") and job(a=5) == (5, 2, "y") and job(7) == (7, 2, "y")
    gin_lite.bind_parameter("job.b", 3)
    assert job() == (10, 3, "y") and gin_lite.query_parameter("job.b") == 3
    gin_lite.parse_config("job.nope = 1\n")
    with pytest.raises(gin_lite.GinLiteError, match="does not match any parameter"):
        job()
    gin_lite.clear_config()
    with pytest.raises(gin_lite.GinLiteError, match="unknown constant"):
        gin_lite.parse_config("job.a = %no.such.Constant\n")
    with pytest.raises(gin_lite.GinLiteError, match="cannot parse"):
        gin_lite.parse_config("this is not gin\n")
    with pytest.raises(gin_lite.GinLiteError, match="unbalanced"):
        gin_lite.parse_config("job.a = [1, 2\n")
    gin_lite.parse_config("import a.module.that.does.not.exist\njob.a = 4\n")      # tolerated like a missing dataset dep
    assert job()[0] == 4
    gin_lite.clear_config()


def test_trainer_signature_covers_every_reference_gin_key():
    """Every key bound by the shipped configs must be a parameter of train() (the reference's names)."""
    import inspect
    from hidvae_b200 import gin_lite
    import train_hidvae
    params = set(inspect.signature(train_hidvae.train.__wrapped__).parameters)
    for cfg in ("h_rqvae_amazon.gin", "h_rqvae_kuairand.gin"):
        gin_lite.clear_config()
        gin_lite.parse_config_file(os.path.join(PKG, "configs", cfg))
        assert set(gin_lite.bindings("train")) <= params
    gin_lite.clear_config()


# ------------------------------------------------------------------------------------------------------------------
# schema / synthetic catalogue / id-table layout
# ------------------------------------------------------------------------------------------------------------------
def test_item_data_schema():
    from data.schemas import TaggedSeqBatch
    from data.tags_processed import ItemData, RecDataset
    from data.utils import batch_to
    ds = ItemData(root="", dataset=RecDataset.AMAZON, n_items=500, train_test_split="all", seed=1)
    tr = ItemData(root="", dataset=RecDataset.AMAZON, n_items=500, train_test_split="train", seed=1)
    ev = ItemData(root="", dataset=RecDataset.AMAZON, n_items=500, train_test_split="eval", seed=1)
    assert len(ds) == 500 and len(tr) + len(ev) == 500 and 0 < len(ev) < 60
    b = ds[torch.arange(7)]
    assert isinstance(b, TaggedSeqBatch) and b.x.shape == (7, 768) and b.tags_emb.shape == (7, 3, 768)
    assert b.tags_indices.shape == (7, 3) and b.tags_indices.dtype == torch.int64 and int(b.tags_indices.max()) < 348
    torch.testing.assert_close(b.x.norm(dim=-1), torch.ones(7))
    assert ds[3].x.shape == (1, 768) and ds[2:5].ids.tolist() == [2, 3, 4]
    moved = batch_to(b, "cpu")
    assert isinstance(moved, TaggedSeqBatch) and torch.equal(moved.x, b.x)


def test_tokenizer_id_table_layout():
    from modules.tokenizer.h_semids import HSemanticIdTokenizer as T
    kw = dict(input_dim=16, output_dim=8, hidden_dims=[12], codebook_size=4, n_layers=3, n_cat_feats=0,
              tag_class_counts=[2, 3, 4], tag_embed_dim=4)
    plain, cat, inter = T(**kw), T(use_concatenated_ids=True, **kw), T(use_interleaved_ids=True, **kw)
    assert plain._columns() == ([0, 1, 2], []) and plain.sem_ids_dim == 3
    assert cat._columns() == ([0, 1, 2], [3, 4, 5]) and cat.sem_ids_dim == 6
    assert inter._columns() == ([0, 2, 4], [1, 3, 5]) and inter.sem_ids_dim == 6
    assert T(use_dedup_dim=True, **kw).sem_ids_dim == 4
    with pytest.raises(ValueError):
        T(use_concatenated_ids=True, use_interleaved_ids=True, **kw)
    hits = plain._get_hits(torch.tensor([[1, 2], [0, 0]]), torch.tensor([[1, 2], [3, 4], [1, 2]]))
    assert hits.tolist() == [[True, False, True], [False, False, False]]


# ------------------------------------------------------------------------------------------------------------------
# data-parallel plumbing on two gloo processes
# ------------------------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out_dir):
    import sys
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from hidvae_b200 import dist as hv
    from init.kmeans import Kmeans
    r, w, _ = hv.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and hv.world_size() == world

    # (1) flat-buffer gradient all-reduce == gradient of the mean loss over both ranks' batches
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    if rank == 1:                                     # ranks start different; broadcast makes them equal
        for p in model.parameters():
            p.data.add_(1.0)
    hv.broadcast_parameters(model)
    grads = hv.FlatGradAllReduce(model.parameters())
    g = torch.Generator().manual_seed(100)
    data = torch.randn(2, 8, 6, generator=g)          # [rank, batch, features], identical on both ranks
    for step in range(2):                             # twice: the views must survive zero() and a second backward
        grads.zero()
        model(data[rank]).pow(2).mean().backward()
        grads.check_views()
        grads.all_reduce()
    flat = grads.flat.clone()
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ref.load_state_dict(model.state_dict())
    (0.5 * (ref(data[0]).pow(2).mean() + ref(data[1]).pow(2).mean())).backward()
    ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    torch.testing.assert_close(flat, ref_flat, rtol=1e-6, atol=1e-7)
    work = grads.all_reduce(async_op=True)            # async variant (overlap with the eval encode in bench.py)
    grads.finish(work)
    torch.testing.assert_close(grads.flat, ref_flat, rtol=1e-6, atol=1e-7)   # mean of two equal buffers

    # (2) k-means row exchange: rank 0 draws global rows, every rank contributes the rows of its shard
    full = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    lo, hi = hv.shard_range(10, rank, world)
    km = Kmeans(k=3, process_group=dist.group.WORLD)
    first, total = km._row_offsets(hi - lo, torch.device("cpu"))
    assert (first, total) == (lo, 10)
    idx = torch.tensor([9, 0, 5]) if rank == 0 else torch.tensor([1, 1, 1])
    idx = km._broadcast_idx(idx, torch.device("cpu"))
    assert idx.tolist() == [9, 0, 5]
    rows = km._fetch_rows(full[lo:hi], idx, first)
    assert torch.equal(rows, full[idx])

    # (3) item sharding covers the catalogue exactly once
    pieces = [hv.shard_range(12101, r_, 8) for r_ in range(8)]
    assert pieces[0][0] == 0 and pieces[-1][1] == 12101 and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_tokenizer_cached_consumers_match_reference_golden(golden_dir):
    """`HSemanticIdTokenizer.forward` (cached branch), `_tokenize_seq_batch_from_cached` and `exists_prefix` against
    outputs recorded from the reference's own class (oracle/make_golden.py::make_tokenizer_cases).  Pure indexing of
    `cached_ids`: runs without a GPU."""
    import numpy as np
    import torch
    from data.schemas import SeqBatch
    from modules.tokenizer.h_semids import HSemanticIdTokenizer

    g = np.load(os.path.join(golden_dir, "tokenizer_cached.npz"))
    for name, kw in (("plain", {}), ("concat", dict(use_concatenated_ids=True, tag_class_counts=[5, 7, 9])),
                     ("interleaved", dict(use_interleaved_ids=True, tag_class_counts=[5, 7, 9]))):
        tok = HSemanticIdTokenizer(input_dim=24, output_dim=8, hidden_dims=[16], codebook_size=16, n_layers=3,
                                   n_cat_feats=0, tag_embed_dim=8, **kw)
        assert tok.sem_ids_dim == int(g[f"{name}/sem_ids_dim"])
        tok.cached_ids = torch.from_numpy(g[f"{name}/cached_ids"])
        ids, ids_fut = torch.from_numpy(g[f"{name}/ids"]), torch.from_numpy(g[f"{name}/ids_fut"])
        mask = torch.from_numpy(g[f"{name}/seq_mask"])
        b, n = ids.shape
        batch = SeqBatch(user_ids=torch.arange(b), ids=ids, ids_fut=ids_fut, x=torch.zeros(b, n, 24), x_fut=torch.zeros(b, 24),
                         seq_mask=mask)
        out = tok(batch)
        for key, val in (("sem_ids", out.sem_ids), ("sem_ids_fut", out.sem_ids_fut), ("out_seq_mask", out.seq_mask),
                         ("token_type_ids", out.token_type_ids), ("token_type_ids_fut", out.token_type_ids_fut),
                         ("from_cached", tok._tokenize_seq_batch_from_cached(ids))):
            assert np.array_equal(val.numpy(), g[f"{name}/{key}"]), (name, key)
        assert np.array_equal(tok.exists_prefix(torch.from_numpy(g[f"{name}/prefixes"])).numpy(), g[f"{name}/prefix_hits"])
        assert np.array_equal(tok.exists_prefix(torch.from_numpy(g[f"{name}/full_rows"])).numpy(), g[f"{name}/full_hits"])
        # ids beyond the cache read row 0 (h_semids.py:251-252)
        far = torch.tensor([[0, 10 ** 6]])
        assert torch.equal(tok._tokenize_seq_batch_from_cached(far)[0, tok.sem_ids_dim:], tok.cached_ids[0])


def test_flat_gradient_backward_adds_like_autograd_accumulation():
    """FlatGradAllReduce.backward (detach the grads, backward, one multi-tensor add into the views) against autograd's own
    accumulation into the views: two micro-steps, a parameter that receives no gradient, views bound again afterwards."""
    from hidvae_b200 import dist as hv

    def build():
        torch.manual_seed(3)
        m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        unused = torch.nn.Parameter(torch.ones(4))
        return m, unused

    data = torch.randn(2, 8, 6, generator=torch.Generator().manual_seed(5))
    flats = []
    for use_backward in (False, True):
        m, unused = build()
        grads = hv.FlatGradAllReduce(list(m.parameters()) + [unused])
        grads.zero()
        for k in range(2):
            loss = m(data[k]).pow(2).mean()
            grads.backward(loss) if use_backward else loss.backward()
        grads.check_views()
        assert float(unused.grad.abs().max()) == 0.0
        flats.append(grads.flat.clone())
    torch.testing.assert_close(flats[1], flats[0], rtol=1e-6, atol=1e-7)
    assert float(flats[0].abs().max()) > 0
