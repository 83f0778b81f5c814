"""torchrun check of the data-parallel paths on real GPUs (NCCL): sharded k-means == single-process k-means, the flat
gradient all-reduce, item-sharded bulk assignment + final gather, and a few iterations of the gin-configured trainer, eager
and as CUDA graphs (gather + forward + backward | NCCL all-reduce | AdamW).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp_nccl.py
Prints one JSON line {"ok": true, ...} on rank 0.
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from hidvae_b200 import dist as hv  # noqa: E402
from init.kmeans import Kmeans  # noqa: E402

rank, world, local = hv.init_from_env("nccl")
dev = torch.device("cuda", local)
report = {}

# (1) distributed Lloyd iterations: every rank holds a contiguous shard of the same global matrix; the run must end with
#     the centroids of the single-process run on the whole matrix (same initial rows: rank 0's NumPy draw, broadcast).
g = torch.Generator().manual_seed(5)
full = F.normalize(torch.randn(6000, 32, generator=g), dim=-1).to(dev)
lo, hi = hv.shard_range(full.shape[0], rank, world)
np.random.seed(11)
torch.manual_seed(11)
single = Kmeans(k=64, max_iters=12).run(full)
np.random.seed(11)
torch.manual_seed(11)
km = Kmeans(k=64, max_iters=12, process_group=dist.group.WORLD)
sharded = km.run(full[lo:hi])
torch.testing.assert_close(sharded.centroids, single.centroids, rtol=1e-4, atol=1e-5)
assert torch.equal(sharded.assignment, single.assignment[lo:hi]), "sharded assignment differs from the single-process run"
everyone = [torch.empty_like(sharded.centroids) for _ in range(world)]
dist.all_gather(everyone, sharded.centroids)
assert all(torch.equal(everyone[0], e) for e in everyone), "ranks ended with different centroids"
report["kmeans_iterations"] = km.n_iters

# (2) flat gradient buffer on NCCL == gradient of the mean loss over both ranks' batches
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.SiLU(), torch.nn.Linear(32, 8)).to(dev)
hv.broadcast_parameters(model)
grads = hv.FlatGradAllReduce(model.parameters())
data = torch.randn(world, 64, 16, generator=torch.Generator().manual_seed(3)).to(dev)
grads.zero()
model(data[rank]).pow(2).mean().backward()
grads.all_reduce()
ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.SiLU(), torch.nn.Linear(32, 8)).to(dev)
ref.load_state_dict(model.state_dict())
sum(ref(data[r]).pow(2).mean() for r in range(world)).div(world).backward()
torch.testing.assert_close(grads.flat, torch.cat([p.grad.reshape(-1) for p in ref.parameters()]), rtol=1e-5, atol=1e-6)

# (3) item-sharded bulk assignment + the final gather == the unsharded table
import bench  # noqa: E402
tok = bench.make_tokenizer(dev)
items = bench.synth_items(70001, seed=9, device=dev)           # a ragged catalogue, identical on every rank
whole = tok.precompute_corpus_ids(items).clone()
tok.reset()
tok.precompute_corpus_ids(items, shard=(rank, world))
gathered = tok.gather_shards()
assert torch.equal(gathered, whole), "gathered shards differ from the unsharded id table"
report["bulk_items"] = int(gathered.shape[0])

# (4) a few iterations of the gin-configured trainer: k-means init across ranks, tagged batches, flat all-reduce, eval
from hidvae_b200 import gin_lite  # noqa: E402
import train_hidvae  # noqa: E402
gin_lite.clear_config()
with tempfile.TemporaryDirectory() as tmp:
    gin_lite.parse_config(f"""
import modules.quantize
train.iterations = 6
train.batch_size = 128
train.vae_input_dim = 768
train.vae_embed_dim = 32
train.vae_hidden_dims = [512, 256, 128]
train.vae_codebook_size = 256
train.vae_codebook_normalize = True
train.vae_codebook_mode = %modules.quantize.QuantizeForwardMode.ROTATION_TRICK
train.vae_n_layers = 3
train.vae_n_cat_feats = 0
train.commitment_weight = 0.5
train.use_kmeans_init = True
train.do_eval = True
train.eval_every = 6
train.log_every = 3
train.synthetic_items = 4096
train.tag_class_counts = [37, 168, 353]
train.dataset_folder = "{tmp}/data"
train.save_dir_root = "{tmp}/out"
""")
    for graphed in (False, True):   # (5) the same run with the step replayed as CUDA graphs around the NCCL exchange
        out = train_hidvae.train(use_cuda_graph=graphed)
        flat = torch.cat([p.detach().reshape(-1) for p in out["model"].parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(flat, ref), f"ranks hold different parameters after data-parallel training (cuda graph: {graphed})"
        losses = [h["loss"] for h in out["history"] if "loss" in h]
        assert all(np.isfinite(losses)), losses
        report["trainer_losses_graphed" if graphed else "trainer_losses"] = losses

dist.barrier()
if rank == 0:
    print(json.dumps(dict(ok=True, world=world, **report)))
dist.destroy_process_group()
