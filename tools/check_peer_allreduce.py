"""torchrun check of the peer-memory all-reduce against NCCL: values, bit-identical results on every rank, CUDA-graph
replay, and latency of both at the codebook-gradient size.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_peer_allreduce.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hidvae_b200.dist import PeerAllReduce, init_from_env  # noqa: E402

rank, world, local = init_from_env("nccl")
dev = torch.device("cuda", local)
n = 3 * 256 * 32
par = PeerAllReduce(n, dev)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
for it in range(20):
    x = torch.randn(n, device=dev, generator=gen)
    ref = x.clone()
    dist.all_reduce(ref)
    out = par(x.clone())
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    everyone = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(everyone, out)
    assert all(torch.equal(everyone[0], e) for e in everyone), "ranks disagree bitwise"
# CUDA-graph replay
buf = torch.randn(n, device=dev, generator=gen)
src = buf.clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        work = src.clone(); par(work)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    work = src.clone()
    par(work)
ref = src.clone(); dist.all_reduce(ref)
for _ in range(10):
    g.replay()
torch.cuda.synchronize()
torch.testing.assert_close(work, ref, rtol=1e-5, atol=1e-5)
gc = torch.cuda.CUDAGraph()          # the copy alone, to take it out of both latencies
with torch.cuda.graph(gc):
    wc = src.clone()


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


t_peer = timed(g.replay)
gn = torch.cuda.CUDAGraph()
with torch.cuda.graph(gn):
    wn = src.clone()
    dist.all_reduce(wn)
t_nccl = timed(gn.replay)
t_copy = timed(gc.replay)
res = torch.tensor([t_peer - t_copy, t_nccl - t_copy], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps(dict(check="peer_allreduce", world=world, floats=n, peer_us=float(res[0]), nccl_us=float(res[1]), ok=True)))
dist.barrier()
os._exit(0)
