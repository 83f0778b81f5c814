"""Timings of the auxiliary hot-path kernels next to the CPU oracle:
k-means codebook init at the C3 shape (20,000 x 32, K = 256) and the semantic-ID uniqueness loss / p_unique_ids.

    python tools/bench_aux.py > gpurun_out/aux.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from hidvae_b200 import ops  # noqa: E402
from init.kmeans import Kmeans  # noqa: E402
from oracle import kmeans as OK  # noqa: E402
from oracle import rq as O  # noqa: E402

torch.cuda.set_device(0)


def ev_time(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ---- k-means (init/kmeans.py:43-77): one Lloyd iteration and a whole run from the same initial rows ----
n, d, k = 20000, 32, 256
g = torch.Generator().manual_seed(0)
x = F.normalize(torch.randn(n, d, generator=g), dim=-1)
xd = x.cuda()
cent = xd[torch.randperm(n, generator=g)[:k].cuda()].clone()


def lloyd():
    a = ops.kmeans_assign(xd, cent)
    s, c, _ = ops.kmeans_accumulate(xd, a, k)
    return ops.kmeans_finalize(s, c, cent.clone())


t_iter = ev_time(lloyd)
np.random.seed(0)
t0 = time.perf_counter()
km = Kmeans(k=k, max_iters=50)
km.run(xd)
torch.cuda.synchronize()
t_run = time.perf_counter() - t0
res = dict(case="kmeans_c3", n=n, d=d, k=k, lloyd_iteration_ms=t_iter, run_s=t_run, run_updates=km.n_iters,
           bytes_per_iteration=n * (4 * d + 8) + k * (d + 1) * 4, hbm_gbs=(n * (4 * d + 8) + k * (d + 1) * 4) / (t_iter * 1e-3) / 1e9)
# the stress shape of the verdict: K = 4096, D = 64, 65,536 rows (segmented update)
n4, d4, k4 = 65536, 64, 4096
x4 = F.normalize(torch.randn(n4, d4, generator=g), dim=-1).cuda()
cent4 = x4[torch.randperm(n4, generator=g)[:k4].cuda()].clone()
a4 = ops.kmeans_assign(x4, cent4)
res["stress_k4096_d64_n65536"] = dict(
    assign_ms=ev_time(lambda: ops.kmeans_assign(x4, cent4)),
    assign_tensor_core_ms=ev_time(lambda: ops.kmeans_assign(x4, cent4, exact_diff_form=False)),
    accumulate_ms=ev_time(lambda: ops.kmeans_accumulate(x4, a4, k4)),
    finalize_ms=ev_time(lambda: ops.kmeans_finalize(*ops.kmeans_accumulate(x4, a4, k4)[:2], cent4.clone())))
# CPU oracle: one iteration of the reference algorithm ([N, K, D] difference form)
torch.set_num_threads(os.cpu_count() or 1)
cb = cent.cpu().clone()
OK.lloyd_update(x, cb, lambda: 0)
t0 = time.perf_counter()
for _ in range(3):
    OK.lloyd_update(x, cb, lambda: 0)
res["cpu_oracle_iteration_ms"] = (time.perf_counter() - t0) / 3 * 1e3
res["cpu_cores"] = os.cpu_count()
print(json.dumps(res))

# ---- uniqueness loss + p_unique_ids (modules/h_rqvae.py:41-105, 645-648) ----
for b in (128, 1024, 8192, 65536):
    gi = torch.Generator().manual_seed(b)
    ids = torch.randint(0, 16, (b, 3), generator=gi).cuda()      # many duplicates: 4096 distinct tuples
    f = F.normalize(torch.randn(b, 32, generator=gi), dim=-1).cuda().requires_grad_(True)

    def fwd_bwd():
        loss = ops.uniqueness_loss(ids, f, 0.0, 1.5)
        f.grad = None
        loss.backward()

    t = ev_time(fwd_bwd, reps=5)
    t_cnt = ev_time(lambda: ops.count_rows_with_later_twin(ids), reps=5)
    print(json.dumps(dict(case="uniqueness", batch=b, fwd_bwd_ms=t, p_unique_ms=t_cnt, pairs=b * (b - 1) // 2,
                          pair_rate_g_per_s=b * (b - 1) / 2 / (t * 1e-3) / 1e9)))

# ---- Gumbel-softmax level (modules/quantize.py:125-130): fused kernels vs the dense PyTorch formulas on the same GPU ----
from distributions.gumbel import gumbel_softmax_sample
for nb in (1024, 8192, 65536):
    gi = torch.Generator().manual_seed(nb)
    xg = F.normalize(torch.randn(nb, 32, generator=gi), dim=-1).cuda().requires_grad_(True)
    cbg = F.normalize(torch.rand(256, 32, generator=gi), dim=-1).cuda().requires_grad_(True)

    def fused():
        emb, _ids, loss = ops.gumbel_apply(xg, cbg, 0.2, 0.25, seed=7)
        xg.grad = cbg.grad = None
        (emb.sum() + loss.sum()).backward()

    def dense():
        dist = (xg ** 2).sum(dim=1, keepdim=True) + (cbg.T ** 2).sum(dim=0, keepdim=True) - 2 * xg @ cbg.T
        emb = gumbel_softmax_sample(-dist, 0.2, xg.device) @ cbg
        loss = ((xg.detach() - emb) ** 2).sum(-1) + 0.25 * ((xg - emb.detach()) ** 2).sum(-1)
        xg.grad = cbg.grad = None
        (emb.sum() + loss.sum()).backward()

    print(json.dumps(dict(case="gumbel_level_fwd_bwd", rows=nb, d=32, k=256, fused_ms=ev_time(fused, reps=5),
                          dense_torch_ms=ev_time(dense, reps=5))))
