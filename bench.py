#!/usr/bin/env python
"""bench.py -- RQ encode+train items/s of the HiD-VAE residual-quantisation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--no-sweep]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "c2"): the Amazon-Beauty-shaped catalogue, 12,101 items, D=32, K=256, L=3,
beta=0.4, ROTATION_TRICK (configs/h_rqvae_amazon.gin).  ONE STEP = one pass of the hot path over the catalogue:
    (1) training forward   hv_rq_forward (ids, emb_out, loss)           modules/quantize.py:100-154 x 3 levels
    (2) fused backward     hv_rq_backward (g_x, g_codebooks)            autograd of the same
    (3) eval encode        hv_rq_forward (ids only)                     modules/tokenizer/h_semids.py:127-130
        (independent of (1)-(2) once the operand image is packed: it runs on a side stream beside them)
`value` = items / second with inputs resident in HBM (C-ABI calls, CUDA events, L2 flushed between steps);
`e2e`   = the same pass through the public autograd API from PINNED HOST buffers: H2D of the step's inputs, the
          three kernels, D2H of ids + loss inside the timed region.
N > 1: every rank owns its own 12,101-item shard (weak scaling, the reference's per-rank batches) and the
codebook gradient [L, K, D] is all-reduced over NCCL, overlapped with the eval encode.
--impl reference times the oracle port (the reference's PyTorch-CPU algorithm) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "hid-vae_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "rq_encode_train_items_per_s"
UNIT = "items/s"
WORKLOAD = dict(workload="c2: Amazon-Beauty-shaped catalogue pass (train fwd + bwd + eval encode)", n_items=12101,
                embed_dim=32, codebook_size=256, n_levels=3, beta=0.4, forward_mode="ROTATION_TRICK",
                l2_between_steps="flushed (256 MiB write)")
MODE_ROT = 3


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tensor=float(p["bf16_tflops"]), tensor_sustained=float(p["bf16_tflops_sustained"]),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="B200_PROFILING.md fallback")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the three C2 kernels, taken from the committed
    `ncu --set full` capture of this same step (profiles/r01_traffic.json); {} when the file is missing."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f).get("c2_bytes_per_launch", {})
    return {}


def synth(n, d, k, n_levels, seed, device="cpu"):
    """unit-norm encoder outputs, uniform(0,1) codebooks with level 0 row-normalised and the
    later levels centred/scaled to the residual magnitude ("trained-like"), random upstream gradients."""
    g = torch.Generator().manual_seed(seed)
    x = F.normalize(torch.randn(n, d, generator=g), dim=-1)
    cbs = torch.rand(n_levels, k, d, generator=g)
    cbs[0] = F.normalize(cbs[0], dim=-1)
    for l in range(1, n_levels):
        cbs[l] = (cbs[l] - 0.5) * (0.7 * 0.5 ** l)
    g_emb = torch.randn(n_levels, n, d, generator=g) * 0.01
    g_loss = torch.full((n,), 1.0 / n)
    return x.to(device), cbs.to(device), g_emb.to(device), g_loss.to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t0 - 0.05 <= ts <= t1 + 0.05] or [r for (_, r) in self.rows]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])), smax.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------------
class NativeStep:
    """The three C-ABI calls of one step on device-resident inputs."""
    LAUNCHES_PER_STEP = 4  # pack image + train forward + backward + eval encode (the g_codebooks memset is torch's)

    def __init__(self, ops, x, cbs, g_emb, g_loss, beta):
        self.ops, self.x, self.cbs, self.g_emb, self.g_loss, self.beta = ops, x, cbs, g_emb, g_loss, beta
        self.side = torch.cuda.Stream()   # the eval encode depends only on x and the packed image: it runs beside
                                          # the training forward / backward (a fork/join the CUDA graph keeps)

    def train_fwd(self, packed):
        return self.ops.rq_forward(self.x, self.cbs, MODE_ROT, True, self.beta, want_emb=True, want_loss=True, packed=packed)

    def bwd(self, ids):
        return self.ops.rq_backward(self.x, self.cbs, ids, MODE_ROT, True, self.beta, self.g_emb, self.g_loss, None)

    def encode(self, packed):
        return self.ops.rq_encode(self.x, self.cbs, packed=packed)

    def __call__(self, comm=None):
        main = torch.cuda.current_stream()
        packed = self.ops.pack_codebooks(self.cbs)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            ids = self.encode(packed)
        out = self.train_fwd(packed)
        g_x, g_cb = self.bwd(out.ids)
        if comm is not None:
            comm.allreduce_async(g_cb, overlap=False)
            comm.wait()
        main.wait_stream(self.side)
        packed.record_stream(self.side)
        return out, g_x, g_cb, ids


class GradComm:
    """Codebook-gradient all-reduce.  Default: the one-kernel exchange over NVLink peer memory
    (hidvae_b200.dist.PeerAllReduce -> hv_peer_allreduce); HIDVAE_BENCH_NCCL=1 (or a node without symmetric-memory
    support) uses NCCL.  Either runs on a side stream so that it overlaps the eval encode and the step's D2H copies."""

    def __init__(self, numel, device):
        import torch.distributed as dist
        self.dist = dist
        self.stream = torch.cuda.Stream()
        self.peer, self.pending = None, False
        if os.environ.get("HIDVAE_BENCH_NCCL", "0") != "1":
            try:
                from hidvae_b200.dist import PeerAllReduce
                self.peer = PeerAllReduce(numel, device)
            except Exception as e:  # noqa: BLE001
                print(f"[bench] peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
        flag = torch.tensor([1 if self.peer is not None else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank must take the same path
        if int(flag.item()) == 0:
            self.peer = None
        self.kind = "peer-memory one-shot kernel (hv_peer_allreduce)" if self.peer is not None else "nccl"

    def allreduce_async(self, t, overlap=True):
        """overlap=True: on a side stream (beside the step's D2H copies); False: the peer kernel in line on the current
        stream (measured 2 us shorter for the device-resident step, where only the eval encode runs beside it)."""
        self.pending = overlap or self.peer is None
        if not self.pending:
            self.peer(t.view(-1))
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            if self.peer is not None:
                self.peer(t.view(-1))
            else:
                self.dist.all_reduce(t)
        t.record_stream(self.stream)

    def wait(self):
        if self.pending:
            torch.cuda.current_stream().wait_stream(self.stream)


class GraphedStep:
    """The step captured once into a CUDA graph and replayed: the C2 step is a handful of microsecond-scale
    launches, so the CPU-side launch path (ctypes + allocator) would otherwise be what is measured."""

    def __init__(self, fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out


def time_region(fn, steps, warmup, flush):
    """W warm-ups, then K steps each bracketed by CUDA events on the current stream, L2 flushed between steps."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs)  # ms


def run_native(args):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py (native arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        # a stuck collective must abort the run, not hang it (the watchdog raises after 3 minutes)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    from hidvae_b200 import ops

    w = WORKLOAD
    n, d, k, L, beta = w["n_items"], w["embed_dim"], w["codebook_size"], w["n_levels"], w["beta"]
    x, cbs, g_emb, g_loss = synth(n, d, k, L, seed=rank, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    step = NativeStep(ops, x, cbs, g_emb, g_loss, beta)
    comm = GradComm(cbs.numel(), cbs.device) if world > 1 else None
    pk = peaks()

    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eager = lambda: step(comm)
    run_step, graphed = eager, False
    if not args.no_graph:
        try:
            run_step, graphed = GraphedStep(eager), True
        except Exception as e:  # e.g. a collective that cannot be captured: fall back to eager launches
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            torch.cuda.synchronize()
    t_mark0 = time.time()
    total_ms = time_region(run_step, args.steps, args.warmup, flush)
    eager_ms = time_region(eager, min(args.steps, 100), 3, flush) / min(args.steps, 100) if graphed else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_mark1 = time.time()
    tm = torch.tensor([total_ms], device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public autograd API from pinned host memory (H2D + kernels + D2H per step) ----
    x_h = x.cpu().pin_memory()
    tgt_h = F.normalize(torch.randn(n, d, generator=torch.Generator().manual_seed(99)), dim=-1).pin_memory()
    ids_h = torch.empty((n, L), dtype=torch.int64).pin_memory()
    enc_h = torch.empty((n, L), dtype=torch.int64).pin_memory()
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()
    cb_param = cbs.clone().requires_grad_(True)

    side = torch.cuda.Stream()
    feed = torch.cuda.Stream()
    # Double-buffered input feed (what a DataLoader with pinned-memory prefetch does): the H2D copy of step i+1's inputs
    # runs on its own stream beside step i's kernels.  EVERY step still issues one H2D of its inputs' size and one D2H of
    # its results inside its timed region; `serial_ms_per_step` below is the same step with the copy in front of the kernels.
    x_dev = [torch.empty((n, d), device="cuda").requires_grad_(True) for _ in range(2)]
    tg_dev = [torch.empty((n, d), device="cuda") for _ in range(2)]

    def e2e_step(j=0, prefetch=False):
        main = torch.cuda.current_stream()
        cur, nxt = j & 1, (j & 1) ^ 1
        if prefetch:
            feed.wait_stream(main)                                          # (the previous step has released buffer `nxt`)
            with torch.cuda.stream(feed), torch.no_grad():
                x_dev[nxt].copy_(x_h, non_blocking=True)
                tg_dev[nxt].copy_(tgt_h, non_blocking=True)
        else:
            with torch.no_grad():
                x_dev[cur].copy_(x_h, non_blocking=True)
                tg_dev[cur].copy_(tgt_h, non_blocking=True)
        xd, tg = x_dev[cur], tg_dev[cur]
        packed = ops.pack_codebooks(cb_param.detach())
        side.wait_stream(main)
        with torch.cuda.stream(side):                                       # eval encode + its D2H beside the training step
            enc = ops.rq_encode(xd.detach(), cb_param.detach(), packed=packed)
            enc_h.copy_(enc, non_blocking=True)
        emb, _res, ids, loss, _ll = ops.RqFunction.apply(xd, cb_param, MODE_ROT, True, beta, "auto")
        total = ((emb.sum(0) - tg) ** 2).sum(-1).mean() + loss.mean()      # h_rqvae.py:607-640 shaped consumer
        cb_param.grad = None
        xd.grad = None
        total.backward()
        if comm is not None:
            comm.allreduce_async(cb_param.grad)
        ids_h.copy_(ids, non_blocking=True)
        loss_h.copy_(total.detach(), non_blocking=True)
        if comm is not None:
            comm.wait()
        main.wait_stream(side)
        if prefetch:
            main.wait_stream(feed)
        for t_ in (packed, enc):
            t_.record_stream(side)

    class Alternating:
        """Steps 0, 1, 0, 1, ... (one captured graph per input buffer)."""

        def __init__(self, even, odd):
            self.fns, self.j = (even, odd), 0

        def __call__(self):
            self.fns[self.j & 1]()
            self.j += 1

    if world > 1:
        dist.barrier()
    # the public API is graph-capturable (no host-side data-dependent branches, every call on the current stream):
    # a training loop with static shapes replays ONE graph per step -- pinned-host H2D, kernels, D2H included
    with torch.no_grad():
        x_dev[0].copy_(x_h)
        tg_dev[0].copy_(tgt_h)
    e2e_run, e2e_graphed = Alternating(lambda: e2e_step(0, True), lambda: e2e_step(1, True)), False
    e2e_serial = lambda: e2e_step(0, False)
    if not args.no_graph:
        try:
            e2e_run = Alternating(GraphedStep(lambda: e2e_step(0, True)), GraphedStep(lambda: e2e_step(1, True)))
            e2e_serial = GraphedStep(lambda: e2e_step(0, False))
            e2e_graphed = True
        except Exception as e:
            print(f"[bench] e2e CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager calls", file=sys.stderr)
            torch.cuda.synchronize()
    ser_steps = min(args.steps, 100)
    e2e_serial_ms = time_region(e2e_serial, ser_steps, 3, flush) / ser_steps
    with torch.no_grad():                                                   # buffer 0 holds the first step's inputs
        x_dev[0].copy_(x_h)
        tg_dev[0].copy_(tgt_h)
    torch.cuda.synchronize()
    e2e_ms = time_region(e2e_run, args.steps, args.warmup, flush)
    e2e_eager_ms = time_region(lambda: e2e_step(0, False), min(args.steps, 100), 3, flush) / min(args.steps, 100) if e2e_graphed else None
    te = torch.tensor([e2e_ms], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e_value = world * n * args.steps / (e2e_ms * 1e-3)
    h2d = x_h.numel() * 4 + tgt_h.numel() * 4
    d2h = ids_h.numel() * 8 + enc_h.numel() * 8 + 4

    # ---- per-kernel durations inside the same workload (CUDA events around each C-ABI call) -> roofline ----
    packed = ops.pack_codebooks(cbs)
    out = step.train_fwd(packed)
    ksteps = max(10, min(args.steps, 200))
    t_fwd = time_region(lambda: step.train_fwd(packed), ksteps, 3, flush) / ksteps
    t_bwd = time_region(lambda: step.bwd(out.ids), ksteps, 3, flush) / ksteps
    t_enc = time_region(lambda: step.encode(packed), ksteps, 3, flush) / ksteps
    flops = 2.0 * k * d * L * n                                  # L*2*K*D per item
    b_enc = n * (4 * d + 8 * L)
    b_fwd = n * (4 * d + 4 * d * L + 8 * L + 4)
    b_bwd = n * (4 * d + 8 * L + 4 * d * L + 4 + 4 * d) + 4 * L * k * d
    kernels = [
        dict(kernel="rq_fwd_tc_v11_kernel<rot> (train forward)", ms=t_fwd, bound="tensor", achieved=flops / (t_fwd * 1e-3) / 1e12,
             peak=pk["tensor"], unit="TFLOP/s", hbm_gbs=b_fwd / (t_fwd * 1e-3) / 1e9),
        dict(kernel="rq_bwd_kernel<32,rot> (backward)", ms=t_bwd, bound="hbm", achieved=b_bwd / (t_bwd * 1e-3) / 1e9, peak=pk["hbm"],
             unit="GB/s"),
        dict(kernel="rq_fwd_tc_v11_kernel (eval encode)", ms=t_enc, bound="tensor", achieved=flops / (t_enc * 1e-3) / 1e12,
             peak=pk["tensor"], unit="TFLOP/s", hbm_gbs=b_enc / (t_enc * 1e-3) / 1e9),
    ]
    traffic = ncu_traffic()
    for kk, key in zip(kernels, ("train_forward", "backward", "eval_encode")):
        kk["frac"] = kk["achieved"] / kk["peak"]
        kk["traffic"] = traffic.get(key)
    dom = max(kernels, key=lambda kk: kk["ms"])
    roofline = dict(bound=dom["bound"], achieved=dom["achieved"], peak=dom["peak"], unit=dom["unit"], frac=dom["frac"],
                    traffic=dom.get("traffic"), traffic_source="profiles/r01_traffic.json (ncu --set full, dram read + write per launch)",
                    kernel=dom["kernel"], kernel_ms=dom["ms"], peak_source=pk["source"] + ", burst",
                    kernels=kernels)

    sweep = None
    if rank == 0 and not args.no_sweep:
        sweep = run_sweep(ops, pk, flush)

    clocks = sampler.stop(t_mark0, t_mark1) if sampler else None
    cpu = None
    if rank == 0:
        cpu = cpu_baseline(n, d, k, L, beta, budget_s=12.0)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32 (argmin scores: bf16x3 split on tcgen05, fp32 accumulate)", data="synthetic",
                    config=dict(WORKLOAD, parallelism=f"dp{world}", items_per_step_per_gpu=n, cuda_graph=graphed,
                                eager_ms_per_step=eager_ms, grad_allreduce=(comm.kind if comm is not None else None)),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=e2e_ms / args.steps, cuda_graph=e2e_graphed, eager_ms_per_step=e2e_eager_ms,
                             input_feed="double-buffered: step i+1's pinned-host H2D on a copy stream beside step i's kernels; one H2D + one D2H per timed step",
                             serial_ms_per_step=e2e_serial_ms),
                    gpu_launches=(NativeStep.LAUNCHES_PER_STEP + (1 if comm is not None and comm.peer is not None else 0)) * args.steps, roofline=roofline, cpu_baseline=cpu,
                    clocks=clocks, impl="native")
        if sweep is not None:
            line["sweep"] = sweep
        OUT.emit(json.dumps(line))
    if world > 1:
        # No collective is pending (every rank passed the last all-reduce before rank 0 started its CPU baseline).
        # The captured graphs still hold NCCL kernels, and tearing the communicator down under them was seen to hang
        # (2 x B200, NCCL 2.28.9): drop the graphs, drain the device and leave without the NCCL teardown.
        del run_step, e2e_run
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_sweep(ops, pk, flush):
    """Supplementary single-kernel numbers on the other BASELINE.json shapes (not bench lines; parity cases whose
    roofline fraction is informative because the launch is long enough to fill the machine)."""
    res = []
    for name, (n, d, k, L) in {"c1_batch1024": (1024, 32, 256, 3), "c5_chunk_4Mi": (1 << 22, 32, 256, 3),
                               "c4_65536x64x4096x4": (65536, 64, 4096, 4)}.items():
        x, cbs, g_emb, g_loss = synth(n, d, k, L, seed=7, device="cuda")
        packed = ops.pack_codebooks(cbs)
        fl = None if n * d * 4 > (128 << 20) else flush
        fb = flush if fl is not None else torch.empty(1, dtype=torch.uint8, device="cuda")
        reps = 20
        t_enc = time_region(lambda: ops.rq_encode(x, cbs, packed=packed), reps, 3, fb) / reps
        ent = dict(case=name, n=n, d=d, k=k, L=L, encode_ms=t_enc, encode_items_per_s=n / (t_enc * 1e-3),
                   encode_tflops=2.0 * k * d * L * n / (t_enc * 1e-3) / 1e12)
        ent["encode_frac_of_bf16_peak"] = ent["encode_tflops"] / pk["tensor"]
        if name != "c4_65536x64x4096x4":
            out = ops.rq_forward(x, cbs, MODE_ROT, True, 0.4, want_emb=True, want_loss=True, packed=packed)
            t_f = time_region(lambda: ops.rq_forward(x, cbs, MODE_ROT, True, 0.4, want_emb=True, want_loss=True, packed=packed),
                              reps, 3, fb) / reps
            t_b = time_region(lambda: ops.rq_backward(x, cbs, out.ids, MODE_ROT, True, 0.4, g_emb, g_loss, None), reps, 3, fb) / reps
            byt = n * (4 * d * (3 + 2 * L) + 16 * L + 8)
            ent.update(train_fwd_ms=t_f, train_bwd_ms=t_b, train_items_per_s=n / ((t_f + t_b) * 1e-3),
                       train_hbm_gbs=byt / ((t_f + t_b) * 1e-3) / 1e9)
            ent["train_frac_of_hbm_peak"] = ent["train_hbm_gbs"] / pk["hbm"]
        res.append(ent)
        del x, cbs, g_emb, g_loss
    return res


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (the one place besides tests/smoke that may execute oracle/)
# ------------------------------------------------------------------------------------------------------------------
def oracle_step(O, x, cb_list, tgt, beta):
    xg = x.clone().requires_grad_(True)
    cbs = [c.clone().requires_grad_(True) for c in cb_list]
    out = O.rq_forward(xg, cbs, O.MODE_ROTATION_TRICK, beta, True)
    total = ((out.embeddings.sum(-1) - tgt) ** 2).sum(-1).mean() + out.quantize_loss.mean()
    total.backward()
    with torch.no_grad():
        enc = O.rq_forward(x, cb_list, O.MODE_ROTATION_TRICK, beta, False)
    return float(total.detach()), enc.sem_ids


def cpu_baseline(n, d, k, L, beta, budget_s):
    from oracle import rq as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, cbs, _, _ = synth(n, d, k, L, seed=0)
    tgt = F.normalize(torch.randn(n, d, generator=torch.Generator().manual_seed(99)), dim=-1)
    cb_list = [cbs[l] for l in range(L)]
    oracle_step(O, x, cb_list, tgt, beta)
    oracle_step(O, x, cb_list, tgt, beta)
    t0, reps = time.perf_counter(), 0
    while reps < 10 or (time.perf_counter() - t0 < budget_s and reps < 400):
        oracle_step(O, x, cb_list, tgt, beta)
        reps += 1
    dt = time.perf_counter() - t0
    return dict(value=n * reps / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"{reps} full steps of the same 12,101-item workload (oracle/rq.py, torch CPU, {cores} threads)")


def run_reference(args):
    """The reference's own CPU algorithm (oracle port) on this box's host cores; rank 0 only."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle import rq as O
    w = WORKLOAD
    n, d, k, L, beta = w["n_items"], w["embed_dim"], w["codebook_size"], w["n_levels"], w["beta"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, cbs, _, _ = synth(n, d, k, L, seed=0)
    tgt = F.normalize(torch.randn(n, d, generator=torch.Generator().manual_seed(99)), dim=-1)
    cb_list = [cbs[l] for l in range(L)]
    steps = min(args.steps, 200)   # bounded sample: every step is the full 12,101-item pass
    for _ in range(min(args.warmup, 5)):
        oracle_step(O, x, cb_list, tgt, beta)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(O, x, cb_list, tgt, beta)
    dt = time.perf_counter() - t0
    value = n * steps / dt
    sample = f"{steps} full steps of the 12,101-item workload (oracle/rq.py, torch {torch.__version__} CPU, {cores} threads)"
    OUT.emit(json.dumps(dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 5),
                          ms_per_step=dt / steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                          dtype="f32", data="synthetic", config=dict(WORKLOAD, parallelism="cpu"), impl="reference",
                          cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                          e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


class QuietStdout:
    """Everything but the final JSON line goes to stderr: libraries print to fd 1 on their own (NCCL's version banner),
    and the contract is ONE line on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a captured CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    with QuietStdout() as OUT:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_native(args)


if __name__ == "__main__":
    main()
