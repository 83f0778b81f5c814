"""GPU check + timing of the fused encoder kernel (hv_encoder_forward) against a torch fp32 matmul chain on the device.
    python tools/check_encoder.py [rows]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hid-vae_b200"))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from hidvae_b200 import ops
from oracle import encoder as OE

torch.backends.cuda.matmul.allow_tf32 = False
dims = [768, 512, 256, 128, 32]
ws = [w.cuda() for w in OE.seeded_weights(dims, 2024)]
image = ops.encoder_pack(ws)
torch.cuda.synchronize()
print("packed", image.data.numel(), "bytes")
for n in [1, 127, 128, 129, 1000, 148 * 128 * 3 + 77]:
    g = torch.Generator().manual_seed(n)
    x = F.normalize(torch.randn(n, 768, generator=g), dim=-1).cuda()
    for norm in (True, False):
        for precise in (False, True):
            z = ops.encoder_forward(x, image, normalize=norm, precise_silu=precise)
            torch.cuda.synchronize()
            ref = OE.mlp_forward(x.double(), [w.double() for w in ws], norm).float()
            err = (z - ref).abs().max().item()
            scale = ref.abs().max().item()
            print(f"n={n:6d} norm={int(norm)} precise={int(precise)} max|err|={err:.3e} (max|ref| {scale:.3e}) rel {err / scale:.3e}", flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
x = F.normalize(torch.randn(n, 768, device="cuda"), dim=-1)
z = torch.empty(n, 32, device="cuda")
for precise in (False, True):
    for _ in range(3):
        ops.encoder_forward(x, image, normalize=True, precise_silu=precise, out=z)
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.encoder_forward(x, image, normalize=True, precise_silu=precise, out=z)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    med = ms[len(ms) // 2]
    flops = 2.0 * (768 * 512 + 512 * 256 + 256 * 128 + 128 * 32) * n
    print(f"TIMING rows={n} precise={int(precise)} median {med:.3f} ms min {ms[0]:.3f} ms  {n / med / 1e3:.1f} M items/s  {flops / med / 1e9:.1f} TFLOP/s  "
          f"x stream {n * 3072 / med / 1e6:.0f} GB/s")
# cuBLAS fp32 / tf32 chain for comparison
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    def chain():
        h = x
        for l, w in enumerate(ws):
            h = h @ w.t()
            if l < 3:
                h = F.silu(h)
        return F.normalize(h, dim=-1)
    for _ in range(2):
        chain()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        r = chain()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"TORCH tf32={int(tf32)} {dt * 1e3:.3f} ms  {n / dt / 1e6:.1f} M items/s   max|z - torch| {float((z - r).abs().max()):.3e}")
