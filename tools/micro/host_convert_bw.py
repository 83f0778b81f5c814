import torch, time, os
print("cpus", os.cpu_count(), "threads", torch.get_num_threads())
os.system("lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA node\\(s\\)|Flags' | cut -c1-300 | head -8")
n = 1 << 20
x = torch.empty(n, 768, dtype=torch.float32).pin_memory()
x.normal_()
y = torch.empty(n, 768, dtype=torch.float16).pin_memory()
for th in (4, 8, 16, 32):
    torch.set_num_threads(th)
    y.copy_(x)
    t0 = time.perf_counter()
    for _ in range(3):
        y.copy_(x)
    dt = (time.perf_counter() - t0) / 3
    print(f"threads {th}: fp32->fp16 {x.numel()*4/dt/1e9:.1f} GB/s read ({dt*1e3:.1f} ms per 1 Mi items)")
# H2D bandwidth fp32 vs fp16 pinned
xd = torch.empty_like(x, device="cuda"); yd = torch.empty_like(y, device="cuda")
for name, s, d in (("fp32", x, xd), ("fp16", y, yd)):
    d.copy_(s, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        d.copy_(s, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"H2D {name}: {s.numel()*s.element_size()/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms per 1 Mi items)")
# concurrent: convert while copying
import threading
def conv():
    for _ in range(3): y.copy_(x)
torch.set_num_threads(16)
th = threading.Thread(target=conv); t0 = time.perf_counter(); th.start()
for _ in range(6): yd.copy_(y, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter(); th.join(); t2 = time.perf_counter()
print(f"concurrent: 6 fp16 H2D in {1e3*(t1-t0):.1f} ms, 3 conversions done at {1e3*(t2-t0):.1f} ms")
