"""Where the host time of HSemanticIdTokenizer.precompute_corpus_ids goes (cProfile + wall/GPU split)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench
import torch

dev = torch.device("cuda", 0)
tok = bench.make_tokenizer(dev)
n = 1 << 22
x = bench.synth_items(n, 1000, dev)
for _ in range(3):
    tok.precompute_corpus_ids(x)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    tok.precompute_corpus_ids(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host {1e3 * (t1 - t0):.2f} ms, host+drain {1e3 * (t2 - t0):.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    tok.precompute_corpus_ids(x)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
