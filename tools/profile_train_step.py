"""Top GPU kernels of one HiD-VAE training step (KuaiRand-shaped model of bench.py's train_dp leg), eager, by total time.
    python tools/profile_train_step.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

from data.schemas import TaggedSeqBatch
from hidvae_b200 import dist as hv_dist
from modules.h_rqvae import HRqVae
from modules.quantize import QuantizeForwardMode

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
counts = [37, 168, 353]
torch.manual_seed(0)
model = HRqVae(input_dim=768, embed_dim=32, hidden_dims=[512, 256, 128], codebook_size=256, codebook_kmeans_init=False,
               codebook_normalize=True, codebook_mode=QuantizeForwardMode.ROTATION_TRICK, n_layers=3, n_cat_features=0,
               commitment_weight=0.5, tag_class_counts=counts, tag_embed_dim=768, sem_id_uniqueness_weight=0.5,
               sem_id_uniqueness_margin=0.5).to(dev).train()
n = 32768
g = torch.Generator(device=dev).manual_seed(7)
x = F.normalize(torch.randn(n, 768, generator=g, device=dev), dim=-1)
tags_emb = torch.randn(n, 3, 768, generator=g, device=dev)
tags_idx = torch.stack([torch.randint(0, c, (n,), generator=g, device=dev) for c in counts], dim=1)
grads = hv_dist.FlatGradAllReduce(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
torch.set_float32_matmul_precision("high")


def step():
    idx = torch.randint(0, n, (bs,), device=dev, generator=g)
    grads.zero()
    out = model(TaggedSeqBatch(None, None, None, x[idx], None, None, tags_emb[idx], tags_idx[idx]), gumbel_t=0.2)
    grads.backward(out.loss)
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
