"""Event-timed column-normalise + transpose + cast of the 1.28M x 1024 fp16 key bank (sc_normalize_cast)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from summer_clip_b200 import ops
n, dim = 1281167, 1024
bank = torch.randn(dim, n, device="cuda", dtype=torch.float16)
out = torch.empty((n, dim), dtype=torch.float16, device="cuda")
for _ in range(3): ops.normalize_cast(bank, True, out=out)
torch.cuda.synchronize()
evs=[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for a,b in evs:
    a.record(); ops.normalize_cast(bank, True, out=out); b.record()
torch.cuda.synchronize()
ts=sorted(a.elapsed_time(b) for a,b in evs)
print("k_norm ms median %.3f best %.3f -> %.2f TB/s" % (ts[5], ts[0], n*dim*4/ts[0]/1e9))
ref = torch.nn.functional.normalize(bank[:, :4096].float(), dim=0).t()
print("max err", (out[:4096].float()-ref).abs().max().item())
