"""Probe: why is the segmented kernel slower for 1 query than for 64?  Times the kernel for a few batch sizes, with the
query rows given as they are (the TMA box of 128 rows is mostly out of bounds) and zero-padded to 256 rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from summer_clip_b200 import build as _b, ops
_b.build_library()
dev = torch.device("cuda")
n, dim, c = 1281167, 1024, 1000
g = torch.Generator(device=dev).manual_seed(1)
Kn = torch.empty((n, dim), dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 17):
    e = min(n, s + (1 << 17))
    ops.normalize_cast(torch.randn(e - s, dim, generator=g, device=dev), False, out=Kn[s:e])
labels = torch.randint(0, c, (n,), generator=g, device=dev).int()
bank = ops.hard_bank_layout(labels, c).gather(Kn)
del Kn


def timeit(Qn, splits):
    for _ in range(3):
        o = ops.attn_fwd_hard(Qn, bank, 5.5, splits=splits, merge=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        o = ops.attn_fwd_hard(Qn, bank, 5.5, splits=splits, merge=False)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


for b in (1, 8, 64, 128, 256):
    q = torch.randn(b, dim, generator=g, device=dev)
    Qn = ops.normalize_cast(q, False)
    pad = torch.zeros((256, dim), dtype=torch.float16, device=dev)
    pad[:b] = Qn
    print(f"batch={b}: as given {timeit(Qn, 73):.3f} ms | padded to 256 rows {timeit(pad, 73):.3f} ms | 74 splits {timeit(Qn, 74):.3f} ms", flush=True)
