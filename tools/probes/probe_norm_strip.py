"""Event-timed column-normalise/transpose/cast of a feature-major 16-bit bank (odd and even N), checked against torch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from bench import read_peaks
from summer_clip_b200 import ops

dev = torch.device("cuda")
hbm = read_peaks()["hbm_gbs"]
for n, d, dt in ((1281167, 1024, torch.float16), (1281168, 1024, torch.float16), (1281167, 768, torch.bfloat16),
                 (1281167, 512, torch.float16), (100003, 1023, torch.float16)):
    bank = torch.randn(d, n, device=dev, dtype=dt)
    out = ops.normalize_cast(bank, True)
    ref = (bank.float() / bank.float().norm(dim=0, keepdim=True)).t()
    err = (out[:n, :d].float() - ref).abs().max().item()
    pad_ok = bool((out[:n, d:] == 0).all()) and bool((out[n:] == 0).all())
    del ref
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        a.record()
        ops.normalize_cast(bank, True, out=out)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[5]
    byts = n * d * bank.element_size() + n * out.shape[1] * out.element_size()
    print(json.dumps({"n": n, "d": d, "dtype": str(dt), "ms": ms, "gbs": byts / ms / 1e6, "frac_of_hbm": byts / ms / 1e6 / hbm,
                      "max_err": err, "padding_zero": pad_ok}), flush=True)
    del bank, out
