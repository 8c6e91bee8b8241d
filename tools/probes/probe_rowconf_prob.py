"""Two PROB row scans of a 1.28M x 1000 fp16 logits bank (for ncu: -k regex:rowconf_reg -s 1 -c 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from summer_clip_b200 import ops

g = torch.Generator(device="cuda").manual_seed(4)
L = (0.25 + 0.02 * torch.randn(1281167, 1000, generator=g, device="cuda")).half()
for _ in range(2):
    ops.rowconf(L, scale=100.00000762939453, prob=True)
torch.cuda.synchronize()
