"""One launch of the key-bank normalise kernel at the headline shape (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from summer_clip_b200 import ops

bank = torch.randn(1024, 1281167, device="cuda", dtype=torch.float16)
out = ops.normalize_cast(bank, True)
ops.normalize_cast(bank, True, out=out)
torch.cuda.synchronize()
