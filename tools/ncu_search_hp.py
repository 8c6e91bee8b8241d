"""Launch list helper: one 1-beta and one 4-beta attention + epilogue at the ImageNet 16-shot Tip-Adapter shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from summer_clip_b200 import build as _build, ops
from summer_clip_b200.tip_adapter import utils as tip_utils
_build.build_library()
dev = torch.device("cuda")
nq, nk, dim, c = 50000, 16000, 1024, 1000
g = torch.Generator(device=dev).manual_seed(2)
protos = torch.nn.functional.normalize(torch.randn(c, dim, generator=g, device=dev), dim=1)
yk = torch.arange(nk, device=dev) % c
keys = torch.nn.functional.normalize(protos[yk] + torch.randn(nk, dim, generator=g, device=dev) / dim ** 0.5, dim=1).half().t()
vals = torch.nn.functional.one_hot(yk, c).half()
yq = torch.randint(0, c, (nq,), generator=g, device=dev)
feats = torch.nn.functional.normalize(protos[yq] + torch.randn(nq, dim, generator=g, device=dev) / dim ** 0.5, dim=1).half()
clip_w = torch.nn.functional.normalize(protos, dim=1).t().contiguous().half()
head = tip_utils.TipAdapterHead(keys, vals, feats, clip_w)
print("splits", ops.attn_hard_splits(nq, head.values.hard_bank(head.k).n_sorted, dev))
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("one_beta")
o = head.cache_logits(5.5)
torch.cuda.nvtx.range_pop()
torch.cuda.nvtx.range_push("four_beta")
os_ = head.cache_logits_many([0.5, 1.5, 3.5, 5.5])
torch.cuda.nvtx.range_pop()
r = ops.epilogue(head.clip_logits, o, [0.1 + 0.15 * i for i in range(20)], labels=yq.int(), want_pred=False)
torch.cuda.synchronize()
print("done")
