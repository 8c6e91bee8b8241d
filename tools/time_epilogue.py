"""Epilogue / merge kernels alone at the ImageNet shape (50 000 x 1000 fp32 tiles): merged vs unmerged input,
1 or 20 alphas, with / without predictions.  One JSON line of median device times (ms)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from summer_clip_b200 import build as _build, ops  # noqa: E402
from tools.time_search_hp import timed  # noqa: E402

_build.build_library()
nq, c = 50000, 1000
g = torch.Generator(device="cuda").manual_seed(1)
Z = torch.randn(nq, c, generator=g, device="cuda") * 3
parts = torch.rand(3, nq, c, generator=g, device="cuda")
merged = ops.merge_partials(parts)
labels = torch.randint(0, c, (nq,), generator=g, device="cuda").int()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
al20 = [0.1 + 0.15 * i for i in range(20)]
for name, fn in {
    "merge_3parts": lambda: ops.merge_partials(parts),
    "epi_merged_1a_pred": lambda: ops.epilogue(Z, merged, [1.0], labels=labels),
    "epi_merged_1a_nopred": lambda: ops.epilogue(Z, merged, [1.0], labels=labels, want_pred=False),
    "epi_parts3_1a_pred": lambda: ops.epilogue(Z, parts, [1.0], labels=labels),
    "epi_parts3_1a_nopred": lambda: ops.epilogue(Z, parts, [1.0], labels=labels, want_pred=False),
    "epi_parts3_20a_nopred": lambda: ops.epilogue(Z, parts, al20, labels=labels, want_pred=False),
    "epi_merged_20a_pred": lambda: ops.epilogue(Z, merged, al20, labels=labels),
}.items():
    def run():
        flush.zero_()
        fn()
    base = timed(lambda: flush.zero_(), iters=7)
    out[name] = round(timed(run, iters=7) - base, 4)
print(json.dumps(out))
