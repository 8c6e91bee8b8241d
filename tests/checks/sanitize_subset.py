"""A reduced pass over every kernel family of the library for compute-sanitizer (memcheck / racecheck / synccheck):
small shapes, one launch each, results checked against fp32 torch so that a sanitizer-clean run is also a correct one.

    compute-sanitizer --tool memcheck python tests/checks/sanitize_subset.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import clip_search_oracle as orc
from summer_clip_b200 import build as _build, ops

_build.build_library()
dev = torch.device("cuda")
ok = True


def check(name, got, ref, tol=3e-3):
    global ok
    err = (got.float() - ref.float()).abs().max().item() / max(ref.float().abs().max().item(), 1e-30)
    good = err <= tol
    ok &= good
    print(f"{name}: rel_err={err:.2e} {'OK' if good else 'FAIL'}", flush=True)


banks = orc.synthetic_banks(300, 2900, 192, 33, seed=3, sigma=1.0, sigma_text=0.8, shared=3.0, dtype=torch.float16)
Q, K, L, T = (banks[n].to(dev) for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
Qn, Kn = ops.normalize_cast(Q, True), ops.normalize_cast(K, True)
A = Qn.float() @ Kn.float().t()
lab = L.float().argmax(1)
onehot = torch.nn.functional.one_hot(lab, 33).float()

# segmented kernel: 1 beta x 1 / 3 splits, 4 betas, e4m3 operands
bank = ops.hard_bank_build(ops.hard_labels(L, 33)[:2900], 33, K, True)
for splits in (1, 3):
    check(f"seg splits={splits}", ops.attn_fwd_hard(Qn, bank, 5.5, splits=splits), torch.exp(5.5 * (A - 1)) @ onehot)
for b, o in zip((0.5, 1.5, 5.5, 9.5), ops.attn_fwd_hard_multi(Qn, bank, [0.5, 1.5, 5.5, 9.5], splits=2)):
    check(f"seg 4-beta b={b}", o, torch.exp(b * (A - 1)) @ onehot)
q8, bank8 = ops.normalize_cast(Q, True, op_dtype=ops.E4M3), ops.hard_bank_build(ops.hard_labels(L, 33)[:2900], 33, K, True, op_dtype=ops.E4M3)
check("seg e4m3", ops.attn_fwd_hard(q8, bank8, 5.5), torch.exp(5.5 * (A - 1)) @ onehot, tol=0.2)
# softmax mode: online maximum in the segmented kernel, row-max pre-pass + shifted dense kernel
ref_sm = torch.softmax(100.0 * A, dim=1) @ onehot
O, m, l = ops.softmax_partials(ops.attn_softmax_hard(Qn, bank, 100.0, splits=2))
check("softmax seg", ops.merge_softmax(O[None], m[None], l[None])[0], ref_sm)
rm = ops.attn_rowmax(Qn, Kn, 2900)
check("rowmax", rm, A.amax(1), tol=1e-5)
Vt = ops.values_prepare(L, 33, softmax_scale=2.0, ones_row=True)
V = torch.softmax(2.0 * L.float(), dim=1)
o = ops.attn_fwd(Qn, Kn, Vt, 2900, 34, 100.0, splits=2, row_shift=rm)
check("softmax dense", o[:, :33] / o[:, 33:34], torch.softmax(100.0 * A, dim=1) @ V)
# dense kernel, Tip weights: 2 narrow slices (33 classes), 4 slices (1000 classes)
check("dense c=33", ops.attn_fwd(Qn, Kn, Vt, 2900, 33, 5.5, splits=2), torch.exp(5.5 * (A - 1)) @ V)
L2 = torch.randn(2900, 1000, device=dev)
Vt2 = ops.values_prepare(L2, 1000, softmax_scale=1.0)
check("dense c=1000", ops.attn_fwd(Qn, Kn, Vt2, 2900, 1000, 5.5), torch.exp(5.5 * (A - 1)) @ torch.softmax(L2, dim=1))
# split-fp16 GEMM, fused row scan
Z = ops.zero_shot_logits(Q, True, T.float(), scale=100.0)
check("gemm split", Z, 100.0 * torch.nn.functional.normalize(Q.float(), dim=0).t() @ T.float(), tol=1e-5)
conf, label = ops.rowconf_from_features(K, True, T.float())
Lk = torch.nn.functional.normalize(K.float(), dim=0).t() @ T.float()
check("rowconf fused", conf, Lk.amax(1), tol=1e-5)
ok &= bool((label.long() == Lk.argmax(1)).float().mean() > 0.99)
# HBM kernels: row scans, per-class top-k, epilogue, merges
cf, lb = ops.rowconf(L, scale=100.0, prob=True)
idx = ops.select_topk_per_label(cf, lb, 33, 4)
ok &= bool((idx.cpu().numpy() == orc.topk_prob_select(banks["cache_image_outs"], 4, scale=100.0)).all())
parts = ops.attn_fwd_hard(Qn, bank, 5.5, splits=3, merge=False)
res = ops.epilogue(Z, parts, [0.5, 1.0], labels=banks["test_labels"].to(dev), want_logits=True)
check("epilogue", res["logits"][1], Z + ops.merge_partials(parts))
check("merge peers", ops.merge_peer_parts([parts[0], parts[1], parts[2]]), parts.sum(0), tol=1e-6)
check("mean normalize", ops.mean_normalize_rows(K.t().contiguous()[None].repeat(2, 1, 1)), torch.nn.functional.normalize(K.float().t(), dim=1), tol=2e-3)
torch.cuda.synchronize()
print("SANITIZE SUBSET", "OK" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)
