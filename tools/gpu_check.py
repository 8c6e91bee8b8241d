"""On-GPU bring-up checks for the CUDA path (run through gpurun).  Each case runs in its own
process so that a trapped kernel cannot poison the CUDA context of the next one.

    python tools/gpu_check.py all            # every case, each under `timeout`
    python tools/gpu_check.py attn 200 1000 1024 397 5.5 0
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _banks(Nq, Nk, D, C, seed=0, device="cuda"):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    protos = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1)
    yq = torch.randint(0, C, (Nq,), generator=g)
    yk = torch.randint(0, C, (Nk,), generator=g)
    sigma = 1.0 / (D ** 0.5)
    Q = protos[yq] + sigma * torch.randn(Nq, D, generator=g)
    K = protos[yk] + sigma * torch.randn(Nk, D, generator=g)
    return Q.to(device), K.to(device), yq.to(device), yk.to(device), protos.to(device)


def case_attn(Nq, Nk, D, C, beta, splits, identity_v=False, dt="fp16"):
    import torch
    from summer_clip_b200 import ops
    torch.manual_seed(0)
    ops.OP_DTYPE = torch.float16 if dt == "fp16" else torch.bfloat16
    Q, K, yq, yk, protos = _banks(Nq, Nk, D, C)
    Qn = ops.normalize_cast(Q, feature_major=False)
    Kn = ops.normalize_cast(K, feature_major=False)
    if identity_v:
        assert C == Nk
        lab = torch.arange(Nk, device="cuda", dtype=torch.int32)
        Vt = ops.values_prepare(None, C, labels=lab)
    else:
        L = (torch.nn.functional.normalize(K, dim=1) @ protos.t()).contiguous()
        Vt = ops.values_prepare(L, C, softmax_scale=100.0 * 0.1)
    torch.cuda.synchronize()
    # reference on the SAME bf16-rounded operands, fp32 math
    A = Qn.float()[:, :] @ Kn.float().t()
    W = torch.exp(beta * (A - 1.0))
    V = Vt.float()[:C, :Nk].t()
    O_ref = W @ V
    t0 = time.time()
    O = ops.attn_fwd(Qn, Kn, Vt, Nk, C, beta, splits=splits)
    torch.cuda.synchronize()
    dt = time.time() - t0
    err = (O - O_ref).abs()
    denom = O_ref.abs().max().item() + 1e-30
    # error against fp32 operands too (what the acceptance criteria see)
    A32 = torch.nn.functional.normalize(Q, dim=1) @ torch.nn.functional.normalize(K, dim=1).t()
    O32 = torch.exp(beta * (A32 - 1.0)) @ V
    e32 = (O - O32).abs().max().item() / (O32.abs().max().item() + 1e-30)
    print(f"attn[{dt}] Nq={Nq} Nk={Nk} D={D} C={C} beta={beta} splits={splits} identV={identity_v}: rel_vs_fp32={e32:.3e} "
          f"max_abs_err={err.max().item():.4e} rel_to_max={err.max().item() / denom:.3e} "
          f"ref_max={denom:.4e} t={dt * 1e3:.1f}ms")
    bad = err.max().item() / denom > 2e-2
    if bad:
        # localise: error by 32-row x 32-col blocks of the first 128 x 128 corner
        e = err[:128, :128]
        r = O_ref[:128, :128]
        nb_r, nb_c = (e.shape[0] + 31) // 32, (e.shape[1] + 31) // 32
        for i in range(nb_r):
            print("  blk", i, " ".join(f"{e[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32].max().item():9.2e}" for j in range(nb_c)))
        print("  O  [0,:8]  ", O[0, :8].tolist())
        print("  ref[0,:8]  ", O_ref[0, :8].tolist())
        print("  O  [1,:8]  ", O[1, :8].tolist())
        print("  ref[1,:8]  ", O_ref[1, :8].tolist())
        print("  O  [:8,0]  ", O[:8, 0].tolist())
        print("  ref[:8,0]  ", O_ref[:8, 0].tolist())
        if identity_v:
            # does O equal the reference under a permutation of columns within 8-blocks?
            o = O[:8, :64]
            rr = r[:8, :64]
            for row in range(2):
                match = [(rr[row] - o[row, j]).abs().argmin().item() for j in range(16)]
                print(f"  row {row}: O col j best matches ref col", match)
    return 1 if bad else 0


def case_attn_hard(Nq, Nk, D, C, beta, splits, dt="fp16"):
    """Hard-label kernel (values synthesised from int16 labels) against fp32 on the same rounded operands AND
    against the dense-Vt kernel on one_hot(labels)."""
    import torch
    from summer_clip_b200 import ops
    ops.OP_DTYPE = torch.float16 if dt == "fp16" else torch.bfloat16
    Q, K, yq, yk, protos = _banks(Nq, Nk, D, C)
    Qn = ops.normalize_cast(Q, feature_major=False)
    Kn = ops.normalize_cast(K, feature_major=False)
    L = (torch.nn.functional.normalize(K, dim=1) @ protos.t()).contiguous()
    lab16 = ops.hard_labels(L, C)
    ref_lab = L.argmax(1)
    lab_ok = bool((lab16[:Nk].long() == ref_lab).all()) and bool((lab16[Nk:] == -1).all())
    bank = ops.hard_bank_layout(lab16[:Nk], C).gather(Kn)
    Vt = ops.values_prepare(L, C)
    torch.cuda.synchronize()
    W = torch.exp(beta * (Qn.float() @ Kn.float().t() - 1.0))
    O_ref = W @ torch.nn.functional.one_hot(ref_lab, C).float()
    t0 = time.time()
    O = ops.attn_fwd_hard(Qn, bank, beta, splits=splits)
    torch.cuda.synchronize()
    el = time.time() - t0
    O_dense = ops.attn_fwd(Qn, Kn, Vt, Nk, C, beta, splits=splits)
    torch.cuda.synchronize()
    denom = O_ref.abs().max().item() + 1e-30
    e_ref = (O - O_ref).abs().max().item() / denom
    e_dense = (O - O_dense).abs().max().item() / denom
    bad = (not lab_ok) or e_ref > 2e-2 or e_dense > 2e-3
    print(f"attn_hard[{dt}] Nq={Nq} Nk={Nk} D={D} C={C} beta={beta} splits={splits}: labels={lab_ok} "
          f"rel_err_vs_fp32={e_ref:.3e} rel_diff_vs_dense_kernel={e_dense:.3e} t={el * 1e3:.1f}ms {'FAIL' if bad else 'OK'}")
    if bad:
        e = (O - O_ref).abs()
        print("  worst rows", e.max(1).values.topk(min(8, Nq)).indices.tolist(), "worst cols", e.max(0).values.topk(min(8, C)).indices.tolist())
        print("  O  [0,:8]", O[0, :8].tolist())
        print("  ref[0,:8]", O_ref[0, :8].tolist())
    return 1 if bad else 0


def case_norm():
    import torch
    from summer_clip_b200 import ops
    rc = 0
    for (D, N, dt, fm) in [(512, 300, torch.float16, True), (1024, 77, torch.float32, True), (768, 129, torch.float32, False),
                           (100, 50, torch.float16, False)]:
        g = torch.Generator().manual_seed(1)
        X = torch.randn((D, N) if fm else (N, D), generator=g).to(dt).cuda()
        idx = torch.randperm(N, generator=g)[: N // 2].cuda()
        for ix in (None, idx):
            out = ops.normalize_cast(X, feature_major=fm, idx=ix)
            Xf = X.float() if not fm else X.float().t()
            ref = torch.nn.functional.normalize(Xf, dim=1)
            if ix is not None:
                ref = ref[ix]
            e = (out[:, :D].float() - ref).abs().max().item()
            padz = out[:, D:].abs().max().item() if out.shape[1] > D else 0.0
            ok = e < 1e-2 and padz == 0.0
            rc |= 0 if ok else 1
            print(f"norm D={D} N={N} {dt} fm={fm} idx={'y' if ix is not None else 'n'}: max_err={e:.3e} pad={padz} {'OK' if ok else 'FAIL'}")
    return rc


def case_select():
    import torch
    from summer_clip_b200 import ops
    rc = 0
    for (N, C, k, dt) in [(5000, 37, 4, torch.float32), (20000, 397, 16, torch.float16), (3000, 1000, 8, torch.float32)]:
        g = torch.Generator().manual_seed(2)
        L = (torch.randn(N, C, generator=g) * 0.05).to(dt).cuda()
        for prob in (False, True):
            conf, label = ops.rowconf(L, scale=100.0 if prob else 1.0, prob=prob)
            Lf = L.float()
            if prob:
                ref_conf, ref_label = torch.softmax(Lf * 100.0, dim=1).max(dim=1)
            else:
                ref_conf, ref_label = Lf.max(dim=1)
            lab_ok = bool((label.long() == ref_label).all())
            cerr = (conf - ref_conf).abs().max().item()
            idx = ops.select_topk_per_label(conf, label, C, k)
            # reference selection on OUR conf/label (tests the top-k kernel in isolation)
            ref = []
            confc, labc = conf.cpu(), label.cpu().long()
            for c in labc.unique():
                m = (labc == c).nonzero().squeeze(1)
                order = sorted(m.tolist(), key=lambda i: (-confc[i].item(), i))[:k]
                ref.extend(order)
            sel_ok = idx.cpu().tolist() == ref
            ok = lab_ok and cerr < 1e-5 and sel_ok
            rc |= 0 if ok else 1
            print(f"select N={N} C={C} k={k} {dt} prob={prob}: labels={lab_ok} conf_err={cerr:.2e} topk={sel_ok} n_sel={idx.numel()} {'OK' if ok else 'FAIL'}")
    return rc


def case_misc():
    import torch
    from summer_clip_b200 import ops
    rc = 0
    g = torch.Generator().manual_seed(3)
    # values
    for (N, C) in [(300, 37), (1000, 397)]:
        L = (torch.randn(N, C, generator=g) * 0.05).cuda()
        Vt = ops.values_prepare(L, C)
        ref = torch.nn.functional.one_hot(L.argmax(1), C).float()
        e1 = (Vt[:C, :N].float().t() - ref).abs().max().item()
        Vs = ops.values_prepare(L, C, softmax_scale=100.0, ones_row=True)
        refs = torch.softmax(L * 100.0, dim=1)
        e2 = (Vs[:C, :N].float().t() - refs).abs().max().item()
        ones_ok = bool((Vs[C, :N] == 1).all()) and float(Vs[C, N:].abs().sum()) == 0.0
        pad_ok = float(Vt[C:].abs().sum()) == 0.0 and float(Vt[:, N:].abs().sum()) == 0.0
        ok = e1 == 0 and e2 < 4e-3 and ones_ok and pad_ok
        rc |= 0 if ok else 1
        print(f"values N={N} C={C}: hard_err={e1} soft_err={e2:.2e} ones={ones_ok} pad={pad_ok} {'OK' if ok else 'FAIL'}")
    # zero-shot logits + epilogue + merge
    D, N, C = 512, 333, 101
    X = torch.randn(D, N, generator=g).cuda()
    T = torch.nn.functional.normalize(torch.randn(D, C, generator=g), dim=0).cuda()
    Z = ops.zero_shot_logits(X, True, T)
    Zr = 100.0 * torch.nn.functional.normalize(X, dim=0).t() @ T
    ez = (Z - Zr).abs().max().item()
    Z2 = ops.zero_shot_logits(X.t().contiguous(), False, T)
    ez2 = (Z2 - Zr).abs().max().item()
    O = torch.rand(N, C, generator=g).cuda() * 3
    labels = torch.randint(0, C, (N,), generator=g).cuda()
    alphas = [0.0, 0.5, 2.0]
    res = ops.epilogue(Z, O, alphas, labels=labels, want_logits=True)
    ok = ez < 1e-3 and ez2 < 1e-3
    for i, a in enumerate(alphas):
        out = Z + O * a
        p = out.argmax(1)
        top5 = out.topk(5, dim=1).indices
        t1 = int((p == labels).sum())
        t5 = int((top5 == labels[:, None]).any(1).sum())
        ok &= bool((res["pred"][i].long() == p).all()) and int(res["top1"][i]) == t1 and int(res["top5"][i]) == t5
        ok &= bool((res["logits"][i] == out).all())
    parts = torch.randn(5, 77, 130, generator=g).cuda()
    em = (ops.merge_partials(parts) - parts.sum(0)).abs().max().item()
    ok &= em < 1e-5
    rc |= 0 if ok else 1
    print(f"zeroshot err={ez:.2e}/{ez2:.2e} epilogue+merge(err={em:.1e}) {'OK' if ok else 'FAIL'}")
    return rc


ATTN_HARD_CASES = [
    # Nq, Nk, D, C, beta, splits  (C > 256: the hard-label kernel needs 2 or 4k class slices)
    (128, 128, 64, 512, 5.5, 1),
    (128, 512, 128, 1000, 5.5, 1),
    (200, 1000, 1024, 397, 5.5, 1),
    (300, 5000, 512, 1000, 1.0, 3),
    (1000, 20000, 768, 1000, 11.5, 0),
    (130, 700, 192, 1000, 3.0, 1),
    (257, 300, 64, 37, 5.5, 1),
    (513, 4099, 320, 100, 2.0, 0),
    (64, 40000, 1024, 1000, 5.5, 0),
]

ATTN_CASES = [
    # Nq, Nk, D, C, beta, splits, identity_v
    (128, 128, 64, 128, 5.5, 1, True),
    (128, 128, 64, 16, 5.5, 1, False),
    (128, 256, 128, 256, 5.5, 1, False),
    (128, 384, 1024, 256, 5.5, 1, False),
    (200, 1000, 1024, 397, 5.5, 1, False),
    (300, 5000, 512, 1000, 1.0, 3, False),
    (1000, 20000, 768, 1000, 11.5, 0, False),
]


def main(argv):
    argv = argv or ["all"]
    if argv[0] == "all":
        rc = 0
        me = os.path.abspath(__file__)
        for name in ("norm", "select", "misc"):
            r = subprocess.run(["timeout", "300", sys.executable, me, name])
            rc |= r.returncode
        for c in ATTN_CASES:
            for dt in ("fp16", "bf16"):
                r = subprocess.run(["timeout", "120", sys.executable, me, "attn", *map(str, c[:6]), str(int(c[6])), dt])
                rc |= (r.returncode != 0)
        for c in ATTN_HARD_CASES:
            for dt in ("fp16", "bf16"):
                r = subprocess.run(["timeout", "120", sys.executable, me, "attnh", *map(str, c), dt])
                rc |= (r.returncode != 0)
        print("ALL", "OK" if rc == 0 else "FAIL")
        return rc
    if argv[0] == "hard":
        rc = 0
        me = os.path.abspath(__file__)
        for c in ATTN_HARD_CASES:
            for dt in ("fp16", "bf16"):
                r = subprocess.run(["timeout", "120", sys.executable, me, "attnh", *map(str, c), dt])
                rc |= (r.returncode != 0)
        print("HARD", "OK" if rc == 0 else "FAIL")
        return rc
    if argv[0] == "attnh":
        Nq, Nk, D, C = map(int, argv[1:5])
        return case_attn_hard(Nq, Nk, D, C, float(argv[5]), int(argv[6]), argv[7] if len(argv) > 7 else "fp16")
    if argv[0] == "attn":
        Nq, Nk, D, C = map(int, argv[1:5])
        return case_attn(Nq, Nk, D, C, float(argv[5]), int(argv[6]), bool(int(argv[7])) if len(argv) > 7 else False,
                         argv[8] if len(argv) > 8 else "fp16")
    return {"norm": case_norm, "select": case_select, "misc": case_misc}[argv[0]]()


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
