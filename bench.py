"""Benchmark of the CLIP-search hot path (BASELINE.json metric: queries/sec on the ImageNet-shaped
search, 50k queries x 1.28M keys x 1024-d, 1000 classes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full pass of the hot path over the query bank: normalise/cast the queries, zero-shot
logits, fused attention against the resident key bank, cross-split/rank merge, alpha epilogue with
accuracy counters.  The key bank (normalised K-major keys + transposed one-hot values) is built once
before the timed region, like the reference's loaded caches; its build time is reported separately.
With N > 1 the KEY bank is sharded across ranks (strong scaling): one NCCL reduce-scatter sums the
partial O tiles and leaves each rank the rows of its query slice, which it finishes alone (zero-shot
logits, epilogue); predictions are all-gathered, counters all-reduced.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

LATENCY_BATCHES = (1, 8, 64, 256, 1024, 4096)

WORKLOADS = {
    # name: (Nq, Nk, D, C, description)
    "imagenet_rn50": (50000, 1281167, 1024, 1000, "cfg3 ImageNet CLIP-search RN50: 50k val x 1.28M train keys x 1024-d, 1000 classes, hard values, beta=5.5, alpha=1"),
    "imagenet_vitl14": (50000, 1281167, 768, 1000, "cfg4 ImageNet CLIP-search ViT-L/14: 768-d"),
    "tip_imagenet_16shot": (50000, 16000, 1024, 1000, "cfg2 Tip-Adapter ImageNet 16-shot cache head"),
    "sun397": (19850, 19850, 1024, 397, "cfg1 SUN397-shaped image attention"),
    "tiny": (2048, 16384, 256, 100, "smoke-sized"),
    "latency": (4096, 1281167, 1024, 1000, "cfg5 online classification latency: query batch 1-4096 against the 1.28M-key RN50 bank"),
}
BETA, ALPHA = 5.5, 1.0


def read_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_tflops": float(p["bf16_tflops"]), "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", 0.0)),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    except Exception:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (NVML, 100 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.power = index, [], set(), None, []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        pw = sorted(self.power)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w": round(pw[len(pw) // 2], 1) if pw else None}


def make_banks(torch, nq, nk_lo, nk_hi, dim, n_classes, seed, device):
    """Synthetic clustered banks in the reference's on-disk layout (noise levels chosen so that the
    zero-shot top-1 is ~65 % and the cache lifts it, i.e. argmax is a meaningful check): fp16 feature-major
    [D, N] image features (save_features.py:36), fp16 logits bank L = K_norm^T T [N, C]
    (save_image_outs.py:25).  Only keys [nk_lo, nk_hi) are generated (this rank's shard)."""
    g = torch.Generator(device=device).manual_seed(seed)
    u0 = torch.nn.functional.normalize(torch.randn(dim, generator=g, device=device), dim=0)
    protos = torch.nn.functional.normalize(u0 + torch.randn(n_classes, dim, generator=g, device=device) / dim ** 0.5, dim=1)
    text = torch.nn.functional.normalize(protos + 1.5 / dim ** 0.5 * torch.randn(n_classes, dim, generator=g, device=device), dim=1)
    yq = torch.randint(0, n_classes, (nq,), generator=g, device=device)
    q = (protos[yq] + 2.5 * torch.randn(nq, dim, generator=g, device=device) / dim ** 0.5)
    q_bank = q.t().contiguous().half()
    n_local = nk_hi - nk_lo
    k_bank = torch.empty((dim, n_local), dtype=torch.float16, device=device)
    outs = torch.empty((n_local, n_classes), dtype=torch.float16, device=device)
    gk = torch.Generator(device=device)
    step = 1 << 16
    # keys are generated in GLOBAL 65536-key chunks seeded by the chunk index, and a shard keeps the part of every
    # chunk it overlaps: the bank (hence top1_count) is the same whatever the number of ranks
    for c in range(nk_lo // step, -(-nk_hi // step)):
        gk.manual_seed(seed * 1000003 + c)
        yk = torch.randint(0, n_classes, (step,), generator=gk, device=device)
        x = protos[yk] + torch.randn(step, dim, generator=gk, device=device) / dim ** 0.5
        a, b = max(nk_lo, c * step), min(nk_hi, (c + 1) * step)
        x = x[a - c * step: b - c * step]
        k_bank[:, a - nk_lo: b - nk_lo] = x.t().half()
        outs[a - nk_lo: b - nk_lo] = (torch.nn.functional.normalize(x, dim=1) @ text.t()).half()
    return q_bank, k_bank, outs, text.t().contiguous(), yq.int()


def parity_check(torch, ops, searcher, q_bank, k_bank, outs, text, labels, pred_timed, rows, soft_scale, key_range=None):
    """Outside the timed region: `rows` sampled query rows of the step just timed against fp32 torch arithmetic on
    the RAW banks (the reference's expressions, cache_weights_strategy.py:18-36, cache_value_strategy.py:14-28,
    image_attention.py:109-111, chunked over the keys).  key_range = (lo, hi) restricts the reference to one key
    shard (N > 1, key-sharded: the rank-local partial tile is what is checked)."""
    dev = q_bank.device
    n_classes = text.shape[1]
    q = q_bank[:, rows].float()
    qn = q / q.norm(dim=0, keepdim=True)
    z = 100.0 * qn.t() @ text.float()
    o_ref = torch.zeros((rows.numel(), n_classes), dtype=torch.float32, device=dev)
    nk = k_bank.shape[1]
    for s0 in range(0, nk, 1 << 16):
        k = k_bank[:, s0:s0 + (1 << 16)].float()
        kn = k / k.norm(dim=0, keepdim=True)
        w = (-1 * BETA * (1 - qn.t() @ kn)).exp()
        l = outs[s0:s0 + (1 << 16)].float()
        v = torch.nn.functional.one_hot(l.max(dim=1)[1], n_classes).float() if soft_scale is None else torch.softmax(soft_scale * l, dim=1)
        o_ref += w @ v
    # the library's result for the same rows (its own normalise / attention kernels, resident bank)
    qn_lib = ops.normalize_cast(q_bank[:, rows].contiguous(), True)
    o_lib = searcher.local_cache_logits(qn_lib, BETA)
    rel = ((o_lib - o_ref).abs().max() / o_ref.abs().max()).item()
    out = {"rows": int(rows.numel()), "cache_logits_max_rel_err": rel, "reference": "fp32 torch on the raw banks, same device"}
    if key_range is None:
        ref = z + ALPHA * o_ref
        z_lib = ops.zero_shot_logits(q_bank[:, rows].contiguous(), True, searcher.text, t_split=searcher.text_split)
        got = z_lib + ALPHA * o_lib
        out["softmax_maxabs"] = (torch.softmax(got, 1) - torch.softmax(ref, 1)).abs().max().item()
        out["argmax_agree"] = (got.argmax(1) == ref.argmax(1)).float().mean().item()
        out["timed_step_pred_agree"] = (pred_timed[rows].long() == ref.argmax(1)).float().mean().item()
        out["ok"] = bool(out["softmax_maxabs"] <= 2e-3 and out["argmax_agree"] >= 1.0 - 1.0 / rows.numel()
                         and out["timed_step_pred_agree"] >= 1.0 - 1.0 / rows.numel())
    else:
        out["key_shard"] = list(key_range)
        out["ok"] = bool(rel <= 2e-3)
    return out


def gpu_eager_baseline(torch, q_bank, k_bank, outs, text, labels, soft_scale, chunk=1024, steps=2):
    """The reference's own tensor expressions in torch eager on the SAME B200 (cuBLAS hgemm + ATen element-wise
    kernels, fp16 like the reference's cached CUDA banks), query-chunked because the [Nq, Nk] matrix it materialises
    does not fit: cache_weights_strategy.py:18-21,33-36, cache_value_strategy.py:14-17 / :26-28,
    image_attention.py:80-83,109-111, clip_searcher/utils.py:15-21.  The honest same-box bar."""
    nq = q_bank.shape[1]
    t_half = text.half()

    def step():
        qn = q_bank / q_bank.norm(dim=0, keepdim=True)
        kn = k_bank / k_bank.norm(dim=0, keepdim=True)
        if soft_scale is None:
            _, ids = outs.max(dim=1)
            v = torch.nn.functional.one_hot(ids, num_classes=outs.shape[1]).half()
        else:
            v = torch.softmax(soft_scale * outs, dim=1)
        z = 100.0 * qn.t() @ t_half
        correct = torch.zeros((), dtype=torch.int64, device=q_bank.device)
        for s0 in range(0, nq, chunk):
            w = (-1 * BETA * (1 - qn[:, s0:s0 + chunk].t() @ kn)).exp()
            o = w @ v
            out = z[s0:s0 + chunk] + o * ALPHA
            pred = out.topk(5, 1, True, True)[1]
            correct += (pred[:, 0] == labels[s0:s0 + chunk]).sum()
        return correct

    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        top1 = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "steps": steps, "query_chunk": chunk,
            "dtype": "f16", "top1_count": int(top1),
            "kind": "reference tensor expressions in torch eager on this GPU (cuBLAS + ATen), [chunk, Nk] weights materialised per query chunk"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from summer_clip_b200 import build as _build, ops
    from summer_clip_b200.searcher import ClipSearcher, query_slice, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CLIP-search path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    if rank == 0:
        _build.build_library()
    if world > 1:
        dist.barrier()
    if args.op_dtype:
        ops.OP_DTYPE = {"bf16": torch.bfloat16, "fp16": torch.float16, "e4m3": ops.E4M3}[args.op_dtype]

    nq, nk, dim, n_classes, desc = WORKLOADS[args.workload]
    if args.nq:
        nq = args.nq
    shard_keys = world > 1 and args.shard == "keys"
    shard_queries = world > 1 and args.shard == "queries"
    lo, hi = shard_range(nk, rank, world) if shard_keys else (0, nk)
    q_bank, k_bank, outs, text, labels = make_banks(torch, nq, lo, hi, dim, n_classes, seed=3, device=device)
    soft_scale = 100.00000762939453 * 0.1 if args.values == "softmax" else None     # conf/cache_value_strategy/softmax_cache.yaml

    searcher = ClipSearcher(device, group=group, shard=args.shard)
    searcher.set_text(text)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # normalise+transpose+cast keys; label-sorted bank / dense values.  Key shards are generated per rank (local_shard)
    searcher.set_cache(k_bank, outs, softmax_scale=soft_scale, local_shard=shard_keys)
    torch.cuda.synchronize()
    bank_build_first_ms = (time.perf_counter() - t0) * 1e3      # first call: library load, allocator growth included
    t0 = time.perf_counter()
    searcher.set_cache(k_bank, outs, softmax_scale=soft_scale, local_shard=shard_keys)
    torch.cuda.synchronize()
    bank_build_ms = (time.perf_counter() - t0) * 1e3            # steady state
    n_local = hi - lo
    c_pad = ops.pad_classes(n_classes)
    # the queries this rank runs attention for: all of them (one GPU, key shards) or its slice (query shards)
    qlo, qhi = query_slice(nq, rank, world) if shard_queries else (0, nq)
    nq_attn = qhi - qlo
    blocks = (int(os.environ["SC_BENCH_BLOCKS"]) if os.environ.get("SC_BENCH_BLOCKS") else None) if shard_keys else None
    nq_launch = nq_attn // (blocks or 4) if shard_keys else nq_attn        # queries per attention launch
    splits = ops.attn_hard_splits(nq_launch, searcher.hard_bank.n_sorted, device, bank=searcher.hard_bank) if searcher.hard_bank is not None \
        else ops.attn_splits(nq_launch, n_local, c_pad, device, D_pad=ops.pad_dim(dim))

    q_host = q_bank.cpu().pin_memory()
    labels_host = labels.cpu().pin_memory()
    labels_dev = labels
    stream = torch.cuda.current_stream()

    def step_device():
        """Inputs resident in HBM: one call of the public API, ClipSearcher.search."""
        r = searcher.search(q_bank, [BETA], [ALPHA], labels=labels_dev, want_cache_logits=False, blocks=blocks)[0]
        return {"pred": r["pred"], "top1": r["top1"], "top5": r["top5"]}

    # e2e: queries and labels come from pinned host memory every step and the predictions + counters go back.  Every
    # rank copies only ITS query slice over its own PCIe link (1/N of the bank); key-sharded ranks, which need every
    # query, all-gather the normalised fp16 rows over NVLink inside search(query_shard=True).  The host->device copy
    # of step i+1 is issued on a side stream while step i computes (two device buffers), as a serving loop would; the
    # first copy of the timed region is fully exposed.
    copy_stream = torch.cuda.Stream(device=device)
    elo, ehi = query_slice(nq, rank, world)
    q_host_mine = q_host[:, elo:ehi].contiguous().pin_memory() if world > 1 else q_host
    lab_host_mine = labels_host[elo:ehi].contiguous().pin_memory() if world > 1 else labels_host
    q_bufs = [torch.empty_like(q_host_mine, device=device) for _ in range(2)]
    lab_bufs = [torch.empty_like(lab_host_mine, device=device) for _ in range(2)]
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]

    computed = [None, None]                        # per input buffer: the event after the search that read it last

    def start_copy(i):
        with torch.cuda.stream(copy_stream):
            if computed[i % 2] is not None:
                copy_stream.wait_event(computed[i % 2])      # the step that read this buffer last has finished
            q_bufs[i % 2].copy_(q_host_mine, non_blocking=True)
            lab_bufs[i % 2].copy_(lab_host_mine, non_blocking=True)
            copy_done[i % 2].record(copy_stream)

    # results leave the same way: the device->host copy of step i's predictions and counters runs on a second copy
    # stream into double-buffered pinned host buffers and is awaited one step later, under step i + 1's compute
    out_stream = torch.cuda.Stream(device=device)
    pred_host = [torch.empty((1, nq), dtype=torch.int32).pin_memory() for _ in range(2)]
    cnt_host = [torch.empty((2, 1), dtype=torch.int32).pin_memory() for _ in range(2)]
    out_done = [torch.cuda.Event(), torch.cuda.Event()]
    out_keep = [None, None]

    def step_e2e(i, n_steps):
        """Host buffers in, host result out: H2D of the query slice, D2H of predictions + counters, both double
        buffered against the compute of the neighbouring steps.  Returns the host tensors of step i - 1 (complete)."""
        if i == 0:
            start_copy(0)
        if i + 1 < n_steps:
            start_copy(i + 1)                       # buffer (i+1) % 2 was last read by step i-1: the copy waits for it
        torch.cuda.current_stream().wait_event(copy_done[i % 2])
        r = searcher.search(q_bufs[i % 2], [BETA], [ALPHA], labels=lab_bufs[i % 2], want_cache_logits=False,
                            query_shard=nq if world > 1 else False, blocks=blocks)[0]
        computed[i % 2] = torch.cuda.Event()
        computed[i % 2].record(torch.cuda.current_stream())
        with torch.cuda.stream(out_stream):
            out_stream.wait_event(computed[i % 2])
            pred_host[i % 2].copy_(r["pred"], non_blocking=True)
            cnt_host[i % 2].copy_(torch.stack([r["top1"], r["top5"]]), non_blocking=True)
            out_done[i % 2].record(out_stream)
        out_keep[i % 2] = r                         # keep the device tensors alive until their copy has been awaited
        if i > 0:
            out_done[(i - 1) % 2].synchronize()     # step i-1's results are on the host now
        if i + 1 == n_steps:
            out_done[i % 2].synchronize()
        return pred_host[i % 2], cnt_host[i % 2]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing
    for _ in range(args.warmup):
        res = step_device()
    sync_all()
    searcher.gpu_launches = 0
    searcher.events = []                                   # (name, start, end) CUDA events on the launching streams
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0.record(stream)
    for i in range(args.steps):
        res = step_device()
    ev1.record(stream)
    sync_all()
    clocks = sampler.stop()
    events, searcher.events = searcher.events, None
    if os.environ.get("SC_ATTN_CLKPROBE"):                 # experiments builds: clock seen by the attention CTAs
        try:
            import ctypes
            from summer_clip_b200 import _lib
            fn = getattr(_lib.load(), "sc_debug_attn_hard_clock_mhz" if searcher.hard_bank is not None else "sc_debug_attn_clock_mhz")
            fn.restype = ctypes.c_double
            clocks["attn_cta_mhz"] = round(float(fn()), 1)
        except Exception as exc:  # noqa: BLE001
            clocks["attn_cta_mhz"] = str(exc)
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    phase_ms = {}
    for name, e0, e1 in events:
        phase_ms[name] = phase_ms.get(name, 0.0) + e0.elapsed_time(e1) / args.steps
    attn_ms = phase_ms["attention"]                        # per step: the sum over the step's attention launches
    attn_ms_ranks = [attn_ms]
    if world > 1:                                          # the step ends with the slowest rank: report the spread
        t = torch.zeros(world, dtype=torch.float64, device=device)
        t[rank] = attn_ms
        dist.all_reduce(t)
        attn_ms_ranks = [round(v, 3) for v in t.tolist()]
    n_attn_launches = sum(1 for name, _, _ in events if name == "attention") // args.steps
    gpu_launches = searcher.gpu_launches
    top1 = int(res["top1"][0])
    pred_timed = res["pred"][0].clone()

    # ---------------- end-to-end timing (host buffers)
    n_warm = max(1, args.warmup // 2)
    for i in range(n_warm):
        step_e2e(i, n_warm)
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pred, counts = step_e2e(i, args.steps)
    sync_all()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_top1 = int(counts[0][0])
    phases = {k: round(v, 4) for k, v in phase_ms.items()} if args.phases else None

    # ---------------- output check of the step just timed (outside every timed region)
    check = None
    if not args.no_parity_check:
        g = torch.Generator(device=device).manual_seed(1234)
        rows = torch.randperm(nq_attn, generator=g, device=device)[:64].sort()[0] + qlo
        check = parity_check(torch, ops, searcher, q_bank, k_bank, outs, text, labels_dev, pred_timed, rows, soft_scale,
                             key_range=(lo, hi) if shard_keys else None)
        check["e2e_top1_equals_device_top1"] = bool(e2e_top1 == top1)
        if world > 1:
            ok = torch.tensor([int(check["ok"])], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            check["ok_all_ranks"] = bool(ok.item())
    eager = None
    if world == 1 and not args.no_eager_baseline:
        try:
            eager = gpu_eager_baseline(torch, q_bank, k_bank, outs, text, labels_dev, soft_scale)
        except Exception as exc:  # noqa: BLE001  (an OOM of the baseline must not lose the measurement)
            eager = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    del k_bank, outs

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = read_peaks()
    ms_per_step = total_ms / args.steps
    qps = nq / (ms_per_step * 1e-3)
    flops = 2.0 * nq_attn * n_local * (dim + n_classes)              # 2*Nq*Nk*(D + C) per launch
    if searcher.hard_bank is not None:
        # label-sorted one-hot bank: GEMM-1 (Q.K^T) on every (padded) key on the tensor cores; W @ one_hot is a
        # per-class segmented sum done in fp32 by the exp warps (no second GEMM exists to execute)
        executed = 2.0 * nq_attn * searcher.hard_bank.n_sorted * dim
    else:
        executed = 2.0 * nq_attn * n_local * (dim + c_pad)
    achieved = executed / (attn_ms * 1e-3) / 1e12                   # tensor-core work actually issued
    dense_equiv = flops / (attn_ms * 1e-3) / 1e12
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "attn_traffic.json")) as f:
            traffic = json.load(f).get((args.workload if args.values == "hard" else args.workload + "_softmax_values") if world == 1 else "", None)
    except Exception:
        pass
    dense = searcher.hard_bank is None
    if dense:
        # dense values: both GEMMs are real tensor-core work; SURVEY §8d's count is the roofline numerator
        achieved = dense_equiv
    sharding = "single GPU"
    if shard_keys:
        sharding = f"key-sharded x{world}: reduce-scatter of the partial tiles, each rank finishes its query slice"
    elif shard_queries:
        sharding = f"query-sharded x{world}: every rank holds the whole bank and scores its query slice; no data-path collective"
    out = {
        "metric": "clip_search_queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": {torch.float16: "f16", torch.bfloat16: "bf16"}.get(ops.OP_DTYPE, "e4m3 (opt-in reduced precision; NOT the headline)"), "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "n_queries": nq, "n_keys": nk, "dim": dim,
                   "n_classes": n_classes, "beta": BETA, "alpha": ALPHA,
                   "values": "hard (one-hot of argmax L)" if soft_scale is None else f"softmax({soft_scale:.6f} * L) (SoftmaxCacheStrategy, scale 0.1)",
                   "sharding": sharding, "key_splits_per_gpu": splits, "attention_launches_per_step": n_attn_launches,
                   "accumulate": "fp32",
                   "l2": "inputs larger than L2: key bank = %.2f GB per GPU (+ %s)" % (
                       2 * n_local * dim / 1e9, "sorted by label" if not dense else "%.2f GB dense values" % (2 * n_local * c_pad / 1e9)),
                   "values_operand": "one-hot, label-sorted bank (W @ V = per-class segmented sum out of tensor memory)" if not dense else "dense Vt",
                   "bank_build_ms": bank_build_ms, "bank_build_first_call_ms": bank_build_first_ms, "top1_count": top1},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                     "kernel": "sc_attn_seg_kernel" if not dense else "sc_attn_t_kernel", "kernel_ms": attn_ms,
                     "kernel_ms_per_rank": attn_ms_ranks, "algorithmic_flops_per_launch": flops,
                     "executed_flops_per_launch": executed, "dense_equivalent_tflops": dense_equiv,
                     "note": ("one-hot values: achieved/frac count the tensor-core FLOPs actually issued (GEMM-1; W @ one_hot is a segmented fp32 "
                              "sum, no GEMM-2 exists); dense_equivalent_tflops = 2*Nq*Nk*(D+C)/time is what a dense-V kernel would need for the "
                              "same queries/s and is NOT a roofline figure") if not dense else
                             "dense values: achieved = SURVEY 8d's algorithmic count 2*Nq*Nk*(D+C) / kernel time (both GEMMs execute)",
                     "peak_source": peaks["source"] + " burst (cuBLAS bf16 8192^3)",
                     "frac_of_sustained": achieved / peaks["bf16_tflops_sustained"] if peaks["bf16_tflops_sustained"] else None},
        "e2e": {"value": nq / (e2e_ms / args.steps * 1e-3), "unit": "queries/s",
                "h2d_bytes_per_step": q_host_mine.numel() * q_host_mine.element_size() + lab_host_mine.numel() * 4,
                "d2h_bytes_per_step": pred.numel() * 4 + counts.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                "note": "bytes are per rank: every rank copies its 1/N slice of the query bank" if world > 1 else None},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
    }
    if check is not None:
        out["parity_check"] = check
    if eager is not None:
        out["gpu_eager_baseline"] = eager
    if phases is not None:
        out["phases_ms"] = phases
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(nq, nk, dim, n_classes, budget_s=args.cpu_budget)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_latency(args):
    """`--workload latency` (BASELINE.json configs[4]): host-observed latency of ClipSearcher.search for query batches
    of 1 .. 4096 against the resident 1.28M-key bank (queries on the device, predictions copied back), eager and
    replayed from a CUDA graph; N > 1 ranks shard the keys.  One JSON line; `value` = graph-replayed p50 at batch 64."""
    import torch
    import torch.distributed as dist

    from summer_clip_b200 import build as _build
    from summer_clip_b200.searcher import ClipSearcher, shard_range

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CLIP-search path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if rank == 0:
        _build.build_library()
    if world > 1:
        dist.barrier()
    nq, nk, dim, n_classes, desc = WORKLOADS["latency"]
    lo, hi = shard_range(nk, rank, world) if world > 1 else (0, nk)
    q_bank, k_bank, outs, text, labels = make_banks(torch, nq, lo, hi, dim, n_classes, seed=5, device=dev)
    s = ClipSearcher(dev, group=group, shard="keys")
    s.set_text(text)
    s.set_cache(k_bank, outs, local_shard=world > 1)
    del k_bank, outs
    bank_bytes = 2.0 * dim * (s.hard_bank.n_sorted if s.hard_bank is not None else (hi - lo))
    iters = max(10, args.steps * 4)
    rows = []

    def timed(fn):
        lat = []
        for _ in range(iters + 3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = sorted(lat[3:])
        return lat[len(lat) // 2], lat[min(len(lat) - 1, int(len(lat) * 0.99))], out

    sampler = ClockSampler(local)
    sampler.start()
    for b in LATENCY_BATCHES:
        qb, lab = q_bank[:, :b].contiguous(), labels[:b].contiguous()
        p50, p99, pred = timed(lambda: s.search(qb, [BETA], [ALPHA], labels=lab)[0]["pred"].cpu())
        g50 = g99 = None
        same = None
        try:
            graph, gres = s.capture_search(qb, [BETA], [ALPHA], labels=lab)

            def replay():
                graph.replay()
                return gres[0]["pred"].cpu()

            g50, g99, gpred = timed(replay)
            same = bool((gpred == pred).all())
            del graph, gres
        except Exception as exc:  # noqa: BLE001
            same = f"{type(exc).__name__}: {exc}"[:120]
        if world > 1:
            t = torch.tensor([p50, p99, g50 if g50 is not None else -1.0, g99 if g99 is not None else -1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p50, p99 = float(t[0]), float(t[1])
            g50, g99 = (float(t[2]), float(t[3])) if g50 is not None else (None, None)
        best = g50 if g50 is not None else p50
        rows.append({"batch": b, "p50_ms": round(p50, 4), "p99_ms": round(p99, 4), "graph_p50_ms": None if g50 is None else round(g50, 4),
                     "graph_p99_ms": None if g99 is None else round(g99, 4), "graph_same_pred": same,
                     "queries_per_s": b / (best * 1e-3), "bank_gbs_per_gpu": bank_bytes / (best * 1e-3) / 1e9})
    clocks = sampler.stop()
    if rank == 0:
        peaks = read_peaks()
        at64 = next(r for r in rows if r["batch"] == 64)
        v = at64["graph_p50_ms"] if at64["graph_p50_ms"] is not None else at64["p50_ms"]
        out = {"metric": "clip_search_latency_p50_ms", "value": v, "unit": "ms", "n_gpus": world, "steps": iters, "warmup": 3,
               "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
               "data": "synthetic",
               "config": {"workload": "latency", "description": desc, "n_keys": nk, "dim": dim, "n_classes": n_classes, "beta": BETA,
                          "alpha": ALPHA, "value_is": "p50 of a CUDA-graph replay of ClipSearcher.search at batch 64, predictions copied to the host",
                          "sharding": f"key-sharded x{world}" if world > 1 else "single GPU"},
               "latency": rows,
               "roofline": {"bound": "hbm", "achieved": at64["bank_gbs_per_gpu"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": at64["bank_gbs_per_gpu"] / peaks["hbm_gbs"], "traffic": None,
                            "note": "bank bytes of this GPU's shard / host-observed p50 at batch 64 (includes launch latency and the D2H of the predictions)"},
               "clocks": clocks}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


CPU_SAMPLE = (1024, 131072)       # queries x keys of the bounded CPU sample, the same in both arms


def _cpu_sample_step(nq, nk, dim, n_classes):
    """One step of the reference path on the host: the oracle port (normalise both banks, Q^T K, exp, one-hot, @,
    Z + alpha*O, top-1/5) in torch fp32 on a CPU_SAMPLE-sized slice of the workload, all host threads."""
    import torch
    from oracle import clip_search_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sq, sk = min(nq, CPU_SAMPLE[0]), min(nk, CPU_SAMPLE[1])
    banks = orc.synthetic_banks(sq, sk, dim, n_classes, seed=3)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"].long()

    def step():
        Z = orc.zero_shot_logits(Q, T)
        O = orc.image_attention(Q, K, orc.hard_values(L), BETA, chunk=256)
        return orc.compute_accuracy(orc.searcher_logits(Z, O, ALPHA), labels)

    return step, sq, sk, cores


def cpu_baseline(nq, nk, dim, n_classes, budget_s=20.0, steps=1):
    """The oracle port timed on the host cores on a bounded sample of the same workload (cost is linear in queries
    and in keys; the measured sample time and the factor to the full workload are reported separately)."""
    step, sq, sk, cores = _cpu_sample_step(nq, nk, dim, n_classes)
    step()                                                      # warm-up
    best = float("inf")
    t_all = time.perf_counter()
    for _ in range(3):
        t0 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    scale = (nk / sk) * (nq / sq)
    return {"value": nq / (best * scale), "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{sq} queries x {sk} keys x {dim}-d, {n_classes} classes (fp32 torch CPU, best of 3)",
            "sample_seconds": best, "scale_factor_to_full_workload": scale,
            "gflops": 2.0 * sq * sk * (dim + n_classes) / best / 1e9}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is a pure Python repo
    without a build or an installable package, and /root/reference is absent on the GPU box, so the timed code is
    the oracle port (kind "port"), all host threads.  Every step is the bounded CPU_SAMPLE slice of the workload:
    `ms_per_step` is the MEASURED time of such a step (steps x ms_per_step is what ran); `value` scales it linearly
    in queries and keys to the full workload, with the factor stated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nq, nk, dim, n_classes, desc = WORKLOADS[args.workload]
    if args.nq:
        nq = args.nq
    step, sq, sk, cores = _cpu_sample_step(nq, nk, dim, n_classes)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sample_s = (time.perf_counter() - t0) / args.steps
    scale = (nk / sk) * (nq / sq)
    qps = nq / (sample_s * scale)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = f"{sq} queries x {sk} keys per step (fp32 torch CPU, {cores} threads); value = the full {nq} x {nk} workload at this rate"
    out = {"impl": "reference", "metric": "clip_search_queries_per_sec", "value": qps, "unit": "queries/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sample_s * 1e3,
           "sample_ms_per_step": sample_s * 1e3, "scale_factor_to_full_workload": scale,
           "full_workload_ms_per_step_extrapolated": sample_s * scale * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "description": desc, "n_queries": nq, "n_keys": nk, "dim": dim,
                      "n_classes": n_classes, "beta": BETA, "alpha": ALPHA},
           "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                            "sample_seconds": sample_s, "scale_factor_to_full_workload": scale},
           "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="imagenet_rn50", choices=sorted(WORKLOADS))
    ap.add_argument("--nq", type=int, default=0, help="override the number of queries (debug)")
    ap.add_argument("--op-dtype", default="", choices=["", "fp16", "bf16", "e4m3"],
                    help="tensor-core operand type of the feature banks (default fp16; e4m3 is the opt-in 8-bit mode)")
    ap.add_argument("--values", default="hard", choices=["hard", "softmax"],
                    help="cache values: hard = HardCacheStrategy (one-hot, the reference default; segmented kernel), "
                         "softmax = SoftmaxCacheStrategy(clip_scale, 0.1) (dense values; dual-GEMM kernel)")
    ap.add_argument("--shard", default="keys", choices=["keys", "queries"],
                    help="what N > 1 ranks split: the key bank (reduce-scatter of partial tiles) or the queries (no data-path collective)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the torch-eager fp16 reference expressions on the GPU")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--phases", action="store_true", help="add phases_ms: rank 0's per-phase device times inside the timed steps")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "latency":
        run_latency(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
