#!/bin/bash
# A/B runs of bench.py over the attention kernels' knobs; produced the r01* "knobs" logs under profiles/.
# SC_ATTN_DEBUG_SKIP / SC_ATTN_CLKPROBE act only in an experiments build (SC_BUILD_EXPERIMENTS=1 python -m
# summer_clip_b200.build; results are wrong by design: work is skipped); SC_ATTN_CLUSTER / SC_ATTN_STAGES selected
# round-1 kernels that no longer exist and are ignored now; SC_ATTN_PREFETCH and SUMMER_CLIP_B200_SPLITS are live.
# usage: tools/bench_knobs.sh "G:NS G:NS ..." [extra bench args]
specs="$1"; shift
for spec in $specs; do
  IFS=: read G NS SP PF SK <<< "$spec"; SP=${SP:-}; PF=${PF:-}; SK=${SK:-0}
  SC_ATTN_CLUSTER=$G SC_ATTN_STAGES=$NS SUMMER_CLIP_B200_SPLITS=$SP SC_ATTN_PREFETCH=$PF SC_ATTN_CLKPROBE=1 SC_ATTN_DEBUG_SKIP=$SK python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('G=$G NS=$NS SP=$SP PF=$PF SKIP=$SK', 'ms_step=%.1f'%d['ms_per_step'], 'qps=%.0f'%d['value'], 'attn_ms=%.1f'%d['roofline']['kernel_ms'], 'frac=%.3f'%d['roofline']['frac'], 'sm_mhz=%s'%d['clocks']['sm_mhz'], 'cta_mhz=%s'%d['clocks'].get('attn_cta_mhz'), 'W=%s'%d['clocks'].get('power_w'), d['clocks']['reasons'], 'top1=%d'%d['config']['top1_count'])"
done
