#!/bin/bash
# round 2, call A: GPU test suite, smoke, headline bench, dense-values split experiments
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest_gpu.log
tail -30 gpurun_out/r02a_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; tail -2 gpurun_out/r02a_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_hard.json 2> gpurun_out/r02a_bench_hard.err; tail -c 1500 gpurun_out/r02a_bench_hard.json
for sp in "" 20 40 80 160; do
  SC_BENCH_SPLITS=$sp python bench.py --values softmax --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02a_bench_soft_sp${sp:-auto}.json 2> gpurun_out/r02a_bench_soft_sp${sp:-auto}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02a_bench_soft_sp${sp:-auto}.json").read().strip().splitlines()[-1])
    print("soft splits=${sp:-auto}", "ms=%.1f attn=%.1f frac=%.3f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]), d["clocks"], d.get("parity_check"))
except Exception as e:
    print("soft splits=${sp:-auto} FAILED", e); print(open("gpurun_out/r02a_bench_soft_sp${sp:-auto}.err").read()[-800:])
PY
done
for sp in 39 78; do
  SC_BENCH_SPLITS=$sp python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02a_bench_hard_sp$sp.json 2> gpurun_out/r02a_bench_hard_sp$sp.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02a_bench_hard_sp$sp.json").read().strip().splitlines()[-1])
    print("hard splits=$sp", "ms=%.1f attn=%.1f frac=%.3f"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]), d["clocks"], d.get("parity_check"))
except Exception as e:
    print("hard splits=$sp FAILED", e); print(open("gpurun_out/r02a_bench_hard_sp$sp.err").read()[-800:])
PY
done
