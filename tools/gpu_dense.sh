#!/bin/bash
# dense-values kernel A/B runs: bench.py --values softmax under the knobs given as "NAME=VAL,NAME=VAL" specs
tag=${1:-r02}; shift
mkdir -p gpurun_out
for spec in "$@"; do
  envs=$(echo "$spec" | tr ',' ' ')
  [ "$spec" = "default" ] && envs=""
  env $envs python bench.py --values softmax --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_dense_${spec//[=,]/_}.json 2> gpurun_out/${tag}_dense_${spec//[=,]/_}.err
  python - "$spec" "gpurun_out/${tag}_dense_${spec//[=,]/_}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-40s ms=%.1f attn=%.1f frac=%.3f mhz=%s W=%s splits=%s parity=%s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["power_w"], d["config"]["key_splits_per_gpu"], d["parity_check"]["ok"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace(".json",".err")).read()[-800:])
PY
done
