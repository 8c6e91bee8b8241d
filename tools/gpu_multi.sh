#!/bin/bash
# multi-GPU checks: N-rank sharded-search check + bench at N ranks (key shards with 1 / 4 query blocks, query shards)
N=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tests/checks/check_sharded.py > gpurun_out/${tag}_check_sharded_n$N.log 2>&1; echo "check_sharded rc=$?"; grep -E "SHARDED|FAIL|Error|error" gpurun_out/${tag}_check_sharded_n$N.log | head -20
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "value=%.0f ms=%.3f e2e=%.0f attn=%.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"]), d.get("phases_ms"), d["clocks"]["sm_mhz"], d["config"]["top1_count"], (d.get("parity_check") or {}).get("ok"), (d.get("parity_check") or {}).get("ok_all_ranks"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
for ex in p2p nccl; do for blk in 1 4; do
  SUMMER_CLIP_B200_EXCHANGE=$ex SC_BENCH_BLOCKS=$blk $TR bench.py --gpus $N --steps 5 --warmup 3 --phases --no-parity-check > gpurun_out/${tag}_bench_n${N}_keys_${ex}_b$blk.json 2> gpurun_out/${tag}_bench_n${N}_keys_${ex}_b$blk.err; summ gpurun_out/${tag}_bench_n${N}_keys_${ex}_b$blk.json
done; done
$TR bench.py --gpus $N --steps 5 --warmup 3 --phases > gpurun_out/${tag}_bench_n${N}_keys.json 2> gpurun_out/${tag}_bench_n${N}_keys.err; summ gpurun_out/${tag}_bench_n${N}_keys.json
$TR bench.py --gpus $N --steps 5 --warmup 3 --phases --shard queries > gpurun_out/${tag}_bench_n${N}_queries.json 2> gpurun_out/${tag}_bench_n${N}_queries.err; summ gpurun_out/${tag}_bench_n${N}_queries.json
python bench.py --steps 5 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; summ gpurun_out/${tag}_bench_n1.json
