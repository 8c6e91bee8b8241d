#!/bin/bash
# 8-GPU verification: sharded-search check, then the key-sharded bench (default pipeline, and the NCCL exchange for
# comparison), the query-sharded bench and the dense-values bench
N=${1:-8}; tag=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tests/checks/check_sharded.py > gpurun_out/${tag}_check_sharded_n$N.log 2>&1; echo "check_sharded rc=$?"; grep -E "SHARDED|FAIL|Error" gpurun_out/${tag}_check_sharded_n$N.log | head -12
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "value=%.0f ms=%.3f e2e=%.0f attn=%.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"]), d.get("phases_ms"), d["clocks"]["sm_mhz"], d["config"]["top1_count"], (d.get("parity_check") or {}).get("ok_all_ranks"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
$TR bench.py --gpus $N --steps 10 --warmup 3 --phases > gpurun_out/${tag}_bench_n${N}_keys.json 2> gpurun_out/${tag}_bench_n${N}_keys.err; summ gpurun_out/${tag}_bench_n${N}_keys.json
SUMMER_CLIP_B200_EXCHANGE=nccl SC_BENCH_BLOCKS=1 $TR bench.py --gpus $N --steps 10 --warmup 3 --phases --no-parity-check > gpurun_out/${tag}_bench_n${N}_keys_nccl_b1.json 2> gpurun_out/${tag}_bench_n${N}_keys_nccl_b1.err; summ gpurun_out/${tag}_bench_n${N}_keys_nccl_b1.json
$TR bench.py --gpus $N --workload latency > gpurun_out/${tag}_latency_n${N}.json 2> gpurun_out/${tag}_latency_n${N}.err; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_latency_n${N}.json').read().strip().splitlines()[-1])
print('latency', d['value'], [(r['batch'], r['p50_ms'], r['graph_p50_ms']) for r in d['latency']])" || tail -5 gpurun_out/${tag}_latency_n${N}.err
$TR bench.py --gpus $N --steps 10 --warmup 3 --phases --shard queries > gpurun_out/${tag}_bench_n${N}_queries.json 2> gpurun_out/${tag}_bench_n${N}_queries.err; summ gpurun_out/${tag}_bench_n${N}_queries.json
$TR bench.py --gpus $N --steps 5 --warmup 3 --phases --values softmax > gpurun_out/${tag}_bench_n${N}_keys_softvalues.json 2> gpurun_out/${tag}_bench_n${N}_keys_softvalues.err; summ gpurun_out/${tag}_bench_n${N}_keys_softvalues.json
