#!/bin/bash
# key-sharded bench at N ranks for several query-block counts of the peer-memory pipeline
N=${1:-8}; tag=${2:-r02}; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for blk in "$@"; do
  SC_BENCH_BLOCKS=$blk $TR bench.py --gpus $N --steps 10 --warmup 3 --phases --no-parity-check > gpurun_out/${tag}_bench_n${N}_keys_p2p_b$blk.json 2> gpurun_out/${tag}_bench_n${N}_keys_p2p_b$blk.err
  python - "gpurun_out/${tag}_bench_n${N}_keys_p2p_b$blk.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ph=d.get("phases_ms") or {}
    print(sys.argv[1].split('/')[-1], "value=%.0f ms=%.3f e2e=%.0f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "tail=%.3f"%ph.get("tail_after_last_attention",-1), d["roofline"]["kernel_ms_per_rank"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
