"""BASELINE.json config 5 (online classification latency): thin wrapper of `bench.py --workload latency`.

    python tools/bench_latency.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_latency.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = [sys.argv[0], "--workload", "latency"] + sys.argv[1:]
    bench.main()
