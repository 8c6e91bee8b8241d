#!/bin/bash
# DRAM bytes and duration of the dense kernel per key-split count (ncu, 3 metrics, one launch each)
tag=${1:-r02}; shift
mkdir -p gpurun_out
python bench.py --values softmax --nq 12544 --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline --no-parity-check > gpurun_out/${tag}_plain.log 2>&1 || exit 1
for sp in "$@"; do
  SUMMER_CLIP_B200_SPLITS=$sp ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sc_attn_t -s 2 -c 1 --csv --log-file gpurun_out/${tag}_dram_sp$sp.csv python bench.py --values softmax --nq 12544 --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline --no-parity-check > /dev/null 2>&1
  echo "splits=$sp $(grep -E 'gpu__time_duration|dram__bytes|hit_rate' gpurun_out/${tag}_dram_sp$sp.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tr '\n' ';')"
done
