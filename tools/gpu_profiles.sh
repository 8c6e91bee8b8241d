#!/bin/bash
# round-end evidence: ncu launch list + full capture of the headline kernel for the bench command, DRAM bytes of the
# dense kernel at full size, final dense-values bench line
tag=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-parity-check"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sc_attn_seg -s 3 -c 1 -o gpurun_out/${tag}_attn_seg $CMD > gpurun_out/${tag}_ncu_seg.log 2>&1
DCMD="python bench.py --values softmax --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-parity-check"
$DCMD > gpurun_out/${tag}_dense_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:sc_attn_t -s 1 -c 1 --csv --log-file gpurun_out/${tag}_dense_full_dram.csv $DCMD > /dev/null 2>&1
python bench.py --values softmax --steps 5 --warmup 3 > gpurun_out/${tag}_bench_softmax_values.json 2> gpurun_out/${tag}_bench_softmax_values.err
tail -c 400 gpurun_out/${tag}_bench_softmax_values.json; echo; grep -E "dram__|gpu__time|tensor" gpurun_out/${tag}_dense_full_dram.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'; ls -la gpurun_out/${tag}_*
