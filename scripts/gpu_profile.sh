#!/bin/bash
# ncu passes for one short bench run (1 GPU).  Outputs under gpurun_out/.
mkdir -p gpurun_out
B=${BENCH_BLOCKS:-20000}
CMD="python bench.py --blocks $B --steps 1 --warmup 1 --no-e2e --cpu-seconds 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit $?"
for K in ${NCU_KERNELS:-k_huf_decode k_rollup k_decode_columns k_series_prepare}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
ls -la gpurun_out | tail -20
