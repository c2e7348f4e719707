#!/bin/bash
# ncu passes for one short bench run (1 GPU).  Outputs under gpurun_out/ (copy the summaries you want to keep into profiles/).
mkdir -p gpurun_out
B=${BENCH_BLOCKS:-20000}
CMD="python bench.py --blocks $B --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 --no-aggr --no-alt-encoder --configs2-series 0 --parity-series 0 ${BENCH_EXTRA:-}"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit $?"
for K in ${NCU_KERNELS:-k_fused_rollup k_huf_decode k_zstd_prepare}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
  python scripts/ncu_summary.py gpurun_out/prof_$K.ncu-rep 16 > gpurun_out/ncu_full_$K.txt 2>&1
  # gpurun brings back at most 64 MiB: keep the report of the first (dominant) kernel only, the summaries of all
  if [ "$K" != "${NCU_KEEP:-k_fused_rollup}" ]; then rm -f gpurun_out/prof_$K.ncu-rep; fi
done
# the sequences path barely runs on the bench's random-increment counters: capture its two kernels on reference-encoded smooth counters
SMOOTH="python scripts/exp_refzstd.py 20000 regular counter_smooth"
for K in ${NCU_SEQ_KERNELS:-k_zstd_seq_decode k_zstd_seq_exec}; do
  S=2; [ "$K" = k_zstd_seq_exec ] && S=1   # (k_zstd_seq_decode is launched twice per step: 256-state tables, then the flagged frames)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_$K $SMOOTH > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
  python scripts/ncu_summary.py gpurun_out/prof_$K.ncu-rep 16 > gpurun_out/ncu_full_$K.txt 2>&1
  rm -f gpurun_out/prof_$K.ncu-rep
done
ls -la gpurun_out | tail -20
