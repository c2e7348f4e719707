#!/bin/bash
# the "Other workloads" table of profiles/README.md: 20 000 blocks x 8192, one GPU, kernel-only
mkdir -p gpurun_out
OUT=gpurun_out/workloads.jsonl
: > $OUT
C="--blocks 20000 --steps 10 --warmup 3 --no-e2e --no-aggr --no-alt-encoder --configs2-series 0 --cpu-seconds 0 --parity-series 200"
run() { echo "### $*" >> $OUT; timeout 300 python bench.py $C "$@" 2>>gpurun_out/workloads.err | tail -1 >> $OUT; }
run --kind counter --func rate
run --kind counter --func rate --ts jitter
run --kind counter --func increase
run --kind gauge --func avg_over_time
run --kind gauge --func max_over_time
run --kind gauge --func quantile_over_time
run --kind mixed --func increase --window-ms 3600000 --step-ms 60000
timeout 200 python scripts/exp_multiblock.py > gpurun_out/multiblock.log 2>&1
python - <<'PY'
import json
lab=None
for l in open('gpurun_out/workloads.jsonl'):
    l=l.strip()
    if l.startswith('###'): lab=l[4:]; continue
    if not l.startswith('{'): continue
    d=json.loads(l)
    st={k: round(v['ms'],2) for k,v in d['roofline']['stages'].items()}
    print("%-75s %.2f ms  %.1f G/s  %s parity %s" % (lab, d['ms_per_step'], d['value']/1e9, st, d.get('parity_check',{}).get('ok')))
PY
tail -3 gpurun_out/multiblock.log
