#!/bin/bash
# quick iteration loop: rollup/decode parity subset + the kernel-only bench at 20 000 blocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
timeout 600 python bench.py --blocks ${BENCH_BLOCKS:-20000} --steps 5 --warmup 3 --no-e2e --cpu-seconds 0 ${BENCH_ARGS} > gpurun_out/bench_quick.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_quick.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms_per_step", round(d["ms_per_step"], 3), "value %.2f G" % (d["value"] / 1e9), {k: v["ms"] for k, v in d["roofline"]["stages"].items()})
        break
else:
    print(open("gpurun_out/bench_quick.log").read()[-2000:])
PY
