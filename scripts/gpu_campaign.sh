#!/bin/bash
# ad-hoc differential campaign: the whole GPU suite under shifted seeds (tests/conftest.py SEED0); logs in gpurun_out/
mkdir -p gpurun_out
for k in ${OFFSETS:-1 2 3}; do
  VMB_SEED_OFFSET=$((k * 1000)) VMB_FUZZ_SEED=$((9000 + k)) timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider \
    > gpurun_out/campaign_$k.log 2>&1
  echo "offset $k: $(tail -1 gpurun_out/campaign_$k.log)"
  grep -E "^FAILED|^ERROR" gpurun_out/campaign_$k.log | head -20
done
