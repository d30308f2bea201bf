#!/bin/bash
# A/B sweep of environment switches on the bench workload (no CPU baseline / roofline legs). Usage: tools/sweep.sh "VAR=a VAR=b ..."
# Each argument is one space-separated set of VAR=value assignments; prints images/s and ms/step per setting.
for cfg in "$@"; do
  out=$(env $cfg timeout ${SWEEP_TIMEOUT:-120} python bench.py $SWEEP_ARGS --steps ${SWEEP_STEPS:-4} --warmup 3 --no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --no-latency 2>/dev/null | tail -1)
  echo "$cfg -> $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],2), "img/s", round(d["ms_per_step"],2), "ms", "e2e", round(d["e2e"]["value"],2))' 2>/dev/null || echo FAIL)"
done
