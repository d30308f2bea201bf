#!/bin/bash
# round-2 session 23: final tree (32-key attention stages by default): full GPU suite, smoke, bench with latency + B = 32 point, C4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s23; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --durations=3 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -n 9 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -n 2 $O/smoke.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-reference > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s23/bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","latency_b1","operating_points","step_roofline","roofline"):
    print(k, json.dumps(d.get(k))[:400])
print(json.dumps(d.get("kernels"))[:1200])
PY
F="--no-cpu-baseline --no-gpu-reference --no-operating-points --no-latency --no-roofline --steps 2 --warmup 3"
timeout 400 python bench.py $F --gpt-model GPT-3B --image-size 384 --batch 16 > $O/bench_c4.json 2> $O/c4.err
python -c "
import json
d=json.loads(open('gpurun_out/s23/bench_c4.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['ms_per_step'])"
