#!/bin/bash
# round-2 session 19: full GPU suite + smoke + the default bench line of the current tree
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s19; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -n 12 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -n 3 $O/smoke.log
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s19/bench_default.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","latency_b1","operating_points","step_roofline","roofline","gpu_launches","clocks"):
    print(k, json.dumps(d.get(k))[:500])
print(json.dumps(d.get("gpu_reference",{}).get("batches",{}).get("64",{}))[:600])
print(json.dumps(d.get("kernels"))[:1500])
PY
