#!/bin/bash
# round-2 session 16: fused sampling tail (embed + layer-0 RMSNorm + counter advance inside sample_kernel)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s16; mkdir -p $O
timeout 900 python -m pytest tests/test_gpt_gpu.py tests/test_sampling_gpu.py tests/test_cli_gpu.py -m gpu -q -x > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log; tail -n 8 $O/pytest_a.log
timeout 600 python -m pytest tests/test_parity_configs_gpu.py -m gpu -q -x -k "c2 or c4" > $O/pytest_b.log 2>&1; echo "rc=$?" >> $O/pytest_b.log; tail -n 5 $O/pytest_b.log
bash tools/sweep.sh "LG_FUSE_TAIL=1" "LG_FUSE_TAIL=0" "LG_FUSE_TAIL=1 LG_TC_CTAS=96" "LG_FUSE_TAIL=1 LG_TC_CTAS=120" "LG_FUSE_TAIL=1 LG_L2_HINT=1" "LG_FUSE_TAIL=1 LG_L2_HINT=2" "LG_FUSE_TAIL=1 LG_L2_HINT=3" > $O/sweep_tail.txt 2>&1; cat $O/sweep_tail.txt
B="--no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --steps 2 --warmup 3"
LG_FUSE_TAIL=1 timeout 300 python bench.py $B > $O/bench_lat_tail1.json 2> $O/l1.err
LG_FUSE_TAIL=0 timeout 300 python bench.py $B > $O/bench_lat_tail0.json 2> $O/l0.err
python - <<'PY'
import json
for f in ("gpurun_out/s16/bench_lat_tail1.json","gpurun_out/s16/bench_lat_tail0.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["latency_b1"])
    except Exception as e: print(f,"ERR",e)
PY
