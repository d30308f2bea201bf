#!/bin/bash
# round-2 session 14: t2i prefill tensor-core attention, 3B defaults, 32-column TMEM drain; full GPU suite + default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s14; mkdir -p $O
timeout 300 python -m pytest tests/test_gpt_gpu.py -m gpu -q -k "prefill_tensor_core or golden or t2i" > $O/pytest_prefill.log 2>&1; echo "rc=$?" >> $O/pytest_prefill.log; tail -n 8 $O/pytest_prefill.log
bash tools/sweep.sh "LG_TC_LD32=1" "LG_TC_LD32=0" > $O/sweep_ld32.txt 2>&1; cat $O/sweep_ld32.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -n 16 $O/pytest_gpu.log
F="--no-cpu-baseline --no-gpu-reference --no-operating-points --no-latency --no-roofline --steps 3 --warmup 3"
timeout 400 python bench.py $F --gpt-model GPT-XL --image-size 512 --batch 8 --t2i --cfg-scale 7.5 --top-k 1000 > $O/bench_c5.json 2> $O/c5.err
LG_ATTN_PREFILL_TC=0 timeout 400 python bench.py $F --gpt-model GPT-XL --image-size 512 --batch 8 --t2i --cfg-scale 7.5 --top-k 1000 > $O/bench_c5_cudacore_prefill.json 2> $O/c5b.err
python - <<'PY'
import json
for f in ("gpurun_out/s14/bench_c5.json","gpurun_out/s14/bench_c5_cudacore_prefill.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], json.dumps(d.get("prefill"))[:300])
    except Exception as e: print(f,"ERR",e)
PY
