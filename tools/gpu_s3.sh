#!/bin/bash
# round-2 session 3 (re-entry): status of HEAD on hardware — full GPU suite, smoke, default bench, persistent path, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s3; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
B="--no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --steps 2 --warmup 3"
LG_PERSIST=1 timeout 300 python bench.py $B > $O/bench_persist_lat.json 2> $O/bench_persist_lat.err
LG_PERSIST=1 LG_PD_COOP=0 timeout 300 python bench.py $B > $O/bench_persist_nocoop_lat.json 2> $O/bench_persist_nocoop_lat.err
LG_SPLIT=1 timeout 300 python bench.py $B --no-latency --steps 5 > $O/bench_split1.json 2> $O/bench_split1.err
timeout 300 python bench.py $B --no-latency --steps 5 --batch 32 > $O/bench_b32.json 2> $O/bench_b32.err
tail -n 25 $O/pytest_gpu.log; tail -n 5 $O/smoke.log; for f in $O/bench_*.json; do echo $f; cut -c1-1500 $f; done; tail -n 5 $O/bench_default.err
