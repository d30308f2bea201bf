#!/bin/bash
# round-2 session 21 (2 GPUs): weak-scaling sanity of the final tree + conv variants test + BASELINE configs on the final tree
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s21; mkdir -p $O
timeout 300 python -m pytest tests/test_vq_gpu.py -m gpu -q -x -k "variants" > $O/pytest_variants.log 2>&1; echo "rc=$?" >> $O/pytest_variants.log; tail -n 3 $O/pytest_variants.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --no-latency > $O/bench_n2.json 2> $O/n2.err; echo "rc=$?" >> $O/n2.err
tail -n 1 $O/bench_n2.json | cut -c1-400
F="--no-cpu-baseline --no-gpu-reference --no-operating-points --no-latency --no-roofline --steps 3 --warmup 3"
timeout 400 python bench.py $F --gpt-model GPT-XL --image-size 384 --batch 32 > $O/bench_c3.json 2> $O/c3.err
timeout 400 python bench.py $F --gpt-model GPT-3B --image-size 384 --batch 16 > $O/bench_c4.json 2> $O/c4.err
timeout 400 python bench.py $F --gpt-model GPT-XL --image-size 512 --batch 8 --t2i --cfg-scale 7.5 --top-k 1000 > $O/bench_c5.json 2> $O/c5.err
for f in $O/bench_c3.json $O/bench_c4.json $O/bench_c5.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), round(d['ms_per_step'],1), d['config'].get('kv_cache_bytes'))"; done
