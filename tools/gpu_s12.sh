#!/bin/bash
# round-2 session 12: persistent conv with a CTA budget (does the VQ decode of batch i now hide behind the AR loop of batch i+1?),
# GPT-3B padded-KV TMA attention parity
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s12; mkdir -p $O
timeout 600 python -m pytest tests/test_vq_gpu.py -m gpu -q -x > $O/pytest_vq.log 2>&1; echo "rc=$?" >> $O/pytest_vq.log; tail -n 4 $O/pytest_vq.log
LG_CONV_CTAS=64 timeout 600 python -m pytest tests/test_vq_gpu.py -m gpu -q -x -k "decoder or decode" > $O/pytest_vq_budget64.log 2>&1; echo "rc=$?" >> $O/pytest_vq_budget64.log; tail -n 4 $O/pytest_vq_budget64.log
timeout 600 python -m pytest tests/test_gpt_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -k "3b or head_dim_100" > $O/pytest_3b.log 2>&1; echo "rc=$?" >> $O/pytest_3b.log; tail -n 6 $O/pytest_3b.log
bash tools/sweep.sh "LG_PIPE_CONV_CTAS=0" "LG_PIPE_CONV_CTAS=296" "LG_PIPE_CONV_CTAS=128" "LG_PIPE_CONV_CTAS=96" "LG_PIPE_CONV_CTAS=64" "LG_PIPE_CONV_CTAS=48" "LG_PIPE_CONV_CTAS=32" \
   "LG_PIPE_CONV_CTAS=64 LG_AR_PRIORITY=0" "LG_PIPE_CONV_CTAS=0 LG_AR_PRIORITY=0" "LG_BENCH_PIPELINE=0" > $O/sweep_budget.txt 2>&1
cat $O/sweep_budget.txt
F="--no-cpu-baseline --no-gpu-reference --no-operating-points --no-latency --no-roofline --steps 3 --warmup 3"
timeout 400 python bench.py $F --gpt-model GPT-3B --image-size 384 --batch 16 > $O/bench_c4_3b_tma.json 2> $O/c4.err
LG_HD_PAD=0 timeout 400 python bench.py $F --gpt-model GPT-3B --image-size 384 --batch 16 > $O/bench_c4_3b_cudacore.json 2> $O/c4b.err
for f in $O/bench_c4*.json; do echo $f; cut -c1-200 $f; done
