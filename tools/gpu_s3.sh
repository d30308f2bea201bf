#!/bin/bash
# round-2 session 3: persistent-kernel phase probe, full GPU suite after the sampler / row-block GEMM / uint8 drain changes, BASELINE configs C3-C5
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s3; mkdir -p $O
timeout 300 python tools/persist_probe.py GPT-L 1 > $O/probe_l_b1.txt 2>&1
timeout 300 python tools/persist_probe.py GPT-L 4 > $O/probe_l_b4.txt 2>&1
LG_PD_COOP=0 timeout 300 python tools/persist_probe.py GPT-B 1 > $O/probe_b_b1.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_parity_configs_gpu.py > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
B="--no-cpu-baseline --no-gpu-reference --no-operating-points --steps 3 --warmup 3"
timeout 600 python bench.py $B --no-roofline --no-latency > $O/bench_default.json 2> $O/bench_default.err
timeout 900 python bench.py $B --gpt-model GPT-XL --image-size 384 --batch 32 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py $B --gpt-model GPT-3B --image-size 384 --batch 16 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 900 python bench.py $B --t2i --gpt-model GPT-XL --image-size 512 --batch 8 --cfg-scale 7.5 --top-k 1000 > $O/bench_c5.json 2> $O/bench_c5.err
cat $O/probe_l_b1.txt; tail -n 5 $O/pytest_all.log; for f in $O/bench_*.json; do echo $f; cut -c1-300 $f; done
