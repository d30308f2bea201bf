#!/bin/bash
# round-2 session 1: validate the experimental cluster-fused decode path and A/B the bench variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
LG_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "cluster" > $O/pytest_cluster_gemm.log 2>&1
echo "rc=$?" >> $O/pytest_cluster_gemm.log
LG_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpt_gpu.py -m gpu -q -k "experimental" > $O/pytest_cluster_gpt.log 2>&1
echo "rc=$?" >> $O/pytest_cluster_gpt.log
B="--no-cpu-baseline --no-latency --no-roofline --steps 5 --warmup 3"
timeout 300 python bench.py $B > $O/bench_default.json 2> $O/bench_default.err
LG_TC_CLUSTER=2 timeout 300 python bench.py $B > $O/bench_cl2.json 2> $O/bench_cl2.err
LG_TC_CLUSTER=2 LG_SPLIT=1 timeout 300 python bench.py $B > $O/bench_cl2_split1.json 2> $O/bench_cl2_split1.err
LG_SPLIT=1 timeout 300 python bench.py $B > $O/bench_split1.json 2> $O/bench_split1.err
LG_TC_CLUSTER=2 timeout 300 python bench.py $B --batch 32 > $O/bench_cl2_b32.json 2> $O/bench_cl2_b32.err
timeout 300 python bench.py $B --batch 32 > $O/bench_b32.json 2> $O/bench_b32.err
tail -n 3 $O/*.log; cat $O/bench_*.json | cut -c1-400
