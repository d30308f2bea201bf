#!/bin/bash
# round-2 session 2: cluster-fused path (after the smem-attribute fix), parity at bench configs, GPU reference leg, L2 hints
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s2; mkdir -p $O
LG_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "cluster" > $O/pytest_cluster_gemm.log 2>&1; echo "rc=$?" >> $O/pytest_cluster_gemm.log
LG_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpt_gpu.py -m gpu -q -k "experimental" > $O/pytest_cluster_gpt.log 2>&1; echo "rc=$?" >> $O/pytest_cluster_gpt.log
B="--no-cpu-baseline --no-latency --no-roofline --no-gpu-reference --no-operating-points --steps 5 --warmup 3"
LG_TC_CLUSTER=2 timeout 300 python bench.py $B > $O/bench_cl2.json 2> $O/bench_cl2.err
LG_TC_CLUSTER=2 LG_SPLIT=1 timeout 300 python bench.py $B > $O/bench_cl2_split1.json 2> $O/bench_cl2_split1.err
LG_TC_CLUSTER=2 LG_L2_HINT=3 timeout 300 python bench.py $B > $O/bench_cl2_hint3.json 2> $O/bench_cl2_hint3.err
LG_L2_HINT=1 timeout 300 python bench.py $B > $O/bench_hint1.json 2> $O/bench_hint1.err
LG_L2_HINT=3 timeout 300 python bench.py $B > $O/bench_hint3.json 2> $O/bench_hint3.err
LG_TC_CLUSTER=2 timeout 300 python bench.py $B --batch 32 > $O/bench_cl2_b32.json 2> $O/bench_cl2_b32.err
timeout 600 python -m pytest tests/test_gpt_gpu.py -m gpu -q -k "persistent" > $O/pytest_persist.log 2>&1; echo "rc=$?" >> $O/pytest_persist.log
LG_PERSIST=1 timeout 300 python bench.py --no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --steps 2 --warmup 3 > $O/bench_persist_lat.json 2> $O/bench_persist_lat.err
LG_PERSIST=1 LG_PD_COOP=0 timeout 300 python bench.py --no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --steps 2 --warmup 3 > $O/bench_persist_nocoop_lat.json 2> $O/bench_persist_nocoop_lat.err
timeout 1200 python -m pytest tests/test_parity_configs_gpu.py -m gpu -q -s > $O/pytest_parity.log 2>&1; echo "rc=$?" >> $O/pytest_parity.log
timeout 300 python -m pytest tests/test_cli_gpu.py -m gpu -q -k "feature_files" > $O/pytest_f3.log 2>&1; echo "rc=$?" >> $O/pytest_f3.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 1000 python bench.py --impl gpu-reference --gpu-reference-batches 64,32 > $O/gpu_reference.json 2> $O/gpu_reference.err
tail -n 4 $O/*.log; for f in $O/bench_*.json; do echo $f; cut -c1-200 $f; done; cat $O/gpu_reference.json
