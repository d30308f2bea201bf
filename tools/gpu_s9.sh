#!/bin/bash
# round-2 session 9: does rotating the k-blocks over independent TMEM accumulators shorten the tcgen05.mma chain?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s9; mkdir -p $O
for n in 1 2 4; do LG_DX_NACC=$n timeout 120 python tools/dx_probe.py 64 > $O/dx_probe_nacc$n.txt 2>&1; done
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "dx" > $O/pytest_dx.log 2>&1; echo "rc=$?" >> $O/pytest_dx.log
tail -n 3 $O/pytest_dx.log
grep -A1 "cold_l2=False" $O/dx_probe_nacc1.txt | head -8; echo; grep -A1 "cold_l2=False" $O/dx_probe_nacc2.txt | head -8; echo; grep -A1 "cold_l2=False" $O/dx_probe_nacc4.txt | head -8
bash tools/sweep.sh "LG_DIRECT=1 LG_DX_NACC=4" "LG_DIRECT=1 LG_DX_NACC=2" > $O/sweep.txt 2>&1; cat $O/sweep.txt
