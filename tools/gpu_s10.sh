#!/bin/bash
# round-2 session 10: TS-form MMA (weights copied smem -> TMEM with tcgen05.cp, A read from TMEM) in gemm_dx: parity + phase stamps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s10; mkdir -p $O
LG_DX_TS=1 timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "dx" > $O/pytest_dx_ts.log 2>&1; echo "rc=$?" >> $O/pytest_dx_ts.log
tail -n 12 $O/pytest_dx_ts.log
LG_DX_TS=1 LG_DX_NACC=1 timeout 120 python tools/dx_probe.py 64 > $O/dx_probe_ts_nacc1.txt 2>&1
LG_DX_TS=1 LG_DX_NACC=4 timeout 120 python tools/dx_probe.py 64 > $O/dx_probe_ts_nacc4.txt 2>&1
grep -A1 "cold_l2=False" $O/dx_probe_ts_nacc1.txt | head -12; echo; grep -A1 "cold_l2=False" $O/dx_probe_ts_nacc4.txt | head -12
bash tools/sweep.sh "LG_DIRECT=1 LG_DX_TS=1 LG_DX_NACC=1" "LG_DIRECT=1 LG_DX_TS=1 LG_DX_NACC=4" > $O/sweep.txt 2>&1; cat $O/sweep.txt
