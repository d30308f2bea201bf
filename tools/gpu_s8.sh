#!/bin/bash
# round-2 session 8: gemm_dx phase stamps; attention with 32-key stages (smaller shared-memory footprint -> chains co-reside?)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s8; mkdir -p $O
timeout 120 python tools/dx_probe.py 64 > $O/dx_probe_r64.txt 2>&1
LG_DX_STAGES=4 timeout 120 python tools/dx_probe.py 64 > $O/dx_probe_r64_st4.txt 2>&1
cat $O/dx_probe_r64.txt
KC32=$PWD/llamagen_b200/lib_kc32/libllamagen_b200.so
bash tools/sweep.sh "LG_DIRECT=0" "LG_DIRECT=0 LG_LIB_PATH=$KC32" "LG_DIRECT=0 LG_LIB_PATH=$KC32 LG_TC_STAGES=3" "LG_DIRECT=0 LG_LIB_PATH=$KC32 LG_SPLIT=4" \
   "LG_DIRECT=0 LG_LIB_PATH=$KC32 LG_SPLIT=4 LG_TC_STAGES=3" "LG_DIRECT=0 LG_LIB_PATH=$KC32 LG_SPLIT=1" > $O/sweep_kc32.txt 2>&1
cat $O/sweep_kc32.txt
