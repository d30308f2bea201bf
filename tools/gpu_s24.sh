#!/bin/bash
# round-2 session 24: batched slab loads in the row kernels / attention epilogue: parity subset + A/B against the previous build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s24; mkdir -p $O
timeout 240 python -m pytest tests/test_gpt_gpu.py -m gpu -q -x -k "golden or teacher or multi_chain or small_row_decode or direct" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -n 4 $O/pytest.log
PREV=$PWD/llamagen_b200/lib_kcprev/libllamagen_b200.so
bash tools/sweep.sh "LG_X=0" "LG_LIB_PATH=$PREV" "LG_X=1" "LG_LIB_PATH=$PREV" > $O/sweep.txt 2>&1; cat $O/sweep.txt
