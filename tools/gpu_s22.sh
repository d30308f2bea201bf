#!/bin/bash
# round-2 session 22: attention stage size on the final tree (GPT-L bench shape and GPT-3B)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s22; mkdir -p $O
KC32=$PWD/llamagen_b200/lib_kc32/libllamagen_b200.so
KC32C12=$PWD/llamagen_b200/lib_kc32c12/libllamagen_b200.so
bash tools/sweep.sh "LG_X=0" "LG_LIB_PATH=$KC32" "LG_LIB_PATH=$KC32C12" "LG_X=1" "LG_LIB_PATH=$KC32" "LG_TC_STAGES=3" > $O/sweep_l.txt 2>&1; cat $O/sweep_l.txt
export SWEEP_ARGS="--gpt-model GPT-3B --image-size 384 --batch 16" SWEEP_STEPS=2 SWEEP_TIMEOUT=200
bash tools/sweep.sh "LG_X=0" "LG_LIB_PATH=$KC32" > $O/sweep_3b.txt 2>&1; cat $O/sweep_3b.txt
export SWEEP_ARGS="--batch 32" SWEEP_STEPS=4
bash tools/sweep.sh "LG_X=0" "LG_LIB_PATH=$KC32" > $O/sweep_b32.txt 2>&1; cat $O/sweep_b32.txt
