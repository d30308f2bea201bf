#!/bin/bash
# round-2 session 13: GPT-3B (C4) and GPT-XL (C3): chain count x stream priority x KV layout
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s13; mkdir -p $O
export SWEEP_ARGS="--gpt-model GPT-3B --image-size 384 --batch 16" SWEEP_STEPS=2 SWEEP_TIMEOUT=200
bash tools/sweep.sh "LG_HD_PAD=0 LG_SPLIT=2 LG_AR_PRIORITY=0" "LG_HD_PAD=0 LG_SPLIT=2 LG_AR_PRIORITY=1" "LG_HD_PAD=0 LG_SPLIT=1" "LG_HD_PAD=1 LG_SPLIT=1" "LG_HD_PAD=1 LG_SPLIT=2 LG_AR_PRIORITY=0" > $O/sweep_c4.txt 2>&1
cat $O/sweep_c4.txt
export SWEEP_ARGS="--gpt-model GPT-XL --image-size 384 --batch 32"
bash tools/sweep.sh "LG_SPLIT=2" "LG_SPLIT=1" "LG_SPLIT=2 LG_AR_PRIORITY=0" > $O/sweep_c3.txt 2>&1
cat $O/sweep_c3.txt
