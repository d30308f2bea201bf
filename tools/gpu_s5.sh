#!/bin/bash
# round-2 session 5: chain-count / ring-depth / CTA-target sweep of the bench workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s5; mkdir -p $O
bash tools/sweep.sh "LG_SPLIT=2" "LG_SPLIT=4" "LG_SPLIT=4 LG_TC_STAGES=3" "LG_SPLIT=4 LG_TC_STAGES=2" "LG_SPLIT=4 LG_TC_CTAS=64" \
  "LG_SPLIT=4 LG_TC_CTAS=64 LG_TC_STAGES=3" "LG_SPLIT=4 LG_TC_CTAS=148" "LG_SPLIT=8 LG_ATTN_DEEP=0" "LG_SPLIT=8 LG_ATTN_DEEP=0 LG_TC_STAGES=2" \
  "LG_SPLIT=8 LG_ATTN_DEEP=0 LG_TC_CTAS=64 LG_TC_STAGES=3" "LG_SPLIT=8" "LG_SPLIT=2 LG_TC_CTAS=148" "LG_SPLIT=4 LG_TC_CTAS=48 LG_TC_STAGES=4" > $O/sweep.txt 2>&1
cat $O/sweep.txt
