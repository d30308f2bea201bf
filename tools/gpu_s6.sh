#!/bin/bash
# round-2 session 6: direct-epilogue GEMM (gemm_dx.cu) — unit parity, model parity, A/B + chain-count sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s6; mkdir -p $O
timeout 400 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "dx" > $O/pytest_dx.log 2>&1; echo "rc=$?" >> $O/pytest_dx.log
tail -n 15 $O/pytest_dx.log
timeout 600 python -m pytest tests/test_gpt_gpu.py -m gpu -q -k "direct or multi_chain or l_bf16_teacher or bf16_teacher_forced_vs_oracle" > $O/pytest_gpt.log 2>&1; echo "rc=$?" >> $O/pytest_gpt.log
tail -n 15 $O/pytest_gpt.log
bash tools/sweep.sh "LG_DIRECT=0" "LG_DIRECT=1" "LG_DIRECT=1 LG_DX_RBLK=16" "LG_DIRECT=1 LG_DX_RBLK=64" "LG_DIRECT=1 LG_DX_STAGES=4" "LG_DIRECT=1 LG_DX_STAGES=6" \
  "LG_DIRECT=0 LG_SPLIT=4" "LG_DIRECT=1 LG_SPLIT=4" "LG_DIRECT=0 LG_SPLIT=4 LG_TC_STAGES=2" "LG_DIRECT=1 LG_SPLIT=4 LG_TC_STAGES=2 LG_DX_STAGES=4" \
  "LG_DIRECT=0 LG_SPLIT=8 LG_ATTN_DEEP=0" "LG_DIRECT=1 LG_SPLIT=8 LG_ATTN_DEEP=0 LG_DX_RBLK=16" "LG_DIRECT=1 LG_SPLIT=1" > $O/sweep.txt 2>&1
cat $O/sweep.txt
