#!/bin/bash
# round-2 session 17: localise the illegal access in the continuous-batching test; GN-stats fusion + RoPE hoist parity; L2 hints
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s17; mkdir -p $O
LG_FUSE_TAIL=0 timeout 300 python -m pytest tests/test_cli_gpu.py -m gpu -q -x -k "continuous_batching" > $O/cb_tail0.log 2>&1; echo "rc=$?" >> $O/cb_tail0.log; tail -n 3 $O/cb_tail0.log
LG_FUSE_TAIL=1 LG_DEBUG_SYNC=1 timeout 600 python -m pytest tests/test_cli_gpu.py -m gpu -q -x -k "continuous_batching" > $O/cb_tail1_sync.log 2>&1; echo "rc=$?" >> $O/cb_tail1_sync.log
grep -n "launched" $O/cb_tail1_sync.log | grep -v "no error" | head -5; tail -n 5 $O/cb_tail1_sync.log
timeout 600 python -m pytest tests/test_vq_gpu.py -m gpu -q -x > $O/pytest_vq.log 2>&1; echo "rc=$?" >> $O/pytest_vq.log; tail -n 5 $O/pytest_vq.log
timeout 300 python tools/bench_aux.py > $O/aux_gnfuse1.json 2> $O/aux1.err
LG_GN_FUSE=0 timeout 300 python tools/bench_aux.py > $O/aux_gnfuse0.json 2> $O/aux0.err
for f in $O/aux_gnfuse1.json $O/aux_gnfuse0.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['decode'], d['encode']['ms'])"; done
timeout 400 python -m pytest tests/test_gpt_gpu.py -m gpu -q -x -k "teacher or multi_chain or small_row" > $O/pytest_gpt.log 2>&1; echo "rc=$?" >> $O/pytest_gpt.log; tail -n 4 $O/pytest_gpt.log
bash tools/sweep.sh "LG_L2_HINT=0" "LG_L2_HINT=1" "LG_L2_HINT=3" "LG_L2_HINT=3 LG_GN_FUSE=0" > $O/sweep.txt 2>&1; cat $O/sweep.txt
