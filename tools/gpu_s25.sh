#!/bin/bash
# round-2 session 25: last sanity of the committed tree (smoke + quick bench)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s25; mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -n 2 $O/smoke.log
bash tools/sweep.sh "LG_X=0" > $O/sweep.txt 2>&1; cat $O/sweep.txt
