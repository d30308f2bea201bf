#!/bin/bash
# round-2 session 15: weights-as-A conv kernel (UMMA N = 256): parity, VQ decode / encode time, bench A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s15; mkdir -p $O
timeout 600 python -m pytest tests/test_vq_gpu.py -m gpu -q -x > $O/pytest_vq.log 2>&1; echo "rc=$?" >> $O/pytest_vq.log; tail -n 12 $O/pytest_vq.log
LG_CONV_SWAP=1 timeout 300 python tools/bench_aux.py > $O/aux_swap1.json 2> $O/aux1.err
LG_CONV_SWAP=0 timeout 300 python tools/bench_aux.py > $O/aux_swap0.json 2> $O/aux0.err
for f in $O/aux_swap1.json $O/aux_swap0.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print({k:d[k] for k in d if k in ('decode','encode')})"; done
bash tools/sweep.sh "LG_CONV_SWAP=1" "LG_CONV_SWAP=0" > $O/sweep_swap.txt 2>&1; cat $O/sweep_swap.txt
