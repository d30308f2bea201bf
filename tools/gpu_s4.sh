#!/bin/bash
# round-2 session 4: information gathering — chain-overlap timeline, persistent-kernel phase stamps, warm-cache DRAM traffic
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s4; mkdir -p $O
timeout 300 python tools/timeline.py --out $O/timeline_b64.json > $O/timeline_b64.log 2>&1
timeout 300 python tools/timeline.py --batch 32 --out $O/timeline_b32.json > $O/timeline_b32.log 2>&1
LG_SPLIT=1 timeout 300 python tools/timeline.py --out $O/timeline_b64_split1.json > $O/timeline_b64_split1.log 2>&1
LG_PERSIST=1 timeout 300 python tools/persist_probe.py GPT-L 1 > $O/persist_probe.txt 2>&1
LG_NO_GRAPH=1 timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2000 -c 10 -f -o $O/ncu_gemm_tc_warm python tools/ncu_target.py 24 64 > $O/ncu_gemm.log 2>&1
timeout 120 ncu -i $O/ncu_gemm_tc_warm.ncu-rep --page raw --csv > $O/ncu_gemm_tc_warm_raw.csv 2>> $O/ncu_gemm.log
LG_NO_GRAPH=1 timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:attn_tma_kernel -s 960 -c 4 -f -o $O/ncu_attn_warm python tools/ncu_target.py 24 64 > $O/ncu_attn.log 2>&1
timeout 120 ncu -i $O/ncu_attn_warm.ncu-rep --page raw --csv > $O/ncu_attn_warm_raw.csv 2>> $O/ncu_attn.log
cat $O/persist_probe.txt; tail -n 2 $O/*.log
