#!/bin/bash
# round-2 session 7: timelines of the direct-epilogue path (why is it slower?)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s7; mkdir -p $O
LG_DIRECT=1 timeout 200 python tools/timeline.py --out $O/timeline_dx_b64.json > $O/t1.log 2>&1
LG_DIRECT=1 LG_SPLIT=1 timeout 200 python tools/timeline.py --batch 32 --out $O/timeline_dx_b32_split1.json > $O/t2.log 2>&1
LG_DIRECT=0 LG_SPLIT=1 timeout 200 python tools/timeline.py --batch 32 --out $O/timeline_slab_b32_split1.json > $O/t3.log 2>&1
LG_DIRECT=1 LG_SPLIT=1 LG_DX_STAGES=2 timeout 200 python tools/timeline.py --batch 32 --out $O/timeline_dx_b32_split1_st2.json > $O/t4.log 2>&1
tail -n 1 $O/t*.log
