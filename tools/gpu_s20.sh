#!/bin/bash
# round-2 session 20: ncu launch list of the bench command, ncu --set full of the N = 256 conv, conv-variant test, b1 latency vs stream priority
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s20; mkdir -p $O
timeout 300 python -m pytest tests/test_vq_gpu.py -m gpu -q -x -k "variants" > $O/pytest_variants.log 2>&1; echo "rc=$?" >> $O/pytest_variants.log; tail -n 6 $O/pytest_variants.log
B="--no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --steps 2 --warmup 3"
LG_AR_PRIORITY=0 timeout 300 python bench.py $B > $O/bench_lat_prio0.json 2> $O/l0.err
python -c "
import json
d=json.loads(open('gpurun_out/s20/bench_lat_prio0.json').read().strip().splitlines()[-1]); print('prio0', d['value'], d['latency_b1'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 2500 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-gpu-reference --no-operating-points --no-latency > $O/ncu_launches.log 2>&1
wc -l $O/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tcw_kernel -s 20 -c 3 -f -o $O/ncu_conv_tcw python tools/bench_aux.py --iters 1 > $O/ncu_conv.log 2>&1
timeout 120 ncu -i $O/ncu_conv_tcw.ncu-rep --page raw --csv > $O/ncu_conv_tcw_raw.csv 2>> $O/ncu_conv.log
timeout 120 ncu -i $O/ncu_conv_tcw.ncu-rep --page details > $O/ncu_conv_tcw_details.txt 2>> $O/ncu_conv.log
grep -E "Duration|Compute \(SM\) Throughput|Memory Throughput|Grid Size|Executed Ipc Active|Achieved Occupancy" $O/ncu_conv_tcw_details.txt | head -20
