#!/bin/bash
# round-2 session 11: attention ring variants; BASELINE configs C3/C4/C5 as bench operating points; single-pass DRAM traffic of the decode loop
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s11; mkdir -p $O
KC32=$PWD/llamagen_b200/lib_kc32/libllamagen_b200.so
bash tools/sweep.sh "LG_ATTN_NST=2" "LG_ATTN_NST=3" "LG_LIB_PATH=$KC32 LG_ATTN_NST=3" "LG_LIB_PATH=$KC32 LG_ATTN_NST=4" > $O/sweep_attn.txt 2>&1
cat $O/sweep_attn.txt
F="--no-cpu-baseline --no-gpu-reference --no-operating-points --no-latency --steps 3 --warmup 3"
timeout 400 python bench.py $F --gpt-model GPT-XL --image-size 384 --batch 32 > $O/bench_c3_xl_384_b32.json 2> $O/c3.err
timeout 400 python bench.py $F --gpt-model GPT-3B --image-size 384 --batch 16 > $O/bench_c4_3b_384_b16.json 2> $O/c4.err
timeout 400 python bench.py $F --gpt-model GPT-XL --image-size 512 --batch 8 --t2i --cfg-scale 7.5 --top-k 1000 > $O/bench_c5_xl_t2i_512_b8.json 2> $O/c5.err
for f in $O/bench_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, json.dumps(d.get("roofline"))[:300], json.dumps(d.get("step_roofline"))[:300], json.dumps(d.get("prefill"))[:200], json.dumps(d.get("kv_cache"))[:200])
except Exception as e:
    print("ERR", e)
PY
done
LG_NO_GRAPH=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -s 6000 -c 800 --csv --log-file $O/dram_single_pass.csv python tools/ncu_target.py 40 64 > $O/ncu_single.log 2>&1
tail -n 3 $O/ncu_single.log; wc -l $O/dram_single_pass.csv
