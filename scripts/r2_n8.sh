# 8-GPU bench (with extras), per-step timeline and the same-box 1-GPU number: gpurun --gpus 8 -- bash scripts/r2_n8.sh
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
(time timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
python - <<PY
import json
l=[x for x in open("gpurun_out/r2_bench_n8.json") if x.startswith("{")]
d=json.loads(l[-1]); print("N=8", d["value"], d["ms_per_step"], d["e2e"]); print(json.dumps(d["extra_keys"])[:1800])
PY
timeout 200 $TR scripts/dp_timeline.py gpurun_out/r2_dp8_timeline.txt > gpurun_out/r2_dp8.log 2>&1; head -8 gpurun_out/r2_dp8_timeline.txt
timeout 120 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 same box', d['ms_per_step'], d['e2e']['ms_per_step'])"
