mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
for mode in bf16 fp32 bf16 fp32; do
  LG_REDUCE_DTYPE=$mode timeout 150 $TR bench.py --gpus 8 --no-extras --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['value']))"
done
LG_REDUCE_DTYPE=bf16 timeout 200 $TR scripts/dp_timeline.py gpurun_out/r2_dp8_timeline.txt > gpurun_out/r2_dp8.log 2>&1; head -8 gpurun_out/r2_dp8_timeline.txt
timeout 120 python bench.py --no-extras --no-cpu-baseline --steps 30 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 same box', d['ms_per_step'], d['e2e']['ms_per_step'])"
