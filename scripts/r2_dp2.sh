mkdir -p gpurun_out
python -m pytest tests/test_dp_nccl_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_dp_test.log 2>&1; grep -n "worst per-tensor\|passed\|failed\|^E  " gpurun_out/r2_dp_test.log | cut -c1-300 | head
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
python bench.py --no-extras --no-cpu-baseline --steps 40 > gpurun_out/r2_b1.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r2_b1.json'));print('N=1', d['ms_per_step'], d['e2e']['ms_per_step'])"
for cta in default 4 8 16; do
  if [ $cta = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$cta; fi
  $TR bench.py --gpus 2 --no-extras --steps 40 > gpurun_out/r2_b2_$cta.json 2>gpurun_out/r2_b2_$cta.err
  python -c "import json;d=json.load(open('gpurun_out/r2_b2_$cta.json'));print('N=2 ctas $cta', d['ms_per_step'], d['e2e']['ms_per_step'], d['extra_keys'])"
done
unset NCCL_MAX_CTAS
$TR scripts/dp_timeline.py gpurun_out/r2_dp2_timeline.txt > gpurun_out/r2_dp2.log 2>&1; head -8 gpurun_out/r2_dp2_timeline.txt
