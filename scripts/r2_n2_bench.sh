mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
(time timeout 420 $TR bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -4 gpurun_out/r2_bench_n2.err
python - <<PY
import json
l=[x for x in open("gpurun_out/r2_bench_n2.json") if x.startswith("{")]
d=json.loads(l[-1]); print("N=2", d["value"], d["ms_per_step"], d["e2e"]); print(json.dumps(d["extra_keys"])[:1500])
PY
(time timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1) > gpurun_out/r2_ref_n2.json 2> gpurun_out/r2_ref_n2.err
tail -3 gpurun_out/r2_ref_n2.err; cut -c1-900 gpurun_out/r2_ref_n2.json
