mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
run() {  # name, env...
  name=$1; shift
  (env "$@" timeout 150 $TR bench.py --gpus 8 --no-extras --steps 30 --warmup 5) > gpurun_out/r2_n8_$name.json 2> gpurun_out/r2_n8_$name.err
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2_n8_$name.json") if x.startswith("{")]
    d=json.loads(l[-1]); print("$name", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), round(d["value"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run default X=1
run simple NCCL_PROTO=Simple
run nvls NCCL_ALGO=NVLS,NVLSTree,Ring
run ll128 NCCL_PROTO=LL128,Simple
grep -h "NVLS\|Algo\|algo" gpurun_out/r2_n8_nvls.err | head -5
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 150 $TR bench.py --gpus 8 --no-extras --steps 3 --warmup 3 2>&1 | grep -i "nvls\|AllReduce.*algo\|Algo" | head -12
