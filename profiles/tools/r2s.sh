# call 13 (4 GPUs): the driver's N=4 bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29681 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2s_bench4.log 2>&1; echo "bench4 rc=$?"
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2s_bench4.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("bench4", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["kernels"].items()}, d.get("single_gpu_same_workload"))
except Exception as e:
    print("bench4 failed", e); print(open("gpurun_out/r2s_bench4.log").read()[-3000:])
PY
