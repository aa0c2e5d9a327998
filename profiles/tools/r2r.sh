# call 11 (8 GPUs): the driver's N=8 bench line with the shared-memory camera board
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29671 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2r_bench8.log 2>&1; echo "bench8 rc=$?"
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2r_bench8.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("bench8", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["kernels"].items()}, d.get("single_gpu_same_workload"), d.get("other_workloads"))
except Exception as e:
    print("bench8 failed", e); print(open("gpurun_out/r2r_bench8.log").read()[-3000:])
PY
timeout 200 $TR --master-port 29672 profiles/tools/steptimes.py > gpurun_out/r2r_st8.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2r_st8.log | cut -c1-400
