# 2-GPU validation of the sharded renderer: distributed tests, step times (peer exchange vs NCCL), bench --gpus 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -q -m gpu -x > gpurun_out/r2j_dist.log 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/r2j_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29641 profiles/tools/steptimes.py > gpurun_out/r2j_st_peer.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2j_st_peer.log | cut -c1-400
B200GS_PEER_EXCHANGE=0 timeout 300 $TR --master-port 29643 profiles/tools/steptimes.py > gpurun_out/r2j_st_nccl.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2j_st_nccl.log | cut -c1-400
timeout 400 $TR --master-port 29644 bench.py --gpus 2 --steps 24 --warmup 4 > gpurun_out/r2j_bench2.log 2>&1
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2j_bench2.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("bench2", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["kernels"].items()}, d.get("single_gpu_same_workload"))
except Exception as e:
    print("bench2 failed", e); print(open("gpurun_out/r2j_bench2.log").read()[-2500:])
PY
