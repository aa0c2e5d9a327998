mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I gaussian-splatting-lightning_b200/csrc -I include profiles/tools/sweep_bench.cu -o /tmp/sweep_bench > gpurun_out/r2h_nvcc.log 2>&1
timeout 300 /tmp/sweep_bench > gpurun_out/r2h_sweep.log 2>&1; echo "sweep rc=$?"
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_optim.py -q -m gpu -x > gpurun_out/r2h_dist.log 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/r2h_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29641 profiles/tools/steptimes.py > gpurun_out/r2h_st_plain.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2h_st_plain.log | cut -c1-400
B200GS_PEER_EXCHANGE=0 timeout 300 $TR --master-port 29643 profiles/tools/steptimes.py > gpurun_out/r2h_st_nccl.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2h_st_nccl.log | cut -c1-400
timeout 300 python profiles/tools/steptimes.py --config 3 --mode gsplat > gpurun_out/r2h_st_n1.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2h_st_n1.log | cut -c1-400
timeout 600 $TR --master-port 29642 profiles/tools/timeline.py --tag n2b > gpurun_out/r2h_tl2.log 2>&1; echo "tl2 rc=$?"
timeout 300 $TR --master-port 29644 bench.py --gpus 2 --steps 24 --warmup 4 > gpurun_out/r2h_bench2.log 2>&1
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2h_bench2.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("bench2", round(d["value"],1), d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}, d.get("single_gpu_same_workload"))
except Exception as e:
    print("bench2 failed", e); print(open("gpurun_out/r2h_bench2.log").read()[-2500:])
PY
cat gpurun_out/r2h_sweep.log
