mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "binning or culling or lazy or rows or end_to_end" > gpurun_out/r2f_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2f_parity.log
timeout 900 python -m pytest tests/test_gpu_distributed.py -q -m gpu -x > gpurun_out/r2f_dist.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r2f_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29631 profiles/tools/steptimes.py > gpurun_out/r2f_st_plain.log 2>&1; grep -A2 "^rank" gpurun_out/r2f_st_plain.log
timeout 300 $TR --master-port 29632 profiles/tools/steptimes.py --sampler > gpurun_out/r2f_st_sampler.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2f_st_sampler.log
B200GS_PEER_EXCHANGE=0 timeout 300 $TR --master-port 29633 profiles/tools/steptimes.py > gpurun_out/r2f_st_nccl.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2f_st_nccl.log
timeout 300 python profiles/tools/steptimes.py --config 3 --mode gsplat > gpurun_out/r2f_st_n1.log 2>&1; grep -A2 "^rank 0" gpurun_out/r2f_st_n1.log
timeout 300 $TR --master-port 29634 bench.py --gpus 2 --steps 24 --warmup 4 --no-extras > gpurun_out/r2f_bench2.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/r2f_bench2.log") if x.startswith("{")][-1]; d=json.loads(l)
print("bench2", round(d["value"],1), d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
