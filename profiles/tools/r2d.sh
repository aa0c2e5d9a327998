# 2-GPU call: distributed tests (peer + NCCL exchange, oracle comparison), sharded bench at N=2, single-GPU same-workload line
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.log 2>&1
timeout 900 python -m pytest tests/test_gpu_distributed.py -q -m gpu -x > gpurun_out/r2d_dist.log 2>&1; echo "dist rc=$?"
tail -c 3000 gpurun_out/r2d_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 24 --warmup 4 > gpurun_out/r2d_bench2.log 2>&1; echo "bench2 rc=$?"
B200GS_PEER_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 24 --warmup 4 --no-extras > gpurun_out/r2d_bench2_nccl.log 2>&1; echo "bench2 nccl rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 4 --warmup 2 --impl reference > gpurun_out/r2d_bench2_ref.log 2>&1; echo "bench2 ref rc=$?"
for f in bench2 bench2_nccl bench2_ref; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2d_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],2), d.get("e2e"), d.get("gpu_launches"), {k:v["ms"] for k,v in d.get("kernels",{}).items()}, d.get("single_gpu_same_workload"), d.get("config"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2d_$f.log").read()[-2500:])
PY
done
