# call 6 (2 GPUs): ballot ranking + fixed-8 look-back in the onesweep passes, wide first look-back step, K1 SH prefetch, cov3D hoisted out of the camera loop,
# smem SH-gradient accumulators in K8-multi; diagnostic of the 1-ulp sharded-vs-single image difference
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_distributed.py > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log | cut -c1-800
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29651 profiles/tools/dbg_shard.py > gpurun_out/r2l_dbg.log 2>&1; echo "dbg rc=$?"; grep "^\[rank" gpurun_out/r2l_dbg.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_distributed.py -q -m gpu -x > gpurun_out/r2l_dist.log 2>&1; echo "dist rc=$?"; grep -E "AssertionError|passed|failed" gpurun_out/r2l_dist.log | head -5 | cut -c1-400
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2l_bench.log 2>&1
B200GS_K1_PREFETCH=0 timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2l_bench_nopf.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2l_bench_c0.log 2>&1
timeout 400 $TR --master-port 29654 bench.py --gpus 2 --steps 24 --warmup 4 > gpurun_out/r2l_bench2.log 2>&1
for f in bench bench_nopf bench_c0 bench2; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2l_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["kernels"].items()}, d.get("single_gpu_same_workload"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2l_$f.log").read()[-2500:])
PY
done
