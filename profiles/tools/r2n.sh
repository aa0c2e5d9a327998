# call 8 (2 GPUs): distributed parity after the explicit-rounding projection; K8 with smem SH-gradient rows; ncells in the sort payload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2n_tests.log | cut -c1-800
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2n_bench.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --mode gsplat > gpurun_out/r2n_bench_gs.log 2>&1
timeout 400 $TR --master-port 29654 bench.py --gpus 2 --steps 24 --warmup 4 > gpurun_out/r2n_bench2.log 2>&1
B200GS_K1_PREFETCH=1 timeout 400 $TR --master-port 29655 bench.py --gpus 2 --steps 24 --warmup 4 --no-extras > gpurun_out/r2n_bench2_pf.log 2>&1
for f in bench bench_gs bench2 bench2_pf; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2n_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["kernels"].items()}, d.get("single_gpu_same_workload"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2n_$f.log").read()[-2500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2n_ncu_launch.log 2>&1
python profiles/tools/launch_list.py gpurun_out/r2n_launches.csv 2>/dev/null | head -45
