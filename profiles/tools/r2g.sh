mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_v1_surface.py tests/test_gpu_optim.py -q -m gpu -x > gpurun_out/r2g_parity.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/r2g_parity.log
timeout 1200 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_reference_dropin.py -q -m gpu -x > gpurun_out/r2g_full.log 2>&1; echo "full rc=$?"; tail -4 gpurun_out/r2g_full.log
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --mode gsplat > gpurun_out/r2g_bench_gs.log 2>&1
timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-extras --config 3 --mode gsplat > gpurun_out/r2g_bench_c3.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2g_bench_c0.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_ncu_launch.log 2>&1
for f in bench bench_gs bench_c3 bench_c0; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2g_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2g_$f.log").read()[-1500:])
PY
done
