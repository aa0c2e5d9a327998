# call 9 (1 GPU): longest-tile-first order for K6/K7 (A/B), emit_cells item table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2o_tests.log | cut -c1-800
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2o_bench.log 2>&1
B200GS_TILE_ORDER=0 timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2o_bench_noorder.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2o_bench_c0.log 2>&1
timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-extras --config 3 --mode gsplat > gpurun_out/r2o_bench_c3.log 2>&1
for f in bench bench_noorder bench_c0 bench_c3; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2o_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2o_$f.log").read()[-2500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2o_ncu_launch.log 2>&1
python profiles/tools/launch_list.py gpurun_out/r2o_launches.csv 2>/dev/null | head -45
