mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2i_tests.log
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2i_bench.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2i_bench_c0.log 2>&1
timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-extras --config 3 --mode gsplat > gpurun_out/r2i_bench_c3.log 2>&1
for f in bench bench_c0 bench_c3; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2i_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2i_$f.log").read()[-1500:])
PY
done
