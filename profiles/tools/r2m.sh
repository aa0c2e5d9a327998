# call 7 (1 GPU): explicit-rounding forward projection (bit-identical K1 variants), v_mean2d written by K8, e2e input wait moved to the loss
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2m_tests.log | cut -c1-800
for k in 1 2; do
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2m_bench$k.log 2>&1
done
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --mode gsplat > gpurun_out/r2m_bench_gs.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2m_bench_c0.log 2>&1
for f in bench1 bench2 bench_gs bench_c0; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2m_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2m_$f.log").read()[-2500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2m_ncu_launch.log 2>&1
python profiles/tools/launch_list.py gpurun_out/r2m_launches.csv 2>/dev/null | head -45
