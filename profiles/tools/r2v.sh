# last call of round 2 (1 GPU): the committed default (row layout -> cp.async staging): smoke, parity tests, one bench line
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2v_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py tests/test_gpu_sharded_kernels.py tests/test_gpu_reference_dropin.py -q -m gpu -x > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r2v_tests.log
timeout 200 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2v_bench.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/r2v_bench.log") if x.startswith("{")][-1]; d=json.loads(l)
print("bench", round(d["value"],1), round(d["e2e"]["value"],1), d["gpu_launches"], {k:v["ms"] for k,v in d["kernels"].items()}, d["roofline"]["frac"], d["roofline_issue"]["frac"])
PY
