mkdir -p gpurun_out
B200GS_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -q -m gpu > gpurun_out/r2a_exp.log 2>&1; echo "exp rc=$?"
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2a_gpu_parity.log 2>&1; echo "parity rc=$?"
timeout 1200 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_reference_dropin.py tests/test_gpu_v1_surface.py -q -m gpu > gpurun_out/r2a_gpu_full.log 2>&1; echo "full rc=$?"
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2a_bench_tr.log 2>&1
B200GS_BWD_BUTTERFLY=1 timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2a_bench_bf.log 2>&1
B200GS_FWD_SYNC=1 timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2a_bench_sync.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2a_bench_c0.log 2>&1
tail -c 1500 gpurun_out/r2a_exp.log; tail -c 1500 gpurun_out/r2a_gpu_parity.log; tail -c 1500 gpurun_out/r2a_gpu_full.log
for f in tr bf sync c0; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2a_bench_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2a_bench_$f.log").read()[-1500:])
PY
done
