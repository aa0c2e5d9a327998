# call 15 (1 GPU): cp.async forward staging at 3 blocks per SM (register cap) vs synchronous staging; GPU tests under both
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2u_tests_sync.log 2>&1; echo "tests(sync) rc=$?"; tail -1 gpurun_out/r2u_tests_sync.log
B200GS_FWD_ASYNC=1 timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2u_tests_async.log 2>&1; echo "tests(async) rc=$?"; tail -1 gpurun_out/r2u_tests_async.log
timeout 200 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2u_bench_sync.log 2>&1
B200GS_FWD_ASYNC=1 timeout 200 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2u_bench_async.log 2>&1
for f in bench_sync bench_async; do python - <<PY
import json
l=[x for x in open("gpurun_out/r2u_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
print("$f", round(d["value"],1), round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["kernels"].items()})
PY
done
