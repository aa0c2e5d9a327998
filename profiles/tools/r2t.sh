# call 14 (1 GPU): forward blend kernel, synchronous vs cp.async staging: ncu --set full of each + timing of each
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd" -s 4 -c 1 -o gpurun_out/r2t_sync python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2t_ncu_sync.log 2>&1; echo "sync rc=$?"
B200GS_FWD_ASYNC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd" -s 4 -c 1 -o gpurun_out/r2t_async python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2t_ncu_async.log 2>&1; echo "async rc=$?"
timeout 200 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2t_bench_sync.log 2>&1
B200GS_FWD_ASYNC=1 timeout 200 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2t_bench_async.log 2>&1
for f in bench_sync bench_async; do python - <<PY
import json
l=[x for x in open("gpurun_out/r2t_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
print("$f", round(d["value"],1), round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["kernels"].items()}, d["roofline"].get("traffic"), d.get("roofline_issue"))
PY
done
