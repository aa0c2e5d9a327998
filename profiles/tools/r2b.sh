mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loss.py tests/test_gpu_optim.py -q -m gpu > gpurun_out/r2b_gpu_parity.log 2>&1; echo "parity rc=$?"
timeout 1200 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_reference_dropin.py tests/test_gpu_v1_surface.py -q -m gpu > gpurun_out/r2b_gpu_full.log 2>&1; echo "full rc=$?"
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2b_bench.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2b_bench_c0.log 2>&1
timeout 300 python bench.py --steps 32 --warmup 6 --no-cpu-baseline --no-extras --mode gsplat > gpurun_out/r2b_bench_gsplat.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd_tr|blend_fwd_kernel|onesweep_pass|depth_keys|emit_cells|project_fwd|project_bwd|chunk_counts|chunk_prefix|scatter_ids|hist4|cell_table" -s 120 -c 22 -o gpurun_out/r2b_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_ncu_full.log 2>&1
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from b200gs.renderers import B200VanillaRenderer, B200GSplatRenderer
from b200gs.scene import SyntheticGaussians, make_ring_cameras, make_scene
dev = "cuda"
model = SyntheticGaussians(make_scene(6000, 3, mean_scale=0.03)).to(dev)
cam = make_ring_cameras(320, 240)[2].to_device(dev)
bg = torch.zeros(3, device=dev)
for R in (B200VanillaRenderer(), B200GSplatRenderer()):
    for _ in range(2):
        out = R(cam, model, bg)
        out["render"].sum().backward()
torch.cuda.synchronize()
print("sanitizer workload done")
PY
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py > gpurun_out/r2b_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -c 1200 gpurun_out/r2b_gpu_parity.log; tail -c 2500 gpurun_out/r2b_gpu_full.log
for f in bench bench_c0 bench_gsplat; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2b_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2b_$f.log").read()[-1500:])
PY
done
tail -5 gpurun_out/r2b_racecheck.log; tail -5 gpurun_out/r2b_memcheck.log; tail -3 gpurun_out/r2b_ncu_full.log
