mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loss.py tests/test_gpu_optim.py tests/test_gpu_v1_surface.py -q -m gpu > gpurun_out/r2c_gpu_parity.log 2>&1; echo "parity rc=$?"
timeout 1200 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_reference_dropin.py -q -m gpu > gpurun_out/r2c_gpu_full.log 2>&1; echo "full rc=$?"
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline > gpurun_out/r2c_bench.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2c_bench_c0.log 2>&1
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras --config 2 > gpurun_out/r2c_bench_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2c_ncu_launch.log 2>&1
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from b200gs.renderers import B200VanillaRenderer, B200GSplatRenderer
from b200gs.v1 import B200GSplatV1Renderer
from b200gs.scene import SyntheticGaussians, make_ring_cameras, make_scene
from b200gs import ops
dev = "cuda"
model = SyntheticGaussians(make_scene(6000, 3, mean_scale=0.03)).to(dev)
cam = make_ring_cameras(320, 240)[2].to_device(dev)
bg = torch.zeros(3, device=dev)
for R in (B200VanillaRenderer(), B200GSplatRenderer(), B200GSplatV1Renderer(tile_based_culling=True).instantiate()):
    for _ in range(2):
        out = R(cam, model, bg)
        out["render"].sum().backward()
img = out["render"].detach().clone().requires_grad_(True)
loss, _ = ops.l1_ssim_loss(img, torch.rand_like(img), 0.2)
loss.backward()
ops.knn_mean_dist2(model.gaussians["means"].detach())
torch.cuda.synchronize()
print("sanitizer workload done")
PY
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py > gpurun_out/r2c_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/r2c_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -c 1500 gpurun_out/r2c_gpu_parity.log; tail -c 1500 gpurun_out/r2c_gpu_full.log
for f in bench bench_c0 bench_c2; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2c_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()}, d.get("loss_stage"), d.get("scaling_base"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2c_$f.log").read()[-1500:])
PY
done
tail -4 gpurun_out/r2c_racecheck.log; tail -3 gpurun_out/r2c_memcheck.log
