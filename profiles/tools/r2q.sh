# final call of round 2 (1 GPU): full GPU test suite, the bench lines the driver asks for, launch list + full ncu capture, sanitizers
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2q_tests.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2q_bench_default.log 2>&1; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2q_bench_reference.log 2>&1; echo "bench reference rc=$?"
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --mode gsplat > gpurun_out/r2q_bench_gs.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2q_bench_c0.log 2>&1
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras --config 2 > gpurun_out/r2q_bench_c2.log 2>&1
for f in bench_default bench_gs bench_c0 bench_c2; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2q_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), d["ms_per_step"], round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()}, d.get("roofline",{}).get("frac"), d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2q_$f.log").read()[-2500:])
PY
done
tail -c 600 gpurun_out/r2q_bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2q_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd_tr|blend_fwd_kernel|onesweep_pass|depth_keys|rank_offsets|emit_cells|project_fwd|project_bwd|chunk_counts|chunk_prefix|scatter_ids|cell_table" -s 72 -c 18 -o gpurun_out/r2q_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2q_ncu_full.log 2>&1; echo "ncu full rc=$?"
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from b200gs.renderers import B200VanillaRenderer, B200GSplatRenderer
from b200gs.v1 import B200GSplatV1Renderer
from b200gs.scene import SyntheticGaussians, make_ring_cameras, make_scene
from b200gs import ops
dev = "cuda"
model = SyntheticGaussians(make_scene(6000, 3, mean_scale=0.03)).to(dev)
cam = make_ring_cameras(320, 240)[2].to_device(dev)
bg = torch.zeros(3, device=dev)
for R in (B200VanillaRenderer(), B200GSplatRenderer(), B200GSplatRenderer(fused_activations=False, absgrad=True), B200GSplatV1Renderer(tile_based_culling=True).instantiate()):
    for _ in range(2):
        for p in model.parameters():
            p.grad = None
        out = R(cam, model, bg)
        out["render"].sum().backward()
img = out["render"].detach().clone().requires_grad_(True)
loss, _ = ops.l1_ssim_loss(img, torch.rand_like(img), 0.2)
loss.backward()
ops.knn_mean_dist2(model.gaussians["means"].detach())
# the multi-view kernels of the sharded renderer (single GPU): forward variants + fused pack
import test_gpu_sharded_kernels as T
T.test_multi_view_kernels_match_single_view_rows_bitwise(True)
T.test_compacted_and_blocked_rows_render_bit_identically()
T.test_multi_view_backward_matches_accumulated_single_view_backward()
torch.cuda.synchronize()
print("sanitizer workload done")
PY
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py > gpurun_out/r2q_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/r2q_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r2q_racecheck.log; tail -3 gpurun_out/r2q_memcheck.log
