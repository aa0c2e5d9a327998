# call 5 of round 2: striped depth_keys, warp-cooperative SH rows in K1/K8, chunk_prefix/cell_table latency trims; K1-variant consistency test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2k_tests.log
timeout 600 python -m pytest tests/test_gpu_sharded_kernels.py -q -m gpu > gpurun_out/r2k_shk.log 2>&1; echo "sharded-kernels rc=$?"; grep -E "differ|passed|failed" gpurun_out/r2k_shk.log | cut -c1-600
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench.log 2>&1
B200GS_K1_COOP=0 B200GS_K8_COOP=0 timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench_nocoop.log 2>&1
timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-extras --config 0 > gpurun_out/r2k_bench_c0.log 2>&1
for f in bench bench_nocoop bench_c0; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2k_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d.get("gpu_launches"), {k:v["ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2k_$f.log").read()[-1500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 90 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd_tr|blend_fwd_kernel|onesweep_pass|depth_keys|rank_offsets|emit_cells|project_fwd|project_bwd|chunk_counts|chunk_prefix|scatter_ids|cell_table" -s 72 -c 18 -o gpurun_out/r2k_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_ncu_full.log 2>&1; echo "ncu full rc=$?"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I gaussian-splatting-lightning_b200/csrc -I include profiles/tools/sweep_bench.cu -o /tmp/sweep_bench > gpurun_out/r2k_nvcc.log 2>&1
timeout 300 /tmp/sweep_bench > gpurun_out/r2k_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r2k_sweep.log | head -60
