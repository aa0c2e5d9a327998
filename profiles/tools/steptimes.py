"""Per-step wall/device times of bench.py's Workload without a profiler (diagnostic).  Usage: [torchrun ...] steptimes.py [--sampler]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=None)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--sampler", action="store_true")
    ap.add_argument("--mode", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k = a.config if a.config is not None else (1 if world == 1 else 3)
    _, n, w, h = bench.CONFIGS[k]
    mode = a.mode or ("vanilla" if world == 1 else "gsplat")
    wl = bench.Workload(n, w, h, mode, rank, world, local, world > 1)
    for i in range(6):
        wl.step(i)
    sampler = None
    if a.sampler and rank == 0:
        sampler = bench.ClockSampler(local)
        sampler.start()
    wl.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    host = []
    ev[0].record()
    t_prev = time.perf_counter()
    for i in range(a.steps):
        wl.step(i)
        wl.step_done[i & 1].record()
        ev[i + 1].record()
        if i > 0:
            wl.step_done[(i - 1) & 1].synchronize()
        t = time.perf_counter()
        host.append((t - t_prev) * 1e3)
        t_prev = t
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.stop()
    dev = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    print(f"rank {rank} config {k} mode {mode} sampler {a.sampler}: total {ev[0].elapsed_time(ev[-1]):.2f} ms for {a.steps} steps = "
          f"{ev[0].elapsed_time(ev[-1]) / a.steps:.3f} ms/step\n  device per step: " + " ".join(f"{x:.2f}" for x in dev) +
          "\n  host per step:   " + " ".join(f"{x:.2f}" for x in host), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
