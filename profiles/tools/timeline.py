"""GPU/host timeline of benchmark steps (there is no nsys in the image): torch.profiler (CUPTI) around a few steps of bench.py's
Workload, reduced to a text table per rank: every kernel / memcpy of ONE step with its start (us, relative to the step's first
kernel), duration and stream, the gaps between consecutive device activities, and the host-side spans (record_function ranges and
the longest CPU ops).  Usage (N = 1 or under torchrun):  python profiles/tools/timeline.py [--config K] [--steps S] [--tag T]
Writes gpurun_out/timeline_<tag>_rank<r>.txt.  Numbers under a profiler are for reading the structure, never bench values."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=None)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--tag", default="t")
    ap.add_argument("--mode", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k = a.config if a.config is not None else (1 if world == 1 else 3)
    _, n, w, h = bench.CONFIGS[k]
    cfg = {"n": n, "w": w, "h": h}
    mode = a.mode or ("vanilla" if world == 1 else "gsplat")
    wl = bench.Workload(n, w, h, mode, rank, world, local, world > 1)
    for i in range(6):
        wl.step(i)
    wl.barrier()
    from torch.profiler import ProfilerActivity, profile, record_function
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(a.steps):
            with record_function(f"STEP{i}"):
                wl.step(i)
                wl.step_done[i & 1].record()
                if i > 0:
                    wl.step_done[(i - 1) & 1].synchronize()
        torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    trace = os.path.join(ROOT, "gpurun_out", f"trace_{a.tag}_rank{rank}.json")
    prof.export_chrome_trace(trace)
    ev = json.load(open(trace))["traceEvents"]
    os.remove(trace)
    dev = sorted((e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e), key=lambda e: e["ts"])
    steps = sorted((e for e in ev if e.get("cat") in ("user_annotation", "cpu_op") and str(e.get("name", "")).startswith("STEP")
                    and e.get("ph") == "X"), key=lambda e: e["ts"])
    steps = [s for s in steps if s.get("cat") == "user_annotation"] or steps
    out = [f"# rank {rank}/{world}, config {k} ({cfg['n']} Gaussians, {cfg['w']}x{cfg['h']}), mode {mode}; {len(dev)} device activities in {a.steps} steps",
           "# host-side step spans (us): " + ", ".join(f"{s['name']} {s['dur']:.0f}" for s in steps)]
    # device activities launched from the host span of the second-to-last step (correlate by time: those that start after the step's
    # first launch; simpler and robust: split the device list at the first project_fwd kernel of every step)
    starts = [i for i, e in enumerate(dev) if "project_fwd" in e["name"]]
    if len(starts) >= 3:
        lo, hi = starts[-2], starts[-1]
        seg = dev[lo:hi]
        t0 = seg[0]["ts"]
        out.append(f"# one step on the device: {seg[-1]['ts'] + seg[-1]['dur'] - t0:.0f} us from its first kernel to the end of its last; "
                   f"next step's first kernel at {dev[hi]['ts'] - t0:.0f} us; busy {sum(e['dur'] for e in seg):.0f} us")
        out.append("start_us  dur_us  gap_before_us  stream  name")
        prev_end = t0
        for e in seg:
            gap = e["ts"] - prev_end
            name = e["name"].replace("void ", "").replace("b200gs::(anonymous namespace)::", "").replace("b200gs::", "")
            name = name.split("(")[0][:100]
            out.append(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} {gap:9.1f}  {e.get('args', {}).get('stream', '?'):>4}  {name}")
            prev_end = max(prev_end, e["ts"] + e["dur"])
    # host side of the same step: the longest CPU ops / runtime calls
    if len(steps) >= 2:
        s = steps[-2]
        h = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("cpu_op", "cuda_runtime", "cuda_driver", "user_annotation", "python_function")
             and s["ts"] <= e["ts"] <= s["ts"] + s["dur"] and e is not s]
        h.sort(key=lambda e: -e["dur"])
        out.append("")
        out.append(f"# host, {s['name']} ({s['dur']:.0f} us): the 40 longest spans (start relative to the step, dur, name)")
        for e in h[:40]:
            out.append(f"{e['ts'] - s['ts']:9.1f} {e['dur']:8.1f}  {e['cat']:14s} {str(e['name'])[:90]}")
    path = os.path.join(ROOT, "gpurun_out", f"timeline_{a.tag}_rank{rank}.txt")
    open(path, "w").write("\n".join(out) + "\n")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    print("written", path)


if __name__ == "__main__":
    main()
