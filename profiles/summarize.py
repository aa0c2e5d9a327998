"""Turn the ncu artefacts of a gpurun call (gpurun_out/launches.csv, gpurun_out/prof_blend_{fwd,bwd}.ncu-rep) into the
tracked summaries under profiles/.  Usage: python profiles/summarize.py <round-tag>
Captured with (B200, one GPU, see /opt/skills/guides/B200_PROFILING.md):
  ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 90 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline
  ncu --set full --clock-control none --import-source on -k regex:blend_bwd -s 3 -c 1 -o gpurun_out/prof_blend_bwd   python bench.py --steps 3 --warmup 2 --no-cpu-baseline
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "round1"

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_global_red.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum"]


TRAFFIC = {}


def ncu_csv(rep, page):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout)))


def kernel_summary(name):
    rep = os.path.join(OUT, f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    lines = [f"# ncu --set full, {name} (one launch of the benchmark step: 1 M Gaussians, 1920x1080, vanilla mode)", "",
             f"kernel: `{vals[hdr.index('Kernel Name')]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f"| {w} | {vals[i]} | {units[i]} |")
    src = ncu_csv(rep, "source")
    h, data = src[1], src[2:]
    ia, isrc = h.index("Instructions Executed"), h.index("Source")
    tot = sum(int(r[ia]) for r in data)
    mix = collections.Counter()
    for r in data:
        t = r[isrc].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        mix[op] += int(r[ia])
    lines += ["", f"SASS opcode mix (share of {tot} executed warp instructions):", "",
              ", ".join(f"{op} {c / tot * 100:.1f}%" for op, c in mix.most_common(20))]
    open(os.path.join(ROOT, "profiles", f"{tag}_{name}_ncu.md"), "w").write("\n".join(lines) + "\n")

    def to_bytes(metric):
        i = hdr.index(metric)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(vals[i]) * scale
    TRAFFIC[name] = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))


def launch_summary():
    p = os.path.join(OUT, "launches.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 10 and r[0].isdigit()]
    names = [(r[4].split("(")[0], float(r[-1]) * (1000 if r[-2] == "ms" else (0.001 if r[-2] == "ns" else 1))) for r in rows]
    idx = [i for i, (n, _) in enumerate(names) if "project_fwd" in n]
    if len(idx) < 2:
        return
    step = names[idx[0]:idx[1]]
    tot = sum(t for _, t in step)
    lines = [f"# ncu launch list of ONE benchmark step (gpu__time_duration.sum, cold-cache + serialised: compare shares, not absolutes)", "",
             f"total {tot:.1f} us, {len(step)} launches", "", "| us | share | kernel |", "|---|---|---|"]
    for n, t in step:
        short = n.replace("void ", "").replace("b200gs::<unnamed>::", "b200gs::")
        lines.append(f"| {t:.1f} | {t / tot * 100:.1f}% | `{short[-110:]}` |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")


def binning_summary():
    """prof_bin*.ncu-rep: one full capture of each of our binning kernels (captured with
    ncu --set full --clock-control none --import-source on -k regex:"emit_cells|scatter_ids|chunk_counts|chunk_prefix" -s 8 -c 4 -o gpurun_out/prof_bin ...)"""
    reps = sorted(f for f in os.listdir(OUT) if f.startswith("prof_bin") and f.endswith(".ncu-rep"))
    if not reps:
        return
    cols = [("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "warp inst"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
            ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr")]
    lines = ["# ncu --set full, binning kernels of one benchmark step (1 M Gaussians, 1920x1080, vanilla mode)", "",
             "| kernel | " + " | ".join(c[1] for c in cols) + " |", "|---|" + "---|" * len(cols)]
    seen = set()
    for rep in reps[::-1]:      # newest capture of a kernel wins
        raw = ncu_csv(os.path.join(OUT, rep), "raw")
        hdr, units = raw[0], raw[1]
        for vals in raw[2:]:
            name = vals[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", "")
            if name in seen:
                continue
            seen.add(name)
            cells = []
            for m, _ in cols:
                if m in hdr:
                    i = hdr.index(m)
                    v = vals[i]
                    try:
                        v = f"{float(v):.4g}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {units[i]}".strip())
                else:
                    cells.append("-")
            lines.append(f"| `{name}` | " + " | ".join(cells) + " |")
    lines += ["", "History of these kernels this round (same capture settings): first hierarchical version — `tile_masks` (per-entry row loops + 128",
              "ballots) 76 us, `scatter_ids` (per-tile ballot loop) 105 us; per-entry bit loops with shared-memory atomics 58 / 60 us at 6 of 32",
              "lanes active; now: masks computed in `emit_cells` from balanced (rank, row) items, counts / scatter from warp bit-matrix transposes."]
    open(os.path.join(ROOT, "profiles", f"{tag}_binning_ncu.md"), "w").write("\n".join(lines) + "\n")


kernel_summary("blend_bwd")
kernel_summary("blend_fwd")
binning_summary()
launch_summary()
import json
json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full), bytes", **TRAFFIC},
          open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
print("written:", [f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith(tag)])
