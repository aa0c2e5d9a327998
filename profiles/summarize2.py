"""Tracked summaries from the ncu artefacts of a gpurun call.  Usage:
    python profiles/summarize2.py <tag> <launches.csv> <prof.ncu-rep>
* <launches.csv>: ncu --metrics gpu__time_duration.sum --clock-control none ... --csv  (one benchmark step is cut out of it)
* <prof.ncu-rep>: ONE report with `--set full --import-source on` captures of the kernels of a step (several kernels per report)
Writes profiles/<tag>_launches.md, profiles/<tag>_kernels_ncu.md (metric table + opcode mix + top stall reasons per kernel) and
profiles/<tag>_traffic.json (DRAM bytes and executed warp instructions per launch, read by bench.py)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]

COLS = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "warp inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ lim regs"), ("launch__occupancy_limit_shared_mem", "occ lim smem"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"), ("lts__t_bytes.sum", "L2 bytes"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("smsp__inst_executed_op_global_red.sum", "RED inst"), ("sm__inst_executed_pipe_xu.sum", "XU inst"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld requests")]


def ncu_csv(page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def short(name):
    return name.split("(")[0].replace("void ", "").replace("b200gs::<unnamed>::", "").replace("b200gs::sweep::", "sweep::").replace("b200gs::", "")


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def kernels():
    raw = ncu_csv("raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    lines = [f"# ncu --set full --clock-control none --import-source on: the kernels of one benchmark step ({tag}; configs[1]: 1 M Gaussians, 1920x1080, vanilla mode)", "",
             "| kernel | " + " | ".join(c[1] for c in COLS) + " |", "|---|" + "---|" * len(COLS)]
    traffic = {}
    seen = {}
    for vals in rows:
        name = short(vals[hdr.index("Kernel Name")])
        if name in seen:      # several launches of one kernel (radix passes): keep them all in the table, numbered
            seen[name] += 1
            label = f"{name} #{seen[name]}"
        else:
            seen[name] = 1
            label = name
        cells = []
        for m, _ in COLS:
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
        lines.append(f"| `{label}` | " + " | ".join(cells) + " |")
        key = {"blend_bwd_tr_kernel": "blend_bwd", "blend_fwd_kernel": "blend_fwd", "blend_fwd_async_kernel": "blend_fwd", "project_fwd_kernel": "project_fwd",
               "project_bwd_kernel": "project_bwd", "depth_keys_kernel": "bin_count", "emit_cells_kernel": "bin_sort"}.get(name.split("::")[-1].split("<")[0])
        if key and key not in traffic:
            ir, iw, ii = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("smsp__inst_executed.sum")
            traffic[key] = int(to_bytes(vals[ir], units[ir]) + to_bytes(vals[iw], units[iw]))
            traffic[key + "_warp_instructions"] = int(float(vals[ii]))
    # per-kernel opcode mix and stall reasons from the source page
    src = ncu_csv("source")
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": short(r[1]) if len(r) > 1 else "?", "hdr": None, "rows": []}
            blocks.append(cur)
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = r
            continue
        cur["rows"].append(r)
    for b in blocks:
        h = b["hdr"]
        if not h or "Instructions Executed" not in h or "Source" not in h:
            continue
        ia, isrc = h.index("Instructions Executed"), h.index("Source")
        tot = sum(int(r[ia]) for r in b["rows"] if len(r) > ia and r[ia].isdigit()) or 1
        mix = collections.Counter()
        for r in b["rows"]:
            if len(r) <= max(ia, isrc) or not r[ia].isdigit():
                continue
            t = r[isrc].strip().split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            mix[op] += int(r[ia])
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        lines += ["", f"### `{b['name']}`: opcode mix (share of {tot} executed warp instructions)", "",
                  ", ".join(f"{op} {c / tot * 100:.1f}%" for op, c in mix.most_common(18))]
        if stall_cols:
            st = collections.Counter()
            for r in b["rows"]:
                for i, c in stall_cols:
                    if len(r) > i:
                        try:
                            st[c] += float(r[i])
                        except ValueError:
                            pass
            s_tot = sum(st.values()) or 1
            lines += ["", "stall samples: " + ", ".join(f"{c.replace('stall_', '')} {v / s_tot * 100:.0f}%" for c, v in st.most_common(6))]
    open(os.path.join(ROOT, "profiles", f"{tag}_kernels_ncu.md"), "w").write("\n".join(lines) + "\n")
    json.dump({"what": "per launch, from the ncu --set full capture of bench.py's configs[1] step: dram__bytes_read.sum + dram__bytes_write.sum (bytes) and "
                       "smsp__inst_executed.sum (<kernel>_warp_instructions)", **traffic},
              open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)


def launch_summary():
    rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 10 and r[0].isdigit()]
    names = [(r[4].split("(")[0], float(r[-1]) * (1000 if r[-2] == "ms" else (0.001 if r[-2] == "ns" else 1))) for r in rows]
    idx = [i for i, (n, _) in enumerate(names) if "project_fwd" in n]
    if len(idx) < 2:
        print("launch list: fewer than two steps captured")
        return
    step = names[idx[0]:idx[1]]
    tot = sum(t for _, t in step)
    lines = ["# ncu launch list of ONE benchmark step (gpu__time_duration.sum, cold-cache + serialised: compare shares, not absolutes)", "",
             f"total {tot:.1f} us, {len(step)} launches", "", "| us | share | kernel |", "|---|---|---|"]
    for n, t in step:
        lines.append(f"| {t:.1f} | {t / tot * 100:.1f}% | `{short(n)[-110:]}` |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")


kernels()
launch_summary()
print("written:", sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith(tag)))
