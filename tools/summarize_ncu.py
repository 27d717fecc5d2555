"""Turns every <name>.raw.csv (+ .details.txt, .sass.csv) of an ncu scratch directory into tracked summaries under
profiles/: <tag>_ncu_<name>.txt (key metrics + the details page), <tag>_sass_opcount_<name>.txt (executed SASS of the
captured launch) and an entry captures_<tag>.<name> in profiles/traffic.json.

    python tools/summarize_ncu.py <scratch dir> <tag>
"""
import collections
import csv
import io
import json
import os
import sys

src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg"] + \
       ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % r for r in
        ("long_scoreboard", "not_selected", "mio_throttle", "short_scoreboard", "wait", "barrier", "lg_throttle",
         "branch_resolving", "math_pipe_throttle", "dispatch_stall", "no_instruction")]
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
tj = os.path.join(out, "traffic.json")
traffic = json.load(open(tj)) if os.path.exists(tj) else {}
for f in sorted(os.listdir(src)):
    if not f.endswith(".raw.csv"):
        continue
    name = f[:-len(".raw.csv")]
    rows = list(csv.reader(open(os.path.join(src, f))))
    if len(rows) < 3:
        print("skip (empty):", f)
        continue
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    with open(os.path.join(out, "%s_ncu_%s.txt" % (tag, name)), "w") as fh:
        fh.write("# ncu --set full --clock-control none --cache-control none: key metrics of the captured launch\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            fh.write("== %s\n" % d.get("Kernel Name", ""))
            for k in KEYS:
                if d.get(k, "") != "":
                    fh.write("   %-86s %s %s\n" % (k, d[k], units.get(k, "")))
        det = os.path.join(src, name + ".details.txt")
        if os.path.exists(det):
            fh.write("\n# ---- ncu --page details ----\n" + open(det).read())
    d = dict(zip(hdr, rows[2]))
    rd = float(d["dram__bytes_read.sum"]) * SCALE[units["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * SCALE[units["dram__bytes_write.sum"]]
    t_us = float(d["gpu__time_duration.sum"]) * TIME.get(units["gpu__time_duration.sum"], 1.0)
    entry = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
             "gpu_time_us_under_ncu": t_us, "kernel": d.get("Kernel Name", "")}
    for k, short in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
                     ("smsp__inst_executed.sum", "warp_instructions")):
        if d.get(k, "") != "":
            entry[short] = float(d[k])
    sass = os.path.join(src, name + ".sass.csv")
    if os.path.exists(sass):
        lines = open(sass).read().splitlines()
        start = next((i for i, l in enumerate(lines) if l.startswith('"Address"')), None)
        if start is not None:
            hist, total = collections.Counter(), 0
            for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
                try:
                    n = int(r["Instructions Executed"])
                except (KeyError, ValueError):
                    continue
                toks = r["Source"].split()
                hist[toks[1] if toks[0].startswith("@") else toks[0]] += n
                total += n
            with open(os.path.join(out, "%s_sass_opcount_%s.txt" % (tag, name)), "w") as fh:
                fh.write("# executed SASS (warp instructions) of the captured launch, %s\n" % d.get("Kernel Name", ""))
                fh.write("total %d\n" % total)
                for op, n in hist.most_common(45):
                    fh.write("   %-22s %12d  %5.1f %%\n" % (op, n, 100.0 * n / max(total, 1)))
    traffic.setdefault("captures_" + tag, {})[name] = entry
    print(name, entry)
json.dump(traffic, open(tj, "w"), indent=2)
