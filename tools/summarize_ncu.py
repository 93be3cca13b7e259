#!/usr/bin/env python3
"""Turn the CSVs tools/gpu_profile.sh brings back (gpurun_out/<tag>_launches.csv, <tag>_full_raw.csv) into the
markdown tables kept under profiles/.  usage: python tools/summarize_ncu.py <tag> [steps_in_launch_list]"""
import csv
import re
import sys
from collections import OrderedDict

tag = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4   # 3 warm-up + 1 timed step of bench.py under ncu


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("h2agg::", "")


def launches():
    rows = list(csv.reader(l for l in open("gpurun_out/%s_launches.csv" % tag) if l.startswith('"')))
    hdr = rows[0]
    i_name, i_metric, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[i_unit]]
        k = short(r[i_name])
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
    return agg


def full():
    rows = list(csv.reader(open("gpurun_out/%s_full_raw.csv" % tag)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [("time", "gpu__time_duration.sum"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
            ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            ("fmaheavy%", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            ("issue%", "sm__inst_issued.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
            ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("grid", "launch__grid_size"),
            ("L2hit%", "lts__t_sector_hit_rate.pct")]
    cols = [(a, b) for a, b in cols if b in idx]
    out = ["| kernel | " + " | ".join(a for a, _ in cols) + " |", "|---|" + "---|" * len(cols)]
    for r in rows[2:]:
        cells = []
        for a, b in cols:
            v, u = r[idx[b]], units[idx[b]]
            try:
                f = float(v.replace(",", ""))
                v = ("%.4g" % f) + ((" " + u) if a in ("time", "dram_rd", "dram_wr") else "")
            except ValueError:
                pass
            cells.append(v)
        out.append("| `%s` | %s |" % (short(r[idx["Kernel Name"]]), " | ".join(cells)))
    return "\n".join(out)


agg = launches()
if len(sys.argv) <= 2 and "msm_accumulate" in agg and agg["msm_accumulate"][0] % 38 == 0:
    steps = agg["msm_accumulate"][0] // 38   # the schedule has 38 MSMs per step
setup = ("msm_build_table", "ntt_gen_full_table", "ntt_gen_tables", "synth_bases_kernel", "synth_scalars_kernel", "quot_gen_tables")
step_rows = [(k, v) for k, v in agg.items() if k not in setup]
tot = sum(v[1] for _, v in step_rows) / steps
print("## Launch list of the bench step (`%s_launches_bench_k22.csv`: %d bench steps under ncu, per-step averages; cold-cache, serialised: compare SHARES)\n" % (tag, steps))
print("| kernel | launches / step | ms / step (serialised) | share |\n|---|---|---|---|")
for k, v in sorted(step_rows, key=lambda kv: -kv[1][1]):
    print("| `%s` | %.0f | %.2f | %.1f%% |" % (k, v[0] / steps, v[1] / steps, 100 * v[1] / steps / tot))
print("| **sum** | %.0f | %.1f | 100%% |" % (sum(v[0] for _, v in step_rows) / steps, tot))
print("\nOne-time set-up in the same capture:\n\n| kernel | launches | total ms |\n|---|---|---|")
for k in setup:
    if k in agg:
        print("| `%s` | %d | %.1f |" % (k, agg[k][0], agg[k][1]))
print("\n## `--set full` highlights (`%s_ncu_full_raw.csv`; one launch each, k = 22)\n" % tag)
print(full())
