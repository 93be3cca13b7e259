#!/usr/bin/env python3
"""Write profiles/r02_ncu_summary.md from the committed round-2 bench lines and the CSVs tools/gpu_profile.sh brings back.
usage: python tools/make_profile_summary_r02.py [tag=r02]   (expects gpurun_out/<tag>_launches.csv, <tag>_full_raw.csv and
profiles/<tag>_bench_k22_n{1,2,4,8}.json; copies the two CSVs into profiles/)"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = os.path.join(ROOT, "profiles")


def load(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("{"):
            return json.loads(line)
    return None


shutil.copy(os.path.join(ROOT, "gpurun_out", "%s_launches.csv" % tag), os.path.join(P, "%s_launches_bench_k22.csv" % tag))
shutil.copy(os.path.join(ROOT, "gpurun_out", "%s_full_raw.csv" % tag), os.path.join(P, "%s_ncu_full_raw.csv" % tag))
body = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), tag], cwd=ROOT, text=True)

out = ["# Round 2 -- ncu evidence and bench lines (B200, sm_100a)", "",
       "Raw files in this directory:",
       "* `%s_launches_bench_k22.csv` -- launch list of `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 --launch-list-only`" % tag,
       "  (`tools/gpu_profile.sh %s`; nothing but the schedule steps is launched in that mode: 4 steps x 38 MSM + 59 NTT);" % tag,
       "* `%s_ncu_full_raw.csv` -- `ncu -i ... --page raw --csv` of one `--set full --clock-control none --import-source on` capture of" % tag,
       "  `tools/prof_once.py 22` (one launch of every hot kernel at k = 22, the witness kernels and the N1/N3 kernels); the .ncu-rep is not kept;",
       "* `%s_bench_k22_n{1,2,4,8}.json`, `%s_bench_k18_n1.json`, `%s_bench_k20_n1.json` -- un-profiled bench lines of the same build;" % (tag, tag, tag),
       "  `%s_sweep_msm_ntt_n{1,2,4,8}.jsonl` -- config-5 sweep (uniform and witness-like scalars, per-size oracle equality, CPU seconds)." % tag,
       "Tables below: `tools/summarize_ncu.py %s`; this file: `tools/make_profile_summary_r02.py %s`." % (tag, tag), "",
       "## Bench lines (not under a profiler)", ""]
rows = []
for n in (1, 2, 4, 8):
    l = load("%s_bench_k22_n%d.json" % (tag, n))
    if not l:
        continue
    e = l.get("e2e") or {}
    rows.append("| %d | %.4f | %s | %s | %s | %s | %s |" % (
        n, l["value"], ("%.4f" % e["value"]) if e else "-", l.get("gpu_launches"), (l.get("parity") or {}).get("ok"),
        (e.get("parity") or {}).get("all_outputs_equal_single_gpu_prover", (e.get("parity") or {}).get("round1_commitments_equal_oracle_best_multiexp")) if e else "-",
        (l.get("clocks") or {}).get("sm_mhz")))
out += ["| GPUs | value (s) | e2e (s) | launches / step | parity (schedule) | parity (e2e) | SM MHz |", "|---|---|---|---|---|---|---|"] + rows + [""]
l1 = load("%s_bench_k22_n1.json" % tag)
if l1:
    r = l1["roofline"]
    out.append("* roofline (`msm_accumulate`, timed alone, weighted over the schedule's scalar kinds): %.3f ms per launch -> %.1f GB/s of algorithmic bytes = **%.2f %%** of the measured %.0f GB/s;"
               % (r["avg_launch_ms"], r["achieved"], 100 * r["frac"], r["peak"]))
    out.append("  uniform column alone: %.2f ms (%.2f %%); DRAM traffic per launch %.2f GB vs %.3f GB algorithmic." % (
        r["launch_ms_uniform_column"], 100 * r["frac_uniform_column"], r["traffic"] / 1e9, r["algorithmic_bytes_per_launch"] / 1e9))
    rm = l1["roofline_multiplier"]
    out.append("* multiplier pipe (ncu fmaheavy, pct of peak sustained active): %s." % ", ".join("%s %.1f %%" % kv for kv in rm["ncu_pipe_fmaheavy_pct_of_peak"].items()))
    cb = l1.get("cpu_baseline")
    if cb:
        out.append("* CPU port on the box's %d host cores (bounded sample scaled by the unit counts): %.1f s; `--impl reference` measures a full replay." % (cb["cores"], cb["value"]))
    w = l1.get("witness")
    if w:
        out.append("* witness: %d rows / %d records: recording %.1f ms on %d host threads, H2D + expansion %.1f ms (kernels %.2f ms) -> %.0f M rows/s." % (
            w["rows"], w["op_records"], w["host_record_s"] * 1e3, w["host_threads"], w["expand_incl_h2d_s"] * 1e3, w["expand_kernel_ms"], w["rows_per_s_total"] / 1e6))
    st = (l1.get("e2e") or {}).get("stage_ms_synchronised_rank0") or (l1.get("e2e") or {}).get("stage_ms_synchronised")
    if st:
        out.append("* e2e stages (synchronised after each): " + ", ".join("%s %.1f ms" % (k, v) for k, v in st.items() if not isinstance(v, dict)) + ".")
l8 = load("%s_bench_k22_n8.json" % tag)
if l8 and l8.get("phases"):
    out += ["", "Phases of the schedule replay at N = 8 (`phases` of the bench line): busy time of each rank's own units and the span including the wait for the slowest rank + the all-gather.", "",
            "| phase | units | busy max (ms) | busy mean (ms) | span (ms) | window-sharded MSMs |", "|---|---|---|---|---|---|"]
    for i, ph in enumerate(l8["phases"]):
        out.append("| %d | %s | %.2f | %.2f | %.2f | %d |" % (i, ", ".join("%d x %s" % (v, k) for k, v in ph["units"].items()), ph["busy_ms_max"], ph["busy_ms_mean"], ph["span_ms_rank0"], len(ph["sharded"])))
out += ["", body]

# stall reasons of the hot kernels
raw = list(csv.reader(open(os.path.join(P, "%s_ncu_full_raw.csv" % tag))))
hdr = raw[0]
ki = hdr.index("Kernel Name")
stall = [(i, h) for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued")]
out += ["", "## Warp stall samples of the hot kernels (share of all stall samples, top 6)", "", "| kernel | time | stalls |", "|---|---|---|"]
ti = hdr.index("gpu__time_duration.sum")
for r in raw[2:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("h2agg::", "")
    if not re.search(r"msm_accumulate|ntt_pass|quot_evaluate_h|witness_expand_kernel<5>", name):
        continue
    vals = []
    for i, h in stall:
        try:
            vals.append((float(r[i].replace(",", "")), h.split("stalled_")[1]))
        except ValueError:
            pass
    tot = sum(v for v, _ in vals) or 1.0
    top = sorted(vals, reverse=True)[:6]
    out.append("| `%s` | %s %s | %s |" % (name, r[ti], raw[1][ti], ", ".join("%s %.0f %%" % (n_, 100 * v / tot) for v, n_ in top)))

# SASS evidence of the TMA path
lib = os.path.join(ROOT, "halo2_snark_aggregator_b200", "libh2agg.so")
try:
    sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True, stderr=subprocess.DEVNULL)
    cnt = {k: len(re.findall(k, sass)) for k in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "UTMALDG", "IMAD.WIDE.U32")}
    out += ["", "## SASS (cuobjdump -sass libh2agg.so)", "",
            "`ntt_pass_kernel<true>` loads its tile with TMA bulk copies: " + ", ".join("%s x %d" % kv for kv in cnt.items() if kv[0] != "IMAD.WIDE.U32") +
            "; the arithmetic of every kernel is IMAD.WIDE.U32 (%d instructions in the library)." % cnt["IMAD.WIDE.U32"]]
except Exception as e:  # pragma: no cover
    out += ["", "(cuobjdump not available: %r)" % (e,)]

open(os.path.join(P, "%s_ncu_summary.md" % tag), "w").write("\n".join(out) + "\n")
print("wrote profiles/%s_ncu_summary.md (%d lines)" % (tag, len(out)))
