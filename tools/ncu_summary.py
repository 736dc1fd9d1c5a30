"""Turn an .ncu-rep (or an ncu --csv launch list) into the small text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_xxx.md [--traffic-key name]
    python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r1_launches_xxx.md
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def summarize_rep(rep, out, traffic_key=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of `%s`" % os.path.basename(rep), "",
             "`ncu --set full --clock-control none --import-source on` (per-launch values; cold-cache, serialised)", ""]
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append("## %s" % name[:110])
        for k in KEYS:
            if k in hdr:
                lines.append("- %s = %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [(num(r[i]) or 0.0, h) for i, h in enumerate(hdr)
              if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        lines.append("- warp stall samples: " + ", ".join(
            "%s %.0f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot) for v, h in st[:7]))
        rd, wr = num(r[hdr.index("dram__bytes_read.sum")]), num(r[hdr.index("dram__bytes_write.sum")])
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        if rd is not None and wr is not None:
            traffic[name] = rd * scale.get(ur, 1) + wr * scale.get(uw, 1)
            lines.append("- dram traffic (read+write) = %.1f MB" % (traffic[name] / 1e6))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    if traffic_key:
        p = os.path.join(os.path.dirname(out), "ncu_traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[traffic_key] = list(traffic.values())[-1]
        json.dump(d, open(p, "w"), indent=1)
    print("wrote", out)


def summarize_launches(csvfile, out):
    rows = [r for r in csv.reader(open(csvfile, errors="ignore")) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    unit = rows[1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else ""
    lines = ["# launch list `%s` (ncu --metrics gpu__time_duration.sum --clock-control none)" % os.path.basename(csvfile),
             "", "| kernel | launches | total %s | share |" % unit, "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f%% |" % (k[:100], a[0], a[1], 100 * a[1] / tot))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        summarize_launches(sys.argv[2], sys.argv[3])
    else:
        tk = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
        summarize_rep(sys.argv[1], sys.argv[2], tk)
