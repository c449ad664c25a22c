#!/usr/bin/env python
"""Turns ncu output brought back in gpurun_out/ into the small tracked summaries under
profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r01_3xtf32.csv  profiles/r01_launches_3xtf32.md
  python profiles/summarize.py full     gpurun_out/prof_r01_tc_gemm_v3.ncu-rep profiles/r01_ncu_tc_gemm.md
"""
import collections
import csv
import re
import subprocess
import sys


def _short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
    return name.strip()


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui, gi = (hdr.index(k) for k in
                          ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "Grid Size"))
    recs = []
    for r in data:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        recs.append((_short(r[ki]), v, r[gi]))
    starts = [i for i, (k, _, _) in enumerate(recs) if k.startswith("reflect_pad_wave")]
    step = recs[starts[-1]:]
    tot = sum(v for _, v, _ in step)
    agg = collections.OrderedDict()
    for k, v, _ in step:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none): last profiled step",
           "", "source: `%s` — %d launches in the log, %d in the last step, %.1f us summed device "
           "time (serialised, cold-cache: compare SHARES, not absolutes)" % (src, len(recs), len(step), tot),
           "", "| kernel | launches | us | share |", "|---|---:|---:|---:|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% |" % (k[:100], n, v, 100 * v / tot))
    out += ["", "## GEMM launches in program order", "", "| # | kernel | grid | us |", "|---|---|---|---:|"]
    i = 0
    for k, v, g in step:
        if "gemm" in k:
            out.append("| %d | `%s` | %s | %.1f |" % (i, k[:60], g, v))
            i += 1
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
]


def _raw_rows(src):
    """rows of `ncu --page raw --csv`: from a .ncu-rep, or from a .csv already exported on the
    GPU box (reports with --import-source are too large to bring back)."""
    if src.endswith(".csv"):
        raw = open(src, errors="ignore").read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "ID")
    return rows[hi:]


def full(src, dst):
    rows = _raw_rows(src)
    hdr, units = rows[0], rows[1]
    col = {}
    for key, _ in KEYS:
        for i, h in enumerate(hdr):
            if h == key:
                col[key] = i
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    out = ["# ncu --set full --clock-control none: `%s`" % src, "",
           "| # | kernel | grid | " + " | ".join(lbl for _, lbl in KEYS) + " |",
           "|---|---|---|" + "---:|" * len(KEYS)]
    for n, r in enumerate(rows[2:]):
        cells = []
        for key, _ in KEYS:
            if key in col:
                v, u = r[col[key]], units[col[key]]
                try:
                    v = "%.3g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append("%s %s" % (v, u) if u not in ("%", "") else v)
            else:
                cells.append("n/a")
        out.append("| %d | `%s` | %s | %s |" % (n, _short(r[ki])[:48], r[gi], " | ".join(cells)))
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


def traffic(src, dst, precision=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of every GEMM launch in an `ncu --set full`
    capture -> profiles/r02_gemm_traffic.json[precision] (bench.py's roofline.traffic)."""
    import json
    import os
    rows = _raw_rows(src)
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    ti = hdr.index("gpu__time_duration.sum")

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per, tot, n = {}, 0.0, 0
    for r in rows[2:]:
        k = _short(r[ki])
        if "gemm" not in k:
            continue
        b = to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
        e = per.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "time_%s" % units[ti]: 0.0})
        e["launches"] += 1
        e["dram_bytes"] += b
        e["time_%s" % units[ti]] += float(r[ti].replace(",", ""))
        tot += b
        n += 1
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    data[precision or "default"] = {
        "source": src, "gemm_launches": n, "dram_bytes_total": tot,
        "dram_bytes_per_launch": tot / max(n, 1), "per_kernel": per,
        "how": "ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum "
               "summed over the GEMM launches captured (whole steps), divided by their count"}
    json.dump(data, open(dst, "w"), indent=1)
    print("wrote", dst, precision, "%.1f MB per launch over %d launches" % (tot / max(n, 1) / 1e6, n))


if __name__ == "__main__":
    fn = {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]]
    fn(*sys.argv[2:])
