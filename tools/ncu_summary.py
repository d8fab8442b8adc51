"""Turn ncu output brought back in gpurun_out/ into the markdown summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv "<command>"  > profiles/launches_rNN_summary.md
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep "<command>"       > profiles/ncu_<what>_rNN_summary.md
"""
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def launches(path, cmd):
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    mu = h.index("Metric Unit")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        t = float(r[mv].replace(",", ""))
        if r[mu] in ("ns", "nsecond"):
            t /= 1e3
        elif r[mu] in ("ms", "msecond"):
            t *= 1e3
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += t
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print("command: " + cmd)
    print("(cold-cache, serialised launches: compare SHARES with bench.py's event-timed kernel_ms_per_step, not absolutes)\n")
    print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
    for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {100 * t / tot:.2f}% | {t:.1f} | {c} | {t / c:.2f} | `{name}` |")


def full(path, cmd):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    stalls = [c for c in h if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("per_issue_active.ratio")]
    print("command: " + cmd + "\nunits as printed by ncu\n")
    for r in rows[2:]:
        print("## " + r[h.index("Kernel Name")])
        for m in FULL_METRICS:
            if m in h:
                print(f"- {m}: {r[h.index(m)]} {units[h.index(m)]}")
        st = sorted(((float(r[h.index(c)]), c.replace("smsp__average_warps_issue_stalled_", "").replace(
            "_per_issue_active.ratio", "")) for c in stalls), reverse=True)[:6]
        print("- top stalls (per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in st) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
