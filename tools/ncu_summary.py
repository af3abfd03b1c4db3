#!/usr/bin/env python
"""Summarise ncu output for profiles/:  tools/ncu_summary.py launches <csv>  |  tools/ncu_summary.py full <ncu-rep>"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {n} launches, {tot:.1f} us total (ncu gpu__time_duration.sum, cold cache, serialised: compare SHARES)")
    print(f"{'kernel':70s} {'n':>6s} {'total_us':>11s} {'share':>7s} {'avg_us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {v[0]:6d} {v[1]:11.1f} {v[1] / tot:7.3f} {v[1] / v[0]:9.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    for row in rows:
        print("kernel:", row[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {row[i]:>16s} {units[i]}")
        print()


def metrics(path):
    """CSV of `ncu --metrics ... --csv`: one block per kernel NAME (mean over its launches): time, DRAM bytes and
    achieved GB/s, %-of-peak numbers, registers, grid."""
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit, met = row["Metric Unit"], row["Metric Name"]
        if met == "gpu__time_duration.sum":
            v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        if met.startswith("dram__bytes") or met == "lts__t_bytes.sum":
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d = per.setdefault(name, collections.defaultdict(list))
        d[met].append(v)
    tot = sum(sum(d["gpu__time_duration.sum"]) for d in per.values())
    print(f"# {len(per)} kernel names, {tot:.1f} us total inside the profiled range (ncu: cold cache, serialised)")
    hdr = ("kernel", "n", "sum_us", "share", "avg_us", "dramMB/l", "GB/s", "dram%", "tensor%", "fma%", "lsu%", "sm%",
           "warps%", "L2hit%", "regs", "grid", "block")
    print("%-58s %4s %9s %6s %8s %9s %8s %6s %7s %6s %6s %6s %6s %6s %5s %6s %5s" % hdr)

    def mean(d, k):
        return sum(d[k]) / len(d[k]) if d.get(k) else float("nan")
    for name, d in sorted(per.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        t = d["gpu__time_duration.sum"]
        n, s = len(t), sum(t)
        mb = (mean(d, "dram__bytes_read.sum") + mean(d, "dram__bytes_write.sum")) / 1e6
        gbs = mb * 1e6 / (s / n * 1e-6) / 1e9 if s > 0 else 0.0
        print("%-58s %4d %9.1f %6.3f %8.1f %9.2f %8.0f %6.1f %7.1f %6.1f %6.1f %6.1f %6.1f %6.1f %5.0f %6.0f %5.0f" % (
            name[:58], n, s, s / tot, s / n, mb, gbs,
            mean(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            mean(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            mean(d, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            mean(d, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            mean(d, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            mean(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            mean(d, "lts__t_sector_hit_rate.pct"), mean(d, "launch__registers_per_thread"),
            mean(d, "launch__grid_size"), mean(d, "launch__block_size")))


if __name__ == "__main__":
    {"launches": launches, "full": full, "metrics": metrics}[sys.argv[1]](sys.argv[2])
