"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list and/or a `--set full` report."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][:80]
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        tot.setdefault(name, [0.0, 0])
        tot[name][0] += ms
        tot[name][1] += 1
    s = sum(v[0] for v in tot.values())
    print(f"# launch list {path}: {sum(v[1] for v in tot.values())} launches, {s:.2f} ms of kernel time "
          "(cold-cache, serialised: compare SHARES)")
    print("| ms | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f"| {v[0]:.3f} | {100 * v[0] / s:.1f}% | {v[1]} | `{k}` |")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"\n# full-set report {path}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0][:90]
        print(f"\n## `{name}`")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"- {w} = {r[i]} {units[i]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (launches if p.endswith(".csv") else full)(p)
