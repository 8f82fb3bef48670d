"""Summarise `ncu --page raw --csv` exports (one row per profiled launch) into a markdown table."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "ms", 1e-6), ("dram__bytes_read.sum", "GB", None), ("dram__bytes_write.sum", "GB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("lts__t_sector_hit_rate.pct", "l2hit%", 1), ("l1tex__t_sector_hit_rate.pct", "l1hit%", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1)]


def conv(v, unit, want_unit):
    v = float(v.replace(",", ""))
    if want_unit == "GB":
        mult = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}.get(unit, 1e-9)
        return v * mult
    if want_unit == "ms":
        mult = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        return v * mult
    return v


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    print("| kernel | " + " | ".join(w[1] for w in WANT) + " |")
    print("|---|" + "---:|" * len(WANT))
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("sgs::", "")[:48]
        cells = []
        for m, u, _ in WANT:
            if m in hdr:
                i = hdr.index(m)
                try:
                    cells.append(f"{conv(r[i], units[i], u):.3g}")
                except ValueError:
                    cells.append(r[i])
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
