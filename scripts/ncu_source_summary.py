"""Summarise `ncu --page source --csv` exports (possibly several kernels per file): instruction mix, stall reasons,
and the top stall locations."""
import collections
import csv
import gzip
import sys


def kernels(path):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    cur, name = None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if cur:
                yield name, cur
            name, cur = r[1][:70], []
        elif cur is not None:
            cur.append(r)
    if cur:
        yield name, cur


def summarise(name, rows, ntop=12):
    hdr = rows[0]
    isrc, iex, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    data = []
    for r in rows[1:]:
        try:
            int(r[iex]); int(r[ism]); data.append(r)
        except (ValueError, IndexError):
            pass
    tot = sum(int(r[ism]) for r in data) or 1
    totn = sum(int(r[iex]) for r in data) or 1
    print(f"## {name}\nwarp instructions {totn:,}, stall samples {tot:,}")
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        t = r[isrc].strip().split()
        o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[o] += int(r[iex]); samp[o] += int(r[ism])
    print("opcode: instr% / samples%:  " + "  ".join(f"{o} {100*n/totn:.1f}/{100*samp[o]/tot:.1f}" for o, n in ops.most_common(16)))
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in data:
        for h in stall:
            try:
                agg[h] += int(r[hdr.index(h)])
            except ValueError:
                pass
    s = sum(agg.values()) or 1
    print("stall reasons %:", {k.replace("stall_", ""): round(100 * v / s, 1) for k, v in agg.most_common(8)})
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][ism]))[:ntop]
    for i in sorted(idx):
        r = data[i]
        top = max(stall, key=lambda h: int(r[hdr.index(h)] or 0))
        print(f"  #{i:5d} {100*int(r[ism])/tot:5.2f}%  x{r[iex]:>10}  {r[isrc].strip()[:60]:60s} {top}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for name, rows in kernels(p):
            summarise(name, rows)
