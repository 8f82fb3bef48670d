"""profiles/r01_traffic.json from the `ncu --set full` raw exports: measured DRAM bytes per launch of every kernel
family bench.py reports a roofline for (bench.py copies them into its `traffic` fields)."""
import csv
import json
import sys

UNIT = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    out = []
    for r in rows[2:]:
        out.append((r[ki], float(r[ri].replace(",", "")) * UNIT[units[ri]], float(r[wi].replace(",", "")) * UNIT[units[wi]]))
    return out


def mean(rows, pred):
    sel = [(a, b) for k, a, b in rows if pred(k)]
    return (sum(a for a, _ in sel) / len(sel), sum(b for _, b in sel) / len(sel)) if sel else (0.0, 0.0)


def main(hot, spmm, dst):
    h, s = load(hot), load(spmm)
    fam = {}
    tag = dst.split("/")[-1].split("_")[0]

    def put(name, kernel, rw):
        fam[name] = {"kernel": kernel, "dram_read_gb": round(rw[0], 3), "dram_write_gb": round(rw[1], 3)}

    put("edge_score_fwd", "edge_score_tc2_kernel<__half,0> (CTA pairs)", mean(h, lambda k: "tc2_kernel<__half, 0>" in k))
    parts = [mean(h, lambda k: "tc2_kernel<__half, 1>" in k), mean(h, lambda k: "bwd_df" in k), mean(h, lambda k: "bwd_dw" in k)]
    put("edge_score_bwd", "edge_score_tc2_kernel<__half,1> + edge_score_bwd_df_kernel + edge_score_bwd_dw_kernel",
        (sum(p[0] for p in parts), sum(p[1] for p in parts)))
    put("loss_fwd", "loss_edges_fused_kernel<3>", mean(h, lambda k: "loss_edges_fused" in k))
    tq = [(a, b) for k, a, b in h if "topq_" in k]
    ndraw = max(1, sum(1 for k, _, _ in h if "topq_write" in k))
    put("sample_topq", "all topq_* kernels of one draw (sample, predict, keys_window, find_window, hist, find, count, "
        "scan, write)", (sum(a for a, _ in tq) / ndraw, sum(b for _, b in tq) / ndraw))
    if any("spmm_h16" in k for k, _, _ in s):     # fp16 gather tables (round 2 default)
        put("spmm_d256", "spmm_h16_kernel (fp16 gather table)", mean(s, lambda k: "spmm_h16" in k))
        put("edge_grad_d256", "edge_grad_sddmm_h16_kernel (fp16 gather table)", mean(s, lambda k: "sddmm_h16" in k))
    else:
        put("spmm_d256", "spmm_kernel<4,2>", mean(s, lambda k: "spmm_kernel<4, 2>" in k))
        put("edge_grad_d256", "edge_grad_sddmm_kernel<4,2>", mean(s, lambda k: "sddmm_kernel<4, 2>" in k))
    put("spmm_d41", "spmm_kernel<1,2>", mean(s, lambda k: "spmm_kernel<1, 2>" in k))
    put("edge_grad_d41", "edge_grad_sddmm_kernel<1,2>", mean(s, lambda k: "sddmm_kernel<1, 2>" in k))
    fam = {k: v for k, v in fam.items() if v["dram_read_gb"] + v["dram_write_gb"] > 0}
    fam["_source"] = (f"ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 (scripts/ncu_{'round2' if tag != 'r01' else 'round1'}.sh; "
                      f"profiles/{hot.split('/')[-1]}, profiles/{spmm.split('/')[-1]}); reddit shape, 1 B200")
    json.dump(fam, open(dst, "w"), indent=1)
    print(json.dumps(fam, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
