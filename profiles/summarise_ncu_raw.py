"""Condenses `ncu -i rep --page raw --csv` of the conv_tc launches of one round into a per-layer table:
duration, tensor-pipe %, DRAM read/write bytes and throughput %, algorithmic bytes/FLOPs.
Round 2: 48 launches per round — the downsample convolution of a stage's first block is fused into that block's conv3
(one concatenated-K GEMM), so those rows carry both convolutions' FLOPs and read two inputs.  Pass `unfused` as the
third argument for captures taken with IVOSW_FUSE_DS=0 (52 launches).
usage: python profiles/summarise_ncu_raw.py raw.csv [B] [unfused]"""
import csv
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ivos-w_b200"))
from ivosw import arch  # noqa: E402

COLS = {
    "t_us": "gpu__time_duration.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)


def fused_specs():
    """48 (spec, extra_flops_per_unit, extra_input_elems_per_unit) rows: downsample folded into the following conv3"""
    out, pend = [], None
    for sp in arch.resnet50_convs():
        if sp.name.endswith("downsample.0"):
            pend = sp
            continue
        if pend is not None and sp.name.endswith("conv3"):
            out.append((sp._replace(name=sp.name + "+ds", residual=""), 2.0 * pend.out_hw ** 2 * pend.cout * pend.cin,
                        pend.in_hw ** 2 * pend.cin, pend.cout * pend.cin))
            pend = None
        else:
            out.append((sp, 0.0, 0, 0))
    return out


def main(path, B=128, unfused=False):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(v) for k, v in COLS.items()}
    name_i = hdr.index("Kernel Name")
    specs = [(sp, 0.0, 0, 0) for sp in arch.resnet50_convs()] if unfused else fused_specs()
    data = rows[2:]
    stem = [i for i, r in enumerate(data) if "stem_tc" in r[name_i]]
    if stem:   # capture window straddles two rounds: rotate so the rows line up with layer 0..51
        k = stem[0]
        srow = data[k]
        data = data[k + 1:] + data[:k]
        tu0 = units[ix["t_us"]]
        t = float(srow[ix["t_us"]].replace(",", ""))
        t = t / 1e3 if tu0 in ("ns", "nsecond") else (t if tu0 in ("us", "usecond") else t * 1e3)
        print("stem_tc_kernel: %.1f us, tensor pipe %s %%, DRAM rd %.1f MB wr %.1f MB, regs %s\n" %
              (t, srow[ix["tensor"]], to_bytes(srow[ix["dram_rd"]], units[ix["dram_rd"]]) / 1e6,
               to_bytes(srow[ix["dram_wr"]], units[ix["dram_wr"]]) / 1e6, srow[ix["regs"]]))
    rows = rows[:2] + data
    print("| # | layer | variant | us | tensor pipe % | DRAM rd MB | DRAM wr MB | DRAM % | alg MB (in+out+res+w) | alg TFLOP/s | traffic/alg |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    tot = {"t": 0.0, "rd": 0.0, "wr": 0.0, "alg": 0.0, "fl": 0.0, "tw": 0.0}
    for i, ((sp, xfl, xin, xw), r) in enumerate(zip(specs, rows[2:])):
        t = float(r[ix["t_us"]].replace(",", ""))
        tu = units[ix["t_us"]]
        t_us = t / 1e3 if tu in ("ns", "nsecond") else (t if tu in ("us", "usecond") else t * 1e3)
        rd = to_bytes(r[ix["dram_rd"]], units[ix["dram_rd"]])
        wr = to_bytes(r[ix["dram_wr"]], units[ix["dram_wr"]])
        fl = 2.0 * B * sp.out_hw ** 2 * sp.cout * sp.cin * sp.k * sp.k + B * xfl
        alg = B * (sp.in_hw ** 2 * sp.cin + xin + sp.out_hw ** 2 * sp.cout * (2 if sp.residual else 1)) * 4 + \
            (sp.cout * sp.cin * sp.k * sp.k + xw) * 4
        tens = float(r[ix["tensor"]])
        var = r[name_i].split("<")[-1].split(">")[0]
        print("| %d | %s | %s | %.1f | %.1f | %.1f | %.1f | %s | %.1f | %.1f | %.2f |" %
              (i, sp.name[8:], var, t_us, tens, rd / 1e6, wr / 1e6, r[ix["dram_pct"]], alg / 1e6, fl / t_us / 1e6,
               (rd + wr) / alg))
        tot["t"] += t_us; tot["rd"] += rd; tot["wr"] += wr; tot["alg"] += alg; tot["fl"] += fl; tot["tw"] += tens * t_us
    print("\nconv stack: %.2f ms under ncu (cold cache, serialised); DRAM read %.2f GB + write %.2f GB = %.2f GB vs "
          "algorithmic %.2f GB (x%.2f); time-weighted tensor pipe %.1f %%; algorithmic %.1f TFLOP/s" %
          (tot["t"] / 1e3, tot["rd"] / 1e9, tot["wr"] / 1e9, (tot["rd"] + tot["wr"]) / 1e9, tot["alg"] / 1e9,
           (tot["rd"] + tot["wr"]) / tot["alg"], tot["tw"] / tot["t"], tot["fl"] / tot["t"] / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 128, len(sys.argv) > 3 and sys.argv[3] == "unfused")
