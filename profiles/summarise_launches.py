"""Summarises an `ncu --metrics gpu__time_duration.sum` launch list (CSV) of bench.py:
per-kernel totals/shares and, for one warm round, the per-layer-class time of the conv stack.
usage: python profiles/summarise_launches.py <launches.csv> [round_index]"""
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ivos-w_b200"))
from ivosw import arch  # noqa: E402


def main(path, round_index=2, B=128):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    idx = {h: i for i, h in enumerate(rows[0])}
    seq = []
    for r in rows[1:]:
        try:
            seq.append((r[idx["Kernel Name"]].split("(")[0], float(r[idx["Metric Value"]].replace(",", ""))))
        except ValueError:
            pass
    agg = collections.OrderedDict()
    for k, v in seq:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("## kernels (all captured launches; ncu times are cold-cache and serialised: compare shares)")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% |" % (k[:70], a[0], a[1] / 1e3, 100 * a[1] / tot))
    starts = [i for i, (k, v) in enumerate(seq) if "split_kernel" in k or "stem_tc_kernel" in k]
    if len(starts) <= round_index:
        return
    s = starts[round_index]
    convs = [(k, v) for k, v in seq[s + 1:s + 1 + 70] if "conv_tc" in k][:52]
    specs = arch.resnet50_convs()
    cls = collections.OrderedDict()
    total = 0.0
    for sp, (k, v) in zip(specs, convs):
        fl = 2 * B * sp.out_hw ** 2 * sp.cout * sp.cin * sp.k * sp.k
        byt = B * (sp.in_hw ** 2 * sp.cin + sp.out_hw ** 2 * sp.cout * (2 if sp.residual else 1)) * 4
        total += v
        key = (sp.cin, sp.cout, sp.k, sp.stride, sp.out_hw, bool(sp.residual), k.split("<")[-1])
        a = cls.setdefault(key, [0, 0.0, fl, byt]); a[0] += 1; a[1] += v
    print("\n## conv stack, round %d (B = %d units): %.2f ms" % (round_index, B, total / 1e6))
    print("| cin | cout | k | stride | out hw | residual | variant | layers | avg us | algorithmic TFLOP/s | activation GB/s |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for key, a in cls.items():
        avg = a[1] / a[0]
        print("| %d | %d | %d | %d | %d | %s | %s | %d | %.1f | %.1f | %.0f |" %
              (key[0], key[1], key[2], key[3], key[4], "y" if key[5] else "n", key[6], a[0], avg / 1e3, a[2] / avg / 1e3, a[3] / avg))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
