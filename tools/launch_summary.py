#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the LAST training step (between the last
two adam launches) and the last inference step (between the last two argmax_finalize launches), per kernel, in
launch order.   python tools/launch_summary.py gpurun_out/x_launches.csv [--seq]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    out = []
    for r in rows[hi + 1:]:
        if len(r) > mv and r[mv] and r[kn] and "spin_kernel" not in r[kn]:   # (bench.py's roofline gate: torch.cuda._sleep)
            try:
                out.append((r[kn].split("(")[0].replace("void ", "").replace("pb::", ""), float(r[mv].replace(",", "")) / 1e3))
            except ValueError:
                pass
    return out


def segment(names, marker, which=-1):
    idx = [i for i, (n, _) in enumerate(names) if marker in n]
    if len(idx) < 2:
        return []
    return names[idx[which - 1] + 1: idx[which] + 1]


def show(title, seg, seq):
    if not seg:
        return
    tot = sum(t for _, t in seg)
    print(f"== {title}: {len(seg)} launches, {tot:.1f} us (cold-cache, serialised)")
    if seq:
        for n, t in seg:
            print(f"   {n[:58]:58s} {t:8.1f}")
    agg = collections.OrderedDict()
    for n, t in seg:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += t
        a[1] += 1
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"   {n[:58]:58s} {t:8.1f} us {c:3d}x  {100 * t / tot:5.1f} %")


if __name__ == "__main__":
    names = load(sys.argv[1])
    seq = "--seq" in sys.argv
    show("training step", segment(names, "adam_kernel"), seq)
    show("inference step", segment(names, "argmax_finalize"), seq)
