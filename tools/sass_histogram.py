#!/usr/bin/env python
"""SASS opcode histogram of every object file of libposeb200.so (no GPU needed): proves which kernels are
tcgen05 / TMEM / TMA code.  Reads the per-source objects pose_estimation_amitai_b200/build/*.o with
`cuobjdump -sass` and counts mnemonics per source file and per kernel.

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt

Blackwell mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), UTCBAR = tcgen05.commit,
LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA; .MULTICAST across a
cluster), UTMAPF = tensor-map prefetch, SYNCS = mbarrier ops, UCGABAR = cluster barrier, ELECT = elect.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "pose_estimation_amitai_b200", "build")
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "SYNCS", "UCGABAR", "ELECT",
       "HMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "FFMA", "RED", "ATOM")
INST = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)")


def histogram(obj: str):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per_fn, fn = collections.OrderedDict(), None
    for line in out.splitlines():
        if line.lstrip().startswith("Function :"):
            fn = line.split(":", 1)[1].strip()
            per_fn[fn] = collections.Counter()
            continue
        m = INST.match(line)
        if m and fn is not None:
            per_fn[fn][m.group(1)] += 1
    return per_fn


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}
    except OSError:
        return {n: n for n in names}


def family(op: str) -> str:
    return op.split(".")[0]


def main() -> None:
    total_all = collections.Counter()
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith(".o"):
            continue
        per_fn = histogram(os.path.join(OBJ, f))
        if not per_fn:
            continue
        tot = collections.Counter()
        for c in per_fn.values():
            tot.update(c)
        total_all.update(tot)
        fam = collections.Counter()
        for op, n in tot.items():
            fam[family(op)] += n
        print(f"=== csrc/{f[:-2]}.cu: {len(per_fn)} kernels, {sum(tot.values())} SASS instructions")
        print("    key mnemonics: " + "  ".join(f"{k} {fam[k]}" for k in KEY if fam[k]))
        variants = sorted(((op, n) for op, n in tot.items() if family(op) in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM")),
                          key=lambda kv: -kv[1])
        if variants:
            print("    tcgen05 / TMA variants: " + "  ".join(f"{op} {n}" for op, n in variants))
        names = demangle(list(per_fn))
        for fn, c in per_fn.items():
            fc = collections.Counter()
            for op, n in c.items():
                fc[family(op)] += n
            hot = "  ".join(f"{k} {fc[k]}" for k in KEY[:11] if fc[k])
            print(f"      {names[fn][:110]:110s} {sum(c.values()):6d} inst  {hot}")
    fam = collections.Counter()
    for op, n in total_all.items():
        fam[family(op)] += n
    print(f"=== libposeb200.so total: {sum(total_all.values())} SASS instructions")
    print("    " + "  ".join(f"{k} {fam[k]}" for k in KEY if fam[k]))


if __name__ == "__main__":
    main()
