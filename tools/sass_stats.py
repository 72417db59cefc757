#!/usr/bin/env python3
"""Static figures of the Ax kernels from the built objects (no GPU needed): registers, spills, instruction mix.
usage: tools/sass_stats.py > profiles/<tag>_static_sass.md      (after python -m libnomp_b200.build)
What `-Xptxas -v` and `cuobjdump -sass` say before any GPU time is spent: does a variant fit its register budget, does
every DFMA still take its D entry from a uniform register, how many shared-memory and global instructions per element."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "libnomp_b200" / "build"
LABELS = {  # template arguments <N, G, W, GPC, GA, PF, ST, MB, DOT, PERSISTENT, TWOBUF, XPAY> -> what it is
    "dot": lambda a: a[8] == "1" and a[11] == "0",
    "xpay+dot": lambda a: a[11] == "1",
}


def resource_usage(obj):
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", str(obj)], capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), local=int(m.group(5)))
    return res


def sass_mix(obj):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    mix, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            mix[name] = dict(total=0, dfma=0, dfma_ur=0, lds=0, sts=0, ldg=0, stg=0, ldcu=0, bar=0)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if name and m:
            op = m.group(1)
            d = mix[name]
            d["total"] += 1
            if op.startswith("DFMA"):
                d["dfma"] += 1
                d["dfma_ur"] += " UR" in line
            for key, prefix in (("lds", "LDS"), ("sts", "STS"), ("ldg", "LDG"), ("stg", "STG"), ("ldcu", "LDCU"), ("bar", "BAR")):
                d[key] += op.startswith(prefix)
    return mix


def template_args(mangled):
    m = re.search(r"ax_kernelI(.*?)EEv", mangled)
    if not m:
        return None
    return [x[1:] if x.startswith("b") else x for x in re.findall(r"L[ib](\d+)E", m.group(1))]


def main():
    print("# Ax kernels: static figures from the built objects (`tools/sass_stats.py`)\n")
    print("Template arguments: n, elements per group, warps per group, groups per CTA, geometric slabs in flight, L2 prefetch "
          "distance, streaming loads, min CTAs/SM, fused dot, persistent grid, two shared buffers, fused direction update, prefetch "
          "mode (kPfMode), pinned ring fill (kPin), even-odd stage mask (kEO: 63 = all six stages).\n")
    print("| n | G,W,GPC | GA | PF | minCTA | dot | persistent | two-buf | xpay | pfmode | pin | EO | registers | stack B | instructions | DFMA | with UR operand | LDS | STS | LDG | STG | LDCU |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for n in (6, 8, 10, 12):
        res, mix = {}, {}
        for obj in (OBJ / f"ax_n{n}.cu.o", OBJ / f"ax_n{n}_p1.cu.o", OBJ / f"ax_n{n}_p2.cu.o"):   # production + profiling shapes
            if obj.exists():
                res.update(resource_usage(obj))
                mix.update(sass_mix(obj))
        rows = []
        for name, r in res.items():
            a = template_args(name)
            if a is None or len(a) < 11:
                continue
            a = a + ["0"] * (15 - len(a))
            s = mix.get(name, {})
            rows.append((a, r, s))
        rows.sort(key=lambda x: [int(v) for v in x[0]])
        for a, r, s in rows:
            print(f"| {a[0]} | {a[1]},{a[2]},{a[3]} | {a[4]} | {a[5]} | {a[7]} | {a[8]} | {a[9]} | {a[10]} | {a[11]} | {a[12]} | {a[13]} | {a[14]} | {r['reg']} | {r['stack']} | "
                  f"{s.get('total', 0)} | {s.get('dfma', 0)} | {s.get('dfma_ur', 0)} | {s.get('lds', 0)} | {s.get('sts', 0)} | {s.get('ldg', 0)} | "
                  f"{s.get('stg', 0)} | {s.get('ldcu', 0)} |")


if __name__ == "__main__":
    main()
