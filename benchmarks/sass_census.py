#!/usr/bin/env python
"""Static census of the built kernels (no GPU needed): registers, shared memory, spills from the ptxas logs and the
SASS instruction mix from cuobjdump, per kernel.  What it is for: checking a kernel change in the build container
before any GPU time is spent (did the instruction count of the hot loop move? did a kernel start spilling? is an
atomic a native RED or a CAS loop?), and recording the evidence next to the ncu captures.

    python benchmarks/sass_census.py [--csv profiles/rN_sass_census.csv] [object names ...]     # default: every object
"""
import argparse
import csv
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "baorec.jl_b200" / "build"

CLASSES = [   # class -> SASS mnemonic stems (the part before the first '.')
    ("fp32", {"FADD", "FMUL", "FFMA", "FMNMX", "FSET", "FSETP", "FSEL", "FCHK", "MUFU", "FADD2", "FMUL2", "FFMA2"}),
    ("fp64", {"DADD", "DMUL", "DFMA", "DSETP", "DMNMX"}),
    ("int", {"IADD", "IADD3", "IMAD", "LEA", "LOP3", "SHF", "ISETP", "IMNMX", "SEL", "PRMT", "POPC", "FLO", "IABS", "BREV", "VIADD",
             "VIMNMX", "UIADD3", "UIMAD", "ULEA", "ULOP3", "USHF", "UISETP", "USEL", "UMOV", "MOV", "UPRMT", "UFLO", "UPOPC"}),
    ("cvt", {"I2F", "F2I", "F2F", "I2I", "FRND", "F2FP", "I2FP"}),
    ("ldg", {"LDG"}), ("stg", {"STG"}), ("lds", {"LDS"}), ("sts", {"STS"}), ("ldgsts", {"LDGSTS"}),
    ("ldc", {"LDC", "ULDC", "LDCU"}), ("local", {"LDL", "STL"}),
    ("red_global", {"REDG", "RED"}), ("atom_global", {"ATOMG", "ATOM"}), ("atom_shared", {"ATOMS"}),
    ("shfl_vote", {"SHFL", "VOTE", "MATCH", "REDUX", "VOTEU"}), ("barrier", {"BAR", "DEPBAR", "SYNCS", "WARPSYNC", "BSSY", "BSYNC", "LDGDEPBAR"}),
    ("branch", {"BRA", "BRX", "EXIT", "RET", "CALL", "JMP", "BRXU"}), ("tma_bulk", {"UBLKCP", "UTMALDG", "UTMASTG"}),
]


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def ptxas_info(log):
    """mangled name -> (registers, static shared bytes, spill bytes, stack bytes)"""
    info, cur = {}, None
    for line in log.splitlines():
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = m.group(1)
            info.setdefault(cur, [None, 0, 0, 0])
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
        if m:
            info[cur][3], info[cur][2] = int(m.group(1)), int(m.group(2))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            info[cur][0] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            info[cur][1] = int(s.group(1)) if s else 0
    return info


def census(obj):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            stem = op.split(".")[0]
            for cls, stems in CLASSES:
                if stem in stems:
                    kernels[cur][cls] += 1
                    break
            if op.startswith("ATOMS.CAS") or op.startswith("ATOMG.CAS") or op.startswith("ATOM.CAS"):
                kernels[cur]["cas_loop"] += 1
    return kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("objects", nargs="*")
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()
    names = args.objects or sorted(p.stem for p in BUILD.glob("*.o"))
    cols = ["object", "kernel", "registers", "smem_static", "spill_bytes", "total"] + [c for c, _ in CLASSES] + ["cas_loop"]
    rows = []
    for name in names:
        obj = BUILD / f"{name}.o"
        if not obj.exists():
            sys.exit(f"{obj} missing: run python baorec.jl_b200/build.py")
        log = (BUILD / f"{name}.ptxas.log")
        info = ptxas_info(log.read_text()) if log.exists() else {}
        for mangled, cnt in census(obj).items():
            reg, smem, spill, _ = info.get(mangled, [None, 0, 0, 0])
            rows.append([name, demangle(mangled), reg, smem, spill, cnt["total"]] + [cnt[c] for c, _ in CLASSES] + [cnt["cas_loop"]])
    if args.csv:
        with open(args.csv, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(cols)
            w.writerows(rows)
        print(f"{len(rows)} kernels -> {args.csv}")
    else:
        for r in rows:
            d = dict(zip(cols, r))
            mix = " ".join(f"{k}={d[k]}" for k in cols[6:] if d[k])
            print(f"{d['object']:9s} {d['kernel'][:70]:70s} regs={d['registers']} spill={d['spill_bytes']} n={d['total']} {mix}")


if __name__ == "__main__":
    main()
