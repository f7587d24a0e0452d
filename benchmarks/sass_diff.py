#!/usr/bin/env python
"""Which kernels did a change touch?  Builds the library of another git revision in a scratch directory and compares
the SASS of every kernel with the working tree's build, kernel by kernel (demangled names, whitespace and addresses
normalised).  A refactor that claims "no kernel changed" is checked in a minute, without a GPU; a variant added next to
a measured kernel is shown to leave the measured one byte-identical.

    python benchmarks/sass_diff.py <git-rev> [--scratch /tmp/baorec_sass_diff]
"""
import argparse
import hashlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    ks, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            ks[cur] = hashlib.sha1()
            continue
        if cur is not None:
            code = re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-fx]+\*/", "", line)).strip()
            if code:
                ks[cur].update(code.encode())
    return {k: v.hexdigest() for k, v in ks.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rev")
    ap.add_argument("--scratch", default="/tmp/baorec_sass_diff")
    args = ap.parse_args()
    scratch = Path(args.scratch) / args.rev.replace("/", "_")
    if not (scratch / "baorec.jl_b200" / "lib" / "libbaorec_b200.so").exists():
        scratch.mkdir(parents=True, exist_ok=True)
        tar = subprocess.run(["git", "-C", str(ROOT), "archive", args.rev, "baorec.jl_b200", "include"], capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", str(scratch)], input=tar, check=True)
        subprocess.run([sys.executable, str(scratch / "baorec.jl_b200" / "build.py")], check=True, capture_output=True)
    new_dir, old_dir = ROOT / "baorec.jl_b200" / "build", scratch / "baorec.jl_b200" / "build"
    total = {"identical": 0, "changed": 0, "added": 0, "removed": 0}
    for obj in sorted(set(p.name for p in new_dir.glob("*.o")) | set(p.name for p in old_dir.glob("*.o"))):
        a = kernels(old_dir / obj) if (old_dir / obj).exists() else {}
        b = kernels(new_dir / obj) if (new_dir / obj).exists() else {}
        same = [k for k in a if b.get(k) == a[k]]
        changed = [k for k in a if k in b and b[k] != a[k]]
        added, removed = [k for k in b if k not in a], [k for k in a if k not in b]
        print(f"{obj}: {len(same)} identical, {len(changed)} changed, {len(added)} added, {len(removed)} removed")
        for tag, names in (("changed", changed), ("added", added), ("removed", removed)):
            for k in names:
                print(f"    {tag}: {k[:140]}")
        for key, val in zip(total, (same, changed, added, removed)):
            total[key] += len(val)
    print(total)


if __name__ == "__main__":
    main()
