#!/usr/bin/env python
"""The file-driven batch pipeline (B.run_batch_files / baorec_batch_files_f32) at a given size: K synthetic NPY catalogs
(x y z w) in a scratch directory -> K NPY files of reconstructed positions, one JSON line with the seconds per catalog
of the whole call, of the reader and writer threads, and of the device pipeline waiting for the reader -- next to
B.run_batch on the same catalogs already in pinned memory (the I/O-free figure).

    python benchmarks/batch_files_bench.py [--mesh 512] [--particles 1e7] [--catalogs 4] [--dir /tmp/baorec_io]"""
import argparse
import json
import shutil
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

B = G.load_package()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=512)
    ap.add_argument("--particles", type=float, default=1e7)
    ap.add_argument("--catalogs", type=int, default=4)
    ap.add_argument("--dir", default="/tmp/baorec_io")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    n, N, K = args.mesh, int(args.particles), args.catalogs
    L = 2500.0 * n / 1024.0
    d = Path(args.dir)
    d.mkdir(parents=True, exist_ok=True)
    ins, outs, cats = [], [], []
    for i in range(K):
        g = torch.Generator(device="cuda").manual_seed(100 + i)
        a = torch.rand((4, N), generator=g, device="cuda", dtype=torch.float32)
        a[:3] *= L * (1 - 1e-6)
        a[3] = 0.5 + a[3]
        h = a.cpu().numpy()
        p = d / f"mock_{i}.npy"
        np.save(p, h.T)                      # (N, 4) in Fortran memory order: what NPZ.jl writes for hcat(x, y, z, w)
        ins.append(p)
        outs.append(d / f"rec_{i}.npy")
        cats.append(tuple(torch.from_numpy(h[c].copy()).pin_memory().numpy() for c in range(4)))
        del a
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, n_iter=3, los=(0.0, 0.0, 1.0), box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32))
    res = {"mesh": n, "particles": N, "catalogs": K}
    for rep in ("warm", "timed"):            # the first call pins the buffer sets and builds the plans
        rec = B.IterativeRecon(**kw)
        t0 = time.perf_counter()
        info = B.run_batch_files(rec, (n, n, n), ins, outs, columns=(0, 1, 2, 3), field="sum", n_threads=args.threads)
        wall = time.perf_counter() - t0
        res[rep] = {"wall_s_per_catalog": wall / K, "read_s_per_catalog": info["read_s"] / K,
                    "write_s_per_catalog": info["write_s"] / K, "device_waited_for_reader_s_per_catalog": info["wait_s"] / K}
    rec = B.IterativeRecon(**kw)
    B.run_batch(rec, (n, n, n), cats, field="sum")
    t0 = time.perf_counter()
    got = B.run_batch(rec, (n, n, n), cats, field="sum")
    res["arrays_in_pinned_memory_s_per_catalog"] = (time.perf_counter() - t0) / K
    back = np.load(outs[-1])
    res["max_abs_diff_files_vs_arrays"] = float(max(np.abs(back[:, a] - got[-1][a]).max() for a in range(3)))
    res["input_bytes_per_catalog"] = ins[0].stat().st_size
    res["output_bytes_per_catalog"] = outs[0].stat().st_size
    print(json.dumps(res))
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
