#!/usr/bin/env python
"""Many mocks in one process -- the workflow the reference's README names ("one process, many reconstructions";
its example scripts are run once per mock): every catalog is reconstructed and read back, with the PCIe transfers of
neighbouring mocks overlapping the solve (BAOrec.run_batch -> baorec_batch_host_f32).  The DESI mocks the reference
reads are not public, so synthetic lognormal boxes stand in (--npy DIR loads DIR/*.npy, one (n, 3) or (n, 4) array
per mock: x, y, z[, w]).

--files DIR is the same loop from catalog FILES to result files, the way the reference's scripts bracket run! with
CSV.jl and NPZ.jl (examples/simulation.jl:12-40): every DIR/*.dat (text: x y z [w]) or DIR/*.npy is read and parsed by
the library's reader thread into pinned memory while the device reconstructs its predecessor, and a writer thread
stores OUT/<name>.rec.npy as npzwrite(fn, hcat(new_pos...)) would (BAOrec.run_batch_files -> baorec_batch_files_f32).

    python examples/many_mocks.py [--mocks 4] [--grid 512] [--particles 5e6] [--npy DIR | --files DIR] [--out DIR]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "benchmarks"))
import __graft_entry__ as G  # noqa: E402
import catalogs  # noqa: E402

BAOrec = G.load_package()


def pinned_columns(arr):
    """(n, 3|4) array -> four pinned float32 columns (x, y, z, w): page-locked memory is what lets the copies overlap."""
    n = len(arr)
    cols = []
    for c in range(4):
        t = torch.empty(n, dtype=torch.float32, pin_memory=True)
        t.copy_(torch.from_numpy(np.ascontiguousarray(arr[:, c], dtype=np.float32)) if c < arr.shape[1] else torch.ones(n))
        cols.append(t.numpy())
    return tuple(cols)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mocks", type=int, default=4)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--particles", type=float, default=5e6)
    ap.add_argument("--npy", default=None)
    ap.add_argument("--files", default=None, help="directory of *.dat / *.npy catalogs: files in, files out")
    ap.add_argument("--weights-column", type=int, default=-1, help="--files: 0-based column of the weights (-1: ones)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    L = 1000.0
    box_size, box_min = np.float32([L, L, L]), np.float32([0.0, 0.0, 0.0])
    if args.files:
        files = sorted(p for p in Path(args.files).iterdir() if p.suffix in (".dat", ".txt", ".npy"))
        out = Path(args.out or args.files)
        out.mkdir(parents=True, exist_ok=True)
        recon = BAOrec.IterativeRecon(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=box_size, box_min=box_min,
                                      los=(0.0, 0.0, 1.0), n_iter=3)
        info = BAOrec.run_batch_files(recon, (args.grid,) * 3, files, [out / f"{f.stem}.rec.npy" for f in files],
                                      columns=(0, 1, 2, args.weights_column), field="sum", positions=True)
        print(f"{len(files)} catalogs ({sum(info['rows'])} rows), {args.grid}^3: {1e3 * info['total_s'] / len(files):.1f} ms per catalog "
              f"(reader {info['read_s']:.2f} s, writer {info['write_s']:.2f} s, device waited {info['wait_s']:.2f} s for input)")
        return
    if args.npy:
        files = sorted(Path(args.npy).glob("*.npy"))
        mocks = [pinned_columns(np.load(f)) for f in files]
        names = [f.stem for f in files]
    else:
        mocks, names = [], []
        for i in range(args.mocks):
            pos, w = catalogs.lognormal_box(int(args.particles), L, seed=100 + i, device="cuda", n_gen=256, f_rsd=0.757)
            mocks.append(pinned_columns(torch.stack([*pos, w], 1).cpu().numpy()))
            names.append(f"mock{i:03d}")
    recon = BAOrec.IterativeRecon(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=box_size, box_min=box_min,
                                  los=(0.0, 0.0, 1.0), n_iter=3)
    outs = [tuple(torch.empty(len(m[0]), dtype=torch.float32, pin_memory=True).numpy() for _ in range(3)) for m in mocks]
    grid = (args.grid,) * 3
    BAOrec.run_batch(recon, grid, mocks[:1], field="sum", out=outs[:1])          # plans, scratch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    BAOrec.run_batch(recon, grid, mocks, field="sum", positions=True, out=outs)
    dt = time.perf_counter() - t0
    print(f"{len(mocks)} mocks, {args.grid}^3: {1e3 * dt / len(mocks):.1f} ms per mock, host catalog in -> reconstructed positions out")
    if args.out:
        out = Path(args.out)
        out.mkdir(parents=True, exist_ok=True)
        for name, o in zip(names, outs):
            np.save(out / f"{name}.dat.rec.npy", np.stack(o, axis=1))


if __name__ == "__main__":
    main()
