"""Multi-rank parity check of the slab-decomposed path (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank builds the same seeded catalog, keeps the particles of its own slab, runs the
distributed reconstruction and compares its mesh slab and its particles' shifts with the CPU
oracle (tolerances of BASELINE.json: rel. rms <= 1e-4, max |ds| <= 1e-3 Mpc/h)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))

import __graft_entry__ as G  # noqa: E402
import baorec_oracle as O  # noqa: E402
from util import clustered_box, rel_rms, maxabs  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = G.load_package()
    ctx = B.Context.get(local)
    B.dist.init_comm(ctx)
    ok = True
    for n, los in ((64, (0.0, 0.0, 1.0)), (96, (0.0, 1.0, 0.0))):
        if n % world:
            continue
        L, N = 1000.0, 400_000
        pos, w = clustered_box(N, L, seed=11)
        pos[2][:64] += np.float32(L)                      # wrap across the last/first slab
        kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
                  box_min=np.zeros(3, np.float32), los=los, n_iter=3)
        orec = O.IterativeRecon(**kw)
        opos = [p.copy() for p in pos]
        omesh = O.run(orec, (n, n, n), *opos, w)
        oshift = O.read_shifts(orec, *opos, omesh, "sum")
        mine = B.dist.owner_of_z(pos[2], 0.0, L, n, world) == rank
        d = [torch.from_numpy(p[mine]).cuda() for p in pos]
        rec = B.IterativeRecon(**kw)
        mesh = B.dist.run_dist(rec, (n, n, n), *d, torch.from_numpy(w[mine]).cuda(), ctx=ctx)
        z_lo, nzl = B.dist.slab_range(ctx)
        e_mesh = rel_rms(mesh.cpu().numpy(), omesh[z_lo:z_lo + nzl])
        s = B.dist.read_shifts_dist(rec, *d, field="sum")
        e_rms = max(rel_rms(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        e_max = max(maxabs(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        good = e_mesh < 1e-4 and e_rms < 1e-4 and e_max < 1e-3
        ok &= good
        print(f"[rank {rank}/{world}] n={n} los={los} particles={int(mine.sum())} slab=[{z_lo},{z_lo + nzl}) "
              f"mesh rel.rms={e_mesh:.2e} shift rel.rms={e_rms:.2e} max={e_max:.2e} {'OK' if good else 'FAIL'}",
              flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
