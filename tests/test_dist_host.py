"""Host-side logic of the multi-GPU path on CPU: slab ownership and the particle redistribution
(torch.distributed all_to_all over gloo, world_size 2)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_owner_of_z_matches_oracle_base_plane(B, O):
    n, L, lo = 64, 1000.0, -250.0
    rng = np.random.default_rng(0)
    z = np.concatenate([(lo + rng.random(20000) * L).astype(np.float32),
                        np.float32([lo, lo + L, lo + L + 2.0, lo + L / 2, lo - 1.0, np.nan])])
    x = np.full_like(z, lo + 1.0)
    bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
    _, i0, _, _, _ = O.cic_cells(x, x, z, (n, n, n), bs, bm, True)
    for world in (1, 2, 4, 8):
        own = B.dist.owner_of_z(z, lo, L, n, world)
        valid = (i0[2] >= 0) & (i0[2] < n)
        assert np.array_equal(own[valid], (i0[2][valid] // (n // world)).astype(np.int32))
        assert (own[~valid] == -1).all() and own[-1] == -1 and own[-2] == -1


def test_shard_catalog_partitions_everything(B):
    n, L = 32, 100.0
    rng = np.random.default_rng(1)
    cols = [(rng.random(5000) * L).astype(np.float32) for _ in range(4)]
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    parts = B.dist.shard_catalog(*cols, bs, bm, n, 4)
    assert sum(len(p[0]) for p in parts) == 5000
    for r, p in enumerate(parts):
        assert (B.dist.owner_of_z(p[2], 0.0, L, n, 4) == r).all()
    with pytest.raises(B.OutOfBoxError):
        B.dist.shard_catalog(cols[0], cols[1], np.float32([-5.0] * 5000), cols[3], bs, bm, n, 4)


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "oracle"))
        import torch
        import torch.distributed as dist
        import __graft_entry__ as G
        B = G.load_package()
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        n, L = 32, 100.0
        rng = np.random.default_rng(100 + rank)
        m = 3000 + 500 * rank
        cols = [torch.from_numpy((rng.random(m) * L).astype(np.float32)) for _ in range(3)]
        tag = torch.arange(m, dtype=torch.float32) + 10000 * rank       # weights double as identity tags
        bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
        x, y, z, w, route = B.dist.exchange_catalog_host(*cols, tag, bs, bm, n)
        own = B.dist.owner_of_z(z.numpy(), 0.0, L, n, world)
        total = torch.tensor([len(x)], dtype=torch.int64)
        dist.all_reduce(total)
        chk = torch.tensor([float(w.double().sum())], dtype=torch.float64)
        dist.all_reduce(chk)
        expect = sum(float((np.arange(3000 + 500 * r) + 10000 * r).sum()) for r in range(world))
        ok = bool((own == rank).all()) and int(total) == sum(3000 + 500 * r for r in range(world)) \
            and abs(float(chk) - expect) < 1e-3 and len(x) == len(y) == len(z) == len(w)
        # the way back: a per-particle result computed in slab order lands on its particle in the caller's order
        bx, bt = B.dist.unshard_host(route, 2.0 * x + z, w)
        ok = ok and torch.equal(bt, tag) and torch.equal(bx, 2.0 * cols[0] + cols[2])
        # an out-of-box particle on ONE rank raises on EVERY rank (no rank is left blocked in the collective)
        bad = cols[2].clone()
        if rank == 1:
            bad[7] = -5.0
        try:
            B.dist.exchange_catalog_host(cols[0], cols[1], bad, tag, bs, bm, n)
            ok = False
        except B.OutOfBoxError:
            pass
        # setup_box over the ranks == setup_box of the concatenated catalog (src/utils.jl:100-109)
        import baorec_oracle as O
        rngs = [np.random.default_rng(100 + r) for r in range(world)]
        allc = [[(g.random(3000 + 500 * r) * L).astype(np.float32) for _ in range(3)] for r, g in enumerate(rngs)]
        full = [np.concatenate([allc[r][a] for r in range(world)]) for a in range(3)]
        obs, obm = O.setup_box(*full, np.float32(500))
        gbs, gbm = B.dist.setup_box_dist(*cols, 500.0)
        ok = ok and np.array_equal(obs.view(np.uint32), gbs.view(np.uint32)) \
            and np.array_equal(obm.view(np.uint32), gbm.view(np.uint32))
        dist.destroy_process_group()
        q.put((rank, ok, ""))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))


def test_exchange_catalog_gloo_world2(B):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
