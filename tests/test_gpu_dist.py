"""Slab-decomposed (multi-GPU) path.  With one GPU the whole distributed code path still runs
(2-D/1-D split FFT, pack + tile transposes, ghost/halo planes, slab-mode scatter/gather) with the
all-to-all degenerating to a copy; tests/multi_gpu_check.py is the same comparison under
torchrun with 2+ ranks (run with gpurun --gpus 2)."""
import numpy as np
import pytest

from util import clustered_box, rel_rms, maxabs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture()
def dctx(B):
    ctx = B.Context.get(0)
    B.dist.init_comm(ctx)
    yield ctx
    ctx.plan_key = None       # force a fresh single-GPU plan for whoever comes next


@pytest.mark.parametrize("shape", [(32, 32, 32), (48, 32, 16), (64, 64, 64)])
def test_dist_fft_roundtrip_and_matches_rfftn(B, dctx, shape):
    nx, ny, nz = shape
    B.dist.plan(dctx, shape, np.full(3, 100.0, np.float32), np.zeros(3, np.float32))
    rng = np.random.default_rng(0)
    a = rng.standard_normal((nz, ny, nx)).astype(np.float32)
    T = B.dist.dist_r2c(dctx, dev(a))                         # [y][x][z]
    ref = np.fft.rfftn(a.astype(np.float64), axes=(0, 1, 2))  # [z][y][x]
    got = T.cpu().numpy().transpose(2, 0, 1)
    assert rel_rms(got.real, ref.real) < 1e-5 and rel_rms(got.imag, ref.imag) < 1e-5
    back = torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda")
    B.dist.dist_c2r(dctx, T, back)
    assert rel_rms(back.cpu().numpy() / a.size, a) < 1e-5


def test_slab_owner_matches_host_logic(B, dctx):
    n, L = 64, 1000.0
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    B.dist.plan(dctx, (n, n, n), bs, bm)
    rng = np.random.default_rng(1)
    z = np.concatenate([(rng.random(5000) * L).astype(np.float32), np.float32([0.0, L, L + 3.0, -1.0])])
    got = B.dist.slab_owner(dctx, dev(z)).cpu().numpy()
    for world in (1,):
        assert np.array_equal(got, B.dist.owner_of_z(z, 0.0, L, n, world))
    assert got[-1] == -1 and got[-2] == 0 and got[-3] == 0


@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0)])
def test_run_dist_matches_single_gpu_and_oracle(B, O, dctx, los):
    n, L, N = 64, 1000.0, 300_000
    pos, w = clustered_box(N, L, seed=5)
    pos[2][:40] += np.float32(L)          # exercises the wrap + write-back in slab mode
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=los, n_iter=3)
    orec = O.IterativeRecon(**kw)
    opos = [p.copy() for p in pos]
    omesh = O.run(orec, (n, n, n), *opos, w)
    oshift = O.read_shifts(orec, *opos, omesh, "sum")
    d = [dev(p) for p in pos]
    rec = B.IterativeRecon(**kw)
    mesh = B.dist.run_dist(rec, (n, n, n), *d, dev(w), ctx=dctx)
    assert mesh.shape == (n, n, n)
    assert rel_rms(mesh.cpu().numpy(), omesh) < 1e-4
    for g, o in zip(d, opos):                                   # wrapped positions written back
        assert np.array_equal(g.cpu().numpy().view(np.uint32), o.view(np.uint32))
    for f in ("disp", "sum"):
        s = B.dist.read_shifts_dist(rec, *d, field=f)
        ref = O.read_shifts(orec, *opos, omesh, f)
        for a in range(3):
            assert rel_rms(s[a].cpu().numpy(), ref[a]) < 1e-4
            assert maxabs(s[a].cpu().numpy(), ref[a]) < 1e-3
    newpos = B.dist.read_shifts_dist(rec, *d, field="sum", positions=True)
    for a in range(3):
        assert maxabs(newpos[a].cpu().numpy(), opos[a] - oshift[a]) < 1e-3


def test_run_dist_rejects_unsupported_modes(B, dctx):
    n, L = 32, 100.0
    e = torch.zeros(4, dtype=torch.float32, device="cuda") + 5
    rec = B.IterativeRecon(bias=2.0, f=0.5, smoothing_radius=5.0, box_size=np.full(3, L, np.float32),
                           box_min=np.zeros(3, np.float32), los=None)
    with pytest.raises(B.BaorecError):
        B.dist.run_dist(rec, (n, n, n), e, e, e, e, ctx=dctx)
