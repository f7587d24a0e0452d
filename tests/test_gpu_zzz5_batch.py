"""GPU: the batched host pipeline (baorec_batch_host_f32 / B.run_batch: run! + reconstructed_positions per catalog
with the PCIe transfers of neighbouring catalogs overlapping the solve) against the one-catalog-at-a-time host
pipeline (baorec_run_host_f32 + baorec_read_host_f32, validated on hardware) and the oracle.
Only stream / event orchestration over functions that have run on a B200 -- but the orchestration itself was
written after this round's GPU budget was spent and has NOT YET RUN ON HARDWARE; hence the file name."""
import numpy as np
import pytest

from util import clustered_box, maxabs, rel_rms

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
f32 = np.float32


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
    t.copy_(torch.from_numpy(a))
    return t.numpy()


@pytest.mark.parametrize("algorithm", ["iterative", "multigrid"])
@pytest.mark.parametrize("field,positions", [("sum", True), ("disp", False)])
def test_batch_equals_one_at_a_time(B, O, algorithm, field, positions):
    n, L = 64, 1000.0
    bs, bm = np.full(3, L, f32), np.zeros(3, f32)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=bs, box_min=bm, los=(0.0, 0.0, 1.0))
    make = (lambda: B.IterativeRecon(n_iter=3, **kw)) if algorithm == "iterative" else (lambda: B.MultigridRecon(**kw))
    cats = []
    for seed, N in ((1, 300_000), (2, 50_000), (3, 400_000), (4, 20_000), (5, 300_000)):   # binned and catalog-order paths, odd/even buffers
        pos, w = clustered_box(N, L, seed=seed)
        if seed == 3:
            pos[2][:25] += f32(L)                            # wrapped by cic! and written back
        cats.append(tuple(pinned(a) for a in (*pos, w)))
    originals = [tuple(a.copy() for a in c) for c in cats]
    got = B.run_batch(make(), (n, n, n), cats, field=field, positions=positions)
    for i, (orig, cat) in enumerate(zip(originals, cats)):
        x, y, z, w = (a.copy() for a in orig)
        rec = make()
        B.run(rec, (n, n, n), x, y, z, w)
        ref = (B.reconstructed_positions(rec, x, y, z, field=field) if positions
               else B.read_shifts(rec, x, y, z, rec.result_cache, field=field))
        for a in range(3):
            assert np.array_equal(cat[a], (x, y, z)[a])      # same write-back of wrapped positions
            # float reductions of the scatter reorder: rounding only (first hardware run: 2.0e-4 Mpc/h on a multigrid
            # shift; a reconstructed position additionally carries 2 ulp of a coordinate near L = 1000)
            assert maxabs(got[i][a], ref[a]) < 5e-4 + (2 * float(np.spacing(f32(L))) if positions else 0.0)
    # and against the oracle for one of them
    x, y, z, w = (a.copy() for a in originals[0])
    orec = (O.IterativeRecon(n_iter=3, **kw) if algorithm == "iterative" else O.MultigridRecon(**kw))
    omesh = O.run(orec, (n, n, n), x, y, z, w)
    oref = O.read_shifts(orec, x, y, z, omesh, field)
    for a in range(3):
        want = (x, y, z)[a] - oref[a] if positions else oref[a]
        assert maxabs(got[0][a], want) < 1e-3


def test_batch_rejects_bad_arguments(B):
    n, L = 32, 100.0
    rec = B.IterativeRecon(bias=2.0, f=0.5, smoothing_radius=5.0, box_size=np.full(3, L, f32), box_min=np.zeros(3, f32),
                           los=(0.0, 0.0, 1.0))
    with pytest.raises(B.BaorecError):
        B.run_batch(rec, (n, n, n), [])
    empty = tuple(np.empty(0, f32) for _ in range(4))
    with pytest.raises(B.BaorecError):
        B.run_batch(rec, (n, n, n), [empty])
