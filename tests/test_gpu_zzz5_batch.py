"""GPU: the batched host pipeline (baorec_batch_host_f32 / B.run_batch: run! + reconstructed_positions per catalog
with the PCIe transfers of neighbouring catalogs overlapping the solve) against the one-catalog-at-a-time host
pipeline (baorec_run_host_f32 + baorec_read_host_f32, validated on hardware) and the oracle.
The same pipeline fed from catalog files (baorec_batch_files_f32 / B.run_batch_files: reader and writer threads, pinned
buffer sets, staging that grows) is held to the array version catalog by catalog."""
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


@pytest.mark.parametrize("slots", [3, 4])
def test_batch_files_equals_batch_arrays(B, tmp_path, slots):
    """Text and NPY catalogs of growing and shrinking sizes through B.run_batch_files: every output file holds what
    B.run_batch returns for the same catalog (same kernels, same order; float reductions of the scatter apart)."""
    n, L = 64, 1000.0
    bs, bm = np.full(3, L, f32), np.zeros(3, f32)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=bs, box_min=bm, los=(0.0, 0.0, 1.0), n_iter=3)
    cats, ins, outs = [], [], []
    for i, (seed, N) in enumerate(((1, 30_000), (2, 300_000), (3, 20_000), (4, 350_000), (5, 280_000))):
        pos, w = clustered_box(N, L, seed=seed)
        a = np.stack([pos[0], pos[1], w, pos[2]], axis=1)                 # x y w z: columns are picked, not assumed
        if i % 2:
            q = tmp_path / f"mock_{i}.npy"
            np.save(q, a)
        else:
            q = tmp_path / f"mock_{i}.dat"
            np.savetxt(q, a, fmt="%.9g")
        ins.append(q)
        outs.append(tmp_path / f"rec_{i}.npy")
        cats.append(tuple(pinned(v) for v in (*pos, w)))
    rec = B.IterativeRecon(**kw)
    B.Context.get(0).set_option("batch_slots", slots)
    try:
        info = B.run_batch_files(rec, (n, n, n), ins, outs, columns=(0, 1, 3, 2), field="sum", positions=True, n_threads=4)
    finally:
        B.Context.get(0).set_option("batch_slots", 4)
    assert info["rows"] == [len(c[0]) for c in cats]
    assert info["total_s"] > 0 and info["read_s"] > 0 and info["write_s"] > 0
    want = B.run_batch(B.IterativeRecon(**kw), (n, n, n), cats, field="sum", positions=True)
    for o, ref in zip(outs, want):
        got = np.load(o)
        assert got.dtype == np.float32 and got.shape == (len(ref[0]), 3)
        for a in range(3):
            assert maxabs(got[:, a], ref[a]) < 5e-4 + 2 * float(np.spacing(f32(L)))
    # weights of one when no column is named; shifts instead of positions; no output files
    ones = [(c[0], c[1], c[2], pinned(np.ones_like(c[3]))) for c in cats[:2]]
    outs2 = [tmp_path / "s0.npy", tmp_path / "s1.npy"]
    B.run_batch_files(B.IterativeRecon(**kw), (n, n, n), ins[:2], outs2, columns=(0, 1, 3, -1), field="disp", positions=False)
    want = B.run_batch(B.IterativeRecon(**kw), (n, n, n), ones, field="disp", positions=False)
    for o, ref in zip(outs2, want):
        got = np.load(o)
        for a in range(3):
            assert maxabs(got[:, a], ref[a]) < 5e-4
    assert B.run_batch_files(B.IterativeRecon(**kw), (n, n, n), ins[:1], None)["rows"] == [30_000]


def test_batch_files_reports_io_errors(B, tmp_path):
    n, L = 32, 100.0
    rec = B.IterativeRecon(bias=2.0, f=0.5, smoothing_radius=5.0, box_size=np.full(3, L, f32), box_min=np.zeros(3, f32),
                           los=(0.0, 0.0, 1.0))
    rng = np.random.default_rng(0)
    good = tmp_path / "good.dat"
    np.savetxt(good, rng.uniform(0, L, (5000, 3)).astype(f32), fmt="%.9g")
    with pytest.raises(B.CatalogIOError, match="missing.dat"):
        B.run_batch_files(rec, (n, n, n), [good, tmp_path / "missing.dat", good], None)
    bad = tmp_path / "bad.dat"
    bad.write_text("1 2 3\n4 five 6\n")
    with pytest.raises(B.CatalogIOError, match="line 2"):
        B.run_batch_files(rec, (n, n, n), [good, bad], None)
    outside = tmp_path / "outside.dat"
    np.savetxt(outside, np.array([[1, 2, 3], [5 * L, 1, 1]], f32), fmt="%.9g")      # a particle far outside the box
    with pytest.raises(B.OutOfBoxError):
        B.run_batch_files(rec, (n, n, n), [good, outside, good], None)
    # and the context still works afterwards
    assert B.run_batch_files(rec, (n, n, n), [good], [tmp_path / "ok.npy"])["rows"] == [5000]
    assert np.load(tmp_path / "ok.npy").shape == (5000, 3)
