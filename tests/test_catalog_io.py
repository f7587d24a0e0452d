"""CPU: catalog files either side of the path -- the library's text / NPY readers and writers and the
row selection, held to numpy (loadtxt, load, save, boolean masks) on the formats the reference's examples use
(examples/simulation.jl:12-13, 38-40; examples/lightcone.jl:22-26).  No GPU, no context."""
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def IO(B):
    return B.catalog_io


def _np(ts):
    return [t.numpy() for t in ts]


@pytest.mark.parametrize("n", [0, 1, 7, 100_003])
@pytest.mark.parametrize("threads", [1, 5])
def test_text_roundtrip_is_bit_exact(B, IO, tmp_path, n, threads):
    """repr(float32) is the shortest string that rounds back to the value: a correctly rounded Float32 parser (CSV.jl's,
    std::from_chars) returns the very bits, whatever the thread count and wherever the chunk boundaries fall."""
    rng = np.random.default_rng(n + threads)
    cols = [rng.uniform(-1000, 1000, n).astype(np.float32), rng.standard_normal(n).astype(np.float32) * 1e-20,
            rng.uniform(0, 1, n).astype(np.float32), (rng.standard_normal(n) * 1e30).astype(np.float32)]
    p = tmp_path / "cat.txt"
    with open(p, "w") as f:
        for row in zip(*cols):
            f.write(" ".join(np.format_float_scientific(v, unique=True) for v in row) + "\n")
    assert IO.scan_text_catalog(p, n_threads=threads) == (n, 4 if n else 0)
    got = _np(IO.read_text_catalog(p, [0, 1, 2, 3], n_threads=threads))
    for g, c in zip(got, cols):
        assert g.dtype == np.float32 and np.array_equal(g.view(np.uint32), c.view(np.uint32))
    # a subset in another order, with a repeated column
    got = _np(IO.read_text_catalog(p, [3, 0, 3], n_threads=threads))
    for g, c in zip(got, (cols[3], cols[0], cols[3])):
        assert np.array_equal(g.view(np.uint32), c.view(np.uint32))


def test_text_matches_loadtxt_on_an_untidy_file(B, IO, tmp_path):
    """Runs of blanks and tabs, leading / trailing blanks, blank lines, comments, CRLF, signs, exponents, integers,
    a last line without a newline: what CSV.File(delim = ' ', ignorerepeated = true) and numpy.loadtxt both accept."""
    body = ("# x y d z\r\n"
            "  1.5   -2.25\t3e2  +4\r\n"
            "\r\n"
            "10 20 30 40   \r\n"
            "   # another comment\r\n"
            "\t-0.0 1E-3 .5 5.\r\n"
            "7 8 9 1e-50\r\n"
            "inf -inf nan 1e39")
    p = tmp_path / "untidy.dat"
    p.write_bytes(body.encode())
    with np.errstate(over="ignore"):
        want = np.loadtxt(p, dtype=np.float64, comments="#").astype(np.float32)
    assert IO.scan_text_catalog(p) == (want.shape[0], 4)
    got = np.stack(_np(IO.read_text_catalog(p, [0, 1, 2, 3])), axis=1)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.signbit(got[2, 0])            # -0.0 keeps its sign
    assert got[3, 3] == 0 and np.isinf(got[4, 3])   # under- and overflow of Float32 like a cast from Float64


def test_text_other_delimiter_and_errors(B, IO, tmp_path):
    p = tmp_path / "c.csv"
    p.write_text("1, 2 ,3\n4,5,6\n")
    assert IO.scan_text_catalog(p, delim=",") == (2, 3)
    a, c = _np(IO.read_text_catalog(p, [0, 2], delim=","))
    assert a.tolist() == [1, 4] and c.tolist() == [3, 6]
    # a header line of names, a short row, an empty field: BAOREC_ERR_IO naming the line
    q = tmp_path / "bad.txt"
    q.write_text("1 2 3\nx y z\n4 5 6\n")
    with pytest.raises(B.lib_loader.CatalogIOError, match=r"line 2: field 1 .*x y z"):
        IO.read_text_catalog(q, [0, 1, 2])
    q.write_text("1 2 3\n4 5\n")
    with pytest.raises(B.lib_loader.CatalogIOError, match=r"line 2: field 3"):
        IO.read_text_catalog(q, [0, 2])
    assert len(IO.read_text_catalog(q, [0, 1])[0]) == 2      # the short row is fine when only its fields are asked for
    q.write_text("1,,3\n")
    with pytest.raises(B.lib_loader.CatalogIOError, match=r"line 1: field 2"):
        IO.read_text_catalog(q, [1], delim=",")
    with pytest.raises(OSError, match="cannot open"):
        IO.read_text_catalog(tmp_path / "missing.txt", [0])
    with pytest.raises(B.BaorecError):
        IO.scan_text_catalog(tmp_path)      # a directory


def test_text_capacity_is_checked(B, IO, tmp_path):
    import ctypes as C
    p = tmp_path / "c.txt"
    p.write_text("1 2\n3 4\n5 6\n")
    lib = B.lib_loader.load()
    buf = torch.empty(2, dtype=torch.float32)
    got = C.c_int64(0)
    rc = lib.baorec_text_catalog_read_f32(os.fsencode(p), b" ", 1, (C.c_int * 1)(0), (C.c_void_p * 1)(buf.data_ptr()), 2,
                                          C.byref(got), 1)
    assert rc == B.lib_loader.ERR_INVALID and b"3 rows" in lib.baorec_last_error()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", ["C", "F"])
def test_npy_read_matches_numpy(B, IO, tmp_path, dtype, order):
    rng = np.random.default_rng(3)
    a = np.asarray(rng.standard_normal((50_001, 5)).astype(dtype), order=order)
    p = tmp_path / "a.npy"
    np.save(p, a)
    info = IO.npy_info(p)
    assert info == {"dtype": "f4" if dtype == np.float32 else "f8", "fortran_order": order == "F", "rows": 50_001, "columns": 5}
    got = _np(IO.read_npy_catalog(p, n_threads=3))
    assert len(got) == 5
    for j, g in enumerate(got):
        assert np.array_equal(g, a[:, j].astype(np.float32))
    x, = _np(IO.read_npy_catalog(p, [4]))
    assert np.array_equal(x, a[:, 4].astype(np.float32))
    with pytest.raises(B.BaorecError, match="5 columns"):
        IO.read_npy_catalog(p, [5])


def test_npy_vector_and_rejections(B, IO, tmp_path):
    v = np.arange(11, dtype=np.float32)
    np.save(tmp_path / "v.npy", v)
    assert IO.npy_info(tmp_path / "v.npy")["columns"] == 1
    assert np.array_equal(IO.read_npy_catalog(tmp_path / "v.npy")[0].numpy(), v)
    np.save(tmp_path / "i.npy", np.arange(4))
    with pytest.raises(OSError, match="dtype"):
        IO.npy_info(tmp_path / "i.npy")
    np.save(tmp_path / "t.npy", np.zeros((2, 2, 2), np.float32))
    with pytest.raises(OSError, match="two dimensions"):
        IO.npy_info(tmp_path / "t.npy")
    (tmp_path / "x.npy").write_bytes(b"not an npy file at all")
    with pytest.raises(OSError, match="not an NPY"):
        IO.npy_info(tmp_path / "x.npy")
    full = (tmp_path / "v.npy").read_bytes()
    (tmp_path / "cut.npy").write_bytes(full[:-8])
    with pytest.raises(OSError, match="truncated"):
        IO.npy_info(tmp_path / "cut.npy")
    # a header whose shape overflows 64 bits when multiplied out must not pass for "fits"
    hd = b"{'descr': '<f4', 'fortran_order': True, 'shape': (4611686018427387904, 4), }"
    hd += b" " * (118 - len(hd) - 1) + b"\n"
    (tmp_path / "huge.npy").write_bytes(b"\x93NUMPY\x01\x00" + len(hd).to_bytes(2, "little") + hd + b"\0" * 64)
    with pytest.raises(OSError, match="truncated"):
        IO.npy_info(tmp_path / "huge.npy")


@pytest.mark.parametrize("k,n", [(3, 0), (3, 1), (3, 12_345), (1, 17), (4, 9)])
def test_npy_write_is_what_npzwrite_of_hcat_stores(B, IO, tmp_path, k, n):
    """NPZ.jl stores a Julia Matrix{Float32}(n, k) as it lies in memory with 'fortran_order': True; numpy.load gives the
    (n, k) matrix back.  The file also equals numpy's own save of the Fortran-ordered matrix byte for byte."""
    rng = np.random.default_rng(k * 100 + n)
    cols = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)) for _ in range(k)]
    p = tmp_path / "out.npy"
    IO.write_npy(p, *cols)
    back = np.load(p)
    want = np.stack([c.numpy() for c in cols], axis=1) if k > 1 else cols[0].numpy()
    assert back.dtype == np.float32 and back.shape == want.shape and np.array_equal(back, want)
    if k > 1:
        assert back.flags.f_contiguous
    if n > 1:       # (numpy calls a matrix of 0 or 1 rows C-ordered, NPZ.jl always says Fortran: same bytes after the header)
        np.save(tmp_path / "np.npy", np.asfortranarray(want) if k > 1 else want)
        assert p.read_bytes() == (tmp_path / "np.npy").read_bytes()
    # and the library reads its own files
    again = _np(IO.read_npy_catalog(p))
    for g, c in zip(again, cols):
        assert np.array_equal(g, c.numpy())


@pytest.mark.parametrize("n,threads", [(0, 1), (1, 1), (1000, 1), (3_000_017, 4)])
def test_select_rows_is_a_stable_mask(B, IO, n, threads):
    rng = np.random.default_rng(n)
    z = rng.uniform(0.5, 1.3, n).astype(np.float32)
    if n > 10:
        z[3] = 0.8          # the bounds are exclusive: (z > 0.8) & (z < 1)
        z[5] = 1.0
        z[7] = np.nan
    others = [rng.standard_normal(n).astype(np.float32) for _ in range(4)]
    cols = [torch.from_numpy(c.copy()) for c in (others[0], others[1], z, others[2], others[3])]
    mask = (z > np.float32(0.8)) & (z < np.float32(1.0))
    kept = IO.select_rows(cols, 2, 0.8, 1.0, n_threads=threads)
    assert all(len(c) == int(mask.sum()) for c in kept)
    for got, src in zip(kept, (others[0], others[1], z, others[2], others[3])):
        assert np.array_equal(got.numpy(), src[mask])


def test_lightcone_example_front_end(B, IO, tmp_path):
    """examples/lightcone.jl:22-26 end to end on a synthetic file: five columns ra dec d z nz, selection 0.8 < z < 1."""
    rng = np.random.default_rng(11)
    n = 20_000
    cat = np.stack([rng.uniform(0, 360, n), rng.uniform(-30, 30, n), rng.uniform(1e3, 3e3, n), rng.uniform(0.6, 1.2, n),
                    rng.uniform(1e-5, 1e-3, n)], axis=1).astype(np.float32)
    p = tmp_path / "lc.dat"
    np.savetxt(p, cat, fmt="%.9g", delimiter=" ")
    ra, dec, z, nz = IO.read_text_catalog(p, [0, 1, 3, 4])
    ra, dec, z, nz = IO.select_rows([ra, dec, z, nz], 2, 0.8, 1.0)
    m = (cat[:, 3] > np.float32(0.8)) & (cat[:, 3] < 1)
    assert np.array_equal(ra.numpy(), cat[m, 0]) and np.array_equal(nz.numpy(), cat[m, 4])
    assert np.array_equal(z.numpy(), cat[m, 3]) and np.array_equal(dec.numpy(), cat[m, 1])


def test_python_mirror_validates_its_arguments(B, IO, tmp_path):
    a = torch.arange(6, dtype=torch.float32)
    with pytest.raises(ValueError, match="different lengths"):
        IO.write_npy(tmp_path / "x.npy", a, a[:3].clone())
    with pytest.raises(ValueError, match="at least one"):
        IO.write_npy(tmp_path / "x.npy")
    with pytest.raises(TypeError, match="contiguous 1-D float32"):
        IO.write_npy(tmp_path / "x.npy", a[::2])                     # a strided view
    with pytest.raises(TypeError, match="contiguous 1-D float32"):
        IO.write_npy(tmp_path / "x.npy", a.double())
    with pytest.raises(ValueError, match="different lengths"):
        IO.select_rows([a, a[:2].clone()], 0, 0.0, 1.0)
    with pytest.raises(B.BaorecError):
        IO.select_rows([a], 3, 0.0, 1.0)                              # key column out of range
    with pytest.raises(ValueError, match="one character"):
        IO.scan_text_catalog(tmp_path / "x.txt", delim="::")
    assert not (tmp_path / "x.npy").exists()
