"""CPU: the reader / writer threads behind baorec_batch_files_f32 (baorec.jl_b200/csrc/batch.cuh: FileBatchSource),
compiled as plain C++ with malloc in place of pinned memory and driven in the call order of the device pipeline by the
stand-in of tests/hostcheck/batch_hostcheck.cpp (result = pos + w * (1, 2, 3)).  Checked: every catalog reaches its own
output file whatever the number of buffer sets and parser threads, catalogs of different sizes and formats mix, weights
default to one, the rows are reported, and failures (a missing file in the middle, a bad line, a failing device step, an
unwritable output) end the call with the right code and message instead of a hang."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "baorec.jl_b200" / "csrc"
ERR_IO, ERR_OUT_OF_BOX = -9, -5


@pytest.fixture(scope="module")
def HC():
    out = ROOT / "tests" / "_build" / "libbatch_hostcheck.so"
    srcs = [ROOT / "tests" / "hostcheck" / "batch_hostcheck.cpp", CSRC / "catalog_io.cu"]
    deps = srcs + [CSRC / "batch.cuh", CSRC / "catalog_io.cuh", CSRC / "internal.cuh", ROOT / "include" / "baorec_b200.h"]
    # -Bsymbolic: the check's own set_error / baorec_last_error, not those of a libbaorec_b200.so loaded RTLD_GLOBAL earlier
    if not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        out.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        cmd = [gxx, "-std=c++17", "-O2", "-shared", "-fPIC", "-pthread", "-Wl,-Bsymbolic", "-x", "c++", "-I", "/usr/local/cuda/include",
               "-I", str(ROOT / "include"), "-I", str(CSRC), "-o", str(out)] + [str(s) for s in srcs]
        subprocess.run(cmd, check=True)
    lib = C.CDLL(str(out))
    lib.hc_batch_files.restype = C.c_int
    lib.hc_batch_files.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char, C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.c_int,
                                   C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    lib.baorec_last_error.restype = C.c_char_p
    return lib


def run(HC, ins, outs, cols=(0, 1, 2, 3), delim=" ", slots=4, threads=2, fail_at=-1):
    n = len(ins)
    a_in = (C.c_char_p * n)(*[os.fsencode(p) for p in ins])
    a_out = None if outs is None else (C.c_char_p * n)(*[None if p is None else os.fsencode(p) for p in outs])
    rows = (C.c_int64 * n)()
    secs = (C.c_double * 4)()
    rc = HC.hc_batch_files(n, a_in, delim.encode(), (C.c_int * 4)(*cols), a_out, slots, threads, fail_at, rows, secs)
    return rc, list(rows), HC.baorec_last_error().decode()


def make_catalogs(tmp_path, sizes, seed=0):
    rng = np.random.default_rng(seed)
    cats, ins = [], []
    for i, n in enumerate(sizes):
        a = np.concatenate([rng.uniform(0, 1000, (n, 3)), rng.uniform(0.5, 1.5, (n, 1))], axis=1).astype(np.float32)
        cats.append(a)
        if i % 3 == 2:       # every third catalog is an NPY matrix
            p = tmp_path / f"mock_{i}.npy"
            np.save(p, a if i % 2 else np.asfortranarray(a))
        else:
            p = tmp_path / f"mock_{i}.dat"
            np.savetxt(p, a, fmt="%.9g")
        ins.append(p)
    return cats, ins


@pytest.mark.parametrize("slots,threads", [(3, 1), (4, 3), (9, 2)])
def test_every_catalog_reaches_its_own_file(HC, tmp_path, slots, threads):
    sizes = [1000, 1, 2500, 40_000, 7, 1200, 39_999, 3]          # grows, shrinks, grows again: slots are re-allocated
    cats, ins = make_catalogs(tmp_path, sizes)
    outs = [tmp_path / f"rec_{i}.npy" for i in range(len(sizes))]
    outs[4] = None                                                 # nothing written for this one
    rc, rows, msg = run(HC, ins, outs, slots=slots, threads=threads)
    assert rc == 0, msg
    assert rows == sizes
    for i, (a, o) in enumerate(zip(cats, outs)):
        if o is None:
            continue
        got = np.load(o)
        want = a[:, :3] + a[:, 3:4] * np.array([1, 2, 3], np.float32)
        assert got.dtype == np.float32 and got.shape == (sizes[i], 3) and np.array_equal(got, want), i
    assert not (tmp_path / "rec_4.npy").exists()


def test_weights_default_to_one_and_columns_are_picked(HC, tmp_path):
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 100, (500, 4)).astype(np.float32)          # x y d z, like the UNIT boxes (examples/simulation.jl:13)
    p = tmp_path / "box.txt"
    np.savetxt(p, a, fmt="%.9g")
    rc, rows, msg = run(HC, [p], [tmp_path / "o.npy"], cols=(0, 1, 3, -1))
    assert rc == 0, msg
    want = a[:, [0, 1, 3]] + np.array([1, 2, 3], np.float32)
    assert np.array_equal(np.load(tmp_path / "o.npy"), want)
    rc, _, msg = run(HC, [p], None)                                # no output files at all
    assert rc == 0, msg


def test_failures_end_the_call(HC, tmp_path):
    cats, ins = make_catalogs(tmp_path, [300] * 6)
    outs = [tmp_path / f"rec_{i}.npy" for i in range(6)]
    # a missing file in the middle of the batch
    broken = list(ins)
    broken[3] = tmp_path / "nowhere.dat"
    rc, _, msg = run(HC, broken, outs, slots=3)
    assert rc == ERR_IO and "nowhere.dat" in msg
    # a bad line, named with its number
    bad = tmp_path / "bad.dat"
    bad.write_text("1 2 3 4\n5 6 seven 8\n")
    rc, _, msg = run(HC, [ins[0], bad, ins[1]], None)
    assert rc == ERR_IO and "line 2" in msg and "field 3" in msg
    # an empty catalog
    empty = tmp_path / "empty.dat"
    empty.write_text("# nothing\n")
    rc, _, msg = run(HC, [ins[0], empty], None)
    assert rc == ERR_IO and "no rows" in msg
    # the device step fails: its code and message survive the shutdown of the I/O threads
    rc, _, msg = run(HC, ins, outs, fail_at=2)
    assert rc == ERR_OUT_OF_BOX and "catalog 2" in msg
    # an output that cannot be written
    outs2 = list(outs)
    outs2[1] = tmp_path / "no_such_dir" / "x.npy"
    rc, _, msg = run(HC, ins, outs2)
    assert rc == ERR_IO and "no_such_dir" in msg


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_io_threads_under_sanitizers(tmp_path, sanitizer):
    """The same class in a binary built with ThreadSanitizer / AddressSanitizer + UBSan: 18 good batches, 18 with a
    failing device step, 18 with a bad line and 18 with a missing file in the middle -- no data race, no leak, no
    use after free reported (any report makes the binary exit non-zero)."""
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    exe = tmp_path / "batch_sanitize"
    cmd = [gxx, "-std=c++17", "-O1", "-g", f"-fsanitize={sanitizer}", "-pthread", "-x", "c++", "-I", "/usr/local/cuda/include",
           "-I", str(ROOT / "include"), "-I", str(CSRC), str(ROOT / "tests" / "hostcheck" / "batch_sanitize_main.cpp"),
           str(ROOT / "tests" / "hostcheck" / "batch_hostcheck.cpp"), str(CSRC / "catalog_io.cu"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("cannot find" in r.stderr or "unrecognized" in r.stderr):
        pytest.skip(f"-fsanitize={sanitizer} is not available here")
    assert r.returncode == 0, r.stderr[-2000:]
    rng = np.random.default_rng(0)
    for i, n in enumerate([5000, 1, 20000, 300, 40000, 7, 12000, 100]):
        a = rng.uniform(0, 1000, (n, 4)).astype(np.float32)
        if i % 3 == 2:
            np.save(tmp_path / f"m{i}.npy", np.asfortranarray(a) if i % 2 else a)
        else:
            np.savetxt(tmp_path / f"m{i}.dat", a, fmt="%.9g")
    (tmp_path / "bad.dat").write_text("1 2 3 4\n5 six 7 8\n")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66", ASAN_OPTIONS="detect_leaks=1 exitcode=66",
               UBSAN_OPTIONS="halt_on_error=1 exitcode=66")
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    if "unexpected memory mapping" in r.stderr or "Shadow memory range interleaves" in r.stderr:
        pytest.skip("the sanitizer runtime cannot map its shadow memory on this kernel / ASLR setting")
    assert r.returncode == 0 and "done bad=0" in r.stdout, (r.stdout + r.stderr)[-3000:]
