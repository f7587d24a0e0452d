"""CPU checks of the C-ABI boundary: the library builds, loads, exports every symbol that
include/baorec_b200.h declares, the header is valid C, the ctypes mirror of baorec_params has the
compiled layout, and calls fail cleanly (error code + message) when no GPU is present."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "baorec_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(baorec_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("baorec_create", "baorec_plan", "baorec_cic_scatter_f32", "baorec_gather_f32", "baorec_smooth_f32",
                 "baorec_setup_overdensity_f32", "baorec_iterate_f32", "baorec_reconstructed_overdensity_f32",
                 "baorec_mg_jacobi_f32", "baorec_mg_residual_f32", "baorec_mg_restrict_f32", "baorec_mg_prolong_f32",
                 "baorec_mg_vcycle_f32", "baorec_mg_fmg_f32", "baorec_reconstructed_potential_f32",
                 "baorec_compute_displacements_f32", "baorec_read_shifts_f32", "baorec_reconstructed_positions_f32",
                 "baorec_run_host_f32", "baorec_read_host_f32", "baorec_comm_init", "baorec_plan_dist"):
        assert must in syms
    assert len(syms) >= 40


def test_library_exports_every_declared_symbol(B):
    lib = B.lib_loader.load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_signatures_cover_the_header(B):
    assert sorted(B.lib_loader.SIGNATURES) == declared_symbols()


def test_header_is_valid_c_and_params_layout_matches(B, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "baorec_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(baorec_params), offsetof(baorec_params, los),'
                   ' offsetof(baorec_params, jacobi_damping_factor), offsetof(baorec_params, mas),'
                   ' offsetof(baorec_params, box_pad)); return 0;}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)],
                   check=True)
    size, o_los, o_jd, o_mas, o_pad = map(int, subprocess.check_output([str(exe)]).split())
    P = B.lib_loader.Params
    assert C.sizeof(P) == size
    assert (P.los.offset, P.jacobi_damping_factor.offset, P.mas.offset, P.box_pad.offset) == (o_los, o_jd, o_mas, o_pad)


def test_version_and_error_reporting_without_gpu(B):
    import torch
    lib = B.lib_loader.load()
    assert lib.baorec_version() == 100
    # NULL / bad arguments never crash: they return BAOREC_ERR_INVALID with a message
    assert lib.baorec_create(0, None) == B.lib_loader.ERR_INVALID
    assert b"invalid argument" in lib.baorec_last_error()
    assert lib.baorec_plan(None, 8, 8, 8, B.lib_loader.f3((1, 1, 1)), B.lib_loader.f3((0, 0, 0))) == B.lib_loader.ERR_INVALID
    assert lib.baorec_destroy(None) == 0
    if not torch.cuda.is_available():
        h = C.c_void_p()
        rc = lib.baorec_create(0, C.byref(h))
        assert rc == B.lib_loader.ERR_CUDA          # fails loudly: there is no CPU fallback
        assert b"CUDA" in lib.baorec_last_error()
        with pytest.raises(B.BaorecError):
            B.Context(0)


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: no product source may import, link or execute it."""
    pkg = ROOT / "baorec.jl_b200"
    files = [f for f in pkg.glob("*.py") if f.name != "build.py"] + list((pkg / "csrc").glob("*.cu*"))
    assert len(files) >= 8
    for f in files:
        text = f.read_text()
        assert "baorec_oracle" not in text and "oracle/" not in text and "import oracle" not in text, f


def test_host_mirror_tables_match_oracle(B, O):
    import numpy as np
    for n, L, lo in ((64, 1000.0, 0.0), (96, 1373.5, -412.25), (256, 2500.0, 0.0)):
        bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
        for a, b in zip(B.k_vec((n, n, n), bs), O.k_vec((n, n, n), bs, np.float32)):
            assert np.array_equal(a, b)
        for a, b in zip(B.x_vec((n, n, n), bs, bm), O.x_vec((n, n, n), bs, bm, np.float32)):
            assert np.array_equal(a, b)


def test_recon_structs_mirror_reference_defaults(B):
    r = B.IterativeRecon(bias=2.2, f=0.757, smoothing_radius=15.0)
    assert r.n_iter == 3 and r.los is None and r.box_size is None and r.fft_plan is None and r.result_cache is None
    assert abs(r.beta - 0.757 / 2.2) < 1e-7
    m = B.MultigridRecon(bias=2.2, f=0.757, smoothing_radius=15.0)
    assert (m.jacobi_damping_factor, m.jacobi_niterations, m.vcycle_niterations) == (0.4, 5, 6)
    p = m._params()
    assert p.has_los == 0 and abs(p.ran_min - 0.01) < 1e-9 and p.box_pad == 500.0 and p.mas == 0
