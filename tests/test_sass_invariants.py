"""CPU: properties of the built kernels that DESIGN.md relies on, read from the SASS (cuobjdump) and the ptxas logs of
the in-tree build -- no GPU needed.  Deposits are fire-and-forget reductions (REDG), not returning atomics; the staged
kernels really stage through cp.async (LDGSTS); nothing on the bench's default path spills to local memory; the
option variants exist with the instruction counts the design quotes."""
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "benchmarks"))
import sass_census as SC  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit's cuobjdump")


@pytest.fixture(scope="module")
def K(B):
    """demangled kernel name -> census row, for the objects the bench's default path and the options use"""
    out = {}
    for name in ("mas", "kspace", "multigrid", "pk", "catalog", "dist"):
        obj = SC.BUILD / f"{name}.o"
        info = SC.ptxas_info((SC.BUILD / f"{name}.ptxas.log").read_text())
        for mangled, cnt in SC.census(obj).items():
            reg, smem, spill, _ = info.get(mangled, [None, 0, 0, 0])
            row = {c: cnt[c] for c, _ in SC.CLASSES}
            row.update(total=cnt["total"], cas_loop=cnt["cas_loop"], registers=reg, smem=smem, spill=spill)
            out[SC.demangle(mangled)] = row
    return out


def find(K, *parts):
    hits = [v for k, v in K.items() if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, [k for k in K if all(p in k for p in parts)])
    return hits[0]


def test_deposits_are_reductions_not_returning_atomics(K):
    for parts, n_red in ((("scatter_records_kernel",), 8), (("scatter_sorted_kernel<0>",), 8), (("scatter_sorted_kernel<1>",), 27)):
        k = find(K, *parts)
        assert k["red_global"] >= n_red and k["atom_global"] == 0 and k["cas_loop"] == 0


def test_vector_reduction_variants(K):
    assert find(K, "scatter_records_pairs_kernel<1>")["red_global"] == 4 + 8 + 1       # 4 pairs | 8 scalars, + the out-of-box counter
    assert find(K, "scatter_records_pairs_kernel<2>")["red_global"] >= 4 + 4
    assert find(K, "scatter_sorted_tsc_vec_kernel")["red_global"] >= 18


def test_staged_kernels_use_cp_async(K):
    assert find(K, "gather_tile_kernel<3, 0>")["ldgsts"] == 108 and find(K, "gather_tile_kernel<3, 1>")["ldgsts"] == 36
    assert find(K, "gather_tile_kernel<3, 1>")["total"] < 0.55 * find(K, "gather_tile_kernel<3, 0>")["total"]
    assert find(K, "mg_stencil_smem<0, true, 6, false>")["ldgsts"] > 0
    assert find(K, "gather_tile_kernel<3, 0>")["smem"] == 28512


def test_no_register_spills_and_known_local_memory_users(K):
    """No kernel of these objects spills registers except the scalar tail instantiation of cartesian_to_sky (<= 3
    particles per call, 12 bytes).  Local memory is otherwise used only where a kernel indexes a table that arrives by
    value in its parameters (the compiler copies it to the stack): listed here so that a new one does not go unnoticed (the slab transpose's peer table was one: 16 local stores per
    thread until it became __grid_constant__)."""
    spills = {k for k, v in K.items() if v["spill"]}
    # (peer_push_kernel: ptxas keeps one 4-byte loop-invariant on the stack at 40 registers -- outside the copy loop)
    assert all("cartesian_to_sky_kernel<1>" in k or "peer_push_kernel" in k for k in spills), spills
    local = {k.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("baorec::", "") for k, v in K.items() if v["local"]}
    assert local == {"cartesian_to_sky_kernel<1>", "sky_to_cartesian_kernel<1>", "sky_to_cartesian_kernel<4>",   # sincos quadrant table
                     "gather_trash_kernel",                      # GatherArgs.o[c] with a run-time c
                     "mg_coarse_kernel",                         # level table of the single-block coarse V-cycle
                     "peer_push_kernel"}, local                  # the 4-byte spill above


def test_multipole_kernel(K):
    for name in ("pk_kernel<false>", "pk_kernel<true>"):          # plain / interlaced (second half mesh + the phase product)
        assert find(K, name)["spill"] == 0
    k = find(K, "pk_kernel<false>")
    assert k["fp64"] > 200 and k["shfl_vote"] > 100 and k["red_global"] == 1 and k["atom_shared"] > 0
