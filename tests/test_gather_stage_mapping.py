"""CPU: the two cp.async staging schedules of gather_tile_kernel (csrc/mas.cu, fast path) restated in Python --
STAGE 0 (rows dealt round-robin over the unrolled loop; the kernel measured in round 1) and STAGE 1 (option
"gather_stage": each warp owns rows py = warp, warp + 4 (, 8) of every (field, plane)) -- must issue exactly the same
set of copies (destination float in the shared-memory window, source element of the field, size).  A restatement, not
the compiled code: it checks the index algebra of the variant (plane bases, the single wrapping row, the edge column,
the shared-memory offsets); the compiled kernels are held to bit-identical shifts on the GPU
(tests/test_gpu_zzz6_gather_stage.py)."""
import pytest

TILE_X, TILE_Y, TILE_WXP, NF = 128, 8, 132, 3


def stage0(x0, y0, iz, nx, ny, nz, slab):
    xe = x0 + TILE_X
    xe = xe - nx if xe >= nx else xe
    zrow = [iz, iz + 1 if slab else (0 if iz + 1 >= nz else iz + 1)]
    out = set()
    for warp in range(4):
        for r in range(NF * 2 * (TILE_Y + 1)):
            if (r & 3) != warp:
                continue
            ry = TILE_Y + 1
            py, pz, f = r % ry, (r // ry) & 1, r // (2 * ry)
            yy = y0 + py
            yy = yy - ny if yy >= ny else yy
            src = (zrow[pz] * ny + yy) * nx
            dst = ((f * 2 + pz) * ry + py) * TILE_WXP
            out.update((f, dst + 4 * lane, src + x0 + 4 * lane, 16) for lane in range(32))
            out.add((f, dst + TILE_X, src + xe, 4))
    return out


def stage1(x0, y0, iz, nx, ny, nz, slab):
    xe = x0 + TILE_X
    xe = xe - nx if xe >= nx else xe
    z1 = iz + 1 if slab else (0 if iz + 1 >= nz else iz + 1)
    y8 = y0 + TILE_Y
    y8 = y8 - ny if y8 >= ny else y8
    out = set()
    for warp in range(4):
        for p in range(NF * 2):
            f = p >> 1
            base = (z1 if (p & 1) else iz) * ny * nx + x0
            for k in range(3):
                py = warp + 4 * k
                if py <= TILE_Y:
                    src = base + (y8 if py == TILE_Y else y0 + py) * nx
                    row = (p * (TILE_Y + 1) + py) * TILE_WXP
                    out.update((f, row + 4 * lane, src + 4 * lane, 16) for lane in range(32))
                    out.add((f, row + TILE_X, src - x0 + xe, 4))
    return out


@pytest.mark.parametrize("nx,ny,nz", [(128, 8, 8), (256, 64, 16), (1024, 1024, 4), (384, 24, 8)])
@pytest.mark.parametrize("slab", [0, 1])
def test_both_schedules_issue_the_same_copies(nx, ny, nz, slab):
    for x0 in range(0, nx, TILE_X):
        for y0 in list(range(0, ny, TILE_Y))[:: max(1, ny // TILE_Y // 6)] + [ny - TILE_Y]:
            for iz in (0, nz // 2, nz - 1):
                a, b = stage0(x0, y0, iz, nx, ny, nz, slab), stage1(x0, y0, iz, nx, ny, nz, slab)
                assert a == b and len(a) == NF * 2 * (TILE_Y + 1) * 33


def test_shift_decode_of_the_tile_index():
    for nxc, nyc, nzt in ((8, 128, 16), (1, 1, 8), (2, 64, 4)):
        sx, sy = nxc.bit_length() - 1, nyc.bit_length() - 1
        for tile in range(0, nxc * nyc * nzt, 7):
            tx, r = tile % nxc, tile // nxc
            assert (tile & (nxc - 1), (tile >> sx) & (nyc - 1), tile >> (sx + sy)) == (tx, r % nyc, r // nyc)
