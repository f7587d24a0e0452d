// TEST INFRASTRUCTURE: compiles the product's cell / weight arithmetic of the mass-assignment kernels
// (baorec.jl_b200/csrc/mas_math.cuh: wrap_pos, cic_axis, gather_axis, tsc_axis -- the device functions every
// scatter / gather kernel of mas.cu calls) as plain C++ and runs it in loops on the CPU, so that the build
// container -- which has no GPU -- can hold the very statements the device executes to the oracle's
// bit-exact cell indices and weights (tests/test_mas_hostcheck.py).  The loops mirror cic_cells_kernel /
// gather_cells_kernel of mas.cu.  Built with g++ -O2 -ffp-contract=off (no FMA contraction == the *_rn
// intrinsics) into tests/_build/.  Never linked into, or loaded by, the product library.
#include <stdint.h>

#include "../../baorec.jl_b200/csrc/mas_math.cuh"

using namespace baorec;

extern "C" {

// i0/i1/w0/w1 are [3][n] like baorec_cic_cells_f32; out-of-box particles get index -1
void hc_cic_cells(const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L, const float* mn,
                  int wrap, int32_t* i0, int32_t* i1, float* w0, float* w1, float* wrapped) {
  const float* p[3] = {x, y, z};
  for (int64_t i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      float v = p[a][i];
      if (wrap) v = wrap_pos(v, mn[0], L[0]);
      if (wrapped) wrapped[a * n + i] = v;
      int c0 = -1, c1 = -1;
      float a0 = 0.f, a1 = 0.f;
      if (!cic_axis(v, mn[a], L[a], ng[a], wrap != 0, c0, c1, a0, a1)) c0 = c1 = -1;
      i0[a * n + i] = c0;
      i1[a * n + i] = c1;
      w0[a * n + i] = a0;
      w1[a * n + i] = a1;
    }
}

void hc_gather_cells(const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L, const float* mn,
                     const float* cell, int gpu_formula, int32_t* id, int32_t* iu, float* wd, float* wu) {
  const float* p[3] = {x, y, z};
  for (int64_t i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      int c0 = -1, c1 = -1;
      float a0 = 0.f, a1 = 0.f;
      if (!gather_axis(p[a][i], mn[a], L[a], cell[a], ng[a], gpu_formula != 0, c0, c1, a0, a1)) c0 = c1 = -1;
      id[a * n + i] = c0;
      iu[a * n + i] = c1;
      wd[a * n + i] = a0;
      wu[a * n + i] = a1;
    }
}

// idx / w are [3 axes][3 stencil points][n]
void hc_tsc_cells(const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L, const float* mn, int wrap,
                  int32_t* idx, float* w, int32_t* ok) {
  const float* p[3] = {x, y, z};
  for (int64_t i = 0; i < n; i++) {
    int good = 1;
    for (int a = 0; a < 3; a++) {
      int id3[3] = {-1, -1, -1};
      float w3[3] = {0.f, 0.f, 0.f};
      if (!tsc_axis(p[a][i], mn[a], L[a], ng[a], wrap != 0, id3, w3)) good = 0;
      for (int o = 0; o < 3; o++) {
        idx[(a * 3 + o) * n + i] = id3[o];
        w[(a * 3 + o) * n + i] = w3[o];
      }
    }
    ok[i] = good;
  }
}

static BoxGeom make_geom(const int* ng, const float* L, const float* mn, int slab, int z_lo, int zoff, int nzp) {
  BoxGeom g;
  for (int a = 0; a < 3; a++) {
    g.mn[a] = mn[a];
    g.L[a] = L[a];
    g.cell[a] = L[a] / (float)ng[a];
    g.n[a] = ng[a];
  }
  g.slab = slab;
  g.z_lo = z_lo;
  g.zoff = zoff;
  g.nzp = nzp;
  return g;
}

// The product's deposit<MAS> (mas_math.cuh) for a list of particles into a local buffer of nzp planes, exactly as
// scatter_direct_kernel calls it (cic! wraps the position first).  slab = 0: whole mesh (nzp = nz); slab = 1 with
// (zoff, nzp) = (0, nz_loc + 1): CIC slab scatter; (1, nz_loc + 3): TSC slab scatter.  Returns the rejected count.
int64_t hc_deposit(int mas, float* buf, const float* x, const float* y, const float* z, const float* w, int64_t n, const int* ng,
                   const float* L, const float* mn, int wrap, int slab, int z_lo, int zoff, int nzp) {
  const BoxGeom g = make_geom(ng, L, mn, slab, z_lo, zoff, nzp);
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    float px = x[i], py = y[i], pz = z[i];
    bool ok;
    if (mas == BAOREC_MAS_CIC) {
      if (wrap) {
        px = wrap_pos(px, g.mn[0], g.L[0]);
        py = wrap_pos(py, g.mn[0], g.L[0]);
        pz = wrap_pos(pz, g.mn[0], g.L[0]);
      }
      ok = deposit<BAOREC_MAS_CIC>(buf, px, py, pz, w[i], g, wrap != 0);
    } else if (mas == BAOREC_MAS_PCS) {
      ok = deposit<BAOREC_MAS_PCS>(buf, px, py, pz, w[i], g, wrap != 0);
    } else {
      ok = deposit<BAOREC_MAS_TSC>(buf, px, py, pz, w[i], g, wrap != 0);
    }
    if (!ok) bad++;
  }
  return bad;
}

// Option "scatter_pairs": the product's deposit_pairs (aligned x pairs through one vector reduction) for a list of
// particles in the given order; on the host the pair is two plain additions, so the result must equal hc_deposit's
// (every cell receives the same values in the same particle order).  Returns (rejected count, pairs used) packed.
int64_t hc_deposit_pairs(int mode, float* buf, const float* x, const float* y, const float* z, const float* w, int64_t n, const int* ng,
                         const float* L, const float* mn, int wrap, int64_t* n_paired) {
  const BoxGeom g = make_geom(ng, L, mn, 0, 0, 0, ng[2]);
  int64_t bad = 0, paired = 0;
  for (int64_t i = 0; i < n; i++) {
    float px = x[i], py = y[i], pz = z[i];
    if (wrap) {
      px = wrap_pos(px, g.mn[0], g.L[0]);
      py = wrap_pos(py, g.mn[0], g.L[0]);
      pz = wrap_pos(pz, g.mn[0], g.L[0]);
    }
    if (!(mode == 2 ? deposit_pairs<2>(buf, px, py, pz, w[i], g, wrap != 0) : deposit_pairs<1>(buf, px, py, pz, w[i], g, wrap != 0))) bad++;
    else {
      int x0, x1;
      float w0, w1;
      cic_axis(px, g.mn[0], g.L[0], g.n[0], wrap != 0, x0, x1, w0, w1);
      if ((mode == 2 && (g.n[0] & 3) == 0) || (((x0 | g.n[0]) & 1) == 0 && x1 == x0 + 1)) paired++;  // took a vector reduction
    }
  }
  *n_paired = paired;
  return bad;
}

// Option "scatter_pairs" for PCS: the product's deposit_pcs_vec (one or two aligned quads per stencil row) against deposit<PCS>.
int64_t hc_deposit_pcs_vec(float* buf, const float* x, const float* y, const float* z, const float* w, int64_t n, const int* ng,
                           const float* L, const float* mn, int wrap, int slab, int z_lo, int zoff, int nzp) {
  const BoxGeom g = make_geom(ng, L, mn, slab, z_lo, zoff, nzp);
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++)
    if (!deposit_pcs_vec(buf, x[i], y[i], z[i], w[i], g, wrap != 0)) bad++;
  return bad;
}

// Option "scatter_pairs" for TSC: the product's deposit_tsc_vec (one aligned quad or two aligned pairs per stencil row)
// against deposit<TSC> -- on the host the vector reductions are plain additions in cell order, the +0 slots included.
int64_t hc_deposit_tsc_vec(float* buf, const float* x, const float* y, const float* z, const float* w, int64_t n, const int* ng,
                           const float* L, const float* mn, int wrap, int slab, int z_lo, int zoff, int nzp) {
  const BoxGeom g = make_geom(ng, L, mn, slab, z_lo, zoff, nzp);
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++)
    if (!deposit_tsc_vec(buf, x[i], y[i], z[i], w[i], g, wrap != 0)) bad++;
  return bad;
}

// TSC gather from a halo'd slab buffer (slab_mode 2: zoff = 1, nzp = nz_loc + 3): the loop of gather_one<1, TSC>
// in mas.cu, with the product's tsc_axis and local_plane1 doing the index work.  Returns the rejected count.
int64_t hc_tsc_gather(const float* buf, const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L,
                      const float* mn, int slab, int z_lo, int zoff, int nzp, float* out) {
  const BoxGeom g = make_geom(ng, L, mn, slab, z_lo, zoff, nzp);
  const size_t nx = g.n[0], ny = g.n[1];
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    int ix[3], iy[3], iz[3];
    float wx[3], wy[3], wz[3];
    bool ok = tsc_axis(x[i], g.mn[0], g.L[0], g.n[0], true, ix, wx);
    ok = tsc_axis(y[i], g.mn[1], g.L[1], g.n[1], true, iy, wy) && ok;
    ok = tsc_axis(z[i], g.mn[2], g.L[2], g.n[2], true, iz, wz) && ok;
    if (ok && g.slab)
      for (int c = 0; c < 3; c++) ok = local_plane1(g, iz[c], iz[c]) && ok;
    float val = 0.f;
    if (ok) {
      for (int oz = 0; oz < 3; oz++)
        for (int oy = 0; oy < 3; oy++) {
          const size_t row = ((size_t)iz[oz] * ny + iy[oy]) * nx;
          for (int ox = 0; ox < 3; ox++)
            val = __fadd_rn(val, __fmul_rn(__fmul_rn(__fmul_rn(buf[row + ix[ox]], wx[ox]), wy[oy]), wz[oz]));
        }
    } else {
      bad++;
    }
    out[i] = val;
  }
  return bad;
}

// idx / w are [3 axes][4 stencil points][n]
void hc_pcs_cells(const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L, const float* mn, int wrap,
                  int32_t* idx, float* w, int32_t* ok) {
  const float* p[3] = {x, y, z};
  for (int64_t i = 0; i < n; i++) {
    int good = 1;
    for (int a = 0; a < 3; a++) {
      int id4[4] = {-1, -1, -1, -1};
      float w4[4] = {0.f, 0.f, 0.f, 0.f};
      if (!pcs_axis(p[a][i], mn[a], L[a], ng[a], wrap != 0, id4, w4)) good = 0;
      for (int o = 0; o < 4; o++) {
        idx[(a * 4 + o) * n + i] = id4[o];
        w[(a * 4 + o) * n + i] = w4[o];
      }
    }
    ok[i] = good;
  }
}

// PCS gather (whole mesh or a halo'd slab buffer, slab_mode 2: zoff = 1, nzp = nz_loc + 3): the loop of
// gather_one<1, PCS> in mas.cu with the product's stencil_axis and local_plane1.  Returns the rejected count.
int64_t hc_pcs_gather(const float* buf, const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L,
                      const float* mn, int slab, int z_lo, int zoff, int nzp, float* out) {
  const BoxGeom g = make_geom(ng, L, mn, slab, z_lo, zoff, nzp);
  const size_t nx = g.n[0], ny = g.n[1];
  constexpr int SW = MasStencil<BAOREC_MAS_PCS>::SW;
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    int ix[4], iy[4], iz[4];
    float wx[4], wy[4], wz[4];
    bool ok = stencil_axis<BAOREC_MAS_PCS>(x[i], g.mn[0], g.L[0], g.n[0], true, ix, wx);
    ok = stencil_axis<BAOREC_MAS_PCS>(y[i], g.mn[1], g.L[1], g.n[1], true, iy, wy) && ok;
    ok = stencil_axis<BAOREC_MAS_PCS>(z[i], g.mn[2], g.L[2], g.n[2], true, iz, wz) && ok;
    if (ok && g.slab)
      for (int c = 0; c < SW; c++) ok = local_plane1(g, iz[c], iz[c]) && ok;
    float val = 0.f;
    if (ok) {
      for (int oz = 0; oz < SW; oz++)
        for (int oy = 0; oy < SW; oy++) {
          const size_t row = ((size_t)iz[oz] * ny + iy[oy]) * nx;
          for (int ox = 0; ox < SW; ox++)
            val = __fadd_rn(val, __fmul_rn(__fmul_rn(__fmul_rn(buf[row + ix[ox]], wx[ox]), wy[oy]), wz[oz]));
        }
    } else {
      bad++;
    }
    out[i] = val;
  }
  return bad;
}

// The deterministic scatter (option "deterministic_scatter"): the product's deposit_fixed for a list of particles in
// the given order into a 64-bit fixed-point accumulator, then from_fixed per cell (fixed_to_float_kernel: rho += ...).
int64_t hc_deposit_fixed(float* rho, unsigned long long* acc, const float* x, const float* y, const float* z, const float* w,
                         int64_t n, const int* ng, const float* L, const float* mn, int wrap) {
  const BoxGeom g = make_geom(ng, L, mn, 0, 0, 0, ng[2]);
  const size_t M = (size_t)ng[0] * ng[1] * ng[2];
  for (size_t i = 0; i < M; i++) acc[i] = 0;
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    float px = x[i], py = y[i], pz = z[i];
    if (wrap) {
      px = wrap_pos(px, g.mn[0], g.L[0]);
      py = wrap_pos(py, g.mn[0], g.L[0]);
      pz = wrap_pos(pz, g.mn[0], g.L[0]);
    }
    if (!deposit_fixed(acc, px, py, pz, w[i], g, wrap != 0)) bad++;
  }
  for (size_t i = 0; i < M; i++)
    if (acc[i]) rho[i] = __fadd_rn(rho[i], from_fixed(acc[i]));
  return bad;
}

// CIC gather from a halo'd slab buffer (slab_mode 2): the index work of gather_one<1, CIC> in mas.cu (gather_axis with
// the CPU formula + local_planes) and its sum of eight products in the reference's order (src/mas.jl:258-265).
int64_t hc_cic_gather(const float* buf, const float* x, const float* y, const float* z, int64_t n, const int* ng, const float* L,
                      const float* mn, int slab, int z_lo, int zoff, int nzp, float* out) {
  const BoxGeom g = make_geom(ng, L, mn, slab, z_lo, zoff, nzp);
  const size_t nx = g.n[0], ny = g.n[1];
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    int xd, xu, yd, yu, zd, zu;
    float dx, ux, dy, uy, dz, uz;
    bool ok = gather_axis(x[i], g.mn[0], g.L[0], g.cell[0], g.n[0], false, xd, xu, dx, ux);
    ok = gather_axis(y[i], g.mn[1], g.L[1], g.cell[1], g.n[1], false, yd, yu, dy, uy) && ok;
    ok = gather_axis(z[i], g.mn[2], g.L[2], g.cell[2], g.n[2], false, zd, zu, dz, uz) && ok;
    ok = ok && local_planes(g, zd, zu, zd, zu);
    float v = 0.f;
    if (ok) {
      const size_t rdd = ((size_t)zd * ny + yd) * nx, rdu = ((size_t)zu * ny + yd) * nx;
      const size_t rud = ((size_t)zd * ny + yu) * nx, ruu = ((size_t)zu * ny + yu) * nx;
      v = __fmul_rn(__fmul_rn(__fmul_rn(buf[rdd + xd], dx), dy), dz);
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(buf[rdu + xd], dx), dy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(buf[rud + xd], dx), uy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(buf[ruu + xd], dx), uy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(buf[rdd + xu], ux), dy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(buf[rdu + xu], ux), dy), uz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(buf[rud + xu], ux), uy), dz));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fmul_rn(buf[ruu + xu], ux), uy), uz));
    } else {
      bad++;
    }
    out[i] = v;
  }
  return bad;
}

// The product's shifts_epilogue<3> (read_shifts: :disp / :rsd / :sum, fixed or per-particle line of sight; with
// positions != 0 reconstructed_positions' pos - shift) applied to given displacement values.
void hc_shifts_epilogue(const float* px, const float* py, const float* pz, const float* dx, const float* dy, const float* dz,
                        int64_t n, int field, int positions, int has_los, const float* los, float fgrowth, float* ox, float* oy,
                        float* oz) {
  GatherArgs a;
  a.f[0] = a.f[1] = a.f[2] = nullptr;
  a.x = px, a.y = py, a.z = pz;
  a.o[0] = ox, a.o[1] = oy, a.o[2] = oz;
  a.n = n;
  a.field = field;
  a.positions = positions;
  a.has_los = has_los;
  for (int c = 0; c < 3; c++) a.los[c] = has_los ? los[c] : 0.f;
  a.fgrowth = fgrowth;
  a.sorted_out = nullptr;
  for (int64_t i = 0; i < n; i++) {
    const float val[3] = {dx[i], dy[i], dz[i]};
    shifts_epilogue<3>(a, val, px[i], py[i], pz[i], i);
  }
}

}  // extern "C"
