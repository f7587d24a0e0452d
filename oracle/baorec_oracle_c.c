/* CPU oracle, C part -- TEST INFRASTRUCTURE ONLY (see oracle/baorec_oracle.py's header).
 *
 * Plain-C restatement of the hot loops of BAOrec.jl's `Array` (CPU) methods, Float32 like the
 * reference's examples, threaded where the reference is threaded and serial where it is serial:
 * its CPU scatter is one sequential loop over the particles (src/mas.jl:5-51), the gather runs
 * under Threads.@threads (src/mas.jl:229), the k-space and multigrid loops under @tturbo /
 * Threads.@threads (src/utils.jl:49-52, src/iterative.jl:9-14,43-62, src/multigrid.jl:37-131).
 * The FFTs stay in Python (scipy.fft = pocketfft on all host threads; FFTW is not in the image).
 *
 * Two jobs: (1) a second, independently written checker that tests/test_oracle_c.py holds against
 * the numpy restatement (scatter and gather bit-identical, the rest to rounding), (2) the CPU
 * baseline bench.py times beside the CUDA path, so that the comparison is against compiled,
 * threaded loops like the Julia code's and not against numpy's interpreter overhead.
 * Built by oracle/Makefile into oracle/_c/ (git-ignored) with -ffp-contract=off: no FMA
 * contraction, every product and sum rounds to Float32 exactly where the numpy port rounds.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may load it.
 *
 * Array convention: Julia A[ix,iy,iz] (x fastest) == C a[(iz*ny+iy)*nx+ix]; complex meshes are
 * interleaved (re, im) float pairs of shape [nz][ny][nx/2+1]. */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline runs on rank 0 alone and wants the host's cores */
EXPORT void oc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

EXPORT int oc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- src/mas.jl:1-52  cic!(rho, x, y, z, w, box_size, box_min; wrap) -------------------------
 * Serial like the reference.  Returns the number of particles that fall outside the mesh (the
 * reference raises BoundsError, :33-49); those are skipped. */
static inline int cic_axis(float p, float mn, float L, int n, int wrap, int64_t *i0, int64_t *i1, float *w0, float *w1) {
    float g = (p - mn) * (float)n;          /* (p - min) * n / L + 1, left to right (:13-15) */
    g = g / L;
    g = g + 1.0f;
    float f0 = floorf(g);
    if (!(f0 >= -1.0e9f && f0 <= 1.0e9f)) return 0;      /* NaN / far out of range */
    int64_t c0 = (int64_t)f0;                             /* 1-based */
    *w1 = g - f0;
    *w0 = 1.0f - *w1;
    if (c0 == n + 1) c0 = 1;                              /* :28-30 */
    int64_t c1 = (wrap && c0 == n) ? 1 : c0 + 1;          /* :33-35 */
    if (c0 < 1 || c0 > n || c1 > n) return 0;
    *i0 = c0 - 1;
    *i1 = c1 - 1;
    return 1;
}

EXPORT int64_t oc_cic_scatter_f32(float *rho, int nx, int ny, int nz, float *x, float *y, float *z, const float *w,
                                  int64_t N, const float *L, const float *mn, int wrap) {
    int64_t bad = 0;
    for (int64_t p = 0; p < N; ++p) {
        float px = x[p], py = y[p], pz = z[p];
        if (wrap) {                                       /* :8-10: axis-1 min/size for all axes, written back */
            if ((px - mn[0]) > L[0]) px = px - L[0];
            if ((py - mn[0]) > L[0]) py = py - L[0];
            if ((pz - mn[0]) > L[0]) pz = pz - L[0];
        }
        int64_t x0, x1, y0, y1, z0, z1;
        float wx0, wx1, wy0, wy1, wz0, wz1;
        int ok = cic_axis(px, mn[0], L[0], nx, wrap, &x0, &x1, &wx0, &wx1);
        ok &= cic_axis(py, mn[1], L[1], ny, wrap, &y0, &y1, &wy0, &wy1);
        ok &= cic_axis(pz, mn[2], L[2], nz, wrap, &z0, &z1, &wz0, &wz1);
        if (!ok) { ++bad; continue; }
        if (wrap) { x[p] = px; y[p] = py; z[p] = pz; }
        wx0 = wx0 * w[p];                                 /* :37-38 */
        wx1 = wx1 * w[p];
#define AT(zz, yy, xx) rho[((size_t)(zz) * ny + (yy)) * nx + (xx)]
        AT(z0, y0, x0) += (wx0 * wy0) * wz0;              /* :42-49, order 000,100,010,001,110,101,011,111 */
        AT(z0, y0, x1) += (wx1 * wy0) * wz0;
        AT(z0, y1, x0) += (wx0 * wy1) * wz0;
        AT(z1, y0, x0) += (wx0 * wy0) * wz1;
        AT(z0, y1, x1) += (wx1 * wy1) * wz0;
        AT(z1, y0, x1) += (wx1 * wy0) * wz1;
        AT(z1, y1, x0) += (wx0 * wy1) * wz1;
        AT(z1, y1, x1) += (wx1 * wy1) * wz1;
#undef AT
    }
    return bad;
}

/* ---- src/mas.jl:218-269  read_cic! (CPU formula d = (p-min)/cell, cell = T(L/n); formula=1:
 * the GPU method's d = (p-min)*n/L, :274).  Threaded over particles like the reference. */
static inline int gather_axis(float p, float mn, float L, int n, int formula, int64_t *id, int64_t *iu, float *wd, float *wu) {
    float d;
    if (formula == 0) { float cell = L / (float)n; d = (p - mn) / cell; }
    else { d = (p - mn) * (float)n; d = d / L; }
    float f = floorf(d);
    if (!(f >= -1.0e9f && f <= 1.0e9f)) return 0;
    *wu = d - f;
    *wd = 1.0f - *wu;
    int64_t i = (int64_t)f + 1;
    if (i > n) i -= n;
    int64_t j = i + 1;
    if (j > n) j -= n;
    if (i < 1 || i > n || j > n) return 0;
    *id = i - 1;
    *iu = j - 1;
    return 1;
}

EXPORT int64_t oc_read_cic_f32(const float *fld, int nx, int ny, int nz, const float *x, const float *y, const float *z,
                               int64_t N, const float *L, const float *mn, int formula, float *out) {
    int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int64_t p = 0; p < N; ++p) {
        int64_t xd, xu, yd, yu, zd, zu;
        float dx, ux, dy, uy, dz, uz;
        int ok = gather_axis(x[p], mn[0], L[0], nx, formula, &xd, &xu, &dx, &ux);
        ok &= gather_axis(y[p], mn[1], L[1], ny, formula, &yd, &yu, &dy, &uy);
        ok &= gather_axis(z[p], mn[2], L[2], nz, formula, &zd, &zu, &dz, &uz);
        if (!ok) { ++bad; out[p] = 0.0f; continue; }
#define F(zz, yy, xx) fld[((size_t)(zz) * ny + (yy)) * nx + (xx)]
        float s = ((F(zd, yd, xd) * dx) * dy) * dz;       /* :258-265: ddd,ddu,dud,duu,udd,udu,uud,uuu (x,y,z) */
        s = s + ((F(zu, yd, xd) * dx) * dy) * uz;
        s = s + ((F(zd, yu, xd) * dx) * uy) * dz;
        s = s + ((F(zu, yu, xd) * dx) * uy) * uz;
        s = s + ((F(zd, yd, xu) * ux) * dy) * dz;
        s = s + ((F(zu, yd, xu) * ux) * dy) * uz;
        s = s + ((F(zd, yu, xu) * ux) * uy) * dz;
        s = s + ((F(zu, yu, xu) * ux) * uy) * uz;
#undef F
        out[p] = s;
    }
    return bad;
}

/* ---- TSC (extension, NOT in the reference; same restatement as baorec_oracle.py tsc_cells / tsc_scatter / read_tsc):
 * g = (p - min) * n / L, centre c = floor(g + 0.5), d = g - c, weights 0.5 (0.5 - d)^2, 0.75 - d^2, 0.5 (0.5 + d)^2.
 * The scatter is serial over the particles (like cic!); the numpy port loops offset-major instead, so the two meshes
 * agree to summation order only.  The gather accumulates in the numpy port's order (oz, oy, ox): bit-identical. */
static inline int tsc_axis(float p, float mn, float L, int n, int wrap, int64_t idx[3], float w[3]) {
    float g = (p - mn) * (float)n;
    g = g / L;
    float c = floorf(g + 0.5f);
    if (!(c >= -1.0e9f && c <= 1.0e9f)) return 0;
    float d = g - c;
    float hm = 0.5f - d, hp = 0.5f + d;
    w[0] = 0.5f * (hm * hm);
    w[1] = 0.75f - (d * d);
    w[2] = 0.5f * (hp * hp);
    int64_t ic = (int64_t)c;
    for (int o = 0; o < 3; ++o) {
        int64_t i = ic + o - 1;
        if (wrap) { i %= n; if (i < 0) i += n; }
        else if (i < 0 || i >= n) return 0;
        idx[o] = i;
    }
    return 1;
}

EXPORT int64_t oc_tsc_scatter_f32(float *rho, int nx, int ny, int nz, const float *x, const float *y, const float *z,
                                  const float *w, int64_t N, const float *L, const float *mn, int wrap) {
    int64_t bad = 0;
    for (int64_t p = 0; p < N; ++p) {
        int64_t ix[3], iy[3], iz[3];
        float wx[3], wy[3], wz[3];
        int ok = tsc_axis(x[p], mn[0], L[0], nx, wrap, ix, wx);
        ok &= tsc_axis(y[p], mn[1], L[1], ny, wrap, iy, wy);
        ok &= tsc_axis(z[p], mn[2], L[2], nz, wrap, iz, wz);
        if (!ok) { ++bad; continue; }
        for (int oz = 0; oz < 3; ++oz)
            for (int oy = 0; oy < 3; ++oy)
                for (int ox = 0; ox < 3; ++ox)
                    rho[((size_t)iz[oz] * ny + iy[oy]) * nx + ix[ox]] += ((wx[ox] * w[p]) * wy[oy]) * wz[oz];
    }
    return bad;
}

EXPORT int64_t oc_read_tsc_f32(const float *fld, int nx, int ny, int nz, const float *x, const float *y, const float *z,
                               int64_t N, const float *L, const float *mn, float *out) {
    int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int64_t p = 0; p < N; ++p) {
        int64_t ix[3], iy[3], iz[3];
        float wx[3], wy[3], wz[3];
        int ok = tsc_axis(x[p], mn[0], L[0], nx, 1, ix, wx);
        ok &= tsc_axis(y[p], mn[1], L[1], ny, 1, iy, wy);
        ok &= tsc_axis(z[p], mn[2], L[2], nz, 1, iz, wz);
        if (!ok) { ++bad; out[p] = 0.0f; continue; }
        float s = 0.0f;
        for (int oz = 0; oz < 3; ++oz)
            for (int oy = 0; oy < 3; ++oy)
                for (int ox = 0; ox < 3; ++ox)
                    s = s + ((fld[((size_t)iz[oz] * ny + iy[oy]) * nx + ix[ox]] * wx[ox]) * wy[oy]) * wz[oz];
        out[p] = s;
    }
    return bad;
}

/* ---- k-space passes over a Complex{Float32} half-mesh [nz][ny][nxh] ----------------------------
 * op 0: src/utils.jl:49-52   F *= exp(-0.5 R^2 k^2), exponent / exp / product in Float64 (a = R^2 as Float32)
 * op 1: src/iterative.jl:9-14  F /= k^2 (k^2 == 0 -> 1), F[k=0] = 0
 * op 2: src/iterative.jl:53    out = ((k_a k_a) * a) * F            (fixed LOS term; a = los_a)
 * op 3: src/iterative.jl:260-271  out = (i k_a) F / k^2, 0 at k = 0  (IterativeRecon displacement)
 * op 4: src/multigrid.jl:761-769  out = (i k_a) F                     (MultigridRecon displacement)
 * op 5: src/iterative.jl:27-30  out = (k_a k_b) F                    (radial LOS pair term) */
EXPORT void oc_kspace_c64(float *out, const float *in, int nxh, int ny, int nz, const float *kx, const float *ky,
                          const float *kz, int op, int axis, int axis_b, float a) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int iz = 0; iz < nz; ++iz)
        for (int iy = 0; iy < ny; ++iy) {
            const float ky2 = ky[iy] * ky[iy], kz2 = kz[iz] * kz[iz];
            size_t row = ((size_t)iz * ny + iy) * nxh;
            for (int ix = 0; ix < nxh; ++ix) {
                const float re = in[2 * (row + ix)], im = in[2 * (row + ix) + 1];
                const float k2 = (kx[ix] * kx[ix] + ky2) + kz2;
                const float kk[3] = {kx[ix], ky[iy], kz[iz]};
                float ore, oim;
                switch (op) {
                case 0: {
                    double g = exp(-0.5 * (double)a * (double)k2);
                    ore = (float)((double)re * g);
                    oim = (float)((double)im * g);
                } break;
                case 1: {
                    float d = (k2 == 0.0f) ? 1.0f : k2;
                    ore = re / d;
                    oim = im / d;
                    if (ix == 0 && iy == 0 && iz == 0) ore = oim = 0.0f;
                } break;
                case 2: {
                    float c = (kk[axis] * kk[axis]) * a;
                    ore = c * re;
                    oim = c * im;
                } break;
                case 3:
                    if (k2 > 0.0f) { ore = (-(kk[axis] * im)) / k2; oim = (kk[axis] * re) / k2; }
                    else ore = oim = 0.0f;
                    break;
                case 4:
                    ore = -(kk[axis] * im);
                    oim = kk[axis] * re;
                    break;
                default: {
                    float c = kk[axis] * kk[axis_b];
                    ore = c * re;
                    oim = c * im;
                } break;
                }
                out[2 * (row + ix)] = ore;
                out[2 * (row + ix) + 1] = oim;
            }
        }
}

/* ---- real-space passes ---------------------------------------------------------------------- */
/* src/recon.jl:54-55  delta = (rho / mean(rho) - 1) / bias; the mean accumulates in Float64 here
 * (Julia's mean is a pairwise Float32 sum: tolerance-level difference). */
EXPORT void oc_overdensity_box_f32(float *delta, int64_t M, float bias) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < M; ++i) s += delta[i];
    const float mean = (float)(s / (double)M);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) delta[i] = ((delta[i] / mean) - 1.0f) / bias;
}

/* src/iterative.jl:56-58  delta_r -= fac * dpsi (fixed LOS) */
EXPORT void oc_axpy_f32(float *delta_r, const float *dpsi, int64_t M, float fac) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) delta_r[i] = delta_r[i] - fac * dpsi[i];
}

/* src/iterative.jl:31-36  delta_r -= fac * dpsi * x_a * x_b / |x|^2, 0 where |x|^2 == 0 (radial LOS) */
EXPORT void oc_radial_update_f32(float *delta_r, const float *dpsi, int nx, int ny, int nz, const float *xv, const float *yv,
                                 const float *zv, int a, int b, float fac) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int iz = 0; iz < nz; ++iz)
        for (int iy = 0; iy < ny; ++iy) {
            size_t row = ((size_t)iz * ny + iy) * nx;
            for (int ix = 0; ix < nx; ++ix) {
                const float c[3] = {xv[ix], yv[iy], zv[iz]};
                const float x2 = (c[0] * c[0] + c[1] * c[1]) + c[2] * c[2];
                const float u = delta_r[row + ix] - (((fac * dpsi[row + ix]) * c[a]) * c[b]) / x2;
                delta_r[row + ix] = (x2 > 0.0f) ? u : 0.0f;
            }
        }
}

/* src/recon.jl:264-306  shifts epilogue: out = disp (field 0), f (disp . r) r (1), disp + that (2);
 * r = los or p/|p| per particle when los == NULL.  In place on (sx, sy, sz). */
EXPORT void oc_shifts_f32(float *sx, float *sy, float *sz, const float *x, const float *y, const float *z, int64_t N,
                          int field, float f, const float *los) {
    if (field == 0) return;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N; ++p) {
        float lx, ly, lz;
        if (los) { lx = los[0]; ly = los[1]; lz = los[2]; }
        else {
            float d = sqrtf((x[p] * x[p] + y[p] * y[p]) + z[p] * z[p]);
            lx = x[p] / d; ly = y[p] / d; lz = z[p] / d;
        }
        const float dot = (sx[p] * lx + sy[p] * ly) + sz[p] * lz;
        const float rx = (f * dot) * lx, ry = (f * dot) * ly, rz = (f * dot) * lz;
        if (field == 1) { sx[p] = rx; sy[p] = ry; sz[p] = rz; }
        else { sx[p] = sx[p] + rx; sy[p] = sy[p] + ry; sz[p] = sz[p] + rz; }
    }
}

/* ---- src/multigrid.jl -------------------------------------------------------------------------
 * One stencil evaluation (jacobi!: :14-135, residual!: :213-326).  mode 0: out = (1-w) v + w * jac,
 * jac = (f + offdiag(v)) / diag; mode 1: out = f - (diag v - offdiag(v)).  xv,yv,zv = cell centres
 * (radial) or NULL with los given.  Periodic neighbours. */
/* Yardstick hook (tests only): with the flag set the stencil below evaluates the SAME formula with its Float32 sums
 * associated differently (terms added in reverse order, the damping and the residual written in their other algebraic
 * form) -- what LoopVectorization's @tturbo (src/multigrid.jl:37...) is free to do, and what any other correct
 * implementation does somewhere.  The distance between the two results is how sharply the Float32 reference defines
 * its own answer: on a lightcone the right-hand side has a non-zero mean, the iterate drifts to a constant ~70x the
 * rms of its fluctuations, and the stencil's cancellation then leaves the potential defined to ~1e-2 only. */
static int oc_reassociate = 0;
EXPORT void oc_set_reassociate(int on) { oc_reassociate = on; }

EXPORT void oc_mg_stencil_f32(float *out, const float *v, const float *f, int nx, int ny, int nz, const float *L,
                              const float *xv, const float *yv, const float *zv, const float *los, float beta, float w,
                              int mode) {
    const float cell[3] = {L[0] / (float)nx, L[1] / (float)ny, L[2] / (float)nz};
    const float c2[3] = {cell[0] * cell[0], cell[1] * cell[1], cell[2] * cell[2]};
    const float ic2[3] = {1.0f / c2[0], 1.0f / c2[1], 1.0f / c2[2]};
    const int radial = (los == NULL);
    const int re = oc_reassociate;
#pragma omp parallel for schedule(static) collapse(2)
    for (int iz = 0; iz < nz; ++iz)
        for (int iy = 0; iy < ny; ++iy) {
            const int zp = (iz + 1) % nz, zm = (iz + nz - 1) % nz, yp = (iy + 1) % ny, ym = (iy + ny - 1) % ny;
            for (int ix = 0; ix < nx; ++ix) {
                const int xp = (ix + 1) % nx, xm = (ix + nx - 1) % nx;
                float px, py, pz;
                if (radial) { px = xv[ix] / cell[0]; py = yv[iy] / cell[1]; pz = zv[iz] / cell[2]; }
                else { px = los[0] / cell[0]; py = los[1] / cell[1]; pz = los[2] / cell[2]; }
                const float den = (c2[0] * (px * px) + c2[1] * (py * py)) + c2[2] * (pz * pz);
                const float g = beta / den;
                const float gx = ic2[0] + g * (px * px), gy = ic2[1] + g * (py * py), gz = ic2[2] + g * (pz * pz);
#define V(zz, yy, xx) v[((size_t)(zz) * ny + (yy)) * nx + (xx)]
                float s = (gx * (V(iz, iy, xp) + V(iz, iy, xm)) + gy * (V(iz, yp, ix) + V(iz, ym, ix))) +
                          gz * (V(zp, iy, ix) + V(zm, iy, ix));
                const float cxy = ((V(iz, yp, xp) + V(iz, ym, xm)) - V(iz, yp, xm)) - V(iz, ym, xp);
                const float cxz = ((V(zp, iy, xp) + V(zm, iy, xm)) - V(zp, iy, xm)) - V(zm, iy, xp);
                const float cyz = ((V(zp, yp, ix) + V(zm, ym, ix)) - V(zp, ym, ix)) - V(zm, yp, ix);
                s = s + (g / 2.0f) * (((px * py) * cxy + (px * pz) * cxz) + (py * pz) * cyz);
                if (radial)
                    s = s + g * ((px * (V(iz, iy, xp) - V(iz, iy, xm)) + py * (V(iz, yp, ix) - V(iz, ym, ix))) +
                                 pz * (V(zp, iy, ix) - V(zm, iy, ix)));
                const float diag = 2.0f * ((gx + gy) + gz);
                const size_t c = ((size_t)iz * ny + iy) * nx + ix;
                if (re) {   /* the same terms, summed from the other end; see oc_set_reassociate */
                    float t = gz * (V(zm, iy, ix) + V(zp, iy, ix)) + (gy * (V(iz, ym, ix) + V(iz, yp, ix)) + gx * (V(iz, iy, xm) + V(iz, iy, xp)));
                    t = (g / 2.0f) * ((py * pz) * cyz + ((px * pz) * cxz + (px * py) * cxy)) + t;
                    if (radial)
                        t = g * (pz * (V(zp, iy, ix) - V(zm, iy, ix)) + (py * (V(iz, yp, ix) - V(iz, ym, ix)) + px * (V(iz, iy, xp) - V(iz, iy, xm)))) + t;
                    if (mode == 0) out[c] = v[c] + w * ((t + f[c]) / diag - v[c]);
                    else out[c] = (f[c] + t) - diag * v[c];
                } else if (mode == 0) out[c] = (1.0f - w) * v[c] + w * ((f[c] + s) / diag);
                else out[c] = f[c] - (diag * v[c] - s);
#undef V
            }
        }
}

/* reduce! src/multigrid.jl:520-584: coarse c <- fine 2c+1, weights 8/4/2/1 over 27 points, /64 */
EXPORT void oc_mg_restrict_f32(float *coarse, const float *fine, int nx, int ny, int nz) {
    const int cx = nx / 2, cy = ny / 2, cz = nz / 2;
#pragma omp parallel for schedule(static) collapse(2)
    for (int kz = 0; kz < cz; ++kz)
        for (int ky = 0; ky < cy; ++ky)
            for (int kx = 0; kx < cx; ++kx) {
                float acc = 0.0f;
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int m = abs(dz) + abs(dy) + abs(dx);
                            const float wgt = (m == 0) ? 8.0f : (m == 1) ? 4.0f : (m == 2) ? 2.0f : 1.0f;
                            const int fz = (2 * kz + 1 + dz + nz) % nz, fy = (2 * ky + 1 + dy + ny) % ny,
                                      fx = (2 * kx + 1 + dx + nx) % nx;
                            acc += wgt * fine[((size_t)fz * ny + fy) * nx + fx];
                        }
                coarse[((size_t)kz * cy + ky) * cx + kx] = acc / 64.0f;
            }
}

/* prolong! src/multigrid.jl:398-453: fine 2c+1 <- coarse c; fine (2c+2) mod n <- mean over the + neighbours */
EXPORT void oc_mg_prolong_f32(float *fine, const float *coarse, int nx, int ny, int nz) {
    const int cx = nx / 2, cy = ny / 2, cz = nz / 2;
#pragma omp parallel for schedule(static) collapse(2)
    for (int kz = 0; kz < cz; ++kz)
        for (int ky = 0; ky < cy; ++ky)
            for (int kx = 0; kx < cx; ++kx) {
                const int k1z = (kz + 1) % cz, k1y = (ky + 1) % cy, k1x = (kx + 1) % cx;
#define C(zz, yy, xx) coarse[((size_t)(zz) * cy + (yy)) * cx + (xx)]
#define FI(zz, yy, xx) fine[((size_t)(zz) * ny + (yy)) * nx + (xx)]
                const int zo = 2 * kz + 1, yo = 2 * ky + 1, xo = 2 * kx + 1;
                const int ze = (zo + 1) % nz, ye = (yo + 1) % ny, xe = (xo + 1) % nx;
                const float c000 = C(kz, ky, kx), c001 = C(kz, ky, k1x), c010 = C(kz, k1y, kx), c100 = C(k1z, ky, kx);
                const float c011 = C(kz, k1y, k1x), c110 = C(k1z, k1y, kx), c101 = C(k1z, ky, k1x), c111 = C(k1z, k1y, k1x);
                FI(zo, yo, xo) = c000;
                FI(zo, yo, xe) = (c000 + c001) / 2.0f;
                FI(zo, ye, xo) = (c000 + c010) / 2.0f;
                FI(ze, yo, xo) = (c000 + c100) / 2.0f;
                FI(zo, ye, xe) = (((c000 + c001) + c010) + c011) / 4.0f;
                FI(ze, ye, xo) = (((c000 + c010) + c100) + c110) / 4.0f;
                FI(ze, yo, xe) = (((c000 + c001) + c100) + c101) / 4.0f;
                FI(ze, ye, xe) = (((((((c000 + c001) + c010) + c100) + c011) + c101) + c110) + c111) / 8.0f;
#undef C
#undef FI
            }
}
