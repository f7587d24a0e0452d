"""CPU oracle with compiled hot loops -- TEST INFRASTRUCTURE ONLY.

`load()` returns a module object with the same functions as `baorec_oracle` (run, read_shifts,
setup_overdensity, iterate, jacobi, fmg ...), in which the Float32 hot loops are replaced by the C
restatement in oracle/baorec_oracle_c.c (oracle/_c/libbaorec_oracle_c.so, built by
`make -C oracle`): serial scatter like the reference's CPU `cic!`, OpenMP-threaded gather, k-space
and multigrid loops like its threaded CPU methods, scipy.fft on all host threads for the
transforms.  The driver logic (set-up, iteration, V-cycle, FMG, read-back) is the numpy oracle's own
code, re-executed in a private copy of that module whose loop functions are rebound -- one
statement of the algorithm, two implementations of the loops, compared by tests/test_oracle_c.py.

Float64 inputs fall through to the numpy loops.  Used by bench.py's `cpu_baseline` / `--impl
reference` legs (the CPU baseline should be compiled threaded code, like the Julia it stands in
for) and by the tests; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_c" / "libbaorec_oracle_c.so"

_F = C.POINTER(C.c_float)


def _p(a):
    return a.ctypes.data_as(_F)


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _bind(lib):
    i, i64, f = C.c_int, C.c_int64, C.c_float
    lib.oc_max_threads.restype = i
    lib.oc_set_num_threads.restype = None
    lib.oc_set_num_threads.argtypes = [i]
    lib.oc_cic_scatter_f32.restype = i64
    lib.oc_cic_scatter_f32.argtypes = [_F, i, i, i, _F, _F, _F, _F, i64, _F, _F, i]
    lib.oc_read_cic_f32.restype = i64
    lib.oc_read_cic_f32.argtypes = [_F, i, i, i, _F, _F, _F, i64, _F, _F, i, _F]
    lib.oc_tsc_scatter_f32.restype = i64
    lib.oc_tsc_scatter_f32.argtypes = [_F, i, i, i, _F, _F, _F, _F, i64, _F, _F, i]
    lib.oc_read_tsc_f32.restype = i64
    lib.oc_read_tsc_f32.argtypes = [_F, i, i, i, _F, _F, _F, i64, _F, _F, _F]
    lib.oc_kspace_c64.restype = None
    lib.oc_kspace_c64.argtypes = [_F, _F, i, i, i, _F, _F, _F, i, i, i, f]
    lib.oc_overdensity_box_f32.restype = None
    lib.oc_overdensity_box_f32.argtypes = [_F, i64, f]
    lib.oc_axpy_f32.restype = None
    lib.oc_axpy_f32.argtypes = [_F, _F, i64, f]
    lib.oc_radial_update_f32.restype = None
    lib.oc_radial_update_f32.argtypes = [_F, _F, i, i, i, _F, _F, _F, i, i, f]
    lib.oc_shifts_f32.restype = None
    lib.oc_shifts_f32.argtypes = [_F, _F, _F, _F, _F, _F, i64, i, f, _F]
    lib.oc_mg_stencil_f32.restype = None
    lib.oc_mg_stencil_f32.argtypes = [_F, _F, _F, i, i, i, _F, _F, _F, _F, _F, f, f, i]
    lib.oc_set_reassociate.restype = None
    lib.oc_set_reassociate.argtypes = [i]
    lib.oc_mg_restrict_f32.restype = None
    lib.oc_mg_restrict_f32.argtypes = [_F, _F, i, i, i]
    lib.oc_mg_prolong_f32.restype = None
    lib.oc_mg_prolong_f32.argtypes = [_F, _F, i, i, i]
    return lib


def available() -> bool:
    return LIB_PATH.exists()


def load(threads=None):
    """A private copy of the numpy oracle with its Float32 loops rebound to the C library.
    threads: OpenMP threads for the loops (default: the runtime's own choice, i.e. OMP_NUM_THREADS or all cores)."""
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `make -C oracle` (or __graft_entry__.build())")
    lib = _bind(C.CDLL(str(LIB_PATH)))
    if threads:
        lib.oc_set_num_threads(int(threads))
    spec = importlib.util.spec_from_file_location("baorec_oracle_fastcopy", HERE / "baorec_oracle.py")
    O = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = O          # dataclasses look their module up while the class body runs
    spec.loader.exec_module(O)
    N = {k: getattr(O, k) for k in ("tsc_scatter", "read_tsc", "cic_scatter", "read_cic", "smooth", "setup_overdensity", "iterate", "jacobi",
                                    "residual", "restrict", "prolong", "displacement_meshes", "read_shifts")}
    f32 = np.float32

    def is32(*arrs):
        return all(isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous for a in arrs)

    def vec3(v):
        return _f32c(O._vec3(v, f32))

    def cic_scatter(rho, x, y, z, w, box_size, box_min, wrap=True):
        if not is32(rho, x, y, z) or not x.flags.writeable:
            return N["cic_scatter"](rho, x, y, z, w, box_size, box_min, wrap)
        nz, ny, nx = rho.shape
        w = _f32c(w)
        L, mn = vec3(box_size), vec3(box_min)
        bad = lib.oc_cic_scatter_f32(_p(rho), nx, ny, nz, _p(x), _p(y), _p(z), _p(w), len(x), _p(L), _p(mn), int(bool(wrap)))
        if bad:
            raise O.OutOfBoxError(f"{bad} particle(s) outside the mesh (reference: BoundsError, src/mas.jl:33-49)")
        return rho

    def read_cic(fld, x, y, z, box_size, box_min, wrap=True, formula="cpu"):
        if not is32(fld, x, y, z) or not wrap:
            return N["read_cic"](fld, x, y, z, box_size, box_min, wrap, formula)
        nz, ny, nx = fld.shape
        out = np.empty(len(x), f32)
        L, mn = vec3(box_size), vec3(box_min)
        bad = lib.oc_read_cic_f32(_p(fld), nx, ny, nz, _p(x), _p(y), _p(z), len(x), _p(L), _p(mn),
                                  0 if formula == "cpu" else 1, _p(out))
        if bad:
            raise O.OutOfBoxError(f"{bad} particle(s) outside the mesh in read_cic")
        return out

    def tsc_scatter(rho, x, y, z, w, box_size, box_min, wrap=True):
        if not is32(rho, x, y, z):
            return N["tsc_scatter"](rho, x, y, z, w, box_size, box_min, wrap)
        nz, ny, nx = rho.shape
        w = _f32c(w)
        L, mn = vec3(box_size), vec3(box_min)
        bad = lib.oc_tsc_scatter_f32(_p(rho), nx, ny, nz, _p(x), _p(y), _p(z), _p(w), len(x), _p(L), _p(mn), int(bool(wrap)))
        if bad:
            raise O.OutOfBoxError(f"{bad} particle(s): TSC stencil leaves the mesh with wrap=False")
        return rho

    def read_tsc(fld, x, y, z, box_size, box_min, wrap=True):
        if not is32(fld, x, y, z) or not wrap:
            return N["read_tsc"](fld, x, y, z, box_size, box_min, wrap)
        nz, ny, nx = fld.shape
        out = np.empty(len(x), f32)
        L, mn = vec3(box_size), vec3(box_min)
        bad = lib.oc_read_tsc_f32(_p(fld), nx, ny, nz, _p(x), _p(y), _p(z), len(x), _p(L), _p(mn), _p(out))
        if bad:
            raise O.OutOfBoxError(f"{bad} particle(s) outside the mesh in read_tsc")
        return out

    def kspace(fk, kv, op, axis=0, axis_b=0, a=0.0, out=None):
        """fk: complex64 [nz][ny][nxh] C-contiguous; returns a complex64 array (in place when out is fk)."""
        nz, ny, nxh = fk.shape
        if out is None:
            out = np.empty_like(fk)
        kx, ky, kz = (_f32c(k) for k in kv)
        lib.oc_kspace_c64(_p(out.view(f32)), _p(fk.view(f32)), nxh, ny, nz, _p(kx), _p(ky), _p(kz), op, axis, axis_b, float(a))
        return out

    def c64(a):
        return np.ascontiguousarray(a, dtype=np.complex64)

    def smooth(fld, smoothing_radius, box_size):
        if not is32(fld):
            return N["smooth"](fld, smoothing_radius, box_size)
        nz, ny, nx = fld.shape
        fk = c64(O.rfft(fld))
        kv = O.k_vec((nx, ny, nz), box_size, f32)
        R2 = f32(f32(smoothing_radius) * f32(smoothing_radius))
        kspace(fk, kv, 0, a=R2, out=fk)
        fld[...] = O.irfft(fk, fld.shape)
        return fld

    def setup_overdensity(delta, recon, x, y, z, w, rx=None, ry=None, rz=None, rw=None, wrap=True, ran_min=0.01,
                          force_mask=None, info=None):
        if rx is not None or not is32(delta):
            return N["setup_overdensity"](delta, recon, x, y, z, w, rx, ry, rz, rw, wrap, ran_min, force_mask, info)
        O._scatter(recon, delta, x, y, z, w, wrap)
        O.smooth(delta, f32(recon.smoothing_radius), recon.box_size)
        lib.oc_overdensity_box_f32(_p(delta), delta.size, float(recon.bias))
        return delta

    def iterate(delta_r, delta_s, kv, it, beta, los=None, xv=None):
        if not is32(delta_r, delta_s):
            return N["iterate"](delta_r, delta_s, kv, it, beta, los, xv)
        beta = f32(beta)
        shape = delta_r.shape
        dk = c64(O.rfft(delta_r))
        kspace(dk, kv, 1, out=dk)
        delta_r[...] = delta_s
        if los is None:
            xs = [_f32c(v) for v in xv]
            for i in range(3):
                for j in range(i, 3):
                    fac = f32((1.0 + float(i != j)) * np.float64(beta))
                    if it == 1:
                        fac = f32(fac / (f32(1) + beta))
                    dpx = _f32c(O.irfft(kspace(dk, kv, 5, i, j), shape))
                    lib.oc_radial_update_f32(_p(delta_r), _p(dpx), shape[2], shape[1], shape[0], _p(xs[0]), _p(xs[1]),
                                             _p(xs[2]), i, j, float(fac))
        else:
            los = np.asarray(los, dtype=f32)
            for i in range(3):
                if los[i] == 0:
                    continue
                fac = f32(beta / (f32(1) + beta)) if it == 1 else beta
                dpx = _f32c(O.irfft(kspace(dk, kv, 2, i, a=los[i]), shape))
                lib.oc_axpy_f32(_p(delta_r), _p(dpx), delta_r.size, float(fac))
        return delta_r

    def displacement_meshes(mesh, recon):
        if not is32(mesh):
            return N["displacement_meshes"](mesh, recon)
        nz, ny, nx = mesh.shape
        kv = O.k_vec((nx, ny, nz), recon.box_size, f32)
        dk = c64(O.rfft(mesh))
        op = 4 if isinstance(recon, O.MultigridRecon) else 3
        return [_f32c(O.irfft(kspace(dk, kv, op, a), mesh.shape)) for a in range(3)]

    def compute_displacements(mesh, x, y, z, recon, formula="cpu"):
        if isinstance(recon, O.MultigridRecon) and getattr(recon, "fd_gradient", False):
            return O.read_grad_cic(mesh, x, y, z, recon.box_size, recon.box_min)
        return [O._read(recon, m, x, y, z, formula) for m in O.displacement_meshes(mesh, recon)]

    def read_shifts(recon, x, y, z, mesh, field="disp", formula="cpu"):
        if not is32(mesh, x, y, z):
            return N["read_shifts"](recon, x, y, z, mesh, field, formula)
        s = O.compute_displacements(mesh, x, y, z, recon, formula)
        code = {"disp": 0, "rsd": 1, "sum": 2}[field]
        los = None if recon.los is None else _f32c(recon.los)
        lib.oc_shifts_f32(_p(s[0]), _p(s[1]), _p(s[2]), _p(x), _p(y), _p(z), len(x), code, float(recon.f),
                          _p(los) if los is not None else None)
        return s

    def _stencil(out, v, f, xv, box_size, beta, w, los, mode):
        nz, ny, nx = v.shape
        L = vec3(box_size)
        if los is None:
            xs = [_f32c(a) for a in xv]
            lib.oc_mg_stencil_f32(_p(out), _p(v), _p(f), nx, ny, nz, _p(L), _p(xs[0]), _p(xs[1]), _p(xs[2]), None,
                                  float(beta), float(w), mode)
        else:
            l3 = _f32c(los)
            lib.oc_mg_stencil_f32(_p(out), _p(v), _p(f), nx, ny, nz, _p(L), None, None, None, _p(l3), float(beta), float(w), mode)

    def jacobi(v, f, xv, box_size, box_min, beta, damping_factor, niterations, los=None):
        if not is32(v, f):
            return N["jacobi"](v, f, xv, box_size, box_min, beta, damping_factor, niterations, los)
        a, b = v, np.empty_like(v)
        for _ in range(niterations):
            _stencil(b, a, f, xv, box_size, beta, damping_factor, los, 0)
            a, b = b, a
        if a is not v:
            v[...] = a
        return v

    def residual(v, f, xv, box_size, box_min, beta, los=None):
        if not is32(v, f):
            return N["residual"](v, f, xv, box_size, box_min, beta, los)
        r = np.empty_like(v)
        _stencil(r, v, f, xv, box_size, beta, 0.0, los, 1)
        return r

    def restrict(v1h):
        if not is32(v1h):
            return N["restrict"](v1h)
        nz, ny, nx = v1h.shape
        c = np.empty((nz // 2, ny // 2, nx // 2), f32)
        lib.oc_mg_restrict_f32(_p(c), _p(v1h), nx, ny, nz)
        return c

    def prolong(v1h, v2h):
        if not is32(v1h, v2h):
            return N["prolong"](v1h, v2h)
        nz, ny, nx = v1h.shape
        lib.oc_mg_prolong_f32(_p(v1h), _p(v2h), nx, ny, nz)
        return v1h

    for name, fn in dict(tsc_scatter=tsc_scatter, read_tsc=read_tsc, cic_scatter=cic_scatter, read_cic=read_cic, smooth=smooth, setup_overdensity=setup_overdensity,
                         iterate=iterate, displacement_meshes=displacement_meshes, compute_displacements=compute_displacements,
                         compute_displacements_iterative=lambda m, x, y, z, r, formula="cpu": compute_displacements(m, x, y, z, r, formula),
                         compute_displacements_multigrid=lambda m, x, y, z, r, formula="cpu": compute_displacements(m, x, y, z, r, formula),
                         read_shifts=read_shifts, jacobi=jacobi, residual=residual, restrict=restrict, prolong=prolong).items():
        setattr(O, name, fn)
    O.set_reassociate = lambda on: lib.oc_set_reassociate(int(bool(on)))    # tests: a differently associated evaluation of the same stencil
    O.threads = lib.oc_max_threads()
    O.kspace_c = kspace
    O.numpy_loops = N
    return O
