"""Host-side mirror of BAOrec.jl's API (src/recon.jl, src/mas.jl, src/utils.jl,
src/iterative.jl, src/multigrid.jl) on top of the C ABI of libbaorec_b200.so.

Same names, argument order and semantics as the reference (Julia's `f!` is `f`
here): the mesh passed in is filled in place and returned, `recon.result_cache`
holds it, `run` with randoms overrides `box_size/box_min` with pad 500, and
`cic` with wrap mutates the caller's position arrays.

Array types select the path exactly like the reference's dispatch on
Array / CuArray:
  * torch CUDA tensors (float32)  -> device entry points, caller's stream;
  * numpy arrays (float32)        -> `run` / `read_shifts` /
    `reconstructed_positions` go through the host-buffer pipeline entry points
    (upload + solve + download inside the library).
Meshes are torch tensors of shape (nz, ny, nx) (== Julia Array (nx,ny,nz) memory);
`grid_size` is given Julia-style as (nx, ny, nz).

torch is used for device memory and streams only.  Every computation happens
in the CUDA library; nothing here falls back to PyTorch or numpy math.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field as _dcfield
from typing import Optional, Sequence

import numpy as np
import torch

from . import lib_loader as L


# --------------------------------------------------------------------------------------------
# context / plan handling (replaces setup_fft!, src/recon.jl:35-37)
# --------------------------------------------------------------------------------------------
class Context:
    """One baorec_ctx per device."""
    _by_device = {}

    def __init__(self, device: int):
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.baorec_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)
        self.plan_key = None

    @classmethod
    def get(cls, device: Optional[int] = None) -> "Context":
        if device is None:
            device = torch.cuda.current_device()
        ctx = cls._by_device.get(device)
        if ctx is None:
            ctx = cls._by_device[device] = Context(device)
        return ctx

    def plan(self, grid_size_xyz, box_size, box_min):
        nx, ny, nz = (int(v) for v in grid_size_xyz)
        bs = tuple(float(np.float32(v)) for v in box_size)
        bm = tuple(float(np.float32(v)) for v in box_min)
        key = (nx, ny, nz, bs, bm)
        if key != self.plan_key:
            L.check(self.lib.baorec_plan(self.handle, nx, ny, nz, L.f3(bs), L.f3(bm)))
            self.plan_key = key
        return self

    def set_box(self, box_size, box_min):
        nx, ny, nz = self.plan_key[:3]
        return self.plan((nx, ny, nz), box_size, box_min)

    def launch_counts(self):
        k, f = C.c_int64(), C.c_int64()
        L.check(self.lib.baorec_launch_counts(self.handle, C.byref(k), C.byref(f)))
        return k.value, f.value

    def stage_ms(self):
        buf = (C.c_float * 8)()
        n = self.lib.baorec_last_stage_ms(self.handle, buf, 8)
        return [buf[i] for i in range(n)]

    def set_option(self, name: str, value: int):
        L.check(self.lib.baorec_set_option(self.handle, name.encode(), int(value)))

    def profile(self, on: bool):
        L.check(self.lib.baorec_profile_enable(self.handle, int(bool(on))))

    def profile_read(self):
        """{kernel name: (total ms, launches)} for the current profiling window."""
        cap, stride = 64, 96
        names = C.create_string_buffer(cap * stride)
        ms = (C.c_float * cap)()
        cnt = (C.c_int32 * cap)()
        n = self.lib.baorec_profile_read(self.handle, names, stride, ms, cnt, cap)
        if n < 0:
            L.check(n)
        out = {}
        for i in range(n):
            nm = names.raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode()
            out[nm] = (float(ms[i]), int(cnt[i]))
        return out

    def scratch_bytes(self):
        return int(self.lib.baorec_scratch_bytes(self.handle))

    def sort_reuse_count(self):
        """Read-backs that reused the tile sort of the preceding run! (same unmodified position arrays)."""
        return int(self.lib.baorec_sort_reuse_count(self.handle))

    def close(self):
        if self.handle:
            self.lib.baorec_destroy(self.handle)
            self.handle = None
            Context._by_device.pop(self.device, None)


@dataclass
class FFTPlan:
    """What `recon.fft_plan` holds: the planned context for a mesh shape."""
    ctx: Context
    grid_size: tuple


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def _is_dev(a):
    return isinstance(a, torch.Tensor) and a.is_cuda


def _chk_vec(*arrs):
    n = None
    for a in arrs:
        if not (_is_dev(a) and a.dtype == torch.float32 and a.dim() == 1 and a.is_contiguous()):
            raise TypeError("particle arrays must be contiguous 1-D float32 CUDA tensors")
        if n is None:
            n = a.numel()
        elif a.numel() != n:
            raise ValueError("particle arrays differ in length")
    return n or 0


def _chk_mesh(m):
    if not (_is_dev(m) and m.dtype == torch.float32 and m.dim() == 3 and m.is_contiguous()):
        raise TypeError("mesh must be a contiguous 3-D float32 CUDA tensor of shape (nz, ny, nx)")
    nz, ny, nx = m.shape
    return nx, ny, nz


def _los3(los):
    if los is None:
        return None
    return L.f3(los)


# --------------------------------------------------------------------------------------------
# parameter structs (src/recon.jl:1-33)
# --------------------------------------------------------------------------------------------
class AbstractRecon:
    algorithm = L.ITERATIVE

    def _params(self) -> L.Params:
        p = L.Params()
        p.bias = float(self.bias)
        p.f = float(self.f)
        p.smoothing_radius = float(self.smoothing_radius)
        p.beta = float(self.beta)
        p.n_iter = int(getattr(self, "n_iter", 0))
        p.has_los = 0 if self.los is None else 1
        los = (0.0, 0.0, 0.0) if self.los is None else self.los
        for i in range(3):
            p.los[i] = float(los[i])
        p.jacobi_damping_factor = float(getattr(self, "jacobi_damping_factor", 0.4))
        p.jacobi_niterations = int(getattr(self, "jacobi_niterations", 5))
        p.vcycle_niterations = int(getattr(self, "vcycle_niterations", 6))
        p.mas = L.MAS[self.mas]
        p.ran_min = 0.01
        p.box_pad = 500.0
        return p

    def _ctx(self, mesh=None, grid_size=None) -> Context:
        if self.fft_plan is None:
            if mesh is None and grid_size is None:
                raise RuntimeError("call setup_fft(recon, mesh) first (recon.fft_plan is nothing)")
            setup_fft(self, mesh if mesh is not None else grid_size)
        gs = self.fft_plan.grid_size
        if self.box_size is None or self.box_min is None:
            raise RuntimeError("recon.box_size / recon.box_min are not set")
        return self.fft_plan.ctx.plan(gs, self.box_size, self.box_min)


@dataclass
class IterativeRecon(AbstractRecon):
    """src/recon.jl:2-16."""
    bias: float
    f: float
    smoothing_radius: float
    box_size: Optional[Sequence[float]] = None
    box_min: Optional[Sequence[float]] = None
    los: Optional[Sequence[float]] = None
    n_iter: int = 3
    beta: Optional[float] = None
    fft_plan: Optional[FFTPlan] = _dcfield(default=None, repr=False)
    result_cache: Optional[object] = _dcfield(default=None, repr=False)
    mas: str = "cic"        # extension; the reference hard-codes cic! (src/recon.jl:52)
    algorithm = L.ITERATIVE

    def __post_init__(self):
        if self.beta is None:
            self.beta = float(np.float32(np.float32(self.f) / np.float32(self.bias)))


@dataclass
class MultigridRecon(AbstractRecon):
    """src/recon.jl:18-33."""
    bias: float
    f: float
    smoothing_radius: float
    box_size: Optional[Sequence[float]] = None
    box_min: Optional[Sequence[float]] = None
    los: Optional[Sequence[float]] = None
    jacobi_damping_factor: float = 0.4
    jacobi_niterations: int = 5
    vcycle_niterations: int = 6
    beta: Optional[float] = None
    fft_plan: Optional[FFTPlan] = _dcfield(default=None, repr=False)
    result_cache: Optional[object] = _dcfield(default=None, repr=False)
    mas: str = "cic"
    algorithm = L.MULTIGRID

    def __post_init__(self):
        if self.beta is None:
            self.beta = float(np.float32(np.float32(self.f) / np.float32(self.bias)))


# --------------------------------------------------------------------------------------------
# utils.jl
# --------------------------------------------------------------------------------------------
def setup_fft(recon: AbstractRecon, field_or_grid_size):
    """setup_fft!(recon, field) src/recon.jl:35-37."""
    if isinstance(field_or_grid_size, torch.Tensor):
        gs = _chk_mesh(field_or_grid_size)
        dev = field_or_grid_size.device.index
    else:
        gs = tuple(int(v) for v in field_or_grid_size)
        dev = None
    recon.fft_plan = FFTPlan(Context.get(dev), gs)
    return recon.fft_plan


def k_vec(grid_size_xyz, box_size):
    """k_vec src/utils.jl:3-10 (host arrays; the device tables live in the plan)."""
    out = []
    for a in range(3):
        n = int(grid_size_xyz[a])
        fs = np.float32(2.0 * np.pi * n / np.float64(np.float32(box_size[a])))
        mult = np.float32(fs / np.float32(n))
        if a == 0:
            idx = np.arange(n // 2 + 1)
        else:
            nn = (n + 1) >> 1
            idx = np.concatenate([np.arange(nn), np.arange(nn - n, 0)])
        out.append(idx.astype(np.float32) * mult)
    return tuple(out)


def x_vec(grid_size_xyz, box_size, box_min):
    """x_vec src/utils.jl:21-25."""
    out = []
    for a in range(3):
        n = int(grid_size_xyz[a])
        cell = np.float32(np.float32(box_size[a]) / np.float32(n))
        start = np.float64(np.float32(box_min[a])) + 0.5 * np.float64(cell)
        out.append((start + np.arange(n, dtype=np.float64) * np.float64(cell)).astype(np.float32))
    return tuple(out)


def setup_box(pos_x, pos_y, pos_z, box_pad):
    """setup_box src/utils.jl:100-109 -> (box_size, box_min)."""
    n = _chk_vec(pos_x, pos_y, pos_z)
    ctx = Context.get(pos_x.device.index)
    bs, bm = (C.c_float * 3)(), (C.c_float * 3)()
    L.check(ctx.lib.baorec_setup_box_f32(ctx.handle, _ptr(pos_x), _ptr(pos_y), _ptr(pos_z), n, float(box_pad),
                                         bs, bm, _stream()))
    return np.array(list(bs), dtype=np.float32), np.array(list(bm), dtype=np.float32)


def _plan_for(mesh, box_size, box_min) -> Context:
    gs = _chk_mesh(mesh)
    return Context.get(mesh.device.index).plan(gs, box_size, box_min)


def _plan_grid(device_index, grid_size_xyz, box_size, box_min) -> Context:
    return Context.get(device_index).plan(grid_size_xyz, np.broadcast_to(np.asarray(box_size, np.float32), 3),
                                          np.broadcast_to(np.asarray(box_min, np.float32), 3))


def smooth(field, smoothing_radius, box_size, fft_plan=None, box_min=(0.0, 0.0, 0.0)):
    """smooth! src/utils.jl:85-96."""
    ctx = _plan_for(field, box_size, box_min)
    L.check(ctx.lib.baorec_smooth_f32(ctx.handle, _ptr(field), float(smoothing_radius), _stream()))
    return field


# --------------------------------------------------------------------------------------------
# mas.jl
# --------------------------------------------------------------------------------------------
def cic(rho, data_x, data_y, data_z, data_w, box_size, box_min, wrap=True, mas="cic"):
    """cic! src/mas.jl:100-107 (accumulates into rho; mutates positions when wrap)."""
    n = _chk_vec(data_x, data_y, data_z, data_w)
    ctx = _plan_for(rho, box_size, box_min)
    L.check(ctx.lib.baorec_cic_scatter_f32(ctx.handle, _ptr(rho), _ptr(data_x), _ptr(data_y), _ptr(data_z),
                                           _ptr(data_w), n, int(bool(wrap)), L.MAS[mas], _stream()))
    return rho


def read_cic(output, field, data_x, data_y, data_z, box_size, box_min, wrap=True, mas="cic"):
    """read_cic! src/mas.jl:320-327 (callers never pass wrap=false; always periodic)."""
    n = _chk_vec(data_x, data_y, data_z, output)
    ctx = _plan_for(field, box_size, box_min)
    L.check(ctx.lib.baorec_gather_f32(ctx.handle, _ptr(field), _ptr(data_x), _ptr(data_y), _ptr(data_z), n,
                                      _ptr(output), L.MAS[mas], _stream()))
    return output


def cic_cells(grid_size_xyz, data_x, data_y, data_z, box_size, box_min, wrap=True):
    """Parity probe: (i0, i1, w0, w1), each [3, n]."""
    n = _chk_vec(data_x, data_y, data_z)
    ctx = Context.get(data_x.device.index).plan(grid_size_xyz, box_size, box_min)
    dev = data_x.device
    i0 = torch.empty((3, n), dtype=torch.int32, device=dev)
    i1 = torch.empty_like(i0)
    w0 = torch.empty((3, n), dtype=torch.float32, device=dev)
    w1 = torch.empty_like(w0)
    L.check(ctx.lib.baorec_cic_cells_f32(ctx.handle, _ptr(data_x), _ptr(data_y), _ptr(data_z), n, int(bool(wrap)),
                                         _ptr(i0), _ptr(i1), _ptr(w0), _ptr(w1), _stream()))
    return i0, i1, w0, w1


def gather_cells(grid_size_xyz, data_x, data_y, data_z, box_size, box_min, gpu_formula=False):
    n = _chk_vec(data_x, data_y, data_z)
    ctx = Context.get(data_x.device.index).plan(grid_size_xyz, box_size, box_min)
    dev = data_x.device
    i0 = torch.empty((3, n), dtype=torch.int32, device=dev)
    i1 = torch.empty_like(i0)
    w0 = torch.empty((3, n), dtype=torch.float32, device=dev)
    w1 = torch.empty_like(w0)
    L.check(ctx.lib.baorec_gather_cells_f32(ctx.handle, _ptr(data_x), _ptr(data_y), _ptr(data_z), n,
                                            int(bool(gpu_formula)), _ptr(i0), _ptr(i1), _ptr(w0), _ptr(w1), _stream()))
    return i0, i1, w0, w1


# --------------------------------------------------------------------------------------------
# recon.jl
# --------------------------------------------------------------------------------------------
def _cat_args(data, rand):
    dx, dy, dz, dw = data
    n = _chk_vec(dx, dy, dz, dw)
    if rand is None or rand[0] is None:
        return (_ptr(dx), _ptr(dy), _ptr(dz), _ptr(dw), n, _ptr(None), _ptr(None), _ptr(None), _ptr(None), 0)
    rx, ry, rz, rw = rand
    nr = _chk_vec(rx, ry, rz, rw)
    return (_ptr(dx), _ptr(dy), _ptr(dz), _ptr(dw), n, _ptr(rx), _ptr(ry), _ptr(rz), _ptr(rw), nr)


def setup_overdensity(delta, recon, data_x, data_y, data_z, data_w, *rest, wrap=True):
    """setup_overdensity! src/recon.jl:42-57 (…, wrap) and :60-91 (…, rand_x, rand_y, rand_z, rand_w)."""
    rand = None
    if len(rest) == 1:
        wrap = rest[0]
    elif len(rest) >= 4:
        rand = rest[:4]
    _chk_mesh(delta)
    ctx = recon._ctx(delta)
    p = recon._params()
    L.check(ctx.lib.baorec_setup_overdensity_f32(ctx.handle, C.byref(p), _ptr(delta),
                                                 *_cat_args((data_x, data_y, data_z, data_w), rand),
                                                 int(bool(wrap)), _stream()))
    return delta


def iterate(delta_r, delta_s, kvec, it, beta, fft_plan: FFTPlan, r_hat=None, x_vec_=None,
            box_size=None, box_min=(0.0, 0.0, 0.0)):
    """iterate! src/iterative.jl:151-211.  kvec / x_vec_ are accepted for signature parity; the
    device tables of the plan (same formulas) are used."""
    gs = _chk_mesh(delta_r)
    ctx = fft_plan.ctx
    if box_size is not None:
        ctx.plan(gs, box_size, box_min)
    L.check(ctx.lib.baorec_iterate_f32(ctx.handle, _ptr(delta_r), _ptr(delta_s), int(it), float(beta),
                                       _los3(r_hat), _stream()))
    return delta_r


def reconstructed_overdensity(delta, recon: IterativeRecon, data_x, data_y, data_z, data_w, *rand):
    """reconstructed_overdensity! src/recon.jl:93-132."""
    _chk_mesh(delta)
    ctx = recon._ctx(delta)
    p = recon._params()
    L.check(ctx.lib.baorec_reconstructed_overdensity_f32(
        ctx.handle, C.byref(p), _ptr(delta), *_cat_args((data_x, data_y, data_z, data_w), rand or None), _stream()))
    return delta


def reconstructed_potential(phi, recon: MultigridRecon, data_x, data_y, data_z, data_w, *rand):
    """reconstructed_potential! src/recon.jl:184-212."""
    _chk_mesh(phi)
    ctx = recon._ctx(phi)
    p = recon._params()
    L.check(ctx.lib.baorec_reconstructed_potential_f32(
        ctx.handle, C.byref(p), _ptr(phi), *_cat_args((data_x, data_y, data_z, data_w), rand or None), _stream()))
    return phi


def _np_ptr(a):
    if a is None:
        return C.c_void_p(0)
    return C.c_void_p(a.ctypes.data)


def _chk_np(*arrs):
    n = None
    for a in arrs:
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 1 and a.flags.c_contiguous):
            raise TypeError("host particle arrays must be contiguous 1-D float32 numpy arrays")
        if n is None:
            n = a.shape[0]
        elif a.shape[0] != n:
            raise ValueError("particle arrays differ in length")
    return n or 0


def run(recon: AbstractRecon, grid_size, data_x, data_y, data_z, data_w, *rand, mesh_out=None):
    """run! src/recon.jl:134-180 (IterativeRecon) / :215-261 (MultigridRecon).
    grid_size = (nx, ny, nz).  Returns the result mesh (also in recon.result_cache)."""
    nx, ny, nz = (int(v) for v in grid_size)
    has_rand = len(rand) >= 4
    if _is_dev(data_x):
        if mesh_out is not None:      # reuse a caller-provided device mesh (zero-filled here, like run!)
            assert _chk_mesh(mesh_out) == (nx, ny, nz)
            mesh = mesh_out.zero_()
        else:
            mesh = torch.zeros((nz, ny, nx), dtype=torch.float32, device=data_x.device)
        if has_rand:
            recon.box_size, recon.box_min = setup_box(rand[0], rand[1], rand[2], 500.0)
        setup_fft(recon, mesh)
        if isinstance(recon, MultigridRecon):
            reconstructed_potential(mesh, recon, data_x, data_y, data_z, data_w, *rand[:4])
        else:
            reconstructed_overdensity(mesh, recon, data_x, data_y, data_z, data_w, *rand[:4])
        recon.result_cache = mesh
        recon._cache_version = mesh._version
        return mesh
    # host arrays: whole pipeline inside the library
    n = _chk_np(data_x, data_y, data_z, data_w)
    nr = _chk_np(*rand[:4]) if has_rand else 0
    setup_fft(recon, (nx, ny, nz))
    if recon.box_size is None:
        if not has_rand:
            raise RuntimeError("recon.box_size / recon.box_min are not set")
        recon.box_size, recon.box_min = (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)   # overridden by setup_box below
    ctx = recon._ctx(grid_size=(nx, ny, nz))
    p = recon._params()
    bs, bm = (C.c_float * 3)(), (C.c_float * 3)()
    r = [_np_ptr(a) for a in rand[:4]] if has_rand else [C.c_void_p(0)] * 4
    if mesh_out is not None:
        assert isinstance(mesh_out, np.ndarray) and mesh_out.dtype == np.float32 and mesh_out.shape == (nz, ny, nx)
    L.check(ctx.lib.baorec_run_host_f32(ctx.handle, C.byref(p), recon.algorithm, _np_ptr(data_x), _np_ptr(data_y),
                                        _np_ptr(data_z), _np_ptr(data_w), n, *r, nr, _np_ptr(mesh_out), bs, bm))
    if has_rand:
        recon.box_size = np.array(list(bs), dtype=np.float32)
        recon.box_min = np.array(list(bm), dtype=np.float32)
        ctx.plan_key = (nx, ny, nz, tuple(float(v) for v in bs), tuple(float(v) for v in bm))
    recon.result_cache = ("device-cache", ctx)
    return mesh_out if mesh_out is not None else recon.result_cache


def run_batch(recon: AbstractRecon, grid_size, catalogs, field="disp", positions=True, out=None):
    """Many catalogs per process (the reference's README: "one process, many reconstructions"; its examples loop over
    mocks): for every (x, y, z, w) in `catalogs` -- HOST float32 arrays, periodic box, no randoms -- `run!` followed by
    `reconstructed_positions` (or `read_shifts` with positions=False) of that catalog, with the PCIe transfers of
    neighbouring catalogs overlapping the solve (baorec_batch_host_f32).  Use pinned arrays (torch `pin_memory`
    + `.numpy()`, or baorec_host_alloc) for the overlap to happen.  `out`: optional list of 3-tuples of host arrays that
    receive the results.  Returns the list of (ox, oy, oz)."""
    nx, ny, nz = (int(v) for v in grid_size)
    K = len(catalogs)
    ns = [_chk_np(*cat) for cat in catalogs]
    if out is None:
        out = [tuple(np.empty(n, dtype=np.float32) for _ in range(3)) for n in ns]
    for n, o in zip(ns, out):
        assert _chk_np(*o) == n
    if recon.box_size is None:
        raise RuntimeError("recon.box_size / recon.box_min are not set")
    setup_fft(recon, (nx, ny, nz))
    ctx = recon._ctx(grid_size=(nx, ny, nz))
    p = recon._params()
    col = lambda arrs: (C.c_void_p * K)(*[a.ctypes.data for a in arrs])
    L.check(ctx.lib.baorec_batch_host_f32(
        ctx.handle, C.byref(p), recon.algorithm, K, col([c[0] for c in catalogs]), col([c[1] for c in catalogs]),
        col([c[2] for c in catalogs]), col([c[3] for c in catalogs]), (C.c_int64 * K)(*ns), L.FIELDS[field],
        0 if positions else 1, col([o[0] for o in out]), col([o[1] for o in out]), col([o[2] for o in out])))
    recon.result_cache = ("device-cache", ctx)
    return out


def run_batch_files(recon: AbstractRecon, grid_size, in_paths, out_paths=None, columns=(0, 1, 2, -1), delim=" ",
                    field="disp", positions=True, n_threads=0):
    """`run_batch` fed from catalog FILES and writing NPY files -- the reference's examples/simulation.jl:12-40 looped
    over mocks (baorec_batch_files_f32): `in_paths[i]` is a delimited text catalog or an .npy matrix, `columns` the
    0-based file columns of x, y, z and the weights (-1: weights of one), `out_paths[i]` (or None) receives what
    `npzwrite(fn, hcat(new_pos...))` stores.  A reader thread parses the next catalogs into pinned memory and a writer
    thread stores the previous results while the device reconstructs the current one.  Returns a dict: rows per catalog
    and the seconds spent reading, writing, waiting for the reader, and in total."""
    import os
    nx, ny, nz = (int(v) for v in grid_size)
    K = len(in_paths)
    if out_paths is not None and len(out_paths) != K:
        raise ValueError("one output path (or None) per catalog")
    if len(columns) != 4:
        raise ValueError("columns = (x, y, z, w); w = -1 for weights of one")
    if recon.box_size is None:
        raise RuntimeError("recon.box_size / recon.box_min are not set")
    setup_fft(recon, (nx, ny, nz))
    ctx = recon._ctx(grid_size=(nx, ny, nz))
    p = recon._params()
    ins = (C.c_char_p * K)(*[os.fsencode(q) for q in in_paths])
    outs = None if out_paths is None else (C.c_char_p * K)(*[None if q is None else os.fsencode(q) for q in out_paths])
    rows, secs = (C.c_int64 * K)(), (C.c_double * 4)()
    d = delim.encode()
    if len(d) != 1:
        raise ValueError("delim is one character")
    L.check(ctx.lib.baorec_batch_files_f32(ctx.handle, C.byref(p), recon.algorithm, K, ins, d,
                                           (C.c_int * 4)(*[int(c) for c in columns]), L.FIELDS[field], 0 if positions else 1,
                                           outs, int(n_threads), rows, secs))
    recon.result_cache = ("device-cache", ctx)
    return {"rows": list(rows), "read_s": secs[0], "write_s": secs[1], "wait_s": secs[2], "total_s": secs[3]}


def compute_displacements(mesh, data_x, data_y, data_z, recon: AbstractRecon):
    """compute_displacements src/iterative.jl:229-250 / src/multigrid.jl:781-799."""
    n = _chk_vec(data_x, data_y, data_z)
    ctx = recon._ctx(mesh)
    out = tuple(torch.empty_like(data_x) for _ in range(3))
    L.check(ctx.lib.baorec_compute_displacements_f32(ctx.handle, _ptr(mesh), recon.algorithm, _ptr(data_x),
                                                     _ptr(data_y), _ptr(data_z), n, _ptr(out[0]), _ptr(out[1]),
                                                     _ptr(out[2]), L.MAS[recon.mas], _stream()))
    return out


def displacement_meshes(mesh, recon: AbstractRecon):
    _chk_mesh(mesh)
    ctx = recon._ctx(mesh)
    out = tuple(torch.empty_like(mesh) for _ in range(3))
    L.check(ctx.lib.baorec_displacement_meshes_f32(ctx.handle, _ptr(mesh), recon.algorithm, _ptr(out[0]),
                                                   _ptr(out[1]), _ptr(out[2]), _stream()))
    return out


def _read(recon, data_x, data_y, data_z, mesh, field, positions, out=None):
    fld = L.FIELDS[field]
    if _is_dev(data_x):
        n = _chk_vec(data_x, data_y, data_z)
        _chk_mesh(mesh)
        ctx = recon._ctx(mesh)
        p = recon._params()
        if out is None:
            out = tuple(torch.empty_like(data_x) for _ in range(3))
        else:
            assert _chk_vec(*out) == n
        # recon.result_cache semantics (src/recon.jl:380-388): when the caller hands back the very mesh
        # run!/reconstructed_overdensity! produced, untouched (torch's in-place version counter), the
        # library reuses the delta_k it kept and skips the forward transform
        if mesh is recon.result_cache and getattr(recon, "_cache_version", None) == mesh._version:
            L.check(ctx.lib.baorec_read_result_cache_f32(
                ctx.handle, C.byref(p), recon.algorithm, _ptr(mesh), _ptr(data_x), _ptr(data_y), _ptr(data_z), n,
                fld, int(bool(positions)), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream()))
            return out
        fn = ctx.lib.baorec_reconstructed_positions_f32 if positions else ctx.lib.baorec_read_shifts_f32
        L.check(fn(ctx.handle, C.byref(p), recon.algorithm, _ptr(mesh), _ptr(data_x), _ptr(data_y), _ptr(data_z), n,
                   fld, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream()))
        return out
    n = _chk_np(data_x, data_y, data_z)
    ctx = recon._ctx()
    p = recon._params()
    out = tuple(np.empty(n, dtype=np.float32) for _ in range(3))
    hmesh = mesh if isinstance(mesh, np.ndarray) else None
    L.check(ctx.lib.baorec_read_host_f32(ctx.handle, C.byref(p), recon.algorithm, _np_ptr(hmesh), _np_ptr(data_x),
                                         _np_ptr(data_y), _np_ptr(data_z), n, fld, 0 if positions else 1,
                                         _np_ptr(out[0]), _np_ptr(out[1]), _np_ptr(out[2])))
    return out


def read_shifts(recon, data_x, data_y, data_z, mesh, field="disp", out=None):
    """read_shifts src/recon.jl:333-364; field in {"disp","rsd","sum"}.  `out` (3 device vectors)
    optionally receives the result instead of freshly allocated arrays."""
    return _read(recon, data_x, data_y, data_z, mesh, field, False, out)


def reconstructed_positions(recon, data_x, data_y, data_z, mesh=None, field="disp", out=None):
    """reconstructed_positions src/recon.jl:366-388 (mesh defaults to recon.result_cache)."""
    if mesh is None:
        mesh = recon.result_cache
    return _read(recon, data_x, data_y, data_z, mesh, field, True, out)


# --------------------------------------------------------------------------------------------
# multigrid.jl primitives
# --------------------------------------------------------------------------------------------
def _mg_ctx(v, box_size, box_min, plan_grid=None):
    gs = _chk_mesh(v)
    return Context.get(v.device.index).plan(plan_grid or gs, box_size, box_min), gs


def jacobi(v, f, x_vec_, box_size, box_min, beta, damping_factor, niterations, los=None, plan_grid=None):
    """jacobi! src/multigrid.jl:193-210."""
    ctx, (nx, ny, nz) = _mg_ctx(v, box_size, box_min, plan_grid)
    L.check(ctx.lib.baorec_mg_jacobi_f32(ctx.handle, _ptr(v), _ptr(f), nx, ny, nz, float(beta), float(damping_factor),
                                         int(niterations), _los3(los), _stream()))
    return v


def residual(r, v, f, x_vec_, box_size, box_min, beta, damping_factor=None, niterations=None, los=None,
             plan_grid=None):
    """residual! src/multigrid.jl:378-395."""
    ctx, (nx, ny, nz) = _mg_ctx(v, box_size, box_min, plan_grid)
    L.check(ctx.lib.baorec_mg_residual_f32(ctx.handle, _ptr(r), _ptr(v), _ptr(f), nx, ny, nz, float(beta),
                                           _los3(los), _stream()))
    return r


def reduce(v2h, v1h, box_size=(1.0, 1.0, 1.0), box_min=(0.0, 0.0, 0.0), plan_grid=None):
    """reduce! (restriction) src/multigrid.jl:641-651."""
    ctx, (nx, ny, nz) = _mg_ctx(v1h, box_size, box_min, plan_grid)
    L.check(ctx.lib.baorec_mg_restrict_f32(ctx.handle, _ptr(v2h), _ptr(v1h), nx, ny, nz, _stream()))
    return v2h


def prolong(v1h, v2h, box_size=(1.0, 1.0, 1.0), box_min=(0.0, 0.0, 0.0), plan_grid=None):
    """prolong! src/multigrid.jl:511-518."""
    ctx, (nx, ny, nz) = _mg_ctx(v1h, box_size, box_min, plan_grid)
    L.check(ctx.lib.baorec_mg_prolong_f32(ctx.handle, _ptr(v1h), _ptr(v2h), nx, ny, nz, _stream()))
    return v1h


def vcycle(v, f, box_size, box_min, beta, damping_factor, niterations, los=None):
    """vcycle! src/multigrid.jl:689-719."""
    ctx, _ = _mg_ctx(v, box_size, box_min)
    L.check(ctx.lib.baorec_mg_vcycle_f32(ctx.handle, _ptr(v), _ptr(f), float(beta), float(damping_factor),
                                         int(niterations), _los3(los), _stream()))
    return v


def fmg(f1h, v1h, box_size, box_min, beta, jacobi_damping_factor, jacobi_niterations, vcycle_niterations, los=None):
    """fmg src/multigrid.jl:722-752."""
    if v1h is None:
        v1h = torch.zeros_like(f1h)
    ctx, _ = _mg_ctx(v1h, box_size, box_min)
    L.check(ctx.lib.baorec_mg_fmg_f32(ctx.handle, _ptr(f1h), _ptr(v1h), float(beta), float(jacobi_damping_factor),
                                      int(jacobi_niterations), int(vcycle_niterations), _los3(los), _stream()))
    return v1h
