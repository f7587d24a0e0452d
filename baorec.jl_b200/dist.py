"""Multi-GPU host side: one process per GPU (torchrun), z-slab decomposition.

torch.distributed is the plumbing (rendezvous, the NCCL unique id broadcast and the particle
redistribution); the data path -- slab FFT transposes, ghost/halo planes -- runs inside the
library over its own NCCL communicator."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib_loader as L
from .host import Context, FFTPlan, _chk_vec, _ptr, _stream


# ---- host logic (pure numpy / torch.distributed; runs on CPU with gloo) --------------------------
def owner_of_z(z, box_min_z, box_size_z, nz, world, box_min_0=None, box_size_0=None):
    """Owner rank of each particle = slab of the base plane cic! deposits into (src/mas.jl:8-30),
    evaluated with the reference's Float32 arithmetic; -1 for out-of-box particles."""
    z = np.asarray(z, np.float32)
    mn, Lz = np.float32(box_min_z), np.float32(box_size_z)
    mn0 = mn if box_min_0 is None else np.float32(box_min_0)
    L0 = Lz if box_size_0 is None else np.float32(box_size_0)
    zw = np.where((z - mn0) > L0, (z - L0).astype(np.float32), z).astype(np.float32)   # upper-face wrap (quirk: axis-1 box)
    g = (((zw - mn) * np.float32(nz)).astype(np.float32) / Lz).astype(np.float32) + np.float32(1)
    c0 = np.floor(g).astype(np.int64)
    c0 = np.where(c0 == nz + 1, 1, c0)
    ok = (g >= 1) & (c0 >= 1) & (c0 <= nz)
    return np.where(ok, (c0 - 1) // (nz // world), -1).astype(np.int32)


def shard_catalog(x, y, z, w, box_size, box_min, nz, world):
    """Splits host arrays into per-rank lists by slab ownership (used before the upload)."""
    own = owner_of_z(z, box_min[2], box_size[2], nz, world, box_min[0], box_size[0])
    if (own < 0).any():
        raise L.OutOfBoxError(L.ERR_OUT_OF_BOX, f"{int((own < 0).sum())} particle(s) outside the box")
    return [tuple(a[own == r] for a in (x, y, z, w)) for r in range(world)]


def exchange_catalog(x, y, z, w, box_size, box_min, nz, group=None):
    """Every rank holds an arbitrary part of the catalog (torch tensors on CPU or GPU); after the
    call it holds exactly the particles of its own slab.  One all_to_all_single per column."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    own = torch.from_numpy(owner_of_z(z.cpu().numpy(), box_min[2], box_size[2], nz, world, box_min[0], box_size[0]))
    if (own < 0).any():
        raise L.OutOfBoxError(L.ERR_OUT_OF_BOX, "particle(s) outside the box")
    order = torch.argsort(own, stable=True).to(x.device)
    counts = torch.bincount(own, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(counts)
    cdev = counts.to(x.device)
    rdev = torch.empty_like(cdev)
    dist.all_to_all_single(rdev, cdev, group=group)
    recv_counts = rdev.cpu()
    outs = []
    for col in (x, y, z, w):
        send = col[order].contiguous()
        recv = torch.empty(int(recv_counts.sum()), dtype=col.dtype, device=col.device)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts.tolist(),
                               input_split_sizes=counts.tolist(), group=group)
        outs.append(recv)
    return tuple(outs)


# ---- library side ---------------------------------------------------------------------------------
def init_comm(ctx: Context = None):
    """Creates the library's NCCL communicator from torch.distributed's ranks."""
    import torch.distributed as dist
    ctx = ctx or Context.get()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        L.check(ctx.lib.baorec_comm_init(ctx.handle, 0, 1, None))
        return ctx
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        L.check(ctx.lib.baorec_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    dev = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    uid = uid.to(dev)
    dist.broadcast(uid, 0)
    raw = bytes(uid.cpu().tolist())
    L.check(ctx.lib.baorec_comm_init(ctx.handle, rank, world, C.create_string_buffer(raw, 128)))
    return ctx


def _enable_p2p(ctx: Context):
    """(Optional, plan(..., p2p=True).)  Measured on 8 x B200 at 1024^3 it does not beat the grouped
    ncclSend/ncclRecv path (22.0 vs 20.4 ms per reconstruction: the per-transpose barrier absorbs the
    rank skew that NCCL hides), so it is off by default.
    Maps every rank's receive buffers into this process (CUDA IPC) so that the pack / transpose
    kernels store straight into peer memory over NVLink instead of staging + ncclSend/Recv."""
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = (C.c_ubyte * 128)()
    L.check(ctx.lib.baorec_dist_ipc_export(ctx.handle, mine))
    t = torch.tensor(list(mine), dtype=torch.uint8, device=torch.device("cuda", ctx.device))
    allh = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allh, t)
    raw = b"".join(bytes(h.cpu().tolist()) for h in allh)
    L.check(ctx.lib.baorec_dist_ipc_open(ctx.handle, C.create_string_buffer(raw, len(raw)), world))


def plan(ctx: Context, grid_size, box_size, box_min, p2p=False):
    import torch.distributed as dist
    nx, ny, nz = (int(v) for v in grid_size)
    key = ("dist", nx, ny, nz, tuple(float(np.float32(v)) for v in box_size),
           tuple(float(np.float32(v)) for v in box_min))
    if ctx.plan_key == key:
        return ctx
    shape_changed = not (isinstance(ctx.plan_key, tuple) and ctx.plan_key[:4] == key[:4])
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if shape_changed and multi:
        # the receive buffers are about to be re-allocated: every rank drops its mappings of the
        # peers' buffers first, and nobody frees before everybody has done so
        L.check(ctx.lib.baorec_dist_ipc_close(ctx.handle))
        dist.barrier()
    L.check(ctx.lib.baorec_plan_dist(ctx.handle, nx, ny, nz, L.f3(box_size), L.f3(box_min)))
    if p2p and shape_changed and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 \
            and dist.get_backend() == "nccl":
        _enable_p2p(ctx)
    ctx.plan_key = ("dist", nx, ny, nz, tuple(float(np.float32(v)) for v in box_size),
                    tuple(float(np.float32(v)) for v in box_min))
    return ctx


def slab_range(ctx: Context):
    a, b = C.c_int(), C.c_int()
    L.check(ctx.lib.baorec_slab_range(ctx.handle, C.byref(a), C.byref(b)))
    return a.value, b.value


def slab_owner(ctx: Context, z):
    n = _chk_vec(z)
    out = torch.empty(n, dtype=torch.int32, device=z.device)
    L.check(ctx.lib.baorec_slab_owner_f32(ctx.handle, _ptr(z), n, _ptr(out), _stream()))
    return out


def dist_r2c(ctx: Context, slab):
    nzl, ny, nx = slab.shape
    z_lo, nz_loc = slab_range(ctx)
    assert nzl == nz_loc
    world = ctx.plan_key[3] // nz_loc
    out = torch.empty((ny // world, nx // 2 + 1, ctx.plan_key[3]), dtype=torch.complex64, device=slab.device)
    L.check(ctx.lib.baorec_dist_r2c_f32(ctx.handle, _ptr(slab), _ptr(out), _stream()))
    return out


def dist_c2r(ctx: Context, kslab_t, out_slab):
    L.check(ctx.lib.baorec_dist_c2r_f32(ctx.handle, _ptr(kslab_t), _ptr(out_slab), _stream()))
    return out_slab


def setup_box_dist(pos_x, pos_y, pos_z, box_pad, group=None):
    """setup_box (src/utils.jl:100-109) of a catalog spread over the ranks: local min/max per axis,
    all-reduced, then the reference's Float32 arithmetic.  Works on CPU (gloo) and CUDA tensors."""
    import torch.distributed as dist
    big = torch.finfo(torch.float32).max
    lo = torch.stack([p.min() if p.numel() else p.new_tensor(big) for p in (pos_x, pos_y, pos_z)])
    hi = torch.stack([p.max() if p.numel() else p.new_tensor(-big) for p in (pos_x, pos_y, pos_z)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    half = np.float32(box_pad) / np.float32(2)
    mn = lo.cpu().numpy().astype(np.float32) - half
    mx = hi.cpu().numpy().astype(np.float32) + half
    return np.full(3, (mx - mn).max(), dtype=np.float32), mn.astype(np.float32)


def run_dist(recon, grid_size, data_x, data_y, data_z, data_w, rand_x=None, rand_y=None, rand_z=None, rand_w=None,
             ctx: Context = None):
    """run!(recon, grid_size, data...[, rand...]) over all ranks: pass this rank's slab particles
    (exchange_catalog / shard_catalog), get this rank's slab of the result mesh, shape
    (nz_loc, ny, nx): delta_r for IterativeRecon, phi for MultigridRecon.  With randoms the caller
    sets recon.box_size / box_min from setup_box_dist(rand..., 500) BEFORE sharding (run! does it
    from the full randoms catalog, src/recon.jl:172,253)."""
    n = _chk_vec(data_x, data_y, data_z, data_w)
    has_ran = rand_x is not None
    nr = _chk_vec(rand_x, rand_y, rand_z, rand_w) if has_ran else 0
    ctx = ctx or Context.get(data_x.device.index)
    plan(ctx, grid_size, recon.box_size, recon.box_min)
    recon.fft_plan = FFTPlan(ctx, tuple(int(v) for v in grid_size))
    _, nzl = slab_range(ctx)
    nx, ny, nz = (int(v) for v in grid_size)
    mesh = torch.empty((nzl, ny, nx), dtype=torch.float32, device=data_x.device)
    p = recon._params()
    L.check(ctx.lib.baorec_run_dist_f32(ctx.handle, C.byref(p), recon.algorithm, _ptr(data_x), _ptr(data_y),
                                        _ptr(data_z), _ptr(data_w), n, _ptr(rand_x), _ptr(rand_y), _ptr(rand_z),
                                        _ptr(rand_w), nr, int(has_ran), _ptr(mesh), _stream()))
    recon.result_cache = mesh
    return mesh


def read_shifts_dist(recon, data_x, data_y, data_z, field="disp", positions=False):
    """read_shifts / reconstructed_positions for this rank's slab particles against the result of
    the last run_dist (IterativeRecon: delta_k; MultigridRecon: phi_k kept by the library)."""
    n = _chk_vec(data_x, data_y, data_z)
    ctx = recon.fft_plan.ctx
    p = recon._params()
    out = tuple(torch.empty_like(data_x) for _ in range(3))
    L.check(ctx.lib.baorec_read_shifts_dist_f32(ctx.handle, C.byref(p), _ptr(data_x), _ptr(data_y), _ptr(data_z), n,
                                                L.FIELDS[field], int(bool(positions)), _ptr(out[0]), _ptr(out[1]),
                                                _ptr(out[2]), _stream()))
    return out
