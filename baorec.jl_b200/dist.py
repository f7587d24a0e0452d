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


def exchange_catalog_host(x, y, z, w, box_size, box_min, nz, group=None):
    """The routing of baorec_shard_catalog_f32 restated with torch.distributed collectives on CPU (or CUDA) tensors --
    the gloo-testable statement of the scheme, not the product path: owner per particle, P x (P+1) count matrix
    all-gathered (last column = out-of-box counts, so that EVERY rank raises when ANY rank holds an out-of-box
    particle instead of leaving the others blocked in the collective), columns grouped by destination, one
    all_to_all_single per column.  Returns (x, y, z, w of this rank's slab, route) with route for unshard_host."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    own = torch.from_numpy(owner_of_z(z.cpu().numpy(), box_min[2], box_size[2], nz, world, box_min[0], box_size[0])).long()
    row = torch.bincount(torch.where(own < 0, torch.full_like(own, world), own), minlength=world + 1)[: world + 1].to(torch.int64)
    rows = [torch.empty_like(row) for _ in range(world)]
    dist.all_gather(rows, row.to(x.device), group=group)
    mat = torch.stack([r.cpu() for r in rows])                       # mat[s, d] = particles rank s sends to rank d
    if int(mat[:, world].sum()) > 0:
        raise L.OutOfBoxError(L.ERR_OUT_OF_BOX, f"{int(mat[:, world].sum())} particle(s) outside the box over all ranks")
    rank = dist.get_rank(group)
    order = torch.argsort(own, stable=True)
    send_cnt, recv_cnt = mat[rank, :world].tolist(), mat[:world, rank].tolist()
    outs = []
    for col in (x, y, z, w):
        send = col[order.to(col.device)].contiguous()
        recv = torch.empty(int(sum(recv_cnt)), dtype=col.dtype, device=col.device)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_cnt, input_split_sizes=send_cnt, group=group)
        outs.append(recv)
    slot_of = torch.empty_like(order)
    slot_of[order] = torch.arange(len(order))                        # original index -> position in the send order
    return (*outs, {"slot_of": slot_of, "send_cnt": send_cnt, "recv_cnt": recv_cnt, "group": group})


def unshard_host(route, *cols):
    """The way back of exchange_catalog_host: columns in slab order -> the order of the arrays that were sharded."""
    import torch.distributed as dist
    outs = []
    for col in cols:
        back = torch.empty(int(sum(route["send_cnt"])), dtype=col.dtype, device=col.device)
        dist.all_to_all_single(back, col.contiguous(), output_split_sizes=route["send_cnt"],
                               input_split_sizes=route["recv_cnt"], group=route["group"])
        outs.append(back[route["slot_of"].to(col.device)])
    return tuple(outs)


def exchange_catalog(x, y, z, w, ctx: Context = None, slot=0):
    """Every rank holds an arbitrary part of the catalog (CUDA tensors); returns this rank's slab particles as four
    tensors that VIEW the library's buffers (valid until the next call with the same slot) -- baorec_shard_catalog_f32:
    device counting sort by owner + one grouped ncclSend/ncclRecv per column.  plan() must have been called.
    Out-of-box particles on any rank raise OutOfBoxError on every rank."""
    n = _chk_vec(x, y, z, w)
    ctx = ctx or Context.get(x.device.index)
    ptrs = [C.c_void_p() for _ in range(4)]
    nl = C.c_int64()
    L.check(ctx.lib.baorec_shard_catalog_f32(ctx.handle, int(slot), _ptr(x), _ptr(y), _ptr(z), _ptr(w), n,
                                             *[C.byref(q) for q in ptrs], C.byref(nl), _stream()))
    return tuple(_view(q.value, nl.value, x.device) for q in ptrs)


def unshard(ctx: Context, cols, like, slot=0):
    """Per-particle columns in slab order (up to three tensors) -> the order of the arrays that were sharded."""
    cols = list(cols) + [None] * (3 - len(cols))
    outs = [torch.empty_like(like) if c is not None else None for c in cols]
    L.check(ctx.lib.baorec_unshard_f32(ctx.handle, int(slot), *[_ptr(c) for c in cols], *[_ptr(o) for o in outs], _stream()))
    return tuple(o for o in outs if o is not None)


def _view(ptr, n, device):
    """A float32 CUDA tensor over library-owned device memory (no copy, no ownership)."""
    if n == 0:
        return torch.empty(0, dtype=torch.float32, device=device)
    class _Ext:                                   # __cuda_array_interface__ v2
        pass
    e = _Ext()
    e.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(e, device=device)


def reconstruct_dist(recon, grid_size, data_x, data_y, data_z, data_w, rand=None, field="sum", positions=False, ctx: Context = None,
                     out=None):
    """run! + read_shifts / reconstructed_positions from this rank's (arbitrary) part of the catalog: sharding by slab,
    the slab reconstruction and the way back happen inside the library (baorec_reconstruct_dist_f32); CUDA tensors in,
    three CUDA tensors in the order of data_x out.  NumPy arrays go through the host entry point instead."""
    p = recon._params()
    if isinstance(data_x, np.ndarray):
        ctx = ctx or Context.get()
        plan(ctx, grid_size, recon.box_size, recon.box_min)
        recon.fft_plan = FFTPlan(ctx, tuple(int(v) for v in grid_size))
        assert rand is None, "the host entry point takes no randoms"
        outs = tuple(out) if out is not None else tuple(np.empty_like(data_x) for _ in range(3))
        L.check(ctx.lib.baorec_reconstruct_dist_host_f32(ctx.handle, C.byref(p), recon.algorithm, *[a.ctypes.data for a in (data_x, data_y, data_z, data_w)],
                                                         len(data_x), L.FIELDS[field], int(bool(positions)), *[o.ctypes.data for o in outs]))
        return outs
    n = _chk_vec(data_x, data_y, data_z, data_w)
    ctx = ctx or Context.get(data_x.device.index)
    plan(ctx, grid_size, recon.box_size, recon.box_min)
    recon.fft_plan = FFTPlan(ctx, tuple(int(v) for v in grid_size))
    nr = _chk_vec(*rand) if rand is not None else 0
    r = list(rand) if rand is not None else [None] * 4
    outs = tuple(out) if out is not None else tuple(torch.empty_like(data_x) for _ in range(3))
    L.check(ctx.lib.baorec_reconstruct_dist_f32(ctx.handle, C.byref(p), recon.algorithm, _ptr(data_x), _ptr(data_y), _ptr(data_z),
                                                _ptr(data_w), n, *[_ptr(a) for a in r], nr, int(rand is not None), L.FIELDS[field],
                                                int(bool(positions)), *[_ptr(o) for o in outs], _stream()))
    return outs


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


def _map_peers(ctx: Context):
    """Maps every rank's three receive buffers, its halo inbox and its flag block into this process (CUDA IPC) -- what the peer-copy
    exchange of the slab transforms (csrc/dist.cu, option "dist_exchange" = 1, the default) writes into: one strided
    copy-engine copy per peer and chunk over NVLink plus a flag store, no pack / transpose kernels, no NCCL.
    (Round 1's version of this exchange ended every transpose with a barrier, which exposed the rank skew that NCCL's
    grouped send/recv hides -- 22.0 vs 20.4 ms at 8 GPUs; per-peer sequence flags removed the barrier.)"""
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = (C.c_ubyte * 320)()
    L.check(ctx.lib.baorec_dist_ipc_export(ctx.handle, mine))      # also resets this rank's flags (new generation)
    t = torch.tensor(list(mine), dtype=torch.uint8, device=torch.device("cuda", ctx.device))
    allh = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allh, t)                                       # the synchronisation point the reset relies on
    raw = b"".join(bytes(h.cpu().tolist()) for h in allh)
    L.check(ctx.lib.baorec_dist_ipc_open(ctx.handle, C.create_string_buffer(raw, len(raw)), world))
    dist.barrier()                                                 # nobody starts an exchange before everybody has mapped


def plan(ctx: Context, grid_size, box_size, box_min, exchange=None):
    """baorec_plan_dist + (several ranks, NCCL backend) the IPC mapping of the peers' receive buffers.
    exchange: "peer" (default) or "nccl" (round 1's pack / transpose kernels + grouped ncclSend/ncclRecv)."""
    import torch.distributed as dist
    nx, ny, nz = (int(v) for v in grid_size)
    if exchange is not None:
        ctx.set_option("dist_exchange", 1 if exchange == "peer" else 0)
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
    if shape_changed and multi and dist.get_backend() == "nccl":
        _map_peers(ctx)
    ctx.plan_key = key
    return ctx


def peer_exchange(ctx: Context) -> bool:
    """True when the slab transforms run the peer-copy exchange (k space is K[z][yl][x]), False for NCCL (T[yl][x][z])."""
    return bool(ctx.lib.baorec_dist_exchange_mode(ctx.handle))


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
    """Distributed R2C of this rank's z slab.  Returns this rank's y slab of k space: K[z][yl][x] with the peer-copy
    exchange (peer_exchange(ctx)), T[yl][x][z] with the NCCL exchange."""
    nzl, ny, nx = slab.shape
    z_lo, nz_loc = slab_range(ctx)
    assert nzl == nz_loc
    nz = ctx.plan_key[3]
    world = nz // nz_loc
    shape = (nz, ny // world, nx // 2 + 1) if peer_exchange(ctx) else (ny // world, nx // 2 + 1, nz)
    out = torch.empty(shape, dtype=torch.complex64, device=slab.device)
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
