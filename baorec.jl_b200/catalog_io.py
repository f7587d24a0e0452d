"""Catalog files either side of the path: the host-side mirror of what the reference's example scripts do with
CSV.jl and NPZ.jl around `run!` / `reconstructed_positions` (examples/simulation.jl:12-13, 38-40;
examples/lightcone.jl:22-26) on top of the C ABI (baorec_text_catalog_*, baorec_npy_*, baorec_catalog_select_f32).

The parsing, the row selection and the NPY layout live in the library (threaded C++); this module only allocates
the SoA Float32 host arrays -- pinned when a GPU is present, so that the uploads of `run` / `run_batch` that follow
are asynchronous -- and hands their addresses over.  Nothing here needs a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import lib_loader as L


def _host_array(n: int, pinned: Optional[bool]) -> torch.Tensor:
    pin = torch.cuda.is_available() if pinned is None else bool(pinned)
    return torch.empty(max(int(n), 0), dtype=torch.float32, pin_memory=pin and n > 0)


def _ptrs(cols: Sequence[torch.Tensor]):
    for c in cols:
        if not (isinstance(c, torch.Tensor) and c.dtype == torch.float32 and c.dim() == 1 and c.is_contiguous()
                and c.device.type == "cpu"):
            raise TypeError("catalog columns are contiguous 1-D float32 host tensors")
    return (C.c_void_p * len(cols))(*[c.data_ptr() for c in cols])


def _delim(delim: str) -> bytes:
    b = delim.encode()
    if len(b) != 1:
        raise ValueError("delim is one character")
    return b


def scan_text_catalog(path, delim: str = " ", n_threads: int = 0) -> Tuple[int, int]:
    """(rows, columns) of a delimited text catalog; blank lines and `#` comments are not rows."""
    n, k = C.c_int64(0), C.c_int(0)
    L.check(L.load().baorec_text_catalog_scan(os.fsencode(path), _delim(delim), C.byref(n), C.byref(k), int(n_threads)))
    return n.value, k.value


def read_text_catalog(path, columns: Sequence[int], delim: str = " ", pinned: Optional[bool] = None,
                      n_threads: int = 0) -> Tuple[torch.Tensor, ...]:
    """`CSV.File(path, delim = ' ', ignorerepeated = true, header = [...], types = [Float32 ...])` followed by picking
    columns (examples/simulation.jl:12-15): the 0-based file columns `columns` as Float32 host arrays."""
    columns = [int(c) for c in columns]
    n, _ = scan_text_catalog(path, delim, n_threads)
    out = [_host_array(n, pinned) for _ in columns]
    got = C.c_int64(0)
    L.check(L.load().baorec_text_catalog_read_f32(os.fsencode(path), _delim(delim), len(columns),
                                                  (C.c_int * len(columns))(*columns), _ptrs(out), n, C.byref(got),
                                                  int(n_threads)))
    assert got.value == n
    return tuple(out)


def npy_info(path) -> dict:
    """dtype ('f4' / 'f8'), memory order and (rows, columns) of an NPY file."""
    dt, fo, n, k = C.c_int(0), C.c_int(0), C.c_int64(0), C.c_int64(0)
    L.check(L.load().baorec_npy_info(os.fsencode(path), C.byref(dt), C.byref(fo), C.byref(n), C.byref(k)))
    return {"dtype": "f4" if dt.value == L.DTYPE_F32 else "f8", "fortran_order": bool(fo.value), "rows": n.value,
            "columns": k.value}


def read_npy_catalog(path, columns: Optional[Sequence[int]] = None, pinned: Optional[bool] = None,
                     n_threads: int = 0) -> Tuple[torch.Tensor, ...]:
    """`npzread(path)` of an (N, K) matrix (or an (N,) vector), split into Float32 columns."""
    info = npy_info(path)
    columns = list(range(info["columns"])) if columns is None else [int(c) for c in columns]
    n = info["rows"]
    out = [_host_array(n, pinned) for _ in columns]
    got = C.c_int64(0)
    L.check(L.load().baorec_npy_read_columns_f32(os.fsencode(path), len(columns), (C.c_int * len(columns))(*columns),
                                                 _ptrs(out), n, C.byref(got), int(n_threads)))
    return tuple(out)


def write_npy(path, *cols: torch.Tensor) -> None:
    """`npzwrite(path, hcat(cols...))` (examples/simulation.jl:38-40): numpy.load returns the (N, len(cols)) matrix."""
    if not cols:
        raise ValueError("at least one column")
    n = cols[0].numel()
    if any(c.numel() != n for c in cols):
        raise ValueError("columns of different lengths")
    L.check(L.load().baorec_npy_write_columns_f32(os.fsencode(path), len(cols), _ptrs(cols), n))


def select_rows(cols: Sequence[torch.Tensor], key: int, lo: float, hi: float, n_threads: int = 0) -> Tuple[torch.Tensor, ...]:
    """`cat[map(z -> ((z > lo) & (z < hi)), cat.z), :]` (examples/lightcone.jl:25-26), in place: the rows with
    lo < cols[key] < hi are moved to the front of every column; returns the views of the kept part."""
    n = cols[0].numel()
    if any(c.numel() != n for c in cols):
        raise ValueError("columns of different lengths")
    kept = C.c_int64(0)
    L.check(L.load().baorec_catalog_select_f32(len(cols), _ptrs(cols), n, int(key), float(lo), float(hi), C.byref(kept),
                                               int(n_threads)))
    return tuple(c[:kept.value] for c in cols)
