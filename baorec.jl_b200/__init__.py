"""baorec.jl_b200 -- B200-native engine behind BAOrec.jl's device hot path.

The directory name carries a dot, so load it with `__graft_entry__.load_package()`
(registers it as the module `baorec_b200`).  Contents:
  csrc/        CUDA kernels (sm_100a) + the C ABI (include/baorec_b200.h)
  lib/         built libbaorec_b200.so (git-ignored)
  lib_loader   ctypes binding (fails loudly if the library is missing)
  host         mirror of the reference's Julia API on top of the C ABI
  catalog      cosmology tables + sky <-> Cartesian, FKP weights, periodic re-wrap (src/cosmo.jl, examples/lightcone.jl)
  catalog_io   text / NPY catalog readers and writers, row selection (what the examples do with CSV.jl / NPZ.jl)
  julia/       the ccall shim a BAOrec.jl maintainer would add
"""
from . import lib_loader
from .lib_loader import BaorecError, OutOfBoxError, OutOfRangeError, CatalogIOError
from .host import (Context, FFTPlan, IterativeRecon, MultigridRecon, setup_fft, k_vec, x_vec, setup_box, smooth, cic,
                   read_cic, cic_cells, gather_cells, setup_overdensity, iterate, reconstructed_overdensity,
                   reconstructed_potential, run, run_batch, run_batch_files, compute_displacements, displacement_meshes, read_shifts,
                   reconstructed_positions, jacobi, residual, reduce, prolong, vcycle, fmg)

from . import dist
from . import catalog_io
from .catalog_io import (scan_text_catalog, read_text_catalog, npy_info, read_npy_catalog, write_npy, select_rows)
from .catalog import Cosmology, DESICosmology, E, H, comoving_distance, comoving_distance_interp, redshift_interp, sky_to_cartesian, cartesian_to_sky, fkp_weights, wrap_positions, power_multipoles, interlace_positions, compute_auto_box

lib_loader.load()   # no library -> ImportError; there is no fallback path
