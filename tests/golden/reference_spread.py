#!/usr/bin/env python
"""Derives the tolerance yardstick the reference itself provides: BAOrec.jl ships the first ten
reconstructed positions of its CPU and of its GPU run on the same (unshipped) DESI mock
(test/bchmk_{cpu,gpu}_{iter,mgrid}.csv).  Their inputs are not available, so they cannot serve as
known-answer vectors, but the CPU-vs-GPU spread fixes the scale of "agreement" the reference's own
two code paths reach.  Run in the build container (needs /root/reference); the derived numbers are
committed as REFERENCE_SPREAD.json so that nothing reads /root/reference at test time."""
import json
from pathlib import Path

import numpy as np

REF = Path("/root/reference/test")
out = {}
for alg in ("iter", "mgrid"):
    cpu = np.loadtxt(REF / f"bchmk_cpu_{alg}.csv", delimiter=",")
    gpu = np.loadtxt(REF / f"bchmk_gpu_{alg}.csv", delimiter=",")
    d = np.abs(cpu - gpu)
    out[alg] = {"rows": int(cpu.shape[0]), "max_abs_diff_Mpc_h": float(d.max()), "mean_abs_diff_Mpc_h": float(d.mean()),
                "position_scale_Mpc_h": float(np.abs(cpu).max())}
(Path(__file__).resolve().parent / "REFERENCE_SPREAD.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
