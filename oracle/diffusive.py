"""TEST INFRASTRUCTURE -- ctypes front end of oracle/libdiffusive_oracle.so (diffusive_oracle.c, `parity unpinned`).

`compute_diffusive(diff_inputs)` takes the dict the reference's diffusive.compute_diffusive takes
(/root/reference/src/troute-routing/troute/routing/fast_reach/diffusive.pyx:126-202) and returns (out_q, out_elv, out_depth),
each (ntss_ev_g, mxncomp_g, nrch_g) float64.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffusive_oracle.so")
_lib = None

# c_diffnw's argument list (pydiffusive.f90:8-52): (name, kind) with kind in 'i' int scalar, 'I' int array, 'D' double array
ARGS = [("timestep_ar_g", "D"), ("nts_ql_g", "i"), ("nts_ub_g", "i"), ("nts_db_g", "i"), ("ntss_ev_g", "i"),
        ("nts_qtrib_g", "i"), ("nts_da_g", "i"), ("mxncomp_g", "i"), ("nrch_g", "i"), ("z_ar_g", "D"), ("bo_ar_g", "D"),
        ("traps_ar_g", "D"), ("tw_ar_g", "D"), ("twcc_ar_g", "D"), ("mann_ar_g", "D"), ("manncc_ar_g", "D"), ("so_ar_g", "D"),
        ("dx_ar_g", "D"), ("iniq", "D"), ("frnw_col", "i"), ("frnw_g", "I"), ("qlat_g", "D"), ("ubcd_g", "D"),
        ("dbcd_g", "D"), ("qtrib_g", "D"), ("paradim", "i"), ("para_ar_g", "D"), ("mxnbathy_g", "i"), ("x_bathy_g", "D"),
        ("z_bathy_g", "D"), ("mann_bathy_g", "D"), ("size_bathy_g", "I"), ("usgs_da_g", "D"), ("usgs_da_reach_g", "I"),
        ("rdx_ar_g", "D"), ("cwnrow_g", "i"), ("cwncol_g", "i"), ("crosswalk_g", "D"), ("z_thalweg_g", "D")]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "diffusive_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-B", "libdiffusive_oracle.so"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.trt_oracle_diffnw.restype = C.c_int
    return _lib


def marshal(diff_inputs):
    """-> (ctypes argument list in c_diffnw order, keep-alive list, (ntss_ev, mxncomp, nrch)); arrays as Fortran-ordered
    copies, exactly what the Cython wrapper does with np.asfortranarray (diffusive.pyx:131-171)."""
    keep, args = [], []
    for name, kind in ARGS:
        v = diff_inputs[name]
        if kind == "i":
            c = C.c_int(int(v)); keep.append(c); args.append(C.byref(c))
        else:
            a = np.asfortranarray(v, dtype=np.int32 if kind == "I" else np.float64)
            if a.size == 0:
                a = np.zeros(1, dtype=a.dtype)
            keep.append(a); args.append(a.ctypes.data_as(C.c_void_p))
    shape = (int(diff_inputs["ntss_ev_g"]), int(diff_inputs["mxncomp_g"]), int(diff_inputs["nrch_g"]))
    return args, keep, shape


POW_LIBM = 0   # x**y = platform pow (what a gfortran build of the reference computes)
POW_DET = 1    # x**y = trt_pow64_det (include/trt_detmath64.h), the bit-specified pow of the CUDA path


def compute_diffusive(diff_inputs, pow_mode=POW_LIBM):
    args, keep, shape = marshal(diff_inputs)
    lib().trt_oracle_diffnw_pow_mode(int(pow_mode))
    outs = [np.zeros(shape, dtype=np.float64, order="F") for _ in range(3)]
    rc = lib().trt_oracle_diffnw(*args, *[o.ctypes.data_as(C.c_void_p) for o in outs])
    if rc != 0:
        raise RuntimeError(f"trt_oracle_diffnw failed with status {rc}")
    return tuple(np.ascontiguousarray(o) for o in outs)
