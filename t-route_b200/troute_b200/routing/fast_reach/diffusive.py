"""Host mirror of troute.routing.fast_reach.diffusive (the Cython wrapper of the Fortran diffusive-wave solver).

    compute_diffusive(diff_inputs) -> (out_q, out_elv, out_depth)

takes the dict diffusive_input_data_v02 builds (/root/reference/src/troute-routing/troute/routing/diffusive_utils_v02.py:1104-1153)
and returns three (ntss_ev_g, mxncomp_g, nrch_g) float64 arrays, exactly like the reference's
fast_reach/diffusive.pyx:126-202 -- but the 42-argument call goes to trt_c_diffnw in libtroute_b200.so (CUDA, one CTA per
domain) instead of the Fortran c_diffnw.  compute_diffusive_batch routes several tailwater domains in one launch; the
reference loops over them (compute.py:1764).  No CPU fallback: without the CUDA library or a device these raise.
"""
import ctypes as C

import numpy as np

from ... import _lib

# c_diffnw's argument list (pydiffusive.f90:8-52): 'i' int scalar, 'I' int array, 'D' double array
_ARGS = [("timestep_ar_g", "D"), ("nts_ql_g", "i"), ("nts_ub_g", "i"), ("nts_db_g", "i"), ("ntss_ev_g", "i"),
         ("nts_qtrib_g", "i"), ("nts_da_g", "i"), ("mxncomp_g", "i"), ("nrch_g", "i"), ("z_ar_g", "D"), ("bo_ar_g", "D"),
         ("traps_ar_g", "D"), ("tw_ar_g", "D"), ("twcc_ar_g", "D"), ("mann_ar_g", "D"), ("manncc_ar_g", "D"), ("so_ar_g", "D"),
         ("dx_ar_g", "D"), ("iniq", "D"), ("frnw_col", "i"), ("frnw_g", "I"), ("qlat_g", "D"), ("ubcd_g", "D"),
         ("dbcd_g", "D"), ("qtrib_g", "D"), ("paradim", "i"), ("para_ar_g", "D"), ("mxnbathy_g", "i"), ("x_bathy_g", "D"),
         ("z_bathy_g", "D"), ("mann_bathy_g", "D"), ("size_bathy_g", "I"), ("usgs_da_g", "D"), ("usgs_da_reach_g", "I"),
         ("rdx_ar_g", "D"), ("cwnrow_g", "i"), ("cwncol_g", "i"), ("crosswalk_g", "D"), ("z_thalweg_g", "D")]


def _marshal(diff_inputs):
    """Fortran-ordered copies of every array (np.asfortranarray, diffusive.pyx:131-171) and by-reference scalars."""
    keep, ptrs = [], []
    for name, kind in _ARGS:
        v = diff_inputs[name]
        if kind == "i":
            c = C.c_int(int(v)); keep.append(c); ptrs.append(C.cast(C.pointer(c), C.c_void_p))
        else:
            a = np.asfortranarray(v, dtype=np.int32 if kind == "I" else np.float64)
            if a.size == 0:
                a = np.zeros(1, dtype=a.dtype)
            keep.append(a); ptrs.append(C.c_void_p(a.ctypes.data))
    shape = (int(diff_inputs["ntss_ev_g"]), int(diff_inputs["mxncomp_g"]), int(diff_inputs["nrch_g"]))
    outs = [np.zeros(shape, dtype=np.float64, order="F") for _ in range(3)]
    ptrs += [C.c_void_p(o.ctypes.data) for o in outs]
    return ptrs, keep, outs


def compute_diffusive(diff_inputs):
    ptrs, keep, outs = _marshal(diff_inputs)
    _lib.check(_lib.lib().trt_c_diffnw(*ptrs))
    return tuple(np.ascontiguousarray(o) for o in outs)


def compute_diffusive_batch(list_of_diff_inputs):
    """[(out_q, out_elv, out_depth), ...] for independent domains, routed concurrently."""
    if not list_of_diff_inputs:
        return []
    all_ptrs, keeps, all_outs = [], [], []
    for d in list_of_diff_inputs:
        ptrs, keep, outs = _marshal(d)
        all_ptrs += ptrs; keeps.append(keep); all_outs.append(outs)
    argv = (C.c_void_p * len(all_ptrs))(*all_ptrs)
    _lib.check(_lib.lib().trt_diffnw_batch(len(list_of_diff_inputs), argv))
    return [tuple(np.ascontiguousarray(o) for o in outs) for outs in all_outs]


def last_run():
    """(table_ms, loop_ms, launches) of the last call."""
    a, b, n = C.c_double(), C.c_double(), C.c_longlong()
    _lib.check(_lib.lib().trt_diffusive_last_run(C.byref(a), C.byref(b), C.byref(n)))
    return a.value, b.value, n.value
