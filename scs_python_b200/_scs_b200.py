"""`_scs_b200`: the per-backend extension-module surface of the reference, over libscsb200.so.

The reference builds one CPython extension per linear-system backend from scs/scspy.c
(scs/scsmodule.h, scs/scsobject.h); each exports

    SCS(shape, Ax, Ai, Ap, Px, Pi, Pp, b, c, cone, **settings)   (scsobject.h:442-913)
        .solve(warm_start, x, y, s) -> {"x","y","s","info"}       (scsobject.h:916-1130)
        .update(b, c)                                             (scsobject.h:1133-1225)
    version(), sizeof_int(), sizeof_float()                       (scsmodule.h:4-23)

This module is that surface for the B200 backend, bound with ctypes to the C ABI declared
in include/scs_b200.h (the same entry points scs/scspy.c binds when it is compiled with the
PY_B200 branch shown in INTEGRATION.md).  Argument meaning, validation messages, error types,
result keys and the per-instance lock (scsobject.h:895,939-949) follow the reference.  There
is no CPU fallback: if libscsb200.so is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCS_B200_LIBPATH: a development build of the same library (tools/spmv_variants.py compares tile geometries)
_LIB_PATH = os.environ.get("SCS_B200_LIBPATH") or os.path.join(_HERE, "libscsb200.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(
        "scs_python_b200: %s not found. Build it with `python -m scs_python_b200.build` "
        "(nvcc, sm_100a). This backend has no CPU fallback." % _LIB_PATH)
lib = C.CDLL(_LIB_PATH)

c_int, c_double = C.c_int, C.c_double
p_int, p_double = C.POINTER(C.c_int), C.POINTER(C.c_double)


class AaStats(C.Structure):
    _fields_ = [("iter", c_int), ("n_accept", c_int), ("n_reject_lapack", c_int), ("n_reject_rank0", c_int),
                ("n_reject_nonfinite", c_int), ("n_reject_weight_cap", c_int), ("n_safeguard_reject", c_int),
                ("last_rank", c_int), ("last_aa_norm", c_double), ("last_regularization", c_double)]


class ScsMatrix(C.Structure):
    _fields_ = [("x", p_double), ("i", p_int), ("p", p_int), ("m", c_int), ("n", c_int)]


class ScsSettings(C.Structure):
    _fields_ = [("normalize", c_int), ("scale", c_double), ("adaptive_scale", c_int), ("rho_x", c_double),
                ("max_iters", c_int), ("eps_abs", c_double), ("eps_rel", c_double), ("eps_infeas", c_double),
                ("alpha", c_double), ("time_limit_secs", c_double), ("verbose", c_int), ("warm_start", c_int),
                ("acceleration_lookback", c_int), ("acceleration_interval", c_int),
                ("acceleration_type_1", c_int), ("acceleration_regularization", c_double),
                ("acceleration_relaxation", c_double), ("write_data_filename", C.c_char_p),
                ("log_csv_filename", C.c_char_p)]


class ScsData(C.Structure):
    _fields_ = [("m", c_int), ("n", c_int), ("A", C.POINTER(ScsMatrix)), ("P", C.POINTER(ScsMatrix)),
                ("b", p_double), ("c", p_double)]


class ScsCone(C.Structure):
    _fields_ = [("z", c_int), ("l", c_int), ("bu", p_double), ("bl", p_double), ("bsize", c_int),
                ("q", p_int), ("qsize", c_int), ("s", p_int), ("ssize", c_int), ("cs", p_int),
                ("cssize", c_int), ("ep", c_int), ("ed", c_int), ("p", p_double), ("psize", c_int)]


class ScsSolution(C.Structure):
    _fields_ = [("x", p_double), ("y", p_double), ("s", p_double)]


class ScsInfo(C.Structure):
    _fields_ = [("iter", c_int), ("status", C.c_char * 128), ("lin_sys_solver", C.c_char * 128),
                ("status_val", c_int), ("scale_updates", c_int), ("pobj", c_double), ("dobj", c_double),
                ("res_pri", c_double), ("res_dual", c_double), ("gap", c_double), ("res_infeas", c_double),
                ("res_unbdd_a", c_double), ("res_unbdd_p", c_double), ("setup_time", c_double),
                ("solve_time", c_double), ("scale", c_double), ("comp_slack", c_double),
                ("rejected_accel_steps", c_int), ("accepted_accel_steps", c_int), ("aa_stats", AaStats),
                ("lin_sys_time", c_double), ("cone_time", c_double), ("accel_time", c_double)]


class ScsB200Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_longlong), ("cg_iters", C.c_longlong), ("admm_iters", C.c_longlong),
                ("spmv_calls", C.c_longlong), ("spmv_ms", c_double), ("algorithmic_bytes", c_double),
                ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
                ("collectives", C.c_longlong), ("collective_bytes", C.c_longlong),
                ("tiled_a", C.c_longlong), ("tiled_g", C.c_longlong), ("tiled_slots", C.c_longlong),
                ("tiled_nnz", C.c_longlong)]


class ScsB200Marks(C.Structure):
    _fields_ = [("ms", c_double), ("iters", C.c_longlong), ("cg_iters", C.c_longlong),
                ("kernel_launches", C.c_longlong), ("algorithmic_bytes", c_double), ("spmv_a_ms", c_double),
                ("spmv_g_ms", c_double), ("spmv_a_launches", C.c_longlong), ("spmv_g_launches", C.c_longlong),
                ("bytes_a", c_double), ("bytes_g", c_double)]


# --- prototypes (include/scs_b200.h) ---
lib.scs_init.restype = C.c_void_p
lib.scs_init.argtypes = [C.POINTER(ScsData), C.POINTER(ScsCone), C.POINTER(ScsSettings)]
lib.scs_update.restype = c_int
lib.scs_update.argtypes = [C.c_void_p, p_double, p_double]
lib.scs_solve.restype = c_int
lib.scs_solve.argtypes = [C.c_void_p, C.POINTER(ScsSolution), C.POINTER(ScsInfo), c_int]
lib.scs_finish.restype = None
lib.scs_finish.argtypes = [C.c_void_p]
lib.scs_set_default_settings.restype = None
lib.scs_set_default_settings.argtypes = [C.POINTER(ScsSettings)]
lib.scs_version.restype = C.c_char_p
lib.scs_get_lin_sys_method.restype = C.c_char_p
lib.scs_init_lin_sys_work.restype = C.c_void_p
lib.scs_init_lin_sys_work.argtypes = [C.POINTER(ScsMatrix), C.POINTER(ScsMatrix), p_double]
lib.scs_free_lin_sys_work.restype = None
lib.scs_free_lin_sys_work.argtypes = [C.c_void_p]
lib.scs_solve_lin_sys.restype = c_int
lib.scs_solve_lin_sys.argtypes = [C.c_void_p, p_double, p_double, c_double]
lib.scs_update_lin_sys_diag_r.restype = c_int
lib.scs_update_lin_sys_diag_r.argtypes = [C.c_void_p, p_double]
lib.scs_b200_dist_unique_id.restype = c_int
lib.scs_b200_dist_unique_id.argtypes = [C.c_void_p]
lib.scs_b200_dist_init.restype = c_int
lib.scs_b200_dist_init.argtypes = [c_int, c_int, C.c_void_p]
lib.scs_b200_dist_finalize.restype = None
lib.scs_b200_dist_rank.restype = c_int
lib.scs_b200_dist_world.restype = c_int
lib.scs_b200_dist_partition.restype = c_int
lib.scs_b200_dist_partition.argtypes = [C.POINTER(ScsData), C.POINTER(ScsCone), c_int, c_int, C.POINTER(c_int * 12)]
lib.scs_b200_dist_local.restype = c_int
lib.scs_b200_dist_local.argtypes = [C.POINTER(ScsData), C.POINTER(ScsCone), c_int, c_int, C.POINTER(c_int * 6), p_int,
                                    p_int, p_int, p_double, p_int, p_int, p_double, p_double]
lib.scs_b200_root_plus.restype = c_double
lib.scs_b200_root_plus.argtypes = [p_double, p_double, p_double, p_double, c_int, c_double, c_double]
lib.scs_b200_lin_sys_cg_its.restype = c_int
lib.scs_b200_lin_sys_cg_its.argtypes = [C.c_void_p]
lib.scs_b200_init_cone.restype = C.c_void_p
lib.scs_b200_init_cone.argtypes = [C.POINTER(ScsCone), c_int]
lib.scs_b200_bench_proj_cone.restype = c_double
lib.scs_b200_bench_proj_cone.argtypes = [C.c_void_p, p_double, p_double, c_double, c_int, c_int, p_double]
lib.scs_b200_proj_dual_cone.restype = c_int
lib.scs_b200_proj_dual_cone.argtypes = [p_double, C.c_void_p, p_double, p_double]
lib.scs_b200_finish_cone.restype = None
lib.scs_b200_finish_cone.argtypes = [C.c_void_p]
for _f in ("scs_b200_accum_by_a", "scs_b200_accum_by_atrans", "scs_b200_accum_by_p"):
    getattr(lib, _f).restype = c_int
    getattr(lib, _f).argtypes = [C.POINTER(ScsMatrix), p_double, p_double]
lib.scs_b200_aa_init.restype = C.c_void_p
lib.scs_b200_aa_init.argtypes = [c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_double, c_int]
lib.scs_b200_aa_apply.restype = c_double
lib.scs_b200_aa_apply.argtypes = [p_double, p_double, C.c_void_p]
lib.scs_b200_aa_safeguard.restype = c_int
lib.scs_b200_aa_safeguard.argtypes = [p_double, p_double, C.c_void_p]
lib.scs_b200_aa_reset.restype = None
lib.scs_b200_aa_reset.argtypes = [C.c_void_p]
lib.scs_b200_aa_get_stats.restype = AaStats
lib.scs_b200_aa_get_stats.argtypes = [C.c_void_p]
lib.scs_b200_aa_finish.restype = None
lib.scs_b200_aa_finish.argtypes = [C.c_void_p]
lib.scs_b200_set_device.restype = c_int
lib.scs_b200_set_device.argtypes = [c_int]
lib.scs_b200_device_count.restype = c_int
lib.scs_b200_get_stats.restype = c_int
lib.scs_b200_get_stats.argtypes = [C.c_void_p, C.POINTER(ScsB200Stats)]
lib.scs_b200_set_marks.restype = c_int
lib.scs_b200_set_marks.argtypes = [C.c_void_p, c_int, c_int]
lib.scs_b200_get_marks.restype = c_int
lib.scs_b200_get_marks.argtypes = [C.c_void_p, C.POINTER(ScsB200Marks)]
lib.scs_b200_bench_spmv.restype = c_double
lib.scs_b200_bench_spmv.argtypes = [C.c_void_p, c_int, c_int, p_double]
lib.scs_b200_write_data.restype = c_int
lib.scs_b200_write_data.argtypes = [C.POINTER(ScsData), C.POINTER(ScsCone), C.POINTER(ScsSettings)]
lib.scs_b200_read_data.restype = c_int
lib.scs_b200_read_data.argtypes = [C.c_char_p, C.POINTER(C.POINTER(ScsData)), C.POINTER(C.POINTER(ScsCone)),
                                   C.POINTER(C.POINTER(ScsSettings))]
lib.scs_b200_free_data.restype = None
lib.scs_b200_free_data.argtypes = [C.POINTER(ScsData), C.POINTER(ScsCone), C.POINTER(ScsSettings)]
lib.scs_b200_csv_header.restype = C.c_char_p
lib.scs_b200_csv_header.argtypes = []
lib.scs_b200_tiled_plan.restype = c_int
lib.scs_b200_tiled_plan.argtypes = [c_int, c_int, p_int, p_int, p_int, c_int, p_int, c_int, p_int, c_int, p_int, p_double,
                                    p_int, p_int]
lib.scs_b200_tiled_geometry.restype = None
lib.scs_b200_tiled_geometry.argtypes = [C.POINTER(c_int * 4)]
lib.scs_b200_tiled_profile.restype = c_int
lib.scs_b200_tiled_profile.argtypes = [C.c_void_p, c_int, p_double, c_int]
lib.scs_b200_solve_batch.restype = c_int
lib.scs_b200_solve_batch.argtypes = [c_int, C.POINTER(C.POINTER(ScsData)), C.POINTER(C.POINTER(ScsCone)),
                                     C.POINTER(ScsSettings), C.POINTER(C.POINTER(ScsSolution)),
                                     C.POINTER(ScsInfo), c_int]
lib.scs_b200_batch_plan.restype = c_int
lib.scs_b200_batch_plan.argtypes = [c_int, C.POINTER(C.POINTER(ScsData)), C.POINTER(C.POINTER(ScsCone)), C.POINTER(ScsSettings),
                                    C.POINTER(c_int), C.POINTER(c_int)]
lib.scs_b200_batch_stats.restype = c_int
lib.scs_b200_batch_stats.argtypes = [C.POINTER(c_double * 18)]


def dist_unique_id():
    """128-byte NCCL id (rank 0 calls this and ships the bytes to the other ranks)."""
    buf = C.create_string_buffer(128)
    if lib.scs_b200_dist_unique_id(buf) != 0:
        raise RuntimeError("scs_b200_dist_unique_id failed (is libnccl.so.2 loadable?)")
    return buf.raw


def dist_init(rank, world, unique_id):
    """Join the row-partitioned communicator: workspaces created afterwards keep a cone-aligned
    block of rows of A on this rank's GPU (include/scs_b200.h)."""
    buf = C.create_string_buffer(bytes(unique_id), 128) if world > 1 else None
    if lib.scs_b200_dist_init(int(rank), int(world), buf) != 0:
        raise RuntimeError("scs_b200_dist_init failed")


def dist_finalize():
    lib.scs_b200_dist_finalize()


def dist_partition(shape, Ax, Ai, Ap, b, c, cone, rank, world):
    """Host-only: what block `rank` of `world` would own (row0, m, nnz, z, l, bsize, qsize, ssize,
    cssize, ep, ed, psize)."""
    m, n = int(shape[0]), int(shape[1])
    Ax = _check_float_1d(Ax, "Ax"); Ai = _check_int_1d(Ai, "Ai"); Ap = _check_int_1d(Ap, "Ap")
    b = _check_float_1d(b, "b"); c = _check_float_1d(c, "c")
    A = make_matrix(Ax, Ai, Ap, m, n)
    k, keep = make_cone(cone)
    d = ScsData(m, n, C.pointer(A), None, _dptr(b), _dptr(c))
    out = (c_int * 12)()
    if lib.scs_b200_dist_partition(C.byref(d), C.byref(k), int(rank), int(world), C.byref(out)) != 0:
        raise ValueError("partition failed")
    keys = ("row0", "m", "nnz", "z", "l", "bsize", "qsize", "ssize", "cssize", "ep", "ed", "psize")
    return dict(zip(keys, list(out)))


def dist_local(shape, Ax, Ai, Ap, Px, Pi, Pp, b, c, cone, rank, world):
    """Host-only: the local problem rank `rank` of `world` builds in the row-partitioned mode (same arguments as
    SCS.__init__).  Returns dict(row0, m, n_sh, loc2glob, A=(Ax, Ai, Ap) local CSC of shape (m, n_loc),
    P=(Px, Pi, Pp) or None, c)."""
    import scipy.sparse as sp
    m, n = int(shape[0]), int(shape[1])
    Ax = _check_float_1d(Ax, "Ax"); Ai = _check_int_1d(Ai, "Ai"); Ap = _check_int_1d(Ap, "Ap")
    b = _check_float_1d(b, "b"); c = _check_float_1d(c, "c")
    A = make_matrix(Ax, Ai, Ap, m, n)
    P = None
    if Px is not None:
        Px = _check_float_1d(Px, "Px"); Pi = _check_int_1d(Pi, "Pi"); Pp = _check_int_1d(Pp, "Pp")
        P = make_matrix(Px, Pi, Pp, n, n)
    k, keep = make_cone(cone)
    d = ScsData(m, n, C.pointer(A), C.pointer(P) if P is not None else None, _dptr(b), _dptr(c))
    sizes = (c_int * 6)()
    nul_i, nul_d = C.cast(None, p_int), C.cast(None, p_double)
    if lib.scs_b200_dist_local(C.byref(d), C.byref(k), int(rank), int(world), C.byref(sizes), nul_i, nul_i, nul_i, nul_d,
                               nul_i, nul_i, nul_d, nul_d) != 0:
        raise ValueError("partition failed")
    row0, ml, nsh, nl, nzA, nzP = list(sizes)
    l2g = np.zeros(nl, dtype=np.int32)
    lAp = np.zeros(nl + 1, dtype=np.int32); lAi = np.zeros(max(nzA, 1), dtype=np.int32); lAx = np.zeros(max(nzA, 1))
    lPp = np.zeros(nl + 1, dtype=np.int32); lPi = np.zeros(max(nzP, 1), dtype=np.int32); lPx = np.zeros(max(nzP, 1))
    cl = np.zeros(nl)
    if lib.scs_b200_dist_local(C.byref(d), C.byref(k), int(rank), int(world), C.byref(sizes), _iptr(l2g), _iptr(lAp),
                               _iptr(lAi), _dptr(lAx), _iptr(lPp), _iptr(lPi), _dptr(lPx), _dptr(cl)) != 0:
        raise ValueError("partition failed")
    out = dict(row0=row0, m=ml, n_sh=nsh, loc2glob=l2g, c=cl,
               A=sp.csc_matrix((lAx[:nzA], lAi[:nzA], lAp), shape=(ml, nl)), P=None)
    if P is not None:
        out["P"] = sp.csc_matrix((lPx[:nzP], lPi[:nzP], lPp), shape=(nl, nl))
    return out


def version():
    return lib.scs_version().decode()


def sizeof_int():
    return C.sizeof(c_int)


def sizeof_float():
    return C.sizeof(c_double)


def _dptr(a):
    return a.ctypes.data_as(p_double)


def _iptr(a):
    return a.ctypes.data_as(p_int)


def make_matrix(x, i, p, m, n):
    """ScsMatrix view over contiguous numpy arrays (kept alive by the caller)."""
    return ScsMatrix(_dptr(x), _iptr(i), _iptr(p), int(m), int(n))


def _check_float_1d(a, name):
    if not isinstance(a, np.ndarray) or a.dtype.kind != "f" or a.ndim != 1:
        raise TypeError("%s must be a 1-D numpy array of floats" % name)
    return np.ascontiguousarray(a, dtype=np.float64)


def _check_int_1d(a, name):
    if not isinstance(a, np.ndarray) or a.dtype.kind not in "iu" or a.ndim != 1:
        raise TypeError("%s must be a 1-D numpy array of ints" % name)
    if a.size and (a.max() > np.iinfo(np.int32).max):
        raise ValueError("%s exceeds the 32-bit index range of the B200 backend" % name)
    return np.ascontiguousarray(a, dtype=np.int32)


def _pos_int(cone, key):
    if key not in cone or cone[key] is None:
        return 0
    v = cone[key]
    if isinstance(v, bool) or not isinstance(v, (int, np.integer)) or v < 0 or v > np.iinfo(np.int32).max:
        raise ValueError("Invalid value for cone field '%s'" % key)
    return int(v)


def _int_arr(cone, key):
    if key not in cone or cone[key] is None:
        return np.zeros(0, dtype=np.int32)
    v = cone[key]
    if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
        v = [v]
    if isinstance(v, np.ndarray):
        if v.dtype.kind not in "iu" or v.ndim != 1:
            raise ValueError("Invalid value for cone field '%s'" % key)
        v = v.tolist()
    if not isinstance(v, (list, tuple)):
        raise ValueError("Invalid value for cone field '%s'" % key)
    for e in v:
        if isinstance(e, bool) or not isinstance(e, (int, np.integer)) or e < 0:
            raise ValueError("Invalid value for cone field '%s'" % key)
    return np.asarray(v, dtype=np.int32).reshape(-1)


def _float_arr(cone, key):
    if key not in cone or cone[key] is None:
        return np.zeros(0, dtype=np.float64)
    v = cone[key]
    if isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool):
        v = [v]
    if isinstance(v, np.ndarray):
        if v.dtype.kind != "f" or v.ndim != 1:
            raise ValueError("Invalid value for cone field '%s'" % key)
    try:
        return np.asarray(v, dtype=np.float64).reshape(-1)
    except (TypeError, ValueError):
        raise ValueError("Invalid value for cone field '%s'" % key)


def make_cone(cone):
    """Python cone dict -> (ScsCone, keepalive) with the parsing rules of scsobject.h:684-794."""
    keep = {}
    k = ScsCone()
    f = _pos_int(cone, "f")
    k.z = _pos_int(cone, "z")
    if f > 0:
        warnings.warn("The 'f' cone field is deprecated; use 'z' (Zero cone) instead. "
                      "If both 'f' and 'z' are set they are summed.", DeprecationWarning, stacklevel=3)
        k.z += f
    k.l = _pos_int(cone, "l")
    bu, bl = _float_arr(cone, "bu"), _float_arr(cone, "bl")
    if len(bu) != len(bl):
        raise ValueError("bu different dimension to bl")
    keep["bu"], keep["bl"] = bu, bl
    if len(bu) > 0:
        k.bsize = len(bu) + 1
        k.bu, k.bl = _dptr(bu), _dptr(bl)
    for key, fld, sz in (("q", "q", "qsize"), ("s", "s", "ssize"), ("cs", "cs", "cssize")):
        arr = _int_arr(cone, key)
        keep[key] = arr
        setattr(k, sz, len(arr))
        if len(arr):
            setattr(k, fld, _iptr(arr))
    p = _float_arr(cone, "p")
    keep["p"] = p
    k.psize = len(p)
    if len(p):
        k.p = _dptr(p)
    k.ep = _pos_int(cone, "ep")
    k.ed = _pos_int(cone, "ed")
    return k, keep


_SETTING_KEYS = ("verbose", "normalize", "adaptive_scale", "max_iters", "scale", "eps_abs", "eps_rel",
                 "eps_infeas", "alpha", "rho_x", "time_limit_secs", "acceleration_lookback",
                 "acceleration_interval", "acceleration_type_1", "acceleration_regularization",
                 "acceleration_relaxation", "write_data_filename", "log_csv_filename")
_BOOL_KEYS = ("verbose", "normalize", "adaptive_scale")
_INT_KEYS = ("max_iters", "acceleration_lookback", "acceleration_interval", "acceleration_type_1")


def make_settings(kwargs):
    """kwargs -> ScsSettings with the type and range checks of scsobject.h:500-868."""
    st = ScsSettings()
    lib.scs_set_default_settings(C.byref(st))
    keep = []
    for key, val in kwargs.items():
        if key not in _SETTING_KEYS:
            raise TypeError("'%s' is an invalid keyword argument for SCS()" % key)
        if key in _BOOL_KEYS:
            if not isinstance(val, (bool, np.bool_)):
                raise TypeError("argument '%s' must be bool, not %s" % (key, type(val).__name__))
            setattr(st, key, 1 if val else 0)
        elif key in _INT_KEYS:
            if isinstance(val, (bool, np.bool_)):
                val = int(val)
            if not isinstance(val, (int, np.integer)):
                raise TypeError("argument '%s' must be int, not %s" % (key, type(val).__name__))
            if abs(int(val)) > np.iinfo(np.int32).max:
                raise OverflowError("signed integer is greater than maximum")
            setattr(st, key, int(val))
        elif key in ("write_data_filename", "log_csv_filename"):
            if val is not None:
                if not isinstance(val, str):
                    raise TypeError("argument '%s' must be str or None" % key)
                b = val.encode()
                keep.append(b)
                setattr(st, key, b)
        else:
            if isinstance(val, (bool, np.bool_)) or not isinstance(val, (int, float, np.integer, np.floating)):
                raise TypeError("argument '%s' must be float, not %s" % (key, type(val).__name__))
            setattr(st, key, float(val))
    if st.max_iters <= 0:
        raise ValueError("max_iters must be positive")
    if st.acceleration_lookback < 0:
        raise ValueError("acceleration_lookback must be nonnegative (use acceleration_type_1=0 for type-II AA)")
    if st.acceleration_interval <= 0:
        raise ValueError("acceleration_interval must be positive")
    if not math.isfinite(st.acceleration_regularization) or st.acceleration_regularization < 0:
        raise ValueError("acceleration_regularization must be a nonnegative finite number")
    if (not math.isfinite(st.acceleration_relaxation) or st.acceleration_relaxation < 0
            or st.acceleration_relaxation > 2):
        raise ValueError("acceleration_relaxation must be in [0, 2]")
    if not math.isfinite(st.scale) or st.scale <= 0:
        raise ValueError("scale must be a positive finite number")
    if math.isnan(st.time_limit_secs) or st.time_limit_secs < 0:
        raise ValueError("time_limit_secs must be nonnegative")
    for key in ("eps_abs", "eps_rel", "eps_infeas"):
        v = getattr(st, key)
        if math.isnan(v) or v < 0:
            raise ValueError("%s must be nonnegative" % key)
    if not math.isfinite(st.alpha) or st.alpha <= 0 or st.alpha >= 2:
        raise ValueError("alpha must be in (0, 2)")
    if not math.isfinite(st.rho_x) or st.rho_x <= 0:
        raise ValueError("rho_x must be a positive finite number")
    st.warm_start = 0
    return st, keep


_INFO_KEYS = ("status_val", "iter", "scale_updates", "scale", "pobj", "dobj", "res_pri", "res_dual", "gap",
              "res_infeas", "res_unbdd_a", "res_unbdd_p", "comp_slack", "solve_time", "setup_time",
              "lin_sys_time", "cone_time", "accel_time", "rejected_accel_steps", "accepted_accel_steps")
_AA_KEYS = ("iter", "n_accept", "n_reject_lapack", "n_reject_rank0", "n_reject_nonfinite",
            "n_reject_weight_cap", "n_safeguard_reject", "last_rank", "last_aa_norm", "last_regularization")


def read_data(filename):
    """SCS(read_data), S/src/rw.c:262-315 (what S/test/run_from_file.c does before solving): a problem
    file in the reference's binary layout -> (data, cone, settings) with data = dict(A, P, b, c) holding
    scipy CSC matrices.  Raises ValueError when the file cannot be read."""
    import scipy.sparse as sp
    d, k, st = C.POINTER(ScsData)(), C.POINTER(ScsCone)(), C.POINTER(ScsSettings)()
    if lib.scs_b200_read_data(os.fsencode(filename), C.byref(d), C.byref(k), C.byref(st)) != 0:
        raise ValueError("could not read SCS data file %r" % (filename,))
    try:
        D, K, S = d.contents, k.contents, st.contents

        def mat(M):
            nnz = M.p[M.n]
            return sp.csc_matrix((np.array(M.x[:nnz], dtype=np.float64), np.array(M.i[:nnz], dtype=np.int32),
                                  np.array(M.p[:M.n + 1], dtype=np.int32)), shape=(M.m, M.n))
        data = dict(A=mat(D.A.contents), b=np.array(D.b[:D.m], dtype=np.float64), c=np.array(D.c[:D.n], dtype=np.float64))
        data["P"] = mat(D.P.contents) if D.P else None
        cone = dict(z=K.z, l=K.l, ep=K.ep, ed=K.ed)
        if K.bsize > 1:
            cone["bl"] = np.array(K.bl[:K.bsize - 1], dtype=np.float64)
            cone["bu"] = np.array(K.bu[:K.bsize - 1], dtype=np.float64)
        if K.qsize > 0:
            cone["q"] = [int(v) for v in K.q[:K.qsize]]
        if K.ssize > 0:
            cone["s"] = [int(v) for v in K.s[:K.ssize]]
        if K.psize > 0:
            cone["p"] = [float(v) for v in K.p[:K.psize]]
        settings = dict(normalize=bool(S.normalize), scale=S.scale, rho_x=S.rho_x, max_iters=S.max_iters,
                        eps_abs=S.eps_abs, eps_rel=S.eps_rel, eps_infeas=S.eps_infeas, alpha=S.alpha,
                        verbose=bool(S.verbose), acceleration_lookback=S.acceleration_lookback,
                        acceleration_interval=S.acceleration_interval, acceleration_type_1=S.acceleration_type_1,
                        acceleration_regularization=S.acceleration_regularization,
                        acceleration_relaxation=S.acceleration_relaxation, adaptive_scale=bool(S.adaptive_scale))
    finally:
        lib.scs_b200_free_data(d, k, st)
    return data, cone, settings


def write_data(filename, shape, Ax, Ai, Ap, Px, Pi, Pp, b, c, cone, **settings):
    """SCS(write_data), S/src/rw.c:240-260, without creating a workspace (no device needed)."""
    m, n = int(shape[0]), int(shape[1])
    if m <= 0 or n <= 0:
        raise ValueError("m and n must be positive integers")
    # same dtype / shape validation as SCS.__init__ (scsobject.h:507-640): the converted arrays are kept
    # alive until the call returns, so the C side never sees a reinterpreted buffer
    Ax = _check_float_1d(Ax, "Ax")
    Ai = _check_int_1d(Ai, "Ai")
    Ap = _check_int_1d(Ap, "Ap")
    if len(Ap) != n + 1:
        raise ValueError("Ap has incompatible dimension with A")
    if len(Ai) != len(Ax) or (len(Ap) and int(Ap[-1]) != len(Ax)):
        raise ValueError("Ai / Ax have incompatible dimension with Ap")
    A = make_matrix(Ax, Ai, Ap, m, n)
    data = ScsData()
    data.m, data.n = m, n
    data.A = C.pointer(A)
    P = None
    if Px is not None and Pi is not None and Pp is not None:
        Px = _check_float_1d(Px, "Px")
        Pi = _check_int_1d(Pi, "Pi")
        Pp = _check_int_1d(Pp, "Pp")
        if len(Pp) != n + 1:
            raise ValueError("Pp has incompatible dimension with P")
        if len(Pi) != len(Px) or int(Pp[-1]) != len(Px):
            raise ValueError("Pi / Px have incompatible dimension with Pp")
        P = make_matrix(Px, Pi, Pp, n, n)
        data.P = C.pointer(P)
    b = _check_float_1d(b, "b")
    c = _check_float_1d(c, "c")
    if b.shape[0] != m:
        raise ValueError("b has incompatible dimension with A")
    if c.shape[0] != n:
        raise ValueError("c has incompatible dimension with A")
    data.b, data.c = _dptr(b), _dptr(c)
    k, keep = make_cone(cone)
    st, keep2 = make_settings(dict(settings, write_data_filename=os.fspath(filename)))
    if lib.scs_b200_write_data(C.byref(data), C.byref(k), C.byref(st)) != 0:
        raise ValueError("could not write SCS data file %r" % (filename,))


def csv_header():
    return lib.scs_b200_csv_header().decode()


def _info_dict(info):
    d = {key: getattr(info, key) for key in _INFO_KEYS}
    d["status"] = info.status.decode()
    d["lin_sys_solver"] = info.lin_sys_solver.decode()
    d["aa_stats"] = {key: getattr(info.aa_stats, key) for key in _AA_KEYS}
    return d


def _mirror(struct):
    """Same layout as `struct` with every pointer field typed c_void_p, so that a field can be set from an
    integer address (a ctypes POINTER object costs ~3 us to build, an integer assignment ~0.2 us; the batch
    wrapper fills ~25 pointer fields per problem)."""
    fields = []
    for name, typ in struct._fields_:
        fields.append((name, C.c_void_p if isinstance(typ, type) and issubclass(typ, C._Pointer) else typ))
    cls = type(struct.__name__ + "Addr", (C.Structure,), {"_fields_": fields})
    assert C.sizeof(cls) == C.sizeof(struct)
    for name, _ in struct._fields_:
        assert getattr(cls, name).offset == getattr(struct, name).offset
    return cls


_MatA, _DataA, _ConeA, _SolA = _mirror(ScsMatrix), _mirror(ScsData), _mirror(ScsCone), _mirror(ScsSolution)
_FAST_CONE_KEYS = frozenset(("z", "l", "q"))


def _addr(a):
    return a.__array_interface__["data"][0]


def _f64(a, name):
    if type(a) is np.ndarray and a.dtype == np.float64 and a.ndim == 1 and a.flags.c_contiguous:
        return a
    return _check_float_1d(a, name)


def _i32(a, name):
    if type(a) is np.ndarray and a.dtype == np.int32 and a.ndim == 1 and a.flags.c_contiguous:
        return a
    return _check_int_1d(a, name)


def _fill_batch(problems):
    """C structures of a batch, filled in place in contiguous arrays (shared by solve_batch and batch_plan)."""
    cnt = len(problems)
    nalloc = max(cnt, 1)
    matsA, matsP = (_MatA * nalloc)(), (_MatA * nalloc)()
    datas, cones = (_DataA * nalloc)(), (_ConeA * nalloc)()
    szM = C.sizeof(_MatA)
    baseA, baseP = C.addressof(matsA), C.addressof(matsP)
    keep = [matsA, matsP]
    dims = []
    for idx, (shape, Ax, Ai, Ap, Px, Pi, Pp, b, c, cone) in enumerate(problems):
        m, n = int(shape[0]), int(shape[1])
        if m <= 0 or n <= 0:
            raise ValueError("m and n must be positive integers")
        if not isinstance(cone, dict):
            raise TypeError("cone must be a dict")
        Ax = _f64(Ax, "Ax"); Ai = _i32(Ai, "Ai"); Ap = _i32(Ap, "Ap")
        if Ap.shape[0] != n + 1:
            raise ValueError("Ap has incompatible dimension with A")
        A = matsA[idx]
        A.x, A.i, A.p, A.m, A.n = _addr(Ax), _addr(Ai), _addr(Ap), m, n
        d = datas[idx]
        d.m, d.n, d.A = m, n, baseA + idx * szM
        if Px is not None and Pi is not None and Pp is not None:
            Px = _f64(Px, "Px"); Pi = _i32(Pi, "Pi"); Pp = _i32(Pp, "Pp")
            P = matsP[idx]
            P.x, P.i, P.p, P.m, P.n = _addr(Px), _addr(Pi), _addr(Pp), n, n
            d.P = baseP + idx * szM
        c = _f64(c, "c"); b = _f64(b, "b")
        if c.shape[0] != n:
            raise ValueError("c has incompatible dimension with A")
        if b.shape[0] != m:
            raise ValueError("b has incompatible dimension with A")
        d.b, d.c = _addr(b), _addr(c)
        k = cones[idx]
        if _FAST_CONE_KEYS.issuperset(cone):   # zero / nonneg / second-order cones only: no further fields
            k.z, k.l = _pos_int(cone, "z"), _pos_int(cone, "l")
            q = _int_arr(cone, "q")
            k.qsize = q.shape[0]
            if k.qsize:
                k.q = _addr(q)
            keep.append((Ax, Ai, Ap, Px, Pi, Pp, b, c, q))
        else:
            ks, keep_cone = make_cone(cone)
            C.memmove(C.addressof(k), C.addressof(ks), C.sizeof(ScsCone))
            keep.append((Ax, Ai, Ap, Px, Pi, Pp, b, c, ks, keep_cone))
        dims.append((n, m))
    return nalloc, datas, cones, keep, dims


def _ptr_array(arr, size, typ, nalloc):
    pa = (C.c_void_p * nalloc)()
    np.frombuffer(pa, dtype=np.uint64)[:] = C.addressof(arr) + size * np.arange(nalloc, dtype=np.uint64)
    return pa, C.cast(pa, C.POINTER(C.POINTER(typ)))


def batch_plan(problems, **settings):
    """Host-only (include/scs_b200.h: scs_b200_batch_plan): which members of the batch the one-CTA kernel takes and
    the shared-memory carve-up they share.  Returns (fused, plan): fused[i] = 1 batch kernel / 0 streaming engine /
    -1 fails validation; plan = dict(smem_bytes, direct, extended_cones, psd_order, psd_workspaces, cones, power_cones,
    box_bounds).  No device is touched."""
    cnt = len(problems)
    stgs, keep_stgs = make_settings(settings)
    nalloc, datas, cones, keep, _ = _fill_batch(problems)
    pd, pdc = _ptr_array(datas, C.sizeof(_DataA), ScsData, nalloc)
    pk, pkc = _ptr_array(cones, C.sizeof(_ConeA), ScsCone, nalloc)
    fused = (c_int * nalloc)()
    out = (c_int * 8)()
    rc = lib.scs_b200_batch_plan(cnt, pdc, pkc, C.byref(stgs), fused, out)
    del keep_stgs, keep, pd, pk
    if rc < 0:
        raise ValueError("scs_b200_batch_plan failed")
    names = ("smem_bytes", "direct", "extended_cones", "psd_order", "psd_workspaces", "cones", "power_cones", "box_bounds")
    return [int(fused[i]) for i in range(cnt)], dict(zip(names, (int(v) for v in out)))


def solve_batch(problems, **settings):
    """Solve independent problems in one call (include/scs_b200.h: scs_b200_solve_batch).

    problems: sequence of (shape, Ax, Ai, Ap, Px, Pi, Pp, b, c, cone) tuples -- the constructor
    arguments of `SCS` -- all solved with the same settings.  Returns a list of
    {"x","y","s","info"} dicts in input order.  The argument checks are those of `SCS`; the C structures
    of the whole batch are filled in place in five contiguous arrays."""
    cnt = len(problems)
    stgs, keep_stgs = make_settings(settings)
    nalloc, datas, cones, keep, dims = _fill_batch(problems)
    sols = (_SolA * nalloc)()
    infos = (ScsInfo * nalloc)()
    # one output block for the whole batch: x | y | s per problem
    offs = np.zeros(cnt + 1, dtype=np.int64)
    if cnt:
        offs[1:] = np.cumsum([n + 2 * m for n, m in dims])
    block = np.zeros(int(offs[-1]))
    base = _addr(block)
    out = []
    for idx, (n, m) in enumerate(dims):
        o = int(offs[idx])
        sl = sols[idx]
        sl.x, sl.y, sl.s = base + 8 * o, base + 8 * (o + n), base + 8 * (o + n + m)
        out.append((block[o:o + n], block[o + n:o + n + m], block[o + n + m:o + n + 2 * m]))

    pd, pdc = _ptr_array(datas, C.sizeof(_DataA), ScsData, nalloc)
    pk, pkc = _ptr_array(cones, C.sizeof(_ConeA), ScsCone, nalloc)
    ps, psc = _ptr_array(sols, C.sizeof(_SolA), ScsSolution, nalloc)
    lib.scs_b200_solve_batch(cnt, pdc, pkc, C.byref(stgs), psc, infos, 0)  # ctypes releases the GIL
    del keep_stgs, keep, pd, pk, ps
    return [{"x": x, "y": y, "s": s, "info": _info_dict(infos[i])} for i, (x, y, s) in enumerate(out)]


def batch_stats():
    buf = (c_double * 18)()
    lib.scs_b200_batch_stats(C.byref(buf))
    keys = ("fused", "streamed", "kernel_ms", "pack_ms", "h2d_bytes", "d2h_bytes", "ctas", "smem_per_cta", "direct",
            "cg_iters", "admm_iters", "_", "clk_equil", "clk_factor", "clk_linsys", "clk_aa", "clk_resid", "clk_total")
    return dict(zip(keys, list(buf)))


class SCS(object):
    """Same constructor / methods as the extension type `scs.SCS` (scsobject.h:1261-1307)."""

    def __init__(self, shape, Ax, Ai, Ap, Px, Pi, Pp, b, c, cone, **settings):
        self._work = None
        self._lock = threading.Lock()
        m, n = int(shape[0]), int(shape[1])
        if m <= 0:
            raise ValueError("m must be a positive integer")
        if n <= 0:
            raise ValueError("n must be a positive integer")
        if not isinstance(cone, dict):
            raise TypeError("cone must be a dict")
        self.m, self.n = m, n
        Ax = _check_float_1d(Ax, "Ax")
        Ai = _check_int_1d(Ai, "Ai")
        Ap = _check_int_1d(Ap, "Ap")
        A = make_matrix(Ax, Ai, Ap, m, n)
        P = None
        if Px is not None and Pi is not None and Pp is not None:
            Px = _check_float_1d(Px, "Px")
            Pi = _check_int_1d(Pi, "Pi")
            Pp = _check_int_1d(Pp, "Pp")
            P = make_matrix(Px, Pi, Pp, n, n)
        c = _check_float_1d(c, "c")
        if c.shape[0] != n:
            raise ValueError("c has incompatible dimension with A")
        b = _check_float_1d(b, "b")
        if b.shape[0] != m:
            raise ValueError("b has incompatible dimension with A")
        if len(Ap) != n + 1:
            raise ValueError("Ap has incompatible dimension with A")
        k, keep_cone = make_cone(cone)
        stgs, keep_stgs = make_settings(settings)
        d = ScsData(m, n, C.pointer(A), C.pointer(P) if P is not None else None, _dptr(b), _dptr(c))
        self._x = np.zeros(n)
        self._y = np.zeros(m)
        self._s = np.zeros(m)
        self._sol = ScsSolution(_dptr(self._x), _dptr(self._y), _dptr(self._s))
        work = lib.scs_init(C.byref(d), C.byref(k), C.byref(stgs))  # ctypes releases the GIL
        del keep_cone, keep_stgs
        if not work:
            raise ValueError("ScsWork allocation error!")
        self._work = C.c_void_p(work)

    def solve(self, warm_start, x, y, s):
        if not isinstance(warm_start, (bool, np.bool_)):
            raise TypeError("argument 1 must be bool, not %s" % type(warm_start).__name__)
        with self._lock:
            if not self._work:
                raise ValueError("Workspace not initialized!")
            if warm_start:
                for name, dst, src, ln in (("x", self._x, x, self.n), ("y", self._y, y, self.m),
                                           ("s", self._s, s, self.m)):
                    if src is None:
                        continue
                    if (not isinstance(src, np.ndarray) or src.dtype.kind != "f" or src.ndim != 1
                            or src.shape[0] != ln):
                        raise ValueError("Warm-start must be a 1-D float array of length %d" % ln)
                    dst[:] = src
            info = ScsInfo()
            lib.scs_solve(self._work, C.byref(self._sol), C.byref(info), 1 if warm_start else 0)
            out_x, out_y, out_s = self._x.copy(), self._y.copy(), self._s.copy()
        return {"x": out_x, "y": out_y, "s": out_s, "info": _info_dict(info)}

    def update(self, b, c):
        bb = cc = None
        if c is not None:
            if not isinstance(c, np.ndarray) or c.dtype.kind != "f" or c.ndim != 1:
                raise TypeError("c_new must be a 1-D numpy array of floats")
            if c.shape[0] != self.n:
                raise ValueError("c_new has incompatible dimension with A")
            cc = np.ascontiguousarray(c, dtype=np.float64)
        if b is not None:
            if not isinstance(b, np.ndarray) or b.dtype.kind != "f" or b.ndim != 1:
                raise TypeError("b_new must be a 1-D numpy array of floats")
            if b.shape[0] != self.m:
                raise ValueError("b_new has incompatible dimension with A")
            bb = np.ascontiguousarray(b, dtype=np.float64)
        with self._lock:
            if not self._work:
                raise ValueError("Workspace not initialized!")
            lib.scs_update(self._work, _dptr(bb) if bb is not None else None, _dptr(cc) if cc is not None else None)

    # --- B200-only extras (no reference counterpart) ---
    def stats(self):
        st = ScsB200Stats()
        with self._lock:
            if not self._work:
                raise ValueError("Workspace not initialized!")
            lib.scs_b200_get_stats(self._work, C.byref(st))
        return {f: getattr(st, f) for f, _ in ScsB200Stats._fields_}

    def set_marks(self, begin_iter, end_iter):
        with self._lock:
            if lib.scs_b200_set_marks(self._work, int(begin_iter), int(end_iter)) != 0:
                raise ValueError("invalid iteration marks")

    def get_marks(self):
        mk = ScsB200Marks()
        with self._lock:
            if lib.scs_b200_get_marks(self._work, C.byref(mk)) != 0:
                return None
        return {f: getattr(mk, f) for f, _ in ScsB200Marks._fields_}

    def bench_spmv(self, which, reps):
        ab = c_double(0.0)
        with self._lock:
            ms = lib.scs_b200_bench_spmv(self._work, int(which), int(reps), C.byref(ab))
        return ms, ab.value

    def tiled_profile(self, which):
        buf = np.zeros(4 * 1024)
        with self._lock:
            n = lib.scs_b200_tiled_profile(self._work, int(which), _dptr(buf), 1024)
        return buf[:4 * max(n, 0)].reshape(-1, 4)

    def finish(self):
        with self._lock:
            if self._work:
                lib.scs_finish(self._work)
                self._work = None

    def __del__(self):
        try:
            self.finish()
        except Exception:
            pass
