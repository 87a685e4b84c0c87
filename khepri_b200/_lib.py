"""ctypes binding of the C ABI declared in include/khepri_b200.h.

The product library is ``khepri_b200/lib/libkhepri_b200.so`` (built in-tree by
``__graft_entry__.build()`` with nvcc for sm_100a).  There is NO CPU fallback: if the library is
missing, or no CUDA device is visible, loading fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libkhepri_b200.so")

KH_EINVAL, KH_ENOMEM, KH_ESTATE = -1, -2, -3
LAYER_UNIFORM, LAYER_PIXMAP, LAYER_HALF_INC, LAYER_HALF_TRN, LAYER_EXTENDED = 0, 1, 3, 4, 5
WANT_STOT, WANT_FLUX, WANT_FIELDS = 1, 2, 4
METHOD_EIG, METHOD_DOUBLING = 0, 1


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("eps_re", C.c_double), ("eps_im", C.c_double), ("depth", C.c_double),
                ("C_dev", C.c_void_p), ("IC_dev", C.c_void_p), ("retain", C.c_int),
                ("ext_base", C.c_int), ("ext_mode", C.c_int)]


class Outputs(C.Structure):
    _fields_ = [("Stot_dev", C.c_void_p), ("RT_dev", C.c_void_p), ("orders_dev", C.c_void_p),
                ("prefix_dev", C.c_void_p), ("suffix_dev", C.c_void_p),
                ("W_dev", C.c_void_p), ("V_dev", C.c_void_p), ("L_dev", C.c_void_p), ("info_dev", C.c_void_p)]


# name -> (restype, argtypes); every symbol declared in include/khepri_b200.h
vp, i32, i64, f64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_size_t
SIGNATURES = {
    "kh_abi_version": (i32, []),
    "kh_last_error": (C.c_char_p, []),
    "kh_convmat_work_bytes": (sz, [i32, i32, i32, i32, i32]),
    "kh_convmat": (i32, [i32, i32, i32, i32, vp, i32, i32, vp, vp, vp, sz, vp]),
    "kh_toeplitz_gather": (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    "kh_zgemm_batched": (i32, [i32, i32, i32, i32, i32, vp, i32, i64, vp, i32, i64, vp, i32, i64, f64, vp]),
    "kh_zinv_work_bytes": (sz, [i32, i32]),
    "kh_zinv_batched": (i32, [i32, i32, vp, vp, vp, vp, sz, vp]),
    "kh_zgeev_work_bytes": (sz, [i32, i32]),
    "kh_zgeev_batched": (i32, [i32, i32, vp, vp, vp, vp, sz, vp, vp]),
    "kh_plan_create": (i32, [C.POINTER(vp), i32, i32, vp, f64, f64, f64, f64, i32, C.POINTER(LayerDesc), i32,
                             C.POINTER(i32), i32, vp, vp]),
    "kh_plan_destroy": (None, [vp]),
    "kh_plan_set_method": (i32, [vp, i32, f64, f64]),
    "kh_solve_workspace_bytes": (sz, [vp, i32, i32]),
    "kh_solve_batch": (i32, [vp, i32, vp, vp, vp, C.POINTER(Outputs), vp, sz, vp]),
    "kh_star_workspace_bytes": (sz, [i32, i32]),
    "kh_star_batch": (i32, [i32, i32, vp, vp, vp, vp, sz, vp]),
    "kh_flux_batch": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp]),
    "kh_fields_workspace_bytes": (sz, [vp, i32, i32, i32]),
    "kh_fields_batch": (i32, [vp, i32, vp, vp, vp, C.POINTER(Outputs), vp, vp, i32, C.POINTER(f64), i32, C.POINTER(f64),
                              vp, vp, sz, vp]),
    "kh_fields_grid_workspace_bytes": (sz, [vp, i32, i32, i32, i32]),
    "kh_fields_grid_batch": (i32, [vp, i32, vp, vp, vp, C.POINTER(Outputs), vp, i32, vp, i32, C.POINTER(f64), i32, C.POINTER(f64),
                                   vp, vp, sz, vp]),
    "kh_fields_fourier_batch": (i32, [vp, i32, vp, vp, vp, C.POINTER(Outputs), C.POINTER(f64), i32, C.POINTER(f64), vp, vp, sz, vp]),
    "kh_idft_work_bytes": (sz, [i32, i32]),
    "kh_idft_batch": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "kh_beam_amplitudes_work_bytes": (sz, [i32, i32, i32]),
    "kh_beam_amplitudes": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, f64, vp, vp, sz, vp]),
    "kh_fp64_peak": (i32, [i32, i32, i32, vp, C.POINTER(f64)]),
    "kh_launch_count": (i64, []),
    "kh_profile_begin": (i32, []),
    "kh_profile_end": (i32, [C.c_char_p, sz]),
}


class KhepriError(RuntimeError):
    pass


def bind(path=None):
    """Load the shared library and attach signatures.  Raises if it is missing."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise KhepriError(
            f"khepri_b200: CUDA library not found at {path}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.kh_abi_version() != 2:
        raise KhepriError("khepri_b200: ABI version mismatch")
    return lib


def check(lib, code, what):
    if code != 0:
        msg = lib.kh_last_error().decode(errors="replace")
        kind = {KH_EINVAL: "invalid argument", KH_ENOMEM: "workspace too small", KH_ESTATE: "bad state"}.get(code, f"CUDA error {code}")
        raise KhepriError(f"{what}: {kind}: {msg}")
