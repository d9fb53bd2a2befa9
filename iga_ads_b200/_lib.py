"""ctypes binding of libadsb200.so -- the C ABI declared in include/adsb200.h.

Nothing here computes: every function forwards to the shared library, and the library has no CPU
fallback for the device entry points.  Importing this module never touches /root/reference or
oracle/.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libadsb200.so")

c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ll = ctypes.c_longlong
dp = ctypes.POINTER(c_dbl)
ip = ctypes.POINTER(c_int)
llp = ctypes.POINTER(c_ll)
vp = ctypes.c_void_p

MAX_SLOTS = 4
MAX_BUFFERS = 8
RHS_COLLAPSED = 0
RHS_QUADRATURE = 1


class AdsbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"adsb error {code}: {message}")
        self.code = code


class Form(ctypes.Structure):
    """adsb_form: rhs = alpha*(u,v) - sum_k beta[k]*(d_k u, d_k v) + gamma*F."""
    _fields_ = [("alpha", c_dbl), ("beta", c_dbl * 3), ("gamma", c_dbl), ("forcing_buf", c_int),
                ("method", c_int), ("source", c_int)]

    @classmethod
    def make(cls, alpha=1.0, beta=(0.0, 0.0, 0.0), gamma=0.0, forcing_buf=-1, method=RHS_COLLAPSED, source=0):
        b = list(beta) + [0.0] * (3 - len(beta))
        return cls(alpha, (c_dbl * 3)(*b), gamma, forcing_buf, method, source)


POINT_LINEAR = 0
POINT_FLOW = 1


class PointForm(ctypes.Structure):
    """adsb_point_form: the integrand of the general quadrature right-hand side (include/adsb200.h)."""
    _fields_ = [("kind", c_int), ("alpha", c_dbl), ("beta", c_dbl * 3), ("adv", c_dbl * 3), ("gamma", c_dbl),
                ("source", c_int), ("source_plain", c_int), ("par", c_dbl * 4), ("forcing_buf", c_int),
                ("forcing_scale", c_dbl)]

    @classmethod
    def linear(cls, alpha=1.0, beta=(0.0, 0.0, 0.0), adv=(0.0, 0.0, 0.0), gamma=0.0, source=0, source_plain=False,
               forcing_buf=-1, forcing_scale=0.0):
        b = list(beta) + [0.0] * (3 - len(beta))
        a = list(adv) + [0.0] * (3 - len(adv))
        return cls(POINT_LINEAR, alpha, (c_dbl * 3)(*b), (c_dbl * 3)(*a), gamma, source, int(source_plain),
                   (c_dbl * 4)(), forcing_buf, forcing_scale)

    @classmethod
    def flow(cls, dt, mi=10.0):
        """examples/flow/flow.hpp:74-101; the permeability table goes in with Context.set_point_coefficient"""
        return cls(POINT_FLOW, 1.0, (c_dbl * 3)(), (c_dbl * 3)(), dt, 2, 0, (c_dbl * 4)(dt, mi, 0.0, 0.0), -1, 0.0)


class Substep(ctypes.Structure):
    _fields_ = [("form", Form), ("slots", c_int * 3), ("fix_axis", c_int), ("fix_buf", c_int)]

    @classmethod
    def make(cls, form, slots=(0, 0, 0), fix_axis=-1, fix_buf=-1):
        s = list(slots) + [0] * (3 - len(slots))
        return cls(form, (c_int * 3)(*s), fix_axis, fix_buf)


class View(ctypes.Structure):
    _fields_ = [("n", c_int * 3), ("s", c_ll * 3)]

    @classmethod
    def make(cls, n, s):
        n = list(n) + [1] * (3 - len(n))
        s = list(s) + [int(np.prod(n))] * (3 - len(s))
        return cls((c_int * 3)(*n), (c_ll * 3)(*s))


class DistArgs(ctypes.Structure):
    """adsb_dist_args (include/adsb200.h): one rank's arguments of the fused distributed sweep"""
    _fields_ = [("rank", c_int), ("nranks", c_int), ("nl", c_int), ("lag", c_int),
                ("dseg_local", vp), ("x_local", vp), ("dseg_next", vp), ("x_prev", vp), ("error_flag", vp),
                ("halo_prev", vp), ("halo_next", vp), ("halo_planes", c_int)]


DIST_SENTINEL_WORD = 0x7FF7A5A5   # both 32-bit halves of the sentinel the fused sweep's state arrays hold


def fill_sentinel(t):
    """fill a float64 torch tensor (a state array of the fused distributed sweep) with the sentinel"""
    import torch

    t.view(torch.int32).fill_(DIST_SENTINEL_WORD)


_SIGNATURES = {
    # name: (restype, argtypes)
    "adsb_abi_version": (c_int, []),
    "adsb_last_error": (ctypes.c_char_p, []),
    "adsb_gauss": (c_int, [c_int, dp, dp]),
    "adsb_knots": (c_int, [c_int, c_int, c_dbl, c_dbl, dp]),
    "adsb_find_span": (c_int, [c_dbl, dp, c_int, c_int]),
    "adsb_basis_ders": (c_int, [c_int, c_dbl, dp, c_int, c_int, dp]),
    "adsb_basis_tables": (c_int, [c_int, c_int, c_dbl, c_dbl, c_int, c_int, dp, dp, dp, dp, ip]),
    "adsb_matrix_1d": (c_int, [c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_int, dp]),
    "adsb_band_factorize": (c_int, [c_int, c_int, c_int, dp, c_int, ip]),
    "adsb_band_unpivot": (c_int, [c_int, c_int, c_int, c_int, dp, ip, dp]),
    "adsb_segment_bounds": (c_int, [c_int, c_int, ip, c_int, c_int, ip]),
    "adsb_segment_plan": (c_int, [c_int, c_int, c_int, c_int, dp, ip, c_int, ip, c_dbl, ip, dp, dp, dp, dp, dp]),
    "adsb_create": (c_int, [c_int, ip, ip, ip, c_int, ctypes.POINTER(vp)]),
    "adsb_destroy": (c_int, [vp]),
    "adsb_set_stream": (c_int, [vp, vp]),
    "adsb_set_sm_limit": (c_int, [vp, c_int]),
    "adsb_copy2d": (c_int, [vp, vp, c_ll, vp, c_ll, c_ll, c_ll]),
    "adsb_synchronize": (c_int, [vp]),
    "adsb_set_axis_tables": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, dp, dp, dp, dp, ip]),
    "adsb_set_axis_factor": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, c_int, dp, ip]),
    "adsb_upload": (c_int, [vp, c_int, dp]),
    "adsb_download": (c_int, [vp, c_int, dp]),
    "adsb_upload_async": (c_int, [vp, c_int, vp, vp]),
    "adsb_download_async": (c_int, [vp, c_int, vp, vp]),
    "adsb_swap": (c_int, [vp, c_int, c_int]),
    "adsb_zero": (c_int, [vp, c_int]),
    "adsb_bind": (c_int, [vp, c_int, vp]),
    "adsb_device_ptr": (vp, [vp, c_int]),
    "adsb_row_pitch": (c_ll, [vp]),
    "adsb_set_plane": (c_int, [vp, c_int, c_int, c_int, dp]),
    "adsb_compute_rhs": (c_int, [vp, ctypes.POINTER(Form), c_int, c_int]),
    "adsb_device_count": (c_int, []),
    "adsb_slabs_create": (c_int, [c_int, ip, ip, ctypes.POINTER(vp)]),
    "adsb_slabs_destroy": (c_int, [vp]),
    "adsb_slabs_set_axis_tables": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, dp, dp, dp, dp, ip]),
    "adsb_slabs_set_axis_factor": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, c_int, dp, ip]),
    "adsb_slabs_commit": (c_int, [vp, ctypes.POINTER(Substep), c_int]),
    "adsb_slabs_upload": (c_int, [vp, dp]),
    "adsb_slabs_download": (c_int, [vp, dp]),
    "adsb_slabs_step": (c_int, [vp, c_int]),
    "adsb_slabs_synchronize": (c_int, [vp]),
    "adsb_slabs_info": (c_int, [vp, ip, ip]),
    "adsb_set_line_factors": (c_int, [vp, c_int, c_int, c_int, dp, ip]),
    "adsb_solve_special": (c_int, [vp, c_int, c_int, ip]),
    "adsb_set_point_coefficient": (c_int, [vp, dp]),
    "adsb_compute_rhs_pointwise": (c_int, [vp, ctypes.POINTER(PointForm), c_int, c_int]),
    "adsb_load_tensor": (c_int, [vp, c_int, c_int, c_int]),
    "adsb_project_init": (c_int, [vp, c_int, c_int]),
    "adsb_sample": (c_int, [vp, c_int, ip, ctypes.POINTER(dp), ctypes.POINTER(dp), dp]),
    "adsb_norm": (c_int, [vp, c_int, c_int, c_int, c_dbl, dp, dp]),
    "adsb_project_values": (c_int, [vp, c_int, c_int, c_int, dp, c_int]),
    "adsb_solve": (c_int, [vp, c_int, ip]),
    "adsb_sweep": (c_int, [vp, c_int, c_int, c_int]),
    "adsb_step": (c_int, [vp, c_int, c_int, ctypes.POINTER(Substep), c_int, c_int]),
    "adsb_enable_timing": (c_int, [vp, c_int]),
    "adsb_stage_times": (c_int, [vp, dp]),
    "adsb_launch_count": (c_ll, [vp]),
    "adsb_sweep_view": (c_int, [vp, c_int, c_int, vp, ctypes.POINTER(View), llp, vp, ctypes.POINTER(View), llp]),
    "adsb_sweep_plan": (c_int, [c_int, c_int, c_int, c_int, dp, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp]),
    "adsb_set_axis_segments": (c_int, [vp, c_int, c_int, c_int, ip, c_int, c_int]),
    "adsb_segment_info": (c_int, [vp, c_int, c_int, ip]),
    "adsb_seg_sweep_view": (c_int, [vp, c_int, c_int, c_int, vp, ctypes.POINTER(View), vp, ctypes.POINTER(View)]),
    "adsb_dist_sweep_view": (c_int, [vp, c_int, c_int, vp, ctypes.POINTER(View), ctypes.POINTER(DistArgs)]),
    "adsb_dist_sweep_check": (c_int, [vp, c_int, c_int, c_int, ctypes.POINTER(View), c_int, c_int]),
    "adsb_neighbor_barrier": (c_int, [vp, vp, vp, vp, vp]),
    "adsb_seg_dseg_view": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, vp, ctypes.POINTER(View),
                                   ctypes.POINTER(vp), c_int]),
    "adsb_seg_din_view": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, vp, ctypes.POINTER(View), vp, vp,
                                  ctypes.POINTER(vp), c_int]),
    "adsb_seg_tin": (c_int, [vp, c_int, c_int, c_int, c_int, c_ll, vp, vp]),
    "adsb_seg_correct_view": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, vp, ctypes.POINTER(View), vp,
                                      ctypes.POINTER(View), vp, vp]),
    "adsb_rhs_view": (c_int, [vp, ctypes.POINTER(Form), vp, ctypes.POINTER(View), ip, vp, vp,
                              ctypes.POINTER(View), ip]),
}

EXPORTS = tuple(_SIGNATURES)
_lib = None


def load():
    """Load libadsb200.so (built by __graft_entry__.build() / make -C iga_ads_b200/csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc < 0:
        raise AdsbError(rc, load().adsb_last_error().decode())
    return rc


def d_(a):
    return a.ctypes.data_as(dp)


def i_(a):
    return a.ctypes.data_as(ip)
