"""Host-side mirror of the reference's setup layer (names follow the reference):

    dim_config, dimension           include/ads/simulation/config.hpp:11-57, dimension.hpp:20-59
    band_matrix / factorize         include/ads/lin/band_matrix.hpp, band_solve.hpp:16-18

All arithmetic is done by libadsb200.so's host entry points (csrc/host_setup.cpp); this module only
holds the arrays.
"""
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check, d_, i_


def gauss(q):
    x, w = np.zeros(q), np.zeros(q)
    check(_lib.load().adsb_gauss(q, d_(x), d_(w)))
    return x, w


def knots(p, elements, a=0.0, b=1.0):
    k = np.zeros(elements + 2 * p + 1)
    n = check(_lib.load().adsb_knots(p, elements, a, b, d_(k)))
    return k[:n]


def find_span(x, knot, p):
    k = np.ascontiguousarray(knot, dtype=np.float64)
    return _lib.load().adsb_find_span(float(x), d_(k), len(k), p)


def basis_ders(span, x, knot, p, ders):
    k = np.ascontiguousarray(knot, dtype=np.float64)
    out = np.zeros((ders + 1, p + 1))
    check(_lib.load().adsb_basis_ders(span, float(x), d_(k), p, ders, d_(out)))
    return out


def basis_tables(p, elements, a=0.0, b=1.0, q=None, ders=1):
    q = q or p + 1
    t = dict(b=np.zeros((elements, q, ders + 1, p + 1)), x=np.zeros((elements, q)), w=np.zeros(q),
             J=np.zeros(elements), first_dof=np.zeros(elements, dtype=np.int32))
    check(_lib.load().adsb_basis_tables(p, elements, a, b, q, ders, d_(t["b"]), d_(t["x"]), d_(t["w"]),
                                        d_(t["J"]), i_(t["first_dof"])))
    return t


def matrix_1d(kind, p, elements, a=0.0, b=1.0, h=0.0, fix=0):
    """Band storage with factor workspace, one row of the array per matrix column: ab[j, 2p+i-j]."""
    ab = np.zeros((elements + p, 3 * p + 1))
    check(_lib.load().adsb_matrix_1d(kind, p, elements, a, b, h, fix, d_(ab)))
    return ab


def band_unpivot(lu, ipiv, kl, ku):
    """the factor the sweeps run with when the interchanges of (lu, ipiv) can be dropped safely, else None"""
    lu = np.ascontiguousarray(lu, dtype=np.float64)
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
    out = np.zeros_like(lu)
    rc = check(_lib.load().adsb_band_unpivot(lu.shape[0], kl, ku, lu.shape[1], d_(lu), i_(ipiv), d_(out)))
    return out if rc == 0 else None


def to_band(dense, kl, ku):
    dense = np.asarray(dense, dtype=np.float64)
    n = dense.shape[0]
    ab = np.zeros((n, 2 * kl + ku + 1))
    for i in range(n):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            ab[j, kl + ku + i - j] = dense[i, j]
    return ab


def band_factorize(ab, kl, ku):
    """lin::factorize: returns (lu, ipiv) -- raises AdsbError(ESINGULAR) on a zero pivot."""
    lu = np.array(ab, dtype=np.float64, order="C")
    n = lu.shape[0]
    ipiv = np.zeros(n, dtype=np.int32)
    check(_lib.load().adsb_band_factorize(n, kl, ku, d_(lu), lu.shape[1], i_(ipiv)))
    return lu, ipiv


def segment_bounds(ipiv, kl, nseg, align=1):
    """Balanced cuts of a line into `nseg` segments that no row interchange of the factor crosses."""
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
    bounds = np.zeros(nseg + 1, dtype=np.int32)
    check(_lib.load().adsb_segment_bounds(len(ipiv), kl, i_(ipiv), nseg, align, i_(bounds)))
    return bounds


def segment_plan(lu, ipiv, kl, ku, bounds, tol=0.0):
    """Tables of the segmented substitution (see adsb_segment_plan in include/adsb200.h)."""
    lu = np.ascontiguousarray(lu, dtype=np.float64)
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
    bounds = np.ascontiguousarray(bounds, dtype=np.int32)
    n, S = lu.shape[0], len(bounds) - 1
    dims = np.zeros(8, dtype=np.int32)
    lib = _lib.load()
    check(lib.adsb_segment_plan(n, kl, ku, lu.shape[1], d_(lu), i_(ipiv), S, i_(bounds), tol, i_(dims),
                                None, None, None, None, None))
    KL, KD, piv, _, DF, DB = (int(v) for v in dims[:6])
    t = dict(KL=KL, KD=KD, piv=piv, S=S, DF=DF, DB=DB, bounds=bounds, E=np.zeros((S, KL, KL)),
             Wf=np.zeros((S, DF, KL, KL)), Vb=np.zeros((S, DB, KD, KD)), XiF=np.zeros((S, KD, KL)),
             cf=np.zeros((n, KD + KL)))
    check(lib.adsb_segment_plan(n, kl, ku, lu.shape[1], d_(lu), i_(ipiv), S, i_(bounds), tol, i_(dims),
                                d_(t["E"]), d_(t["Wf"]), d_(t["Vb"]), d_(t["XiF"]), d_(t["cf"])))
    return t


@dataclass
class dim_config:
    """include/ads/simulation/config.hpp:11-31"""
    p: int
    elements: int
    a: float = 0.0
    b: float = 1.0
    quad_order: int = 0
    repeated_nodes: int = 0

    def __post_init__(self):
        if self.quad_order == 0:
            self.quad_order = self.p + 1
        if self.repeated_nodes:
            raise NotImplementedError("repeated knots are outside the ADS-step path")


@dataclass
class timesteps_config:
    step_count: int
    dt: float


class dimension:
    """ads::dimension: one axis = B-spline basis + Gram matrix M + its quadrature tables.

    M is the band matrix in factor layout; fix_left()/fix_right() edit it exactly like
    dimension::fix_dof; factorize_matrix() returns and caches (lu, ipiv)."""

    def __init__(self, config, derivatives=1):
        self.p, self.elements, self.a, self.b = config.p, config.elements, config.a, config.b
        self.quad_order = config.quad_order
        self.derivatives = derivatives
        self.knot = knots(self.p, self.elements, self.a, self.b)
        self.basis = basis_tables(self.p, self.elements, self.a, self.b, self.quad_order, derivatives)
        self.M = matrix_1d(0, self.p, self.elements, self.a, self.b)
        self.lu = None
        self.ipiv = None

    def dofs(self):
        return self.elements + self.p

    def fix_dof(self, k):
        p, last = self.p, self.dofs() - 1
        for i in range(max(k - p, 0), min(k + p, last) + 1):
            self.M[i, 2 * p + k - i] = 0.0
        self.M[k, 2 * p] = 1.0

    def fix_left(self):
        self.fix_dof(0)

    def fix_right(self):
        self.fix_dof(self.dofs() - 1)

    def factorize_matrix(self):
        self.lu, self.ipiv = band_factorize(self.M, self.p, self.p)
        return self.lu, self.ipiv

    def matrix(self, kind, h=0.0):
        """Another 1-D quadrature matrix on this axis (1 stiffness, 2 advection, 3 M + h*S)."""
        return matrix_1d(kind, self.p, self.elements, self.a, self.b, h)
