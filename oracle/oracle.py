"""ctypes loaders for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``  : oracle/libads_oracle.so, the plain-C restatement (oracle/ads_oracle.c)
* ``Ref``     : oracle/_ref/libads_ref.so, the unmodified reference sources compiled by
                oracle/Makefile (needs /root/reference at BUILD time only)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (iga_ads_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_c_int = ctypes.c_int
_c_dbl = ctypes.c_double


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def build(ref=True):
    """Compile the C restatement and (when /root/reference exists) the reference .so."""
    subprocess.run(["make", "-s", "-C", HERE, "libads_oracle.so"], check=True)
    if ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


PROBLEMS = {"heat_3d": 0, "heat_2d": 1, "implicit_2d": 2, "scalability_3d": 3,
            "scalability_2d": 4, "implicit_3d": 5}
NDIM = {0: 3, 1: 2, 2: 2, 3: 3, 4: 2, 5: 3}


class _Base:
    def _ptr_arrays(self, mats, ipivs):
        nd = len(mats)
        ab_arr = (_dp * nd)(*[_d(m) for m in mats])
        ip_arr = (_ip * nd)(*[_i(p) for p in ipivs])
        return ab_arr, ip_arr


class Oracle(_Base):
    """The C restatement."""

    def __init__(self):
        path = os.path.join(HERE, "libads_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = ctypes.CDLL(path)
        L.orc_gauss.argtypes = [_c_int, _dp, _dp]
        L.orc_knots.argtypes = [_c_int, _c_int, _c_dbl, _c_dbl, _dp]
        L.orc_find_span.argtypes = [_c_dbl, _dp, _c_int, _c_int]
        L.orc_basis_ders.argtypes = [_c_int, _c_dbl, _dp, _c_int, _c_int, _dp]
        L.orc_basis_ders.restype = None
        L.orc_basis_tables.argtypes = [_c_int, _c_int, _c_dbl, _c_dbl, _c_int, _c_int, _dp, _dp,
                                       _dp, _dp, _ip]
        L.orc_matrix_1d.argtypes = [_c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _c_int, _dp]
        L.orc_dgbtrf.argtypes = [_c_int, _c_int, _c_int, _dp, _c_int, _ip]
        L.orc_dgbtrs.argtypes = [_c_int, _c_int, _c_int, _c_int, _dp, _c_int, _ip, _dp, _c_int]
        L.orc_cyclic_transpose.argtypes = [_c_int, _ip, _dp, _dp]
        L.orc_cyclic_transpose.restype = None
        L.orc_ads_solve.argtypes = [_c_int, _ip, _ip, _ip, ctypes.POINTER(_dp),
                                    ctypes.POINTER(_ip), _dp, _dp]
        L.orc_run.argtypes = [_c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int, _dp, _dp]
        L.orc_project_init.argtypes = [_c_int, _c_int, _c_int, _dp]

    def gauss(self, q):
        x, w = np.zeros(q), np.zeros(q)
        assert self.lib.orc_gauss(q, _d(x), _d(w)) == 0
        return x, w

    def knots(self, p, elements, a=0.0, b=1.0):
        k = np.zeros(elements + 2 * p + 1)
        n = self.lib.orc_knots(p, elements, a, b, _d(k))
        return k[:n]

    def find_span(self, x, knots, p):
        k = np.ascontiguousarray(knots, dtype=np.float64)
        return self.lib.orc_find_span(float(x), _d(k), len(k), p)

    def basis_ders(self, span, x, knots, p, ders):
        k = np.ascontiguousarray(knots, dtype=np.float64)
        out = np.zeros((ders + 1, p + 1))
        self.lib.orc_basis_ders(span, float(x), _d(k), p, ders, _d(out))
        return out

    def basis_tables(self, p, elements, a=0.0, b=1.0, q=None, ders=1):
        q = q or p + 1
        bt = np.zeros((elements, q, ders + 1, p + 1))
        xq = np.zeros((elements, q))
        w = np.zeros(q)
        J = np.zeros(elements)
        fd = np.zeros(elements, dtype=np.int32)
        assert self.lib.orc_basis_tables(p, elements, a, b, q, ders, _d(bt), _d(xq), _d(w), _d(J),
                                         _i(fd)) == 0
        return dict(b=bt, x=xq, w=w, J=J, first_dof=fd)

    def matrix_1d(self, kind, p, elements, a=0.0, b=1.0, h=0.0, fix=0):
        n = elements + p
        ab = np.zeros((n, 3 * p + 1))  # column j at ab[j, :]
        self.lib.orc_matrix_1d(kind, p, elements, a, b, h, fix, _d(ab))
        return ab

    def factorize(self, ab, kl, ku):
        ab = np.array(ab, dtype=np.float64, order="C")
        n = ab.shape[0]
        ipiv = np.zeros(n, dtype=np.int32)
        info = self.lib.orc_dgbtrf(n, kl, ku, _d(ab), ab.shape[1], _i(ipiv))
        return ab, ipiv, info

    def solve_factorized(self, ab, ipiv, kl, ku, b):
        """b: array whose memory is nrhs contiguous columns of length n (modified copy returned)."""
        n = ab.shape[0]
        x = np.array(b, dtype=np.float64, order="C").ravel().copy()
        nrhs = x.size // n
        ab = np.ascontiguousarray(ab)
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        self.lib.orc_dgbtrs(n, kl, ku, nrhs, _d(ab), ab.shape[1], _i(ipiv), _d(x), n)
        return x

    def cyclic_transpose(self, shape, data):
        sh = np.array(shape, dtype=np.int32)
        src = np.ascontiguousarray(data, dtype=np.float64).ravel()
        out = np.zeros_like(src)
        self.lib.orc_cyclic_transpose(len(shape), _i(sh), _d(src), _d(out))
        return out

    def ads_solve(self, shape, mats, ipivs, kls, kus, rhs):
        nd = len(shape)
        sh = np.array(shape, dtype=np.int32)
        kl = np.array(kls, dtype=np.int32)
        ku = np.array(kus, dtype=np.int32)
        mats = [np.ascontiguousarray(m) for m in mats]
        ipivs = [np.ascontiguousarray(p, dtype=np.int32) for p in ipivs]
        ab_arr, ip_arr = self._ptr_arrays(mats, ipivs)
        x = np.array(rhs, dtype=np.float64).ravel().copy()
        buf = np.zeros_like(x)
        assert self.lib.orc_ads_solve(nd, _i(sh), _i(kl), _i(ku), ab_arr, ip_arr, _d(x), _d(buf)) == 0
        return x

    def run(self, problem, p, elements, dt, nsteps, u0=None, stage=0):
        pid = PROBLEMS[problem] if isinstance(problem, str) else problem
        n = elements + p
        N = n ** NDIM[pid]
        tm = np.zeros(8)
        if u0 is None:
            u = np.zeros(N)
            mode = 1
        else:
            u = np.array(u0, dtype=np.float64).ravel().copy()
            assert u.size == N
            mode = 0
        rc = self.lib.orc_run(pid, p, elements, dt, nsteps, mode, stage, _d(u), _d(tm))
        assert rc == 0
        return u, tm

    def project_init(self, problem, p, elements):
        pid = PROBLEMS[problem]
        n = elements + p
        u = np.zeros(n ** NDIM[pid])
        assert self.lib.orc_project_init(pid, p, elements, _d(u)) == 0
        return u


class Ref(_Base):
    """The compiled, unmodified reference."""

    @staticmethod
    def path():
        return os.path.join(HERE, "_ref", "libads_ref.so")

    @classmethod
    def available(cls):
        if not os.path.exists(cls.path()):
            return False
        try:
            ctypes.CDLL(cls.path())
            return True
        except OSError:
            return False

    def __init__(self):
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        self.lib = L = ctypes.CDLL(self.path())
        L.ref_gauss.argtypes = [_c_int, _dp, _dp]
        L.ref_dofs.argtypes = [_c_int, _c_int]
        L.ref_basis_tables.argtypes = [_c_int, _c_int, _c_dbl, _c_dbl, _c_int, _c_int, _dp, _dp,
                                       _dp, _dp, _ip, _dp]
        L.ref_matrix_1d.argtypes = [_c_int, _c_int, _c_int, _c_dbl, _c_dbl, _c_dbl, _c_int, _dp]
        L.ref_factorize.argtypes = [_c_int, _c_int, _c_int, _dp, _ip]
        L.ref_solve_factorized.argtypes = [_c_int, _c_int, _c_int, _dp, _ip, _dp, _c_int]
        L.ref_cyclic_transpose.argtypes = [_c_int, _ip, _dp, _dp]
        L.ref_ads_solve.argtypes = [_c_int, _ip, _ip, _ip, ctypes.POINTER(_dp),
                                    ctypes.POINTER(_ip), _dp]
        L.ref_set_threads.argtypes = [_c_int]
        for name in ("ref_heat3d", "ref_heat2d", "ref_implicit2d", "ref_scalability3d",
                     "ref_scalability2d", "ref_implicit3d"):
            getattr(L, name).argtypes = [_c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int, _dp, _dp]
        L.ref_time_heat3d_hoisted.argtypes = [_c_int, _c_int, _c_dbl, _c_int, _dp, _dp]
        if hasattr(L, "ref_flow"):
            L.ref_flow.argtypes = [_c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int, _dp, _dp, _dp]
        self._run = {0: L.ref_heat3d, 1: L.ref_heat2d, 2: L.ref_implicit2d,
                     3: L.ref_scalability3d, 4: L.ref_scalability2d, 5: L.ref_implicit3d}

    def set_threads(self, n):
        self.lib.ref_set_threads(n)

    def gauss(self, q):
        x, w = np.zeros(q), np.zeros(q)
        assert self.lib.ref_gauss(q, _d(x), _d(w)) == 0
        return x, w

    def basis_tables(self, p, elements, a=0.0, b=1.0, q=None, ders=1):
        q = q or p + 1
        bt = np.zeros((elements, q, ders + 1, p + 1))
        xq = np.zeros((elements, q))
        w = np.zeros(q)
        J = np.zeros(elements)
        fd = np.zeros(elements, dtype=np.int32)
        knots = np.zeros(elements + 2 * p + 1)
        self.lib.ref_basis_tables(p, elements, a, b, q, ders, _d(bt), _d(xq), _d(w), _d(J), _i(fd),
                                  _d(knots))
        return dict(b=bt, x=xq, w=w, J=J, first_dof=fd, knots=knots)

    def matrix_1d(self, kind, p, elements, a=0.0, b=1.0, h=0.0, fix=0):
        n = elements + p
        ab = np.zeros((n, 3 * p + 1))
        self.lib.ref_matrix_1d(kind, p, elements, a, b, h, fix, _d(ab))
        return ab

    def factorize(self, ab, kl, ku):
        ab = np.array(ab, dtype=np.float64, order="C")
        n = ab.shape[0]
        ipiv = np.zeros(n, dtype=np.int32)
        info = self.lib.ref_factorize(n, kl, ku, _d(ab), _i(ipiv))
        return ab, ipiv, info

    def solve_factorized(self, ab, ipiv, kl, ku, b):
        n = ab.shape[0]
        x = np.array(b, dtype=np.float64, order="C").ravel().copy()
        nrhs = x.size // n
        ab = np.ascontiguousarray(ab)
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        self.lib.ref_solve_factorized(n, kl, ku, _d(ab), _i(ipiv), _d(x), nrhs)
        return x

    def cyclic_transpose(self, shape, data):
        sh = np.array(shape, dtype=np.int32)
        src = np.ascontiguousarray(data, dtype=np.float64).ravel()
        out = np.zeros_like(src)
        assert self.lib.ref_cyclic_transpose(len(shape), _i(sh), _d(src), _d(out)) == 0
        return out

    def ads_solve(self, shape, mats, ipivs, kls, kus, rhs):
        nd = len(shape)
        sh = np.array(shape, dtype=np.int32)
        kl = np.array(kls, dtype=np.int32)
        ku = np.array(kus, dtype=np.int32)
        mats = [np.ascontiguousarray(m) for m in mats]
        ipivs = [np.ascontiguousarray(p, dtype=np.int32) for p in ipivs]
        ab_arr, ip_arr = self._ptr_arrays(mats, ipivs)
        x = np.array(rhs, dtype=np.float64).ravel().copy()
        assert self.lib.ref_ads_solve(nd, _i(sh), _i(kl), _i(ku), ab_arr, ip_arr, _d(x)) == 0
        return x

    def run(self, problem, p, elements, dt, nsteps, u0=None, stage=0):
        pid = PROBLEMS[problem] if isinstance(problem, str) else problem
        n = elements + p
        N = n ** NDIM[pid]
        tm = np.zeros(8)
        if u0 is None:
            u = np.zeros(N)
            mode = 1
        else:
            u = np.array(u0, dtype=np.float64).ravel().copy()
            assert u.size == N
            mode = 0
        rc = self._run[pid](p, elements, dt, nsteps, mode, stage, _d(u), _d(tm))
        assert rc == 0
        return u, tm


    def flow(self, p, elements, dt, nsteps, u0=None, stage=0):
        """examples/flow/flow.hpp through the compiled reference: (u, permeability table at the quadrature
        points, x fastest).  u0 None: the shipped before(); stage 1: compute_rhs alone."""
        n, nq = elements + p, elements * (p + 1)
        kq = np.zeros(nq ** 3)
        tm = np.zeros(8)
        if u0 is None:
            u, mode = np.zeros(n ** 3), 1
        else:
            u, mode = np.array(u0, dtype=np.float64).ravel().copy(), 0
            assert u.size == n ** 3
        assert self.lib.ref_flow(p, elements, dt, nsteps, mode, stage, _d(u), _d(kq), _d(tm)) == 0
        return u, kq


def pointwise_rhs(tables, u_prev, form, plain=None):
    """numpy restatement of the reference's element loop for a general pointwise form -- zero(rhs); for e, q:
    u = eval_fun(u_prev, e, q); for a: rhs(a) += (k0 v + k . grad v) w J  (examples/scalability/test3d.hpp:66-95,
    examples/flow/flow.hpp:74-101; eval_fun / eval_basis: include/ads/simulation/simulation_3d.hpp:64-136) --
    written with dense per-axis evaluation matrices B[point, dof], so it shares nothing with the kernels.
    tables: basis_tables() of every axis; form(u, ux, uy[, uz], x, y[, z]) -> (k0, k1, k2[, k3]) on arrays shaped
    [x points, y points(, z points)]; plain (optional): the same signature, one array added to every DOF of the
    point's element WITHOUT the test function (test3d.hpp:86-88).  Returns rhs, first index fastest."""
    nd = len(tables)
    B, dB, S, wJ, X = [], [], [], [], []
    for t in tables:
        ne, q, _, m = t["b"].shape
        n = ne + m - 1
        b, db, sup = np.zeros((ne * q, n)), np.zeros((ne * q, n)), np.zeros((ne * q, n))
        for e in range(ne):
            f = int(t["first_dof"][e])
            b[e * q:(e + 1) * q, f:f + m] = t["b"][e, :, 0, :]
            db[e * q:(e + 1) * q, f:f + m] = t["b"][e, :, 1, :]
            sup[e * q:(e + 1) * q, f:f + m] = 1.0
        B.append(b), dB.append(db), S.append(sup)
        wJ.append((t["w"][None, :] * t["J"][:, None]).ravel())
        X.append(t["x"].ravel())
    shape = tuple(b.shape[1] for b in B)
    U = np.asarray(u_prev, dtype=np.float64).reshape(shape, order="F")
    if nd == 3:
        ev = lambda a, b, c, T: np.einsum("ai,bj,ck,ijk->abc", a, b, c, T, optimize=True)      # noqa: E731
        tr = lambda a, b, c, K: np.einsum("ai,bj,ck,abc->ijk", a, b, c, K, optimize=True)      # noqa: E731
        W = wJ[0][:, None, None] * wJ[1][None, :, None] * wJ[2][None, None, :]
        pts = (X[0][:, None, None], X[1][None, :, None], X[2][None, None, :])
        vals = (ev(B[0], B[1], B[2], U), ev(dB[0], B[1], B[2], U), ev(B[0], dB[1], B[2], U), ev(B[0], B[1], dB[2], U))
        k = form(*vals, *pts)
        rhs = (tr(B[0], B[1], B[2], k[0] * W) + tr(dB[0], B[1], B[2], k[1] * W) + tr(B[0], dB[1], B[2], k[2] * W)
               + tr(B[0], B[1], dB[2], k[3] * W))
        if plain is not None:
            rhs = rhs + tr(S[0], S[1], S[2], np.broadcast_to(plain(*vals, *pts), W.shape) * W)
    else:
        ev = lambda a, b, T: np.einsum("ai,bj,ij->ab", a, b, T, optimize=True)      # noqa: E731
        tr = lambda a, b, K: np.einsum("ai,bj,ab->ij", a, b, K, optimize=True)      # noqa: E731
        W = wJ[0][:, None] * wJ[1][None, :]
        pts = (X[0][:, None], X[1][None, :])
        vals = (ev(B[0], B[1], U), ev(dB[0], B[1], U), ev(B[0], dB[1], U))
        k = form(*vals, *pts)
        rhs = tr(B[0], B[1], k[0] * W) + tr(dB[0], B[1], k[1] * W) + tr(B[0], dB[1], k[2] * W)
        if plain is not None:
            rhs = rhs + tr(S[0], S[1], np.broadcast_to(plain(*vals, *pts), W.shape) * W)
    return rhs.ravel(order="F").copy()


def kronecker_heat_rhs(oracle, p, elements, dt, u):
    """the explicit heat right-hand side (examples/heat/heat_3d.hpp:49-67) WITHOUT any element loop:
    rhs = (Mx (x) My (x) Mz) u - dt (Sx (x) My (x) Mz + Mx (x) Sy (x) Mz + Mx (x) My (x) Sz) u with the oracle's own 1-D
    Gram / stiffness matrices (src/ads/form_matrix.cpp:8-42) as scipy sparse matrices applied axis by axis -- an
    independent route to the same numbers (equal to the oracle's element loop to 8e-16 at 12^3, see
    tests/test_oracle.py) that is cheap enough for the full 514^3 tensor (~30 s, ~9 GB)."""
    import scipy.sparse as sp

    n = elements + p

    def csr(ab):
        rows, cols, vals = [], [], []
        for j in range(n):
            for i in range(max(0, j - p), min(n, j + p + 1)):
                rows.append(i), cols.append(j), vals.append(ab[j, 2 * p + i - j])
        return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))

    M, S = csr(oracle.matrix_1d(0, p, elements)), csr(oracle.matrix_1d(1, p, elements))
    U = np.asarray(u, dtype=np.float64).reshape(n, n, n)      # [z][y][x]: x fastest in memory

    def apply(A, T, axis):
        if axis == 0:
            return (A @ T.reshape(n, -1)).reshape(T.shape)
        if axis == 2:
            return (T.reshape(-1, n) @ A.T).reshape(T.shape)
        out = np.empty_like(T)
        for k in range(n):
            out[k] = A @ T[k]
        return out

    X, XS = apply(M, U, 2), apply(S, U, 2)
    Y, YS, YX = apply(M, X, 1), apply(S, X, 1), apply(M, XS, 1)
    del X, XS
    out = apply(M, Y, 0) - dt * (apply(M, YX, 0) + apply(M, YS, 0) + apply(S, Y, 0))
    return out.ravel()


def flow_form(dt, kq, mi=10.0):
    """integrand of examples/flow/flow.hpp:74-101 for pointwise_rhs; kq: permeability at the points [x, y, z]"""
    def form(u, ux, uy, uz, x, y, z):
        h = 1 + np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.sin(2 * np.pi * z)
        e = -dt * kq * np.exp(mi * u)
        return u + dt * h, e * ux, e * uy, e * uz
    return form


def synthetic_state(shape, seed=20260101):
    """Synthetic initial coefficient tensor: smooth part + 0.1 * noise,
    in memory order (first index fastest).  numpy's PCG64 stands in for std::mt19937_64 -- the
    same array is handed to the oracle and to the GPU path, so only determinism matters."""
    rng = np.random.default_rng(seed)
    nd = len(shape)
    grids = np.meshgrid(*[np.linspace(0.0, 1.0, n) for n in shape], indexing="ij")
    if nd == 3:
        smooth = np.sin(3 * grids[0]) * np.cos(2 * grids[1]) * (1 + grids[2] ** 2)
    else:
        smooth = np.sin(3 * grids[0]) * np.cos(2 * grids[1])
    u = smooth + 0.1 * rng.uniform(-1, 1, size=shape)
    return np.asfortranarray(u).ravel(order="F").copy()


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
