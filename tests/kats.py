"""Known-answer tests restated from the reference's own Catch2 suite (data only).

Each entry cites the reference test it comes from.  Matrices are given as dense rows and
converted to the LAPACK band layout of include/ads/lin/band_matrix.hpp:31-40,:69-73
(ldab = 2kl+ku+1, A(i,j) at ab[j, kl+ku+i-j]); tensors are in memory order (first index fastest).
"""
import numpy as np


def to_band(dense, kl, ku):
    dense = np.asarray(dense, dtype=np.float64)
    n = dense.shape[0]
    ab = np.zeros((n, 2 * kl + ku + 1))
    for i in range(n):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            ab[j, kl + ku + i - j] = dense[i, j]
    return ab


# tests/ads/solver_test.cpp:63-99, :101-162, :164-255 -- Mx requires pivoting
MX = np.array([[1, 2, 0, 0], [2, 3, 1, 0], [0, -1, 4, 0], [0, 0, 1, 3]], dtype=float)
MY = np.array([[2, 3, 0], [1, 2, -3], [0, 2, 2]], dtype=float)
MZ = np.array([[1, 2], [3, 4]], dtype=float)

ADS_1D = dict(shape=(4,), mats=[MX], rhs=[5, 11, 10, 15], expected=[1, 2, 3, 4])
ADS_2D = dict(shape=(4, 3), mats=[MX, MY],
              rhs=[61, 127, 86, 123, -48, -96, -48, -64, 92, 188, 112, 156],
              expected=list(range(1, 13)))
ADS_3D = dict(shape=(4, 3, 2), mats=[MX, MY, MZ],
              rhs=[543, 1101, 618, 849, -144, -288, -144, -192, 564, 1140, 624, 852,
                   1147, 2329, 1322, 1821, -336, -672, -336, -448, 1220, 2468, 1360, 1860],
              expected=list(range(1, 25)))


def band_solve_kat():
    """tests/ads/lin/band_solve_test.cpp:16-46: kl=1, ku=2, n=6, 4 right-hand sides, abs 1e-5."""
    kl, ku, n, d = 1, 2, 6, 4
    dense = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            dense[i, j] = (i + 1) * 10 + j + 1
    b = np.array([[(j + 1) * (i + 1) for i in range(n)] for j in range(d)], dtype=float)  # [rhs][i]
    sol = np.array([0.230377, -0.126052, -0.0016554, -0.00111222, 0.203603, -0.109609])
    x = np.array([(j + 1) * sol for j in range(d)])
    return kl, ku, dense, b, x


def rotation_kat():
    """tests/ads/lin/tensor_test.cpp:49-80: a(k=2, n=3, m=2) -> e(n, m, k)."""
    k, n, m = 2, 3, 2
    a = np.zeros(k * n * m)
    e = np.zeros(n * m * k)
    for i0 in range(k):
        for i1 in range(n):
            for i2 in range(m):
                v = 100 * (i0 + 1) + 10 * (i1 + 1) + (i2 + 1)
                a[i0 + k * (i1 + n * i2)] = v
                e[i1 + n * (i2 + m * i0)] = v
    return (k, n, m), a, e
