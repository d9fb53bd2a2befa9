"""Output path of the reference, fed from the device: sampling of the spline on a regular grid and the text
writers either side of it.

    output_manager<2>.write / output_manager<3>.write   include/ads/output_manager.hpp:66-73,:101-118
    output::axis (linspace of the basis' range)          include/ads/output/axis.hpp:23-26, include/ads/util.hpp:28-34
    VTK ImageData, ASCII                                  include/ads/output/vtk.hpp:45-82
    gnuplot rows  x y value                               include/ads/output/gnuplot.hpp:45-56
    values: fixed, precision 10, width 18                 include/ads/output_manager.hpp:23 (DEFAULT_FMT)

The values come from adsb_sample (spans / basis values on the host as bspline::eval does, contraction on the
device); nothing here touches the coefficient tensor on the host.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, d_, dp, i_


def linspace(a, b, intervals):
    """ads::linspace (include/ads/util.hpp:28-34): lerp(i, n, a, b) = (1 - i/n) a + (i/n) b, intervals + 1 points"""
    t = np.arange(intervals + 1, dtype=np.float64) / np.float64(intervals)
    return (1 - t) * a + t * b


class output_manager:
    """output_manager<Dim>(bases..., n): samples on (n+1)^Dim points of the domain and writes the reference's files"""

    def __init__(self, sim, intervals):
        self.sim = sim
        n = [intervals] * len(sim.dims) if np.isscalar(intervals) else list(intervals)
        self.points = [linspace(d.a, d.b, k) for d, k in zip(sim.dims, n)]

    def evaluate(self, buf=0):
        """values[i, j(, k)] = u_h(x_i, y_j(, z_k)) as a Fortran-ordered array (first index fastest, the reference's tensor)"""
        ctx = self.sim._context()
        nd = len(self.points)
        npts = np.array([len(p) for p in self.points], dtype=np.int32)
        pts = [np.ascontiguousarray(p, dtype=np.float64) for p in self.points]
        kn = [np.ascontiguousarray(d.knot, dtype=np.float64) for d in self.sim.dims]
        parr = (dp * nd)(*[d_(p) for p in pts])
        karr = (dp * nd)(*[d_(k) for k in kn])
        out = np.zeros(int(np.prod(npts)))
        check(ctx.lib.adsb_sample(ctx.h, buf, i_(npts), parr, karr, d_(out)))
        return out.reshape(tuple(int(v) for v in npts), order="F")

    def write(self, stream, buf=0):
        vals = self.evaluate(buf)
        if vals.ndim == 3:
            write_vtk(stream, vals)
        else:
            write_gnuplot_2d(stream, self.points[0], self.points[1], vals)

    def to_file(self, filename, buf=0):
        with open(filename, "w") as f:
            self.write(f, buf)


def write_vtk(stream, vals):
    """include/ads/output/vtk.hpp:45-82: ImageData header, one value per row in memory order (first index fastest)"""
    ext = " ".join(f"0 {n - 1}" for n in vals.shape)
    stream.write('<?xml version="1.0"?>\n')
    stream.write('<VTKFile type="ImageData" version="0.1">\n')
    stream.write(f'  <ImageData WholeExtent="{ext}" origin="0 0 0" spacing="1 1 1">\n')
    stream.write(f'    <Piece Extent="{ext}">\n')
    stream.write('      <PointData Scalars="Result">\n')
    stream.write('        <DataArray Name="Result"  type="Float32" format="ascii" NumberOfComponents="1">\n')
    np.savetxt(stream, vals.ravel(order="F"), fmt="%18.10f")
    stream.write("        </DataArray>\n      </PointData>\n    </Piece>\n  </ImageData>\n</VTKFile>\n")


def write_gnuplot_2d(stream, xs, ys, vals):
    """include/ads/output/gnuplot.hpp:45-56: rows `x y value`, x outermost"""
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    np.savetxt(stream, np.column_stack([X.ravel(), Y.ravel(), vals.ravel()]), fmt="%18.10f", delimiter="")
