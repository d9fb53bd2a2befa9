"""Device context + the simulation layer of the reference, mirrored on top of the C ABI.

    Context                     adsb_ctx (include/adsb200.h)
    simulation_2d / _3d         include/ads/simulation/simulation_2d.hpp, simulation_3d.hpp:23-150
                                 (x, y, z dimensions, shape(), prepare_matrices(), solve())
    heat_3d, heat_2d,           examples/heat/heat_3d.hpp, heat_2d.hpp
    implicit_2d, implicit_3d,   examples/implicit/implicit.hpp (+ its 3-axis extension)
    scalability_2d/_3d          examples/scalability/test2d.hpp, test3d.hpp

The coefficient tensors live on the GPU (managed buffers U, U_PREV, ...); `u` is downloaded only
when the caller asks.  There is no CPU path: constructing a Context without a CUDA device raises.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import Form, Substep, View, check, d_, i_
from .host import band_factorize, dimension, dim_config, timesteps_config  # noqa: F401

U, U_PREV, FORCING, FIXROW, SCRATCH = 0, 1, 2, 3, 4


class Context:
    def __init__(self, n_global, lo=None, cnt=None, device=0):
        self.lib = _lib.load()
        self.ndim = len(n_global)
        self.n_global = tuple(int(v) for v in n_global)
        self.lo = tuple(lo) if lo is not None else (0,) * self.ndim
        self.cnt = tuple(cnt) if cnt is not None else self.n_global
        ng = np.array(self.n_global, dtype=np.int32)
        lo_a = np.array(self.lo, dtype=np.int32)
        cn = np.array(self.cnt, dtype=np.int32)
        h = ctypes.c_void_p()
        check(self.lib.adsb_create(self.ndim, i_(ng), i_(lo_a), i_(cn), device, ctypes.byref(h)))
        self.h = h
        self.local_size = int(np.prod(self.cnt))

    def close(self):
        if self.h:
            self.lib.adsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def set_stream(self, stream_ptr):
        check(self.lib.adsb_set_stream(self.h, ctypes.c_void_p(stream_ptr)))

    def set_sm_limit(self, sms):
        check(self.lib.adsb_set_sm_limit(self.h, int(sms)))

    def copy2d(self, dst_ptr, dst_pitch, src_ptr, src_pitch, width, height):
        """strided device copy on this context's stream; all sizes in bytes except `height` (rows)"""
        check(self.lib.adsb_copy2d(self.h, ctypes.c_void_p(dst_ptr), dst_pitch, ctypes.c_void_p(src_ptr), src_pitch,
                                   width, height))

    def synchronize(self):
        check(self.lib.adsb_synchronize(self.h))

    def set_axis(self, axis, dim):
        t = dim.basis
        check(self.lib.adsb_set_axis_tables(self.h, axis, dim.p, dim.elements, dim.quad_order,
                                            dim.derivatives, d_(t["b"]), d_(t["x"]), d_(t["w"]),
                                            d_(t["J"]), i_(t["first_dof"])))

    def set_factor(self, axis, slot, lu, ipiv, kl, ku):
        lu = np.ascontiguousarray(lu, dtype=np.float64)
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        check(self.lib.adsb_set_axis_factor(self.h, axis, slot, lu.shape[0], kl, ku, lu.shape[1],
                                            d_(lu), i_(ipiv)))

    # ---- tensors
    def upload(self, buf, host):
        a = np.ascontiguousarray(host, dtype=np.float64).ravel()
        assert a.size == self.local_size
        check(self.lib.adsb_upload(self.h, buf, d_(a)))

    def download(self, buf):
        out = np.empty(self.local_size)
        check(self.lib.adsb_download(self.h, buf, d_(out)))
        return out

    def upload_async(self, buf, host_ptr, stream_ptr):
        """host_ptr: pinned host memory holding local_size doubles; enqueued on the given CUDA stream"""
        check(self.lib.adsb_upload_async(self.h, buf, ctypes.c_void_p(host_ptr), ctypes.c_void_p(stream_ptr)))

    def download_async(self, buf, host_ptr, stream_ptr):
        check(self.lib.adsb_download_async(self.h, buf, ctypes.c_void_p(host_ptr), ctypes.c_void_p(stream_ptr)))

    def swap(self, a, b):
        check(self.lib.adsb_swap(self.h, a, b))

    def zero(self, buf):
        check(self.lib.adsb_zero(self.h, buf))

    def bind(self, buf, device_ptr):
        check(self.lib.adsb_bind(self.h, buf, ctypes.c_void_p(device_ptr)))

    def device_ptr(self, buf):
        p = self.lib.adsb_device_ptr(self.h, buf)
        if not p:
            raise _lib.AdsbError(-5, self.lib.adsb_last_error().decode())
        return p

    def set_plane(self, buf, axis, idx, values):
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        check(self.lib.adsb_set_plane(self.h, buf, axis, idx, d_(v)))

    # ---- the hot path
    def compute_rhs(self, form, src, dst):
        check(self.lib.adsb_compute_rhs(self.h, ctypes.byref(form), src, dst))

    def set_point_coefficient(self, values):
        """coefficient table of the pointwise forms: one value per quadrature point of the domain, x fastest"""
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        check(self.lib.adsb_set_point_coefficient(self.h, d_(v)))

    def compute_rhs_pointwise(self, form, src, dst):
        """general pointwise form by brick quadrature (adsb_compute_rhs_pointwise)"""
        check(self.lib.adsb_compute_rhs_pointwise(self.h, ctypes.byref(form), src, dst))

    def load_tensor(self, source, with_test_function, dst):
        check(self.lib.adsb_load_tensor(self.h, source, int(with_test_function), dst))

    def project_init(self, state, dst):
        check(self.lib.adsb_project_init(self.h, state, dst))

    def norm(self, buf, kind="L2", ref=0, t=0.0, ref_values=None):
        """(norm of u_h - ref, norm of ref) by element quadrature: basic_simulation_Nd::norm / error
        (include/ads/simulation/basic_simulation_3d.hpp:281-398).  ref: 0 none, 1 validation solution at time t,
        2 `ref_values` tabulated at the quadrature points (L2 only)"""
        out = np.zeros(2)
        tab = None
        if ref_values is not None:
            tab = np.ascontiguousarray(ref_values, dtype=np.float64).ravel()
            ref = 2
        check(self.lib.adsb_norm(self.h, buf, {"L2": 0, "H1": 1}[kind], ref, float(t), d_(tab) if tab is not None else None,
                                 d_(out)))
        return float(out[0]), float(out[1])

    def project_values(self, dst, values, ez_lo=0, ez_cnt=None, accumulate=False):
        """L2-projection right-hand side of a function tabulated at the quadrature points (adsb_project_values)"""
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        check(self.lib.adsb_project_values(self.h, dst, ez_lo, ez_cnt or 0, d_(v), int(accumulate)))

    def solve(self, buf, slots=None):
        s = np.array(list(slots or ()) + [0] * (3 - len(slots or ())), dtype=np.int32)
        check(self.lib.adsb_solve(self.h, buf, i_(s)))

    def set_line_factors(self, axis, lus, ipivs, kl, ku):
        """generalised ADS: lus[l] / ipivs[l] = band_factorize() output of line l's matrix along `axis`"""
        ab = np.ascontiguousarray(lus, dtype=np.float64)
        pv = np.ascontiguousarray(ipivs, dtype=np.int32)
        check(self.lib.adsb_set_line_factors(self.h, axis, kl, ku, d_(ab), i_(pv)))

    def solve_special(self, buf, special_axis, slots=None):
        s = np.array(list(slots or ()) + [0] * (3 - len(slots or ())), dtype=np.int32)
        check(self.lib.adsb_solve_special(self.h, buf, special_axis, i_(s)))

    def sweep(self, buf, axis, slot=0):
        check(self.lib.adsb_sweep(self.h, buf, axis, slot))

    def step(self, substeps, nsteps, u=U, u_prev=U_PREV):
        arr = (Substep * len(substeps))(*substeps)
        check(self.lib.adsb_step(self.h, u, u_prev, arr, len(substeps), nsteps))

    def sweep_view(self, axis, slot, in_ptr, vin, out_ptr, vout, off_in=None, off_out=None):
        oi = np.ascontiguousarray(off_in, dtype=np.int64) if off_in is not None else None
        oo = np.ascontiguousarray(off_out, dtype=np.int64) if off_out is not None else None
        check(self.lib.adsb_sweep_view(
            self.h, axis, slot, ctypes.c_void_p(in_ptr), ctypes.byref(vin),
            oi.ctypes.data_as(_lib.llp) if oi is not None else None,
            ctypes.c_void_p(out_ptr), ctypes.byref(vout),
            oo.ctypes.data_as(_lib.llp) if oo is not None else None))

    def rhs_view(self, form, in_ptr, vin, in_lo, out_ptr, vout, out_lo, forcing_ptr=None):
        il = np.array(list(in_lo) + [0] * (3 - len(in_lo)), dtype=np.int32)
        ol = np.array(list(out_lo) + [0] * (3 - len(out_lo)), dtype=np.int32)
        check(self.lib.adsb_rhs_view(self.h, ctypes.byref(form), ctypes.c_void_p(in_ptr), ctypes.byref(vin),
                                     i_(il), ctypes.c_void_p(forcing_ptr) if forcing_ptr else None,
                                     ctypes.c_void_p(out_ptr), ctypes.byref(vout), i_(ol)))

    # ---- segmented substitution (long lines, slab-sharded sweeps); see include/adsb200.h
    def set_segments(self, axis, slot, bounds, local_lo=0, local_cnt=None):
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        nseg = len(b) - 1
        check(self.lib.adsb_set_axis_segments(self.h, axis, slot, nseg, i_(b), local_lo,
                                              nseg - local_lo if local_cnt is None else local_cnt))

    def segment_info(self, axis, slot):
        v = np.zeros(8, dtype=np.int32)
        check(self.lib.adsb_segment_info(self.h, axis, slot, i_(v)))
        return dict(zip(("KL", "KD", "DF", "DB", "S", "local_lo", "local_cnt", "n"), (int(x) for x in v)))

    def seg_sweep_view(self, axis, slot, seg, in_ptr, vin, out_ptr, vout):
        check(self.lib.adsb_seg_sweep_view(self.h, axis, slot, seg, ctypes.c_void_p(in_ptr), ctypes.byref(vin),
                                           ctypes.c_void_p(out_ptr), ctypes.byref(vout)))

    def neighbor_barrier(self, flags_local, flags_prev, flags_next, error_flag=None):
        """barrier with the two neighbouring ranks on this context's stream (device pointers; None at the ends)"""
        vp = ctypes.c_void_p
        check(self.lib.adsb_neighbor_barrier(self.h, vp(flags_local), vp(flags_prev) if flags_prev else None,
                                             vp(flags_next) if flags_next else None, vp(error_flag) if error_flag else None))

    def dist_sweep_check(self, axis, slot, rank, view, nl, lag):
        return check(self.lib.adsb_dist_sweep_check(self.h, axis, slot, rank, ctypes.byref(view), nl, lag)) == 1

    def dist_sweep_view(self, axis, slot, data_ptr, view, args):
        """fused distributed sweep of this rank's slab (args: _lib.DistArgs), in place"""
        check(self.lib.adsb_dist_sweep_view(self.h, axis, slot, ctypes.c_void_p(data_ptr), ctypes.byref(view),
                                            ctypes.byref(args)))

    @staticmethod
    def _ptr_list(ptrs):
        return (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(int(p)) for p in ptrs]), len(ptrs)

    def seg_dseg_view(self, axis, slot, s_lo, s_hi, row_base, xhat_ptr, vin, dst_ptrs):
        arr, n = self._ptr_list(dst_ptrs)
        check(self.lib.adsb_seg_dseg_view(self.h, axis, slot, s_lo, s_hi, row_base, ctypes.c_void_p(xhat_ptr),
                                          ctypes.byref(vin), arr, n))

    def seg_din_view(self, axis, slot, s_lo, s_hi, row_base, xhat_ptr, vin, dseg_ptr, din_ptr, x_dst_ptrs):
        arr, n = self._ptr_list(x_dst_ptrs)
        check(self.lib.adsb_seg_din_view(self.h, axis, slot, s_lo, s_hi, row_base, ctypes.c_void_p(xhat_ptr),
                                         ctypes.byref(vin), ctypes.c_void_p(dseg_ptr), ctypes.c_void_p(din_ptr), arr, n))

    def seg_tin(self, axis, slot, s_lo, s_hi, lines, x_ptr, tin_ptr):
        check(self.lib.adsb_seg_tin(self.h, axis, slot, s_lo, s_hi, lines, ctypes.c_void_p(x_ptr),
                                    ctypes.c_void_p(tin_ptr)))

    def seg_correct_view(self, axis, slot, s_lo, s_hi, row_base, in_ptr, vin, out_ptr, vout, din_ptr, tin_or_x_ptr):
        check(self.lib.adsb_seg_correct_view(self.h, axis, slot, s_lo, s_hi, row_base, ctypes.c_void_p(in_ptr),
                                             ctypes.byref(vin), ctypes.c_void_p(out_ptr), ctypes.byref(vout),
                                             ctypes.c_void_p(din_ptr), ctypes.c_void_p(tin_or_x_ptr)))

    # ---- measurement
    def enable_timing(self, on=True):
        check(self.lib.adsb_enable_timing(self.h, int(on)))

    def stage_times(self):
        ms = np.zeros(5)
        check(self.lib.adsb_stage_times(self.h, d_(ms)))
        return dict(rhs=ms[0], sweep_x=ms[1], sweep_y=ms[2], sweep_z=ms[3], other=ms[4])

    def launch_count(self):
        return int(self.lib.adsb_launch_count(self.h))


class _simulation:
    """Common part of simulation_2d / simulation_3d: dimensions, shape(), prepare_matrices(), solve()."""

    def __init__(self, dims, steps, device=0):
        self.dims = list(dims)
        self.steps = steps
        self.device = device
        self.ctx = None

    def shape(self):
        return tuple(d.dofs() for d in self.dims)

    def _context(self):
        if self.ctx is None:
            self.ctx = Context(self.shape(), device=self.device)
            for ax, d in enumerate(self.dims):
                self.ctx.set_axis(ax, d)
        return self.ctx

    def prepare_matrices(self):
        """x.factorize_matrix(); y...; (simulation_3d.hpp:58-62) + upload of the factors to slot 0."""
        ctx = self._context()
        for ax, d in enumerate(self.dims):
            lu, ipiv = d.factorize_matrix()
            ctx.set_factor(ax, 0, lu, ipiv, d.p, d.p)

    def add_factor(self, axis, slot, ab):
        """Factorise another matrix of this axis (K = M + h S ...) into `slot`."""
        d = self.dims[axis]
        lu, ipiv = band_factorize(ab, d.p, d.p)
        self._context().set_factor(axis, slot, lu, ipiv, d.p, d.p)

    def solve(self, buf=U, slots=None):
        self._context().solve(buf, slots)

    # simulation_base::run (src/ads/simulation/simulation_base.cpp:11-20)
    def run(self):
        self.before()
        self.advance(self.steps.step_count)
        self.after()

    def before(self):
        pass

    def after(self):
        pass

    def set_state(self, u):
        self._context().upload(U, u)

    def state(self):
        return self._context().download(U)

    def advance(self, nsteps):
        self._context().step(self.substeps(), nsteps)

    def substeps(self):
        raise NotImplementedError


class simulation_2d(_simulation):
    def __init__(self, config_x, config_y, steps, derivatives=1, device=0):
        super().__init__([dimension(config_x, derivatives), dimension(config_y, derivatives)], steps, device)
        self.x, self.y = self.dims


class simulation_3d(_simulation):
    def __init__(self, config_x, config_y, config_z, steps, derivatives=1, device=0):
        super().__init__([dimension(c, derivatives) for c in (config_x, config_y, config_z)], steps, device)
        self.x, self.y, self.z = self.dims


# ------------------------------------------------------------------------------------ problems
class heat_3d(simulation_3d):
    """examples/heat/heat_3d.hpp: rhs = (u,v) - dt (grad u, grad v); solve with Mx (x) My (x) Mz."""

    def __init__(self, p, elements, steps, method=_lib.RHS_COLLAPSED, device=0):
        c = dim_config(p, elements)
        super().__init__(c, c, c, steps, device=device)
        self.method = method

    def before(self):
        self.prepare_matrices()
        self._context().project_init(0, U)
        self.solve(U)

    def substeps(self):
        dt = self.steps.dt
        return [Substep.make(Form.make(1.0, (dt, dt, dt), method=self.method))]


class heat_2d(simulation_2d):
    """examples/heat/heat_2d.hpp: x.fix_left(); row x=0 of the rhs is overwritten with the 1-D
    projection of sin(pi y) before every solve (heat_2d.hpp:40-52)."""

    def __init__(self, p, elements, steps, method=_lib.RHS_COLLAPSED, device=0):
        c = dim_config(p, elements)
        super().__init__(c, c, steps, device=device)
        self.method = method

    def prepare_matrices(self):
        self.x.fix_left()
        super().prepare_matrices()
        row = self.dirichlet_row()
        padded = np.zeros(self._context().local_size)
        padded[:row.size] = row
        self._context().upload(FIXROW, padded)

    def dirichlet_row(self):
        """compute_projection(buf, y.basis, sin(pi y)) (include/ads/projection.hpp:12-36), host, O(n)."""
        t = self.y.basis
        ne, q, p = self.y.elements, self.y.quad_order, self.y.p
        buf = np.zeros(self.y.dofs())
        for e in range(ne):
            for k in range(q):
                wj = t["w"][k] * t["J"][e]
                fx = np.sin(t["x"][e, k] * np.pi)
                for a in range(p + 1):
                    buf[e + a] += fx * t["b"][e, k, 0, a] * wj
        return buf

    def before(self):
        self.prepare_matrices()
        ctx = self._context()
        ctx.zero(U)  # init_state == 0 (heat_2d.hpp:32-38)
        ctx.set_plane(U, 0, 0, self.dirichlet_row())
        self.solve(U)

    def substeps(self):
        dt = self.steps.dt
        return [Substep.make(Form.make(1.0, (dt, dt), method=self.method), fix_axis=0, fix_buf=FIXROW)]


class implicit_2d(simulation_2d):
    """examples/implicit/implicit.hpp: two half steps, K = M + dt/2 S implicit along one axis."""

    def __init__(self, p, elements, steps, method=_lib.RHS_COLLAPSED, device=0):
        c = dim_config(p, elements)
        super().__init__(c, c, steps, device=device)
        self.method = method

    def prepare_matrices(self):
        super().prepare_matrices()
        h = 0.5 * self.steps.dt
        self.add_factor(0, 1, self.x.matrix(3, h))
        self.add_factor(1, 1, self.y.matrix(3, h))

    def before(self):
        self.prepare_matrices()
        self._context().project_init(1, U)
        self.solve(U)

    def substeps(self):
        h = 0.5 * self.steps.dt
        return [Substep.make(Form.make(1.0, (0.0, h), method=self.method), slots=(1, 0)),
                Substep.make(Form.make(1.0, (h, 0.0), method=self.method), slots=(0, 1))]


class implicit_3d(simulation_3d):
    """3-axis extension of implicit_2d (oracle/ref_driver.cpp implicit_3d_ref):
    sub-step d is implicit along axis d with K_d = M_d + (dt/3) S_d."""

    def __init__(self, p, elements, steps, method=_lib.RHS_COLLAPSED, device=0):
        c = dim_config(p, elements)
        super().__init__(c, c, c, steps, device=device)
        self.method = method

    def prepare_matrices(self):
        super().prepare_matrices()
        tau = self.steps.dt / 3.0
        for ax, d in enumerate(self.dims):
            self.add_factor(ax, 1, d.matrix(3, tau))

    def before(self):
        self.prepare_matrices()
        self._context().project_init(1, U)
        self.solve(U)

    def substeps(self):
        tau = self.steps.dt / 3.0
        out = []
        for d in range(3):
            beta = [tau] * 3
            beta[d] = 0.0
            slots = [0] * 3
            slots[d] = 1
            out.append(Substep.make(Form.make(1.0, beta, method=self.method), slots=slots))
        return out


class _scalability:
    def prepare_matrices(self):
        self.x.fix_left()
        super().prepare_matrices()
        # time-independent load: dt * sum_q f(x_q) w J added to every DOF of the element
        # (the quadrature method evaluates the source at the Gauss points itself)
        if self.method == _lib.RHS_COLLAPSED:
            self._context().load_tensor(1, False, FORCING)

    def before(self):
        self.prepare_matrices()
        ctx = self._context()
        ctx.upload(U, np.ones(ctx.local_size))
        self.solve(U)

    def substeps(self):
        dt = self.steps.dt
        beta = (dt,) * len(self.dims)
        if self.method == _lib.RHS_QUADRATURE:
            return [Substep.make(Form.make(1.0, beta, gamma=dt, method=self.method, source=1))]
        return [Substep.make(Form.make(1.0, beta, gamma=dt, forcing_buf=FORCING, method=self.method))]


class scalability_3d(_scalability, simulation_3d):
    """examples/scalability/test3d.hpp"""

    def __init__(self, p, elements, steps, method=_lib.RHS_COLLAPSED, device=0):
        c = dim_config(p, elements)
        simulation_3d.__init__(self, c, c, c, steps, device=device)
        self.method = method


class scalability_2d(_scalability, simulation_2d):
    """examples/scalability/test2d.hpp"""

    def __init__(self, p, elements, steps, method=_lib.RHS_COLLAPSED, device=0):
        c = dim_config(p, elements)
        simulation_2d.__init__(self, c, c, steps, device=device)
        self.method = method


class flow(simulation_3d):
    """examples/flow/flow.hpp: nonlinear flow, rhs = (u v + dt (-k(x) exp(mi u) grad u . grad v + h v)) w J with the
    permeability k tabulated at the quadrature points (fill_permeability_map, flow.hpp:53-60) -- a general
    pointwise form: adsb_compute_rhs_pointwise(ADSB_POINT_FLOW) + ads_solve per step (flow.hpp:62-72).
    `permeability` / `init_state`: callables f(x, y, z) on numpy arrays (the reference's are
    environment::permeability, environment.hpp:78-88, and ads::bump(0.1, 0.5, .), flow.hpp:36-41), or tables."""

    def __init__(self, p, elements, steps, permeability, init_state=None, mi=10.0, device=0):
        c = dim_config(p, elements)
        super().__init__(c, c, c, steps, device=device)
        self.permeability, self.init_state, self.mi = permeability, init_state, mi

    def quadrature_grid(self):
        """coordinates of all quadrature points, shaped for broadcasting to [nqz, nqy, nqx]"""
        x, y, z = (d.basis["x"].ravel() for d in self.dims)
        return x[None, None, :], y[None, :, None], z[:, None, None]

    def tabulate(self, f):
        if callable(f):
            x, y, z = self.quadrature_grid()
            f = np.broadcast_to(f(x, y, z), (z.size, y.size, x.size))
        return np.ascontiguousarray(f, dtype=np.float64)

    def before(self):
        self.prepare_matrices()
        ctx = self._context()
        ctx.set_point_coefficient(self.tabulate(self.permeability))
        if self.init_state is not None:
            ctx.project_values(U, self.tabulate(self.init_state))
            self.solve(U)

    def advance(self, nsteps):
        ctx = self._context()
        form = _lib.PointForm.flow(self.steps.dt, self.mi)
        for _ in range(nsteps):
            ctx.swap(U, U_PREV)                           # before_step
            ctx.compute_rhs_pointwise(form, U_PREV, U)    # compute_rhs
            self.solve(U)


PROBLEMS = {"heat_3d": heat_3d, "heat_2d": heat_2d, "implicit_2d": implicit_2d,
            "implicit_3d": implicit_3d, "scalability_3d": scalability_3d,
            "scalability_2d": scalability_2d}
