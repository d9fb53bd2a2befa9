"""iga_ads_b200 -- B200-native ADS time step (RHS assembly + batched banded sweeps) behind the
reference's simulation / dimension / ads_solve surface.  See DESIGN.md and include/adsb200.h."""
from . import _lib
from ._lib import AdsbError, Form, PointForm, Substep, View, RHS_COLLAPSED, RHS_QUADRATURE  # noqa: F401
from .host import (band_factorize, basis_ders, basis_tables, dim_config, dimension, find_span,  # noqa: F401
                   gauss, knots, matrix_1d, timesteps_config, to_band)
from .simulation import (Context, PROBLEMS, U, U_PREV, FORCING, FIXROW, SCRATCH, flow, heat_2d, heat_3d,  # noqa: F401
                         implicit_2d, implicit_3d, scalability_2d, scalability_3d, simulation_2d,
                         simulation_3d)
