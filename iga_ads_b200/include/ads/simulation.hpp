// ads/simulation.hpp -- umbrella header, as in the reference (include/ads/simulation.hpp).
#ifndef ADSB_ADS_SIMULATION_HPP
#define ADSB_ADS_SIMULATION_HPP

#include <cmath>
#include <utility>

#include "ads/basis_data.hpp"
#include "ads/bspline/bspline.hpp"
#include "ads/executor/galois.hpp"
#include "ads/executor/sequential.hpp"
#include "ads/lin/band_matrix.hpp"
#include "ads/lin/tensor.hpp"
#include "ads/simulation/config.hpp"
#include "ads/simulation/dimension.hpp"
#include "ads/simulation/simulation_base.hpp"
#include "ads/simulation/simulation_nd.hpp"
#include "ads/solver.hpp"

#endif
