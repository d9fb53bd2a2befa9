// quadbrick_linear.cu -- instantiates the brick quadrature kernel (quadbrick.cuh) for FormLinear<false>, p = 1 ... 5, 2-D and 3-D.
#include "quadbrick.cuh"

namespace adsb {
namespace qb {

ADSB_BRICK_DISPATCH(FormLinear<false>, 5)

}  // namespace qb
}  // namespace adsb
