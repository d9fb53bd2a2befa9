// quadbrick_plain.cu -- instantiates the brick quadrature kernel (quadbrick.cuh) for FormLinear<true>, p = 1 ... 5, 2-D and 3-D.
#include "quadbrick.cuh"

namespace adsb {
namespace qb {

ADSB_BRICK_DISPATCH(FormLinear<true>, 5)

}  // namespace qb
}  // namespace adsb
