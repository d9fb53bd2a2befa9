// quadbrick_flow.cu -- instantiates the brick quadrature kernel (quadbrick.cuh) for FormFlow, p = 1 ... 3, 2-D and 3-D.
#include "quadbrick.cuh"

namespace adsb {
namespace qb {

ADSB_BRICK_DISPATCH(FormFlow, 3)

}  // namespace qb
}  // namespace adsb
