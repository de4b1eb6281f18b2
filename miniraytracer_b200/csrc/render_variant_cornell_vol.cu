// Megakernel instantiation for the scene-feature mask MRT_VARIANT_CORNELL_VOL (see render_variants.h).
#include "render_kernels.cuh"
#include "render_variants.h"

namespace mrt {
const void *variant_cornell_vol(bool pixel_per_warp, int minb) { return variant_kernel<MRT_VARIANT_CORNELL_VOL>(pixel_per_warp, minb); }
}  // namespace mrt
