// Megakernel instantiation for the scene-feature mask MRT_VARIANT_LISTS (see render_variants.h).
#include "render_kernels.cuh"
#include "render_variants.h"

namespace mrt {
const void *variant_lists(bool pixel_per_warp, int minb) { return variant_kernel<MRT_VARIANT_LISTS>(pixel_per_warp, minb); }
}  // namespace mrt
