// Megakernel instantiation for the scene-feature mask MRT_VARIANT_LISTS_VOL (see render_variants.h).
#include "render_kernels.cuh"
#include "render_variants.h"

namespace mrt {
const void *variant_lists_vol(int kind, int minb) { return variant_kernel<MRT_VARIANT_LISTS_VOL>(kind, minb); }
}  // namespace mrt
