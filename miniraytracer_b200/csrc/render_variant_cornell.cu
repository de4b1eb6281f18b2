// Megakernel instantiation for the scene-feature mask MRT_VARIANT_CORNELL (see render_variants.h).
#include "render_kernels.cuh"
#include "render_variants.h"

namespace mrt {
const void *variant_cornell(int kind, int minb) { return variant_kernel<MRT_VARIANT_CORNELL>(kind, minb); }
}  // namespace mrt
