// Megakernel instantiation for the scene-feature mask MRT_VARIANT_TREES_TEX (see render_variants.h).
#include "render_kernels.cuh"
#include "render_variants.h"

namespace mrt {
const void *variant_trees_tex(int kind, int minb) { return variant_kernel<MRT_VARIANT_TREES_TEX>(kind, minb); }
}  // namespace mrt
