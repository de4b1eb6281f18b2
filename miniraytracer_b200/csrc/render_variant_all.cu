// Megakernel instantiation for the scene-feature mask MRT_VARIANT_STOCK (see render_variants.h).
#include "render_kernels.cuh"
#include "render_variants.h"

namespace mrt {
const void *variant_all(int kind, int minb) { return variant_kernel<MRT_VARIANT_STOCK>(kind, minb); }
}  // namespace mrt
