// Megakernel instantiation for the scene-feature mask MRT_FEAT_ALL (see render_variants.h): every feature, including lone
// triangle_scene_objects.
#include "render_kernels.cuh"
#include "render_variants.h"

namespace mrt {
const void *variant_full(int kind, int minb) { return variant_kernel<MRT_FEAT_ALL>(kind, minb); }
}  // namespace mrt
