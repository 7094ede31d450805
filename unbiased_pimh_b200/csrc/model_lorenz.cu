// model_lorenz.cu -- kernel instantiations for get_lorenz with one particle per thread (reference d = 8,
// R/model_get_lorenz.R:50; d = 4 and 16 for tests).  d = 64 uses the warp-per-particle path in lorenz_wide.cu.
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_lorenz(int* n) {
  static const ModelEntry e[] = {make_entry<LorenzModel<4>>(CPIMH_MODEL_LORENZ), make_entry<LorenzModel<8>>(CPIMH_MODEL_LORENZ),
                                 make_entry<LorenzModel<16>>(CPIMH_MODEL_LORENZ)};
  *n = (int)(sizeof(e) / sizeof(e[0]));
  return e;
}
}  // namespace cpimh
