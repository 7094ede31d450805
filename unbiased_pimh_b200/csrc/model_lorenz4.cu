// model_lorenz4.cu -- kernel instantiations for get_lorenz with d = 4 (thread-per-particle; d = 32 / 64 live in wide.cu)
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_lorenz4(int* n) {
  static const ModelEntry e[] = {make_entry<LorenzModel<4>>(CPIMH_MODEL_LORENZ)};
  *n = 1;
  return e;
}
}  // namespace cpimh
