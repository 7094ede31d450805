// model_sv.cu -- kernel instantiations for LevySvModel (see models.cuh, pf_fused.cuh, chains.cuh)
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_sv(int* n) {
  static const ModelEntry e[] = {make_entry<LevySvModel>(CPIMH_MODEL_LEVY_SV)};
  *n = 1;
  return e;
}
}  // namespace cpimh
