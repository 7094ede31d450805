// model_sk.cu -- kernel instantiations for SkModel (see models.cuh, pf_fused.cuh, chains.cuh)
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_sk(int* n) {
  static const ModelEntry e[] = {make_entry<SkModel>(CPIMH_MODEL_SK)};
  *n = 1;
  return e;
}
}  // namespace cpimh
