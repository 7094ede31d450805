// model_gss.cu -- kernel instantiations for GssModel (see models.cuh, pf_fused.cuh, chains.cuh)
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_gss(int* n) {
  static const ModelEntry e[] = {make_entry<GssModel>(CPIMH_MODEL_GSS)};
  *n = 1;
  return e;
}
}  // namespace cpimh
