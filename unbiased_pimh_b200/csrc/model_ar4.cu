// model_ar4.cu -- kernel instantiations for get_ar(4) (one translation unit per dimension: they compile in parallel)
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_ar4(int* n) {
  static const ModelEntry e[] = {make_entry<ArModel<4>>(CPIMH_MODEL_AR)};
  *n = 1;
  return e;
}
}  // namespace cpimh
