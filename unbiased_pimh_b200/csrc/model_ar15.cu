// model_ar15.cu -- kernel instantiations for get_ar(15) (one translation unit per dimension: they compile in parallel)
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_ar15(int* n) {
  static const ModelEntry e[] = {make_entry<ArModel<15>>(CPIMH_MODEL_AR)};
  *n = 1;
  return e;
}
}  // namespace cpimh
