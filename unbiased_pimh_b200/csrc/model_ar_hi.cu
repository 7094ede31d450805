// model_ar_hi.cu -- kernel instantiations for get_ar(d), d = 5, 8, 10 (BASELINE config 4), 16
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_ar_hi(int* n) {
  static const ModelEntry e[] = {make_entry<ArModel<5>>(CPIMH_MODEL_AR), make_entry<ArModel<8>>(CPIMH_MODEL_AR),
                                 make_entry<ArModel<10>>(CPIMH_MODEL_AR), make_entry<ArModel<16>>(CPIMH_MODEL_AR)};
  *n = (int)(sizeof(e) / sizeof(e[0]));
  return e;
}
}  // namespace cpimh
