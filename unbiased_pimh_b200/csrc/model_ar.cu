// model_ar.cu -- kernel instantiations for get_ar(d), low dimensions (d = 1: inst/reproduce/lgssm scripts)
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_ar(int* n) {
  static const ModelEntry e[] = {make_entry<ArModel<1>>(CPIMH_MODEL_AR), make_entry<ArModel<2>>(CPIMH_MODEL_AR),
                                 make_entry<ArModel<3>>(CPIMH_MODEL_AR), make_entry<ArModel<4>>(CPIMH_MODEL_AR)};
  *n = (int)(sizeof(e) / sizeof(e[0]));
  return e;
}
}  // namespace cpimh
