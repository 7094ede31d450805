// model_ar.cu -- kernel instantiations for get_ar(d), d compiled for the dimensions the reference's scripts use
// (d = 1: inst/reproduce/lgssm; d = 10: BASELINE config 4) plus a few common sizes.
#include "registry.h"
namespace cpimh {
const ModelEntry* entries_ar(int* n) {
  static const ModelEntry e[] = {make_entry<ArModel<1>>(CPIMH_MODEL_AR), make_entry<ArModel<2>>(CPIMH_MODEL_AR),
                                 make_entry<ArModel<3>>(CPIMH_MODEL_AR), make_entry<ArModel<4>>(CPIMH_MODEL_AR),
                                 make_entry<ArModel<5>>(CPIMH_MODEL_AR), make_entry<ArModel<8>>(CPIMH_MODEL_AR),
                                 make_entry<ArModel<10>>(CPIMH_MODEL_AR), make_entry<ArModel<16>>(CPIMH_MODEL_AR)};
  *n = (int)(sizeof(e) / sizeof(e[0]));
  return e;
}
}  // namespace cpimh
