/* oracle/tree.c -- restatement of the Jacob-Murray-Rubenthaler ancestor tree, src/tree.cpp:6-141.
 * TEST INFRASTRUCTURE (see oracle.h).  x_star is d x M column-major as in the reference. */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

struct orc_tree { int N, M, dimx, nsteps; int *a_star, *o_star, *l_star; double* x_star; };

orc_tree* orc_tree_new(int N, int M, int dimx) {              /* src/tree.cpp:6-16 */
  orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
  t->N = N; t->M = M < 3 * N ? 3 * N : M; t->dimx = dimx;     /* :7-9 */
  t->a_star = (int*)calloc((size_t)t->M, sizeof(int));
  t->o_star = (int*)calloc((size_t)t->M, sizeof(int));
  t->x_star = (double*)calloc((size_t)t->M * dimx, sizeof(double));
  t->l_star = (int*)calloc((size_t)N, sizeof(int));
  orc_tree_reset(t);
  return t;
}
void orc_tree_free(orc_tree* t) { if (!t) return; free(t->a_star); free(t->o_star); free(t->x_star); free(t->l_star); free(t); }
void orc_tree_reset(orc_tree* t) {                            /* :22-26 */
  t->nsteps = 0;
  memset(t->a_star, 0, sizeof(int) * (size_t)t->M);
  memset(t->o_star, 0, sizeof(int) * (size_t)t->M);
}
void orc_tree_init(orc_tree* t, const double* x0) {           /* :28-34 */
  for (int i = 0; i < t->N; i++) {
    t->a_star[i] = -2;
    memcpy(t->x_star + (size_t)i * t->dimx, x0 + (size_t)i * t->dimx, sizeof(double) * t->dimx);
    t->l_star[i] = i;
  }
}
static void double_size(orc_tree* t) {                        /* :101-116 */
  int old_M = t->M;
  t->M = 2 * old_M;
  t->a_star = (int*)realloc(t->a_star, sizeof(int) * (size_t)t->M);
  t->o_star = (int*)realloc(t->o_star, sizeof(int) * (size_t)t->M);
  t->x_star = (double*)realloc(t->x_star, sizeof(double) * (size_t)t->M * t->dimx);
  memset(t->a_star + old_M, 0, sizeof(int) * (size_t)old_M);
  memset(t->o_star + old_M, 0, sizeof(int) * (size_t)old_M);
  memset(t->x_star + (size_t)old_M * t->dimx, 0, sizeof(double) * (size_t)old_M * t->dimx);
}
void orc_tree_insert(orc_tree* t, const double* x, const int* a) {   /* :36-68 */
  int N = t->N;
  t->nsteps++;
  int* b = (int*)malloc(sizeof(int) * (size_t)N);
  for (int i = 0; i < N; i++) b[i] = t->l_star[a[i]];         /* :39-42 */
  int slot = 0, i = 0;
  while (slot < N && i < t->M) {                              /* :45-53 first N free slots from 0 */
    if (t->o_star[i] == 0) { t->l_star[slot] = i; slot++; }
    i++;
  }
  if (slot < N) {                                             /* :54-61 */
    int old_M = t->M;
    double_size(t);
    for (i = slot; i < N; i++) t->l_star[i] = old_M + i - slot;
  }
  for (i = 0; i < N; i++) {                                   /* :64-67 */
    t->a_star[t->l_star[i]] = b[i];
    memcpy(t->x_star + (size_t)t->l_star[i] * t->dimx, x + (size_t)i * t->dimx, sizeof(double) * t->dimx);
  }
  free(b);
}
void orc_tree_prune(orc_tree* t, const int* o) {              /* :70-86 */
  for (int i = 0; i < t->N; i++) t->o_star[t->l_star[i]] = o[i];
  for (int i = 0; i < t->N; i++) {
    int j = t->l_star[i];
    while (j >= 0 && t->o_star[j] == 0) {
      j = t->a_star[j];
      if (j >= 0) t->o_star[j] = t->o_star[j] - 1;
    }
  }
}
void orc_tree_update(orc_tree* t, const double* x, const int* a) {   /* :88-99 */
  int* o = (int*)calloc((size_t)t->N, sizeof(int));
  for (int i = 0; i < t->N; i++) o[a[i]]++;
  orc_tree_prune(t, o);
  orc_tree_insert(t, x, a);
  free(o);
}
void orc_tree_get_path(const orc_tree* t, int n, double* path) {     /* :118-129 */
  int d = t->dimx;
  int j = t->l_star[n];
  memcpy(path + (size_t)t->nsteps * d, t->x_star + (size_t)j * d, sizeof(double) * d);
  int step = t->nsteps - 1;
  while (j >= 0 && step >= 0) {
    j = t->a_star[j];
    memcpy(path + (size_t)step * d, t->x_star + (size_t)j * d, sizeof(double) * d);
    step--;
  }
}
void orc_tree_retrieve_xgeneration(const orc_tree* t, int lag, double* xgen) {  /* :131-141 */
  for (int p = 0; p < t->N; p++) {
    int j = t->l_star[p];
    for (int s = 0; s < lag; s++) j = t->a_star[j];
    memcpy(xgen + (size_t)p * t->dimx, t->x_star + (size_t)j * t->dimx, sizeof(double) * t->dimx);
  }
}
int orc_tree_M(const orc_tree* t) { return t->M; }
int orc_tree_nsteps(const orc_tree* t) { return t->nsteps; }
int orc_tree_live_nodes(const orc_tree* t) {
  /* nodes reachable from the current leaves (for capacity studies; not in the reference) */
  char* seen = (char*)calloc((size_t)t->M, 1);
  int cnt = 0;
  for (int i = 0; i < t->N; i++) { int j = t->l_star[i]; while (j >= 0 && !seen[j]) { seen[j] = 1; cnt++; j = t->a_star[j]; } }
  free(seen);
  return cnt;
}
const int* orc_tree_a_star(const orc_tree* t) { return t->a_star; }
const int* orc_tree_o_star(const orc_tree* t) { return t->o_star; }
const int* orc_tree_l_star(const orc_tree* t) { return t->l_star; }
const double* orc_tree_x_star(const orc_tree* t) { return t->x_star; }
