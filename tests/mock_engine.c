/* mock_engine.c -- TEST INFRASTRUCTURE ONLY: the subset of include/theia_ba_b200.h that the C++ adapters call, backed by the CPU
 * oracle, so that adapter_test's end-to-end modes (solve / tracks / micro / twoview) can exercise the ADAPTERS' own logic -- problem
 * flattening, scatter back into the Reconstruction, residency / generation handling, status mapping -- on a machine without a GPU.
 * Linked into tests/adapter_test_mock only (adapter/Makefile target `mock`); the product library is never replaced by this.  */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/theia_ba_b200.h"

void oracle_options_init(tba_options* o);
int oracle_solve(const tba_options* opt, tba_problem* p, tba_summary* s);
int oracle_filter_tracks(const tba_problem* p, double max_err, double min_angle, uint8_t* status, double* mean_sq_error);
int oracle_adjust_tracks(const tba_options* opt, tba_problem* p, uint8_t* status, double* initial_cost, double* final_cost);
int oracle_estimate_tracks(const tba_options* opt, tba_problem* p, double max_px, double min_angle, int ba, uint8_t* status, int32_t counts[5]);

struct tba_context {
  int uploaded;
  tba_options opt;
  tba_problem p; /* deep copy */
  char err[256];
};

static void* dup_mem(const void* src, size_t n) { void* d = malloc(n ? n : 1); if (src && n) memcpy(d, src, n); return d; }
static void free_problem(tba_problem* p) {
  free(p->ext); free((void*)p->ext_const); free((void*)p->cam_group); free((void*)p->group_model); free(p->intr); free((void*)p->group_const_mask);
  free(p->pt); free((void*)p->pt_const); free((void*)p->obs_cam); free((void*)p->obs_pt); free((void*)p->obs_xy);
  memset(p, 0, sizeof *p);
}

void tba_options_init(tba_options* o) { oracle_options_init(o); }
int tba_device_count(void) { return 1; }
int tba_create(int device, int rank, int world, const void* id, tba_context** out) {
  (void)device; (void)rank; (void)world; (void)id;
  *out = calloc(1, sizeof(tba_context));
  return TBA_OK;
}
void tba_destroy(tba_context* c) { if (c) { if (c->uploaded) free_problem(&c->p); free(c); } }
const char* tba_last_error(tba_context* c) { return c ? c->err : "null context"; }

int tba_upload(tba_context* c, const tba_options* o, const tba_problem* p) {
  if (o->linear_solver_type == 6) { snprintf(c->err, sizeof c->err, "mock: CGNR unsupported"); return TBA_ERR_UNSUPPORTED; }
  if (c->uploaded) free_problem(&c->p);
  c->opt = *o;
  c->p = *p;
  c->p.ext = dup_mem(p->ext, (size_t)p->n_cam * 48); c->p.ext_const = dup_mem(p->ext_const, p->n_cam); c->p.cam_group = dup_mem(p->cam_group, (size_t)p->n_cam * 4);
  c->p.group_model = dup_mem(p->group_model, (size_t)p->n_group * 4); c->p.intr = dup_mem(p->intr, (size_t)p->n_group * 80);
  c->p.group_const_mask = dup_mem(p->group_const_mask, (size_t)p->n_group * 4);
  c->p.pt = dup_mem(p->pt, (size_t)p->n_pt * 32); c->p.pt_const = dup_mem(p->pt_const, p->n_pt);
  c->p.obs_cam = dup_mem(p->obs_cam, (size_t)p->n_obs * 4); c->p.obs_pt = dup_mem(p->obs_pt, (size_t)p->n_obs * 4); c->p.obs_xy = dup_mem(p->obs_xy, (size_t)p->n_obs * 16);
  c->uploaded = 1;
  return TBA_OK;
}
int tba_download(tba_context* c, tba_problem* p) {
  if (!c->uploaded || p->n_cam != c->p.n_cam || p->n_pt != c->p.n_pt || p->n_group != c->p.n_group) return TBA_ERR_INVALID_ARGUMENT;
  memcpy(p->ext, c->p.ext, (size_t)p->n_cam * 48); memcpy(p->intr, c->p.intr, (size_t)p->n_group * 80); memcpy(p->pt, c->p.pt, (size_t)p->n_pt * 32);
  return TBA_OK;
}
int tba_solve(tba_context* c, const tba_options* o, tba_problem* p, tba_summary* s) {
  int rc = tba_upload(c, o, p);
  if (rc) { s->success = 0; s->termination_type = TBA_FAILURE; return rc; }
  rc = oracle_solve(o, &c->p, s);
  if (rc) { snprintf(c->err, sizeof c->err, "%s", s->message); return rc; }
  return tba_download(c, p);
}
int tba_solve_multi(const tba_options* o, tba_problem* p, tba_summary* s, int n) {
  (void)n;
  tba_context* c; tba_create(0, 0, 1, NULL, &c);
  const int rc = tba_solve(c, o, p, s);
  tba_destroy(c);
  return rc;
}
int tba_filter_tracks(tba_context* c, double max_err, double min_angle, uint8_t* status, double* mean_sq, int32_t* nb, int32_t* ni) {
  if (!c->uploaded) return TBA_ERR_INVALID_ARGUMENT;
  oracle_filter_tracks(&c->p, max_err, min_angle, status, mean_sq);
  int b = 0, i = 0;
  for (int q = 0; q < c->p.n_pt; ++q) { b += status[q] == 1; i += status[q] == 2; }
  if (nb) *nb = b;
  if (ni) *ni = i;
  return TBA_OK;
}
int tba_adjust_tracks(tba_context* c, const tba_options* o, uint8_t* status, double* ic, double* fc, int32_t* nf) {
  if (!c->uploaded) return TBA_ERR_INVALID_ARGUMENT;
  const int f = oracle_adjust_tracks(o, &c->p, status, ic, fc);
  if (nf) *nf = f;
  return TBA_OK;
}
int tba_estimate_tracks(tba_context* c, const tba_options* o, double max_px, double min_angle, int32_t ba, uint8_t* status, int32_t counts[5]) {
  if (!c->uploaded) return TBA_ERR_INVALID_ARGUMENT;
  oracle_estimate_tracks(o, &c->p, max_px, min_angle, ba, status, counts);
  return TBA_OK;
}
