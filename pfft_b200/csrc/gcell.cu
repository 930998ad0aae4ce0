// gcell.cu -- ghost cells (stub: sizes only; exchange/reduce come next).
#include <cuda_runtime.h>

#include "gcell.h"
#include "plan.h"

namespace pfb {

// reference gcell/gcells_plan.c:51-76
INT gc_local_size(int d, const INT *ln, const INT *ls, INT howmany, const INT *gb, const INT *ga, INT *ngc, INT *gcs) {
  INT mem = howmany;
  for (int t = 0; t < d; t++) {
    const INT b = gb ? gb[t] : 0, a = ga ? ga[t] : 0;
    ngc[t] = b + ln[t] + a;
    gcs[t] = ls[t] - b;
    mem *= ngc[t];
  }
  return mem;
}

struct GcPlan {
  GcTimer exg, red;
};

GcPlan *gc_plan_create(int, int, const INT *, INT, const INT *, const INT *, const INT *, void *, MPI_Comm, unsigned, bool) {
  set_error("ghost-cell plans are not implemented yet");
  return nullptr;
}
void gc_exchange(GcPlan *) {}
void gc_reduce(GcPlan *) {}
void gc_plan_destroy(GcPlan *g) { delete g; }
void gc_reset_timers(GcPlan *g) { if (g) g->exg = g->red = GcTimer(); }
GcTimer *gc_get_timer(GcPlan *g, int which) { return g ? new GcTimer(which ? g->red : g->exg) : nullptr; }
void gc_print_timers(GcPlan *, MPI_Comm, FILE *, bool) {}
void gc_write_timers(GcPlan *, const char *, MPI_Comm, bool) {}
GcTimer *gctimer_copy(const GcTimer *t) { return t ? new GcTimer(*t) : nullptr; }
void gctimer_average(GcTimer *t) {
  if (!t || t->iter <= 0) return;
  t->whole /= t->iter; t->pad_zeros /= t->iter; t->exchange /= t->iter; t->iter = 1;
}
GcTimer *gctimer_add(const GcTimer *a, const GcTimer *b) {
  GcTimer *r = new GcTimer(*a);
  r->iter += b->iter; r->whole += b->whole; r->pad_zeros += b->pad_zeros; r->exchange += b->exchange;
  return r;
}
GcTimer *gctimer_reduce_max(const GcTimer *t, MPI_Comm comm) {
  double v[4], m[4];
  gctimer_to_vec(t, v);
  MPI_Allreduce(v, m, 4, MPI_DOUBLE, MPI_MAX, comm);
  return gctimer_from_vec(m);
}
void gctimer_to_vec(const GcTimer *t, double *v) { v[0] = t->iter; v[1] = t->whole; v[2] = t->pad_zeros; v[3] = t->exchange; }
GcTimer *gctimer_from_vec(const double *v) {
  GcTimer *t = new GcTimer;
  t->iter = (int)v[0]; t->whole = v[1]; t->pad_zeros = v[2]; t->exchange = v[3];
  return t;
}

}  // namespace pfb
