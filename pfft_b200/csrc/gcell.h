// gcell.h -- ghost-cell (halo) plans: pfft_plan_*gc / pfft_exchange / pfft_reduce
// (reference gcell/gcells_plan.c, gcell/gcells_sendrecv.c).
#pragma once
#include <mpi.h>
#include <stdio.h>

#include "core.h"

namespace pfb {

// fields of the reference's gctimer (kernel/ipfft.h:221-226)
struct GcTimer {
  int iter = 0;
  double whole = 0, pad_zeros = 0, exchange = 0;
};

struct GcPlan;

INT gc_local_size(int rnk_n, const INT *local_n, const INT *local_start, INT howmany, const INT *gc_below,
                  const INT *gc_above, INT *local_ngc, INT *local_gc_start);
GcPlan *gc_plan_create(int prec, int rnk_n, const INT *n, INT howmany, const INT *block, const INT *gc_below,
                       const INT *gc_above, void *data, MPI_Comm comm, unsigned gc_flags, bool is_complex);
void gc_exchange(GcPlan *g);
void gc_reduce(GcPlan *g);
void gc_plan_destroy(GcPlan *g);
void gc_reset_timers(GcPlan *g);
GcTimer *gc_get_timer(GcPlan *g, int which);
void gc_print_timers(GcPlan *g, MPI_Comm comm, FILE *f, bool adv);
void gc_write_timers(GcPlan *g, const char *name, MPI_Comm comm, bool adv);
GcTimer *gctimer_copy(const GcTimer *t);
void gctimer_average(GcTimer *t);
GcTimer *gctimer_add(const GcTimer *a, const GcTimer *b);
GcTimer *gctimer_reduce_max(const GcTimer *t, MPI_Comm comm);
void gctimer_to_vec(const GcTimer *t, double *v);
GcTimer *gctimer_from_vec(const double *v);

}  // namespace pfb
