// api.cu -- the exported PFFT C API (include/pfft.h), both precisions, plus the
// pfftb200_* extensions (include/pfft_b200.h).  Thin: argument marshalling into a
// pfb::Problem, then the planner / plan engine.  Reference counterparts:
// api/api-basic.c:198-251,482-581,736-1183 and api/api-adv.c:30-385.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <complex>
#include <vector>

#include "gcell.h"
#include "plan.h"
#include "pfft_b200.h"

// the public header's opaque types, declared here without pulling in pfft.h's C99 complex typedefs
using namespace pfb;

namespace {

struct CommInfo {
  int rnk_pm = 1;
  int np[kMaxMesh] = {1, 1, 1};
  int rank = 0;
};

// what the reference does with a non-Cartesian communicator: treat it as a 1-D mesh
// (kernel/procmesh.c:114-129)
CommInfo comm_info(MPI_Comm comm) {
  CommInfo ci;
  int status = MPI_UNDEFINED;
  MPI_Topo_test(comm, &status);
  MPI_Comm_rank(comm, &ci.rank);
  if (status == MPI_CART) {
    int nd = 1;
    MPI_Cartdim_get(comm, &nd);
    ci.rnk_pm = nd;
    int dims[8], per[8], co[8];
    MPI_Cart_get(comm, 8, dims, per, co);
    for (int t = 0; t < nd && t < kMaxMesh; t++) ci.np[t] = dims[t];
  } else {
    ci.rnk_pm = 1;
    MPI_Comm_size(comm, &ci.np[0]);
  }
  return ci;
}

Problem make_problem(Kind kind, int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
                     const INT *iblock, const INT *oblock, MPI_Comm comm, int sign, const int *kinds,
                     const int *skip, unsigned flags, int *rank_out) {
  CommInfo ci = comm_info(comm);
  Problem p;
  fill_problem(&p, (int)kind, rnk_n, n, ni, no, howmany, iblock, oblock, ci.rnk_pm, ci.np, sign, kinds, skip, flags);
  if (rank_out) *rank_out = ci.rank;
  return p;
}

INT local_size_impl(Kind kind, int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
                    const INT *iblock, const INT *oblock, MPI_Comm comm, unsigned flags, INT *lni, INT *lis,
                    INT *lno, INT *los) {
  int rank = 0;
  Problem p = make_problem(kind, rnk_n, n, ni, no, howmany, iblock, oblock, comm, -1, nullptr, nullptr, flags, &rank);
  if (p.rnk_n < 1 || p.rnk_n > kMaxDims - 1 || p.rnk_pm < 1 || p.rnk_pm > kMaxMesh) {
    // fixed-size arrays behind the integer layer: refuse rather than overrun them
    set_error("pfft_local_size: rnk_n must be 1..7 and the process mesh at most 3-dimensional");
    for (int t = 0; t < rnk_n; t++) lni[t] = lis[t] = lno[t] = los[t] = 0;
    return 0;
  }
  LocalSizes ls;
  local_block(p, rank, &ls);
  for (int t = 0; t < rnk_n; t++) {
    lni[t] = ls.lni[t];
    lis[t] = ls.lis[t];
    lno[t] = ls.lno[t];
    los[t] = ls.los[t];
  }
  return alloc_local(p, rank);
}

void local_block_impl(Kind kind, int rnk_n, const INT *ni, const INT *no, const INT *iblock, const INT *oblock,
                      MPI_Comm comm, int pid, unsigned flags, INT *lni, INT *lis, INT *lno, INT *los) {
  Problem p = make_problem(kind, rnk_n, ni, ni, no, 1, iblock, oblock, comm, -1, nullptr, nullptr, flags, nullptr);
  if (p.rnk_n < 1 || p.rnk_n > kMaxDims - 1 || p.rnk_pm < 1 || p.rnk_pm > kMaxMesh) {
    set_error("pfft_local_block: rnk_n must be 1..7 and the process mesh at most 3-dimensional");
    for (int t = 0; t < rnk_n; t++) lni[t] = lis[t] = lno[t] = los[t] = 0;
    return;
  }
  LocalSizes ls;
  local_block(p, pid, &ls);
  for (int t = 0; t < rnk_n; t++) {
    lni[t] = ls.lni[t];
    lis[t] = ls.lis[t];
    lno[t] = ls.lno[t];
    los[t] = ls.los[t];
  }
}

void *plan_impl(int prec, Kind kind, int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
                const INT *iblock, const INT *oblock, const int *skip, void *in, void *out, MPI_Comm comm, int sign,
                const int *kinds, unsigned flags) {
  set_error("");
  Problem p = make_problem(kind, rnk_n, n, ni, no, howmany, iblock, oblock, comm, sign, kinds, skip, flags, nullptr);
  std::string why;
  if (!problem_is_legal(p, &why)) {   // the reference returns NULL silently (kernel/partrafo.c:337-373)
    set_error(why);
    return nullptr;
  }
  return plan_create(prec, p, in, out, comm);
}

void execute_impl(void *plan, void *in, void *out) {
  if (!plan) {   // reference api/api-basic.c:1050-1054
    int r = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &r);
    if (r == 0) fprintf(stderr, "!!! Error: Can not execute PFFT Plan == NULL !!!\n");
    return;
  }
  plan_execute(static_cast<PlanBase *>(plan), in, out, true);
}

// ---- test-data contract, reference api/api-basic.c:60-189,254-276 -------------------
template <typename R>
std::complex<R> pattern_value(int d, const INT *n, const INT *g) {
  // plain_index: k += k*n[t] + kvec[t], on indices reduced with C's %
  INT k = 0;
  for (int t = 0; t < d; t++) k += k * n[t] + (g[t] % n[t]);
  const R m = (R)k;
  if (m == 0) return std::complex<R>((R)1500.0, (R)1250.0);
  const double re = 1000.0 / (double)(2 * m);
  const double im = 1000.0 / (double)(2 * m + 1);
  return std::complex<R>((R)re, (R)im);
}

enum ArrayType { AT_REAL, AT_COMPLEX, AT_HERMITIAN };

template <typename R>
void pattern_at(int d, const INT *n, const INT *ln, const INT *ls, INT k, ArrayType at, std::complex<R> *val,
                bool *in_range) {
  INT g[kMaxDims], gm[kMaxDims];
  INT rem = k;
  for (int t = d - 1; t >= 0; t--) {
    INT kt = rem % ln[t];
    rem = (rem - kt) / ln[t];
    g[t] = kt + ls[t];
    gm[t] = n[t] - g[t];
  }
  std::complex<R> d1 = pattern_value<R>(d, n, g);
  *in_range = g[d - 1] < n[d - 1];
  if (at == AT_HERMITIAN) {
    std::complex<R> d2 = pattern_value<R>(d, n, gm);
    const double re = 0.5 * ((double)d1.real() + (double)d2.real());
    const double im = 0.5 * ((double)d1.imag() - (double)d2.imag());
    *val = std::complex<R>((R)re, (R)im);
  } else {
    *val = d1;
  }
}

template <typename R>
void init_array(int d, const INT *n, const INT *ln, const INT *ls, ArrayType at, void *data) {
  INT tot = 1;
  for (int t = 0; t < d; t++) tot *= ln[t];
  cudaDeviceSynchronize();   // managed memory: make sure no kernel is touching it
  for (INT k = 0; k < tot; k++) {
    std::complex<R> v;
    bool in_range;
    pattern_at<R>(d, n, ln, ls, k, at, &v, &in_range);
    if (at == AT_REAL) static_cast<R *>(data)[k] = in_range ? v.real() : (R)0;
    else static_cast<std::complex<R> *>(data)[k] = v;
  }
}

template <typename R>
void clear_array(int d, const INT *ln, ArrayType at, void *data) {
  INT tot = 1;
  for (int t = 0; t < d; t++) tot *= ln[t];
  cudaDeviceSynchronize();
  memset(data, 0, (size_t)tot * sizeof(R) * (at == AT_REAL ? 1 : 2));
}

template <typename R>
R check_array(int d, const INT *n, const INT *ln, const INT *ls, ArrayType at, const void *data, MPI_Comm comm) {
  INT tot = 1;
  for (int t = 0; t < d; t++) tot *= ln[t];
  cudaDeviceSynchronize();
  R maxerr = 0;
  for (INT k = 0; k < tot; k++) {
    std::complex<R> v;
    bool in_range;
    pattern_at<R>(d, n, ln, ls, k, at, &v, &in_range);
    R err;
    if (at == AT_REAL) {
      err = in_range ? (R)fabs((double)(static_cast<const R *>(data)[k] - v.real())) : (R)0;
    } else {
      std::complex<R> df = static_cast<const std::complex<R> *>(data)[k] - v;
      err = (R)hypot((double)df.real(), (double)df.imag());
    }
    if (err > maxerr) maxerr = err;
  }
  R glob = 0;
  MPI_Allreduce(&maxerr, &glob, 1, sizeof(R) == 8 ? MPI_DOUBLE : MPI_FLOAT, MPI_MAX, comm);
  return glob;
}

// ---- rank-0 printing (reference util/util.c:175-204) ----------------------------------
void vfprintf_rank0(MPI_Comm comm, FILE *stream, const char *format, va_list ap) {
  int r = 0;
  MPI_Comm_rank(comm, &r);
  if (r == 0) {
    vfprintf(stream, format, ap);
    fflush(stream);
  }
}

void fprintf_rank0(MPI_Comm comm, FILE *stream, const char *format, ...) {
  va_list ap;
  va_start(ap, format);
  vfprintf_rank0(comm, stream, format, ap);
  va_end(ap);
}

// ---- command line helper (reference util/getargs.c:41-66) --------------------------------
void get_args_impl(int argc, char **argv, const char *name, int needed, unsigned type, void *param) {
  for (int i = 0; i < argc; i++) {
    if (strcmp(argv[i], name) != 0) continue;
    if (type == 7u) {   // PFFT_SWITCH
      *static_cast<int *>(param) = 1;
      return;
    }
    int given = 0;
    while (i + 1 + given < argc && argv[i + 1 + given] && argv[i + 1 + given][0] != '-') given++;
    // negative numbers start with '-': accept "-<digit>" as a value
    given = 0;
    while (i + 1 + given < argc && argv[i + 1 + given]) {
      const char *a = argv[i + 1 + given];
      if (a[0] == '-' && !(a[1] >= '0' && a[1] <= '9') && a[1] != '.') break;
      given++;
    }
    if (given < needed) {
      fprintf_rank0(MPI_COMM_WORLD, stdout, "!!! Warning: Not enough command line arguments for %s !!!\n", name);
      return;
    }
    for (int t = 0; t < needed; t++) {
      const char *a = argv[i + 1 + t];
      switch (type) {
        case 1u: static_cast<int *>(param)[t] = (int)strtol(a, nullptr, 0); break;
        case 2u: static_cast<ptrdiff_t *>(param)[t] = (ptrdiff_t)strtol(a, nullptr, 0); break;
        case 3u: static_cast<float *>(param)[t] = strtof(a, nullptr); break;
        case 4u: static_cast<double *>(param)[t] = strtod(a, nullptr); break;
        case 5u: static_cast<long double *>(param)[t] = strtold(a, nullptr); break;
        case 6u: static_cast<unsigned *>(param)[t] = (unsigned)strtoul(a, nullptr, 0); break;
        default:
          fprintf_rank0(MPI_COMM_WORLD, stderr, "!!! Error: PFFT_DATATYPE of %s not supported. !!!\n", name);
          return;
      }
    }
    return;
  }
}

// ---- array printers (reference api/api-basic.c:585-713): rank after rank ---------------
template <typename R>
void apr_3d(const void *data, bool is_complex, const INT *ln, const INT *ls, const int *perm, const char *name,
            MPI_Comm comm) {
  int np = 1, me = 0;
  MPI_Comm_size(comm, &np);
  MPI_Comm_rank(comm, &me);
  cudaDeviceSynchronize();
  for (int turn = 0; turn < np; turn++) {
    MPI_Barrier(comm);
    if (turn != me) continue;
    printf("%s (rank %d): local_n = [%td, %td, %td], local_start = [%td, %td, %td]\n", name, me, ln[0], ln[1], ln[2],
           ls[0], ls[1], ls[2]);
    const INT e[3] = {ln[perm[0]], ln[perm[1]], ln[perm[2]]};
    INT l = 0;
    for (INT a = 0; a < e[0]; a++) {
      for (INT b = 0; b < e[1]; b++) {
        printf("%s[%td][%td][:] = ", name, a, b);
        for (INT c = 0; c < e[2]; c++, l++) {
          if (is_complex) {
            const std::complex<R> v = static_cast<const std::complex<R> *>(data)[l];
            printf("  %.2e + I* %.2e,", (double)v.real(), (double)v.imag());
          } else {
            printf("  %.2e,", (double)static_cast<const R *>(data)[l]);
          }
        }
        printf("\n");
      }
      printf("\n");
    }
    fflush(stdout);
  }
  MPI_Barrier(comm);
}

// ---- timers (reference kernel/timer.c) ---------------------------------------------------
TimerData *timer_copy(const TimerData *o) { return o ? new TimerData(*o) : nullptr; }

void timer_average(TimerData *t) {
  if (!t || t->iter <= 0) return;
  const double s = 1.0 / t->iter;
  t->whole *= s;
  for (auto &x : t->trafo) x *= s;
  for (auto &x : t->remap) x *= s;
  t->remap_3dto2d[0] *= s;
  t->remap_3dto2d[1] *= s;
  t->itwiddle *= s;
  t->otwiddle *= s;
  t->iter = 1;
}

std::vector<double> timer_to_vec(const TimerData *t) {
  std::vector<double> v;
  v.push_back(t->rnk_pm);
  v.push_back(t->rnk_trafo);
  v.push_back(t->rnk_remap);
  v.push_back(t->iter);
  v.push_back(t->whole);
  for (double x : t->trafo) v.push_back(x);
  for (double x : t->remap) v.push_back(x);
  v.push_back(t->remap_3dto2d[0]);
  v.push_back(t->remap_3dto2d[1]);
  v.push_back(t->itwiddle);
  v.push_back(t->otwiddle);
  return v;
}

TimerData *timer_from_vec(const double *v) {
  TimerData *t = new TimerData;
  t->shape((int)v[0]);
  t->iter = (int)v[3];
  t->whole = v[4];
  size_t k = 5;
  for (auto &x : t->trafo) x = v[k++];
  for (auto &x : t->remap) x = v[k++];
  t->remap_3dto2d[0] = v[k++];
  t->remap_3dto2d[1] = v[k++];
  t->itwiddle = v[k++];
  t->otwiddle = v[k++];
  return t;
}

TimerData *timer_add(const TimerData *a, const TimerData *b) {
  std::vector<double> va = timer_to_vec(a), vb = timer_to_vec(b);
  for (size_t k = 3; k < va.size() && k < vb.size(); k++) va[k] += vb[k];
  return timer_from_vec(va.data());
}

TimerData *timer_reduce_max(const TimerData *t, MPI_Comm comm) {
  std::vector<double> v = timer_to_vec(t), m(v.size());
  MPI_Allreduce(v.data(), m.data(), (int)v.size(), MPI_DOUBLE, MPI_MAX, comm);
  return timer_from_vec(m.data());
}

void timer_print(const PlanBase *pl, MPI_Comm comm, FILE *f, bool adv) {
  if (!pl) return;
  TimerData *mt = timer_reduce_max(&pl->timer, comm);
  timer_average(mt);
  int np = 1;
  MPI_Comm_size(comm, &np);
  const int idx = (int)(log((double)np) / log(2.0)) + 1;   // MATLAB index by log2 of the process count
  const char *prefix = "pfft";
  if (!adv) {
    fprintf_rank0(comm, f, "%s_iter(%d)    = %d;  ", prefix, idx, pl->timer.iter);
    fprintf_rank0(comm, f, "%s(%d)   = %.3e;\n", prefix, idx, mt->whole);
  } else {
    fprintf_rank0(comm, f, "%s_itwiddle(%d)   = %.3e;\n", prefix, idx, mt->itwiddle);
    fprintf_rank0(comm, f, "%s_remap_3dto2d(%d, 2)   = %.3e;\n", prefix, idx, mt->remap_3dto2d[0]);
    for (size_t k = 0; k < mt->trafo.size(); k++) {
      fprintf_rank0(comm, f, "%s_trafo%d(%d, 2)   = %.3e;", prefix, (int)k + 1, idx, mt->trafo[k]);
      if (k < mt->remap.size()) fprintf_rank0(comm, f, "  %s_remap%d(%d, 1) = %.3e;", prefix, (int)k + 1, idx, mt->remap[k]);
      fprintf_rank0(comm, f, "\n");
    }
    fprintf_rank0(comm, f, "%s_remap_2dto3d(%d, 2)   = %.3e;\n", prefix, idx, mt->remap_3dto2d[1]);
    fprintf_rank0(comm, f, "%s_otwiddle(%d)   = %.3e;\n", prefix, idx, mt->otwiddle);
  }
  delete mt;
}

void timer_write(const PlanBase *pl, const char *name, MPI_Comm comm, bool adv) {
  int r = 0;
  MPI_Comm_rank(comm, &r);
  FILE *f = r == 0 ? fopen(name, "w") : nullptr;
  timer_print(pl, comm, f ? f : stdout, adv);
  if (f) fclose(f);
}

void *managed_alloc(size_t bytes) {
  ensure_device();
  void *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocManaged(&p, bytes, cudaMemAttachGlobal);
  if (e != cudaSuccess) {
    fprintf(stderr, "pfft_b200: cudaMallocManaged(%zu) failed: %s\n", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}

}  // namespace

// =========================================================================================
// exported symbols, stamped for both precisions
// =========================================================================================
#define LOCAL_OUT INT *local_ni, INT *local_i_start, INT *local_no, INT *local_o_start
#define LOCAL_ARGS local_ni, local_i_start, local_no, local_o_start
#define MANY_ARGS_DECL int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany, const INT *iblock, const INT *oblock
#define BLOCK_DECL const INT *n, const INT *local_n, const INT *local_start

#define PFFT_B200_DEFINE_API(P, R, PREC)                                                                          \
  struct P(plan_s);                                                                                               \
  struct P(gcplan_s);                                                                                             \
  struct P(timer_s);                                                                                              \
  struct P(gctimer_s);                                                                                            \
  extern "C" {                                                                                                    \
  void P(init)(void) { int f = 0; MPI_Initialized(&f); if (!f) MPI_Init(nullptr, nullptr); }                      \
  void P(cleanup)(void) {}                                                                                        \
  void P(plan_with_nthreads)(int) {}                                                                              \
  int P(get_nthreads)(void) { return 1; }                                                                         \
  void *P(malloc)(size_t nbytes) { return managed_alloc(nbytes); }                                                \
  R *P(alloc_real)(size_t cnt) { return static_cast<R *>(managed_alloc(cnt * sizeof(R))); }                       \
  void *P(alloc_complex)(size_t cnt) { return managed_alloc(cnt * 2 * sizeof(R)); }                               \
  void P(free)(void *p) { if (p) { cudaDeviceSynchronize(); cudaFree(p); } }                                      \
  void P(execute)(P(plan_s) * plan) { execute_impl(plan, nullptr, nullptr); }                                     \
  void P(execute_dft)(P(plan_s) * plan, void *in, void *out) { execute_impl(plan, in, out); }                     \
  void P(execute_dft_r2c)(P(plan_s) * plan, R *in, void *out) { execute_impl(plan, in, out); }                    \
  void P(execute_dft_c2r)(P(plan_s) * plan, void *in, R *out) { execute_impl(plan, in, out); }                    \
  void P(execute_r2r)(P(plan_s) * plan, R *in, R *out) { execute_impl(plan, in, out); }                           \
  void P(destroy_plan)(P(plan_s) * plan) {                                                                        \
    if (!plan) { fprintf_rank0(MPI_COMM_WORLD, stderr, "!!! Error: Can not destroy PFFT Plan == NULL !!!\n"); return; } \
    plan_destroy(reinterpret_cast<PlanBase *>(plan));                                                             \
  }                                                                                                               \
  void P(init_input_complex_3d)(BLOCK_DECL, void *data) { init_array<R>(3, n, local_n, local_start, AT_COMPLEX, data); } \
  void P(init_input_complex)(int d, BLOCK_DECL, void *data) { init_array<R>(d, n, local_n, local_start, AT_COMPLEX, data); } \
  void P(init_input_complex_hermitian_3d)(BLOCK_DECL, void *data) { init_array<R>(3, n, local_n, local_start, AT_HERMITIAN, data); } \
  void P(init_input_complex_hermitian)(int d, BLOCK_DECL, void *data) { init_array<R>(d, n, local_n, local_start, AT_HERMITIAN, data); } \
  void P(init_input_real_3d)(BLOCK_DECL, R *data) { init_array<R>(3, n, local_n, local_start, AT_REAL, data); }   \
  void P(init_input_real)(int d, BLOCK_DECL, R *data) { init_array<R>(d, n, local_n, local_start, AT_REAL, data); } \
  void P(clear_input_complex_3d)(BLOCK_DECL, void *data) { (void)n; (void)local_start; clear_array<R>(3, local_n, AT_COMPLEX, data); } \
  void P(clear_input_complex)(int d, BLOCK_DECL, void *data) { (void)n; (void)local_start; clear_array<R>(d, local_n, AT_COMPLEX, data); } \
  void P(clear_input_complex_hermitian_3d)(BLOCK_DECL, void *data) { (void)n; (void)local_start; clear_array<R>(3, local_n, AT_HERMITIAN, data); } \
  void P(clear_input_complex_hermitian)(int d, BLOCK_DECL, void *data) { (void)n; (void)local_start; clear_array<R>(d, local_n, AT_HERMITIAN, data); } \
  void P(clear_input_real_3d)(BLOCK_DECL, R *data) { (void)n; (void)local_start; clear_array<R>(3, local_n, AT_REAL, data); } \
  void P(clear_input_real)(int d, BLOCK_DECL, R *data) { (void)n; (void)local_start; clear_array<R>(d, local_n, AT_REAL, data); } \
  R P(check_output_complex_3d)(BLOCK_DECL, const void *data, MPI_Comm comm) { return check_array<R>(3, n, local_n, local_start, AT_COMPLEX, data, comm); } \
  R P(check_output_complex)(int d, BLOCK_DECL, const void *data, MPI_Comm comm) { return check_array<R>(d, n, local_n, local_start, AT_COMPLEX, data, comm); } \
  R P(check_output_complex_hermitian_3d)(BLOCK_DECL, const void *data, MPI_Comm comm) { return check_array<R>(3, n, local_n, local_start, AT_HERMITIAN, data, comm); } \
  R P(check_output_complex_hermitian)(int d, BLOCK_DECL, const void *data, MPI_Comm comm) { return check_array<R>(d, n, local_n, local_start, AT_HERMITIAN, data, comm); } \
  R P(check_output_real_3d)(BLOCK_DECL, const R *data, MPI_Comm comm) { return check_array<R>(3, n, local_n, local_start, AT_REAL, data, comm); } \
  R P(check_output_real)(int d, BLOCK_DECL, const R *data, MPI_Comm comm) { return check_array<R>(d, n, local_n, local_start, AT_REAL, data, comm); } \
  /* ---- data distribution */                                                                                    \
  INT P(local_size_many_dft)(MANY_ARGS_DECL, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::C2C, rnk_n, n, ni, no, howmany, iblock, oblock, c, fl, LOCAL_ARGS); } \
  INT P(local_size_many_dft_r2c)(MANY_ARGS_DECL, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::R2C, rnk_n, n, ni, no, howmany, iblock, oblock, c, fl, LOCAL_ARGS); } \
  INT P(local_size_many_dft_c2r)(MANY_ARGS_DECL, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::C2R, rnk_n, n, ni, no, howmany, iblock, oblock, c, fl, LOCAL_ARGS); } \
  INT P(local_size_many_r2r)(MANY_ARGS_DECL, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::R2R, rnk_n, n, ni, no, howmany, iblock, oblock, c, fl, LOCAL_ARGS); } \
  INT P(local_size_dft)(int d, const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::C2C, d, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  INT P(local_size_dft_r2c)(int d, const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::R2C, d, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  INT P(local_size_dft_c2r)(int d, const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::C2R, d, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  INT P(local_size_r2r)(int d, const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::R2R, d, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  INT P(local_size_dft_3d)(const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::C2C, 3, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  INT P(local_size_dft_r2c_3d)(const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::R2C, 3, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  INT P(local_size_dft_c2r_3d)(const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::C2R, 3, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  INT P(local_size_r2r_3d)(const INT *n, MPI_Comm c, unsigned fl, LOCAL_OUT) { return local_size_impl(Kind::R2R, 3, n, n, n, 1, nullptr, nullptr, c, fl, LOCAL_ARGS); } \
  void P(local_block_many_dft)(int d, const INT *ni, const INT *no, const INT *ib, const INT *ob, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::C2C, d, ni, no, ib, ob, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_many_dft_r2c)(int d, const INT *ni, const INT *no, const INT *ib, const INT *ob, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::R2C, d, ni, no, ib, ob, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_many_dft_c2r)(int d, const INT *ni, const INT *no, const INT *ib, const INT *ob, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::C2R, d, ni, no, ib, ob, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_many_r2r)(int d, const INT *ni, const INT *no, const INT *ib, const INT *ob, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::R2R, d, ni, no, ib, ob, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_dft)(int d, const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::C2C, d, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_dft_r2c)(int d, const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::R2C, d, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_dft_c2r)(int d, const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::C2R, d, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_r2r)(int d, const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::R2R, d, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_dft_3d)(const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::C2C, 3, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_dft_r2c_3d)(const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::R2C, 3, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_dft_c2r_3d)(const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::C2R, 3, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  void P(local_block_r2r_3d)(const INT *n, MPI_Comm c, int pid, unsigned fl, LOCAL_OUT) { local_block_impl(Kind::R2R, 3, n, n, nullptr, nullptr, c, pid, fl, LOCAL_ARGS); } \
  /* ---- planners */                                                                                             \
  P(plan_s) * P(plan_many_dft_skipped)(MANY_ARGS_DECL, const int *skip, void *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2C, rnk_n, n, ni, no, howmany, iblock, oblock, skip, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_many_dft_r2c_skipped)(MANY_ARGS_DECL, const int *skip, R *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2C, rnk_n, n, ni, no, howmany, iblock, oblock, skip, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_many_dft_c2r_skipped)(MANY_ARGS_DECL, const int *skip, void *in, R *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2R, rnk_n, n, ni, no, howmany, iblock, oblock, skip, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_many_r2r_skipped)(MANY_ARGS_DECL, const int *skip, R *in, R *out, MPI_Comm c, const int *kinds, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2R, rnk_n, n, ni, no, howmany, iblock, oblock, skip, in, out, c, -1, kinds, fl); } \
  P(plan_s) * P(plan_many_dft)(MANY_ARGS_DECL, void *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2C, rnk_n, n, ni, no, howmany, iblock, oblock, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_many_dft_r2c)(MANY_ARGS_DECL, R *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2C, rnk_n, n, ni, no, howmany, iblock, oblock, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_many_dft_c2r)(MANY_ARGS_DECL, void *in, R *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2R, rnk_n, n, ni, no, howmany, iblock, oblock, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_many_r2r)(MANY_ARGS_DECL, R *in, R *out, MPI_Comm c, const int *kinds, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2R, rnk_n, n, ni, no, howmany, iblock, oblock, nullptr, in, out, c, -1, kinds, fl); } \
  P(plan_s) * P(plan_dft)(int d, const INT *n, void *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2C, d, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_dft_r2c)(int d, const INT *n, R *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2C, d, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_dft_c2r)(int d, const INT *n, void *in, R *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2R, d, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_r2r)(int d, const INT *n, R *in, R *out, MPI_Comm c, const int *kinds, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2R, d, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, -1, kinds, fl); } \
  P(plan_s) * P(plan_dft_3d)(const INT *n, void *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2C, 3, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_dft_r2c_3d)(const INT *n, R *in, void *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2C, 3, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_dft_c2r_3d)(const INT *n, void *in, R *out, MPI_Comm c, int sign, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::C2R, 3, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, sign, nullptr, fl); } \
  P(plan_s) * P(plan_r2r_3d)(const INT *n, R *in, R *out, MPI_Comm c, const int *kinds, unsigned fl) { return (P(plan_s) *)plan_impl(PREC, Kind::R2R, 3, n, n, n, 1, nullptr, nullptr, nullptr, in, out, c, -1, kinds, fl); } \
  /* ---- helpers */                                                                                              \
  INT P(prod_INT)(int d, const INT *v) { INT p = 1; for (int t = 0; t < d; t++) p *= v[t]; return p; }             \
  INT P(sum_INT)(int d, const INT *v) { INT s = 0; for (int t = 0; t < d; t++) s += v[t]; return s; }              \
  int P(equal_INT)(int d, const INT *a, const INT *b) { for (int t = 0; t < d; t++) if (a[t] != b[t]) return 0; return 1; } \
  void P(vcopy_INT)(int d, const INT *a, INT *b) { for (int t = 0; t < d; t++) b[t] = a[t]; }                      \
  void P(vadd_INT)(int d, const INT *a, const INT *b, INT *s) { for (int t = 0; t < d; t++) s[t] = a[t] + b[t]; }  \
  void P(vsub_INT)(int d, const INT *a, const INT *b, INT *s) { for (int t = 0; t < d; t++) s[t] = a[t] - b[t]; }  \
  void P(apr_complex_3d)(const void *data, const INT *ln, const INT *ls, const char *name, MPI_Comm c) { const int pm[3] = {0, 1, 2}; apr_3d<R>(data, true, ln, ls, pm, name, c); } \
  void P(apr_complex_permuted_3d)(const void *data, const INT *ln, const INT *ls, int p0, int p1, int p2, const char *name, MPI_Comm c) { const int pm[3] = {p0, p1, p2}; apr_3d<R>(data, true, ln, ls, pm, name, c); } \
  void P(apr_real_3d)(const R *data, const INT *ln, const INT *ls, const char *name, MPI_Comm c) { const int pm[3] = {0, 1, 2}; apr_3d<R>(data, false, ln, ls, pm, name, c); } \
  void P(apr_real_permuted_3d)(const R *data, const INT *ln, const INT *ls, int p0, int p1, int p2, const char *name, MPI_Comm c) { const int pm[3] = {p0, p1, p2}; apr_3d<R>(data, false, ln, ls, pm, name, c); } \
  void P(get_args)(int argc, char **argv, const char *name, int needed, unsigned type, void *param) { get_args_impl(argc, argv, name, needed, type, param); } \
  /* ---- timers */                                                                                               \
  void P(reset_timer)(P(plan_s) * pl) { if (pl) { PlanBase *b = reinterpret_cast<PlanBase *>(pl); int r = b->timer.rnk_pm; b->timer = TimerData(); b->timer.shape(r); } } \
  P(timer_s) * P(get_timer)(P(plan_s) * pl) { return pl ? (P(timer_s) *)timer_copy(&reinterpret_cast<PlanBase *>(pl)->timer) : nullptr; } \
  void P(print_average_timer)(P(plan_s) * pl, MPI_Comm c) { timer_print(reinterpret_cast<PlanBase *>(pl), c, stdout, false); } \
  void P(print_average_timer_adv)(P(plan_s) * pl, MPI_Comm c) { timer_print(reinterpret_cast<PlanBase *>(pl), c, stdout, true); } \
  void P(write_average_timer)(P(plan_s) * pl, const char *name, MPI_Comm c) { timer_write(reinterpret_cast<PlanBase *>(pl), name, c, false); } \
  void P(write_average_timer_adv)(P(plan_s) * pl, const char *name, MPI_Comm c) { timer_write(reinterpret_cast<PlanBase *>(pl), name, c, true); } \
  P(timer_s) * P(copy_timer)(P(timer_s) * o) { return (P(timer_s) *)timer_copy(reinterpret_cast<TimerData *>(o)); } \
  void P(average_timer)(P(timer_s) * t) { timer_average(reinterpret_cast<TimerData *>(t)); }                       \
  P(timer_s) * P(add_timers)(P(timer_s) * a, P(timer_s) * b) { return (P(timer_s) *)timer_add(reinterpret_cast<TimerData *>(a), reinterpret_cast<TimerData *>(b)); } \
  P(timer_s) * P(reduce_max_timer)(P(timer_s) * t, MPI_Comm c) { return (P(timer_s) *)timer_reduce_max(reinterpret_cast<TimerData *>(t), c); } \
  double *P(convert_timer2vec)(P(timer_s) * t) { std::vector<double> v = timer_to_vec(reinterpret_cast<TimerData *>(t)); double *o = (double *)malloc(v.size() * sizeof(double)); memcpy(o, v.data(), v.size() * sizeof(double)); return o; } \
  P(timer_s) * P(convert_vec2timer)(const double *v) { return (P(timer_s) *)timer_from_vec(v); }                   \
  void P(destroy_timer)(P(timer_s) * t) { delete reinterpret_cast<TimerData *>(t); }                               \
  /* ---- printing */                                                                                             \
  void P(vfprintf)(MPI_Comm c, FILE *stream, const char *format, va_list ap) { vfprintf_rank0(c, stream, format, ap); } \
  void P(fprintf)(MPI_Comm c, FILE *stream, const char *format, ...) { va_list ap; va_start(ap, format); vfprintf_rank0(c, stream, format, ap); va_end(ap); } \
  void P(printf)(MPI_Comm c, const char *format, ...) { va_list ap; va_start(ap, format); vfprintf_rank0(c, stdout, format, ap); va_end(ap); } \
  /* ---- process meshes (reference kernel/procmesh.c:36-99) */                                                   \
  int P(create_procmesh)(int rnk, MPI_Comm comm, const int *np, MPI_Comm *cart) {                                 \
    int size = 0, prod = 1, periods[8];                                                                           \
    MPI_Comm_size(comm, &size);                                                                                   \
    for (int t = 0; t < rnk; t++) { prod *= np[t]; periods[t] = 1; }                                              \
    if (prod != size) return 1;                                                                                   \
    return MPI_Cart_create(comm, rnk, np, periods, 1, cart);                                                      \
  }                                                                                                               \
  int P(create_procmesh_1d)(MPI_Comm comm, int np0, MPI_Comm *cart) { return P(create_procmesh)(1, comm, &np0, cart); } \
  int P(create_procmesh_2d)(MPI_Comm comm, int np0, int np1, MPI_Comm *cart) { int np[2] = {np0, np1}; return P(create_procmesh)(2, comm, np, cart); } \
  /* ---- ghost cells */                                                                                          \
  INT P(local_size_many_gc)(int d, const INT *ln, const INT *ls, INT howmany, const INT *gb, const INT *ga, INT *ngc, INT *gcs) { return gc_local_size(d, ln, ls, howmany, gb, ga, ngc, gcs); } \
  INT P(local_size_gc)(int d, const INT *ln, const INT *ls, const INT *gb, const INT *ga, INT *ngc, INT *gcs) { return gc_local_size(d, ln, ls, 1, gb, ga, ngc, gcs); } \
  INT P(local_size_gc_3d)(const INT *ln, const INT *ls, const INT *gb, const INT *ga, INT *ngc, INT *gcs) { return gc_local_size(3, ln, ls, 1, gb, ga, ngc, gcs); } \
  P(gcplan_s) * P(plan_many_rgc)(int d, const INT *n, INT howmany, const INT *blk, const INT *gb, const INT *ga, R *data, MPI_Comm c, unsigned fl) { return (P(gcplan_s) *)gc_plan_create(PREC, d, n, howmany, blk, gb, ga, data, c, fl, false); } \
  P(gcplan_s) * P(plan_many_cgc)(int d, const INT *n, INT howmany, const INT *blk, const INT *gb, const INT *ga, void *data, MPI_Comm c, unsigned fl) { return (P(gcplan_s) *)gc_plan_create(PREC, d, n, howmany, blk, gb, ga, data, c, fl, true); } \
  P(gcplan_s) * P(plan_rgc)(int d, const INT *n, const INT *gb, const INT *ga, R *data, MPI_Comm c, unsigned fl) { return (P(gcplan_s) *)gc_plan_create(PREC, d, n, 1, nullptr, gb, ga, data, c, fl, false); } \
  P(gcplan_s) * P(plan_cgc)(int d, const INT *n, const INT *gb, const INT *ga, void *data, MPI_Comm c, unsigned fl) { return (P(gcplan_s) *)gc_plan_create(PREC, d, n, 1, nullptr, gb, ga, data, c, fl, true); } \
  P(gcplan_s) * P(plan_rgc_3d)(const INT *n, const INT *gb, const INT *ga, R *data, MPI_Comm c, unsigned fl) { return (P(gcplan_s) *)gc_plan_create(PREC, 3, n, 1, nullptr, gb, ga, data, c, fl, false); } \
  P(gcplan_s) * P(plan_cgc_3d)(const INT *n, const INT *gb, const INT *ga, void *data, MPI_Comm c, unsigned fl) { return (P(gcplan_s) *)gc_plan_create(PREC, 3, n, 1, nullptr, gb, ga, data, c, fl, true); } \
  void P(exchange)(P(gcplan_s) * g) { gc_exchange(reinterpret_cast<GcPlan *>(g)); }                                \
  void P(reduce)(P(gcplan_s) * g) { gc_reduce(reinterpret_cast<GcPlan *>(g)); }                                    \
  void P(destroy_gcplan)(P(gcplan_s) * g) { gc_plan_destroy(reinterpret_cast<GcPlan *>(g)); }                      \
  void P(reset_gctimers)(P(gcplan_s) * g) { gc_reset_timers(reinterpret_cast<GcPlan *>(g)); }                      \
  P(gctimer_s) * P(get_gctimer_exg)(P(gcplan_s) * g) { return (P(gctimer_s) *)gc_get_timer(reinterpret_cast<GcPlan *>(g), 0); } \
  P(gctimer_s) * P(get_gctimer_red)(P(gcplan_s) * g) { return (P(gctimer_s) *)gc_get_timer(reinterpret_cast<GcPlan *>(g), 1); } \
  void P(print_average_gctimer)(P(gcplan_s) * g, MPI_Comm c) { gc_print_timers(reinterpret_cast<GcPlan *>(g), c, stdout, false); } \
  void P(print_average_gctimer_adv)(P(gcplan_s) * g, MPI_Comm c) { gc_print_timers(reinterpret_cast<GcPlan *>(g), c, stdout, true); } \
  void P(write_average_gctimer)(P(gcplan_s) * g, const char *name, MPI_Comm c) { gc_write_timers(reinterpret_cast<GcPlan *>(g), name, c, false); } \
  void P(write_average_gctimer_adv)(P(gcplan_s) * g, const char *name, MPI_Comm c) { gc_write_timers(reinterpret_cast<GcPlan *>(g), name, c, true); } \
  P(gctimer_s) * P(copy_gctimer)(P(gctimer_s) * t) { return (P(gctimer_s) *)gctimer_copy(reinterpret_cast<GcTimer *>(t)); } \
  void P(average_gctimer)(P(gctimer_s) * t) { gctimer_average(reinterpret_cast<GcTimer *>(t)); }                   \
  P(gctimer_s) * P(add_gctimers)(P(gctimer_s) * a, P(gctimer_s) * b) { return (P(gctimer_s) *)gctimer_add(reinterpret_cast<GcTimer *>(a), reinterpret_cast<GcTimer *>(b)); } \
  P(gctimer_s) * P(reduce_max_gctimer)(P(gctimer_s) * t, MPI_Comm c) { return (P(gctimer_s) *)gctimer_reduce_max(reinterpret_cast<GcTimer *>(t), c); } \
  void P(convert_gctimer2vec)(P(gctimer_s) * t, double *v) { gctimer_to_vec(reinterpret_cast<GcTimer *>(t), v); }  \
  P(gctimer_s) * P(convert_vec2gctimer)(const double *v) { return (P(gctimer_s) *)gctimer_from_vec(v); }           \
  void P(destroy_gctimer)(P(gctimer_s) * t) { delete reinterpret_cast<GcTimer *>(t); }                             \
  }

#define PFFT_NAME_D(name) pfft_##name
#define PFFT_NAME_F(name) pfftf_##name
PFFT_B200_DEFINE_API(PFFT_NAME_D, double, PREC_F64)
PFFT_B200_DEFINE_API(PFFT_NAME_F, float, PREC_F32)

// ---- FFTW allocator names (include/fftw3.h): same managed memory as pfft_malloc -------------
extern "C" {
void *fftw_malloc(size_t n) { return managed_alloc(n); }
double *fftw_alloc_real(size_t n) { return static_cast<double *>(managed_alloc(n * sizeof(double))); }
void *fftw_alloc_complex(size_t n) { return managed_alloc(n * 2 * sizeof(double)); }
void fftw_free(void *p) { if (p) { cudaDeviceSynchronize(); cudaFree(p); } }
void *fftwf_malloc(size_t n) { return managed_alloc(n); }
float *fftwf_alloc_real(size_t n) { return static_cast<float *>(managed_alloc(n * sizeof(float))); }
void *fftwf_alloc_complex(size_t n) { return managed_alloc(n * 2 * sizeof(float)); }
void fftwf_free(void *p) { if (p) { cudaDeviceSynchronize(); cudaFree(p); } }
}

// ---- the slab interface of FFTW-MPI that PFFT programs use for comparisons ---------------------
// (reference tests/bench_c2c.c:272-398 `-pfft_cmp_fftw`).  Not FFTW: the same stage kernels behind a
// 1-D process mesh -- FFTW-MPI's slab decomposition [n0/P][n1][n2] and its transposed output
// [n1/P][n0][n2] are exactly PFFT's layouts on a 1-D mesh (reference doc/intro.tex, api/pfft.h:423).
namespace {
struct FftwCompatPlan {
  pfft_plan_s *plan;
  MPI_Comm cart;
};
unsigned fftw_flags_to_pfft(unsigned f) {
  unsigned fl = 0;
  if (f & (1u << 29)) fl |= F_TRANSPOSED_IN;    // FFTW_MPI_TRANSPOSED_IN
  if (f & (1u << 30)) fl |= F_TRANSPOSED_OUT;   // FFTW_MPI_TRANSPOSED_OUT
  if (f & (1u << 0)) fl |= F_DESTROY_INPUT;     // FFTW_DESTROY_INPUT
  return fl;
}
}  // namespace

extern "C" {
ptrdiff_t fftw_mpi_local_size_3d_transposed(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm, ptrdiff_t *local_n0,
                                            ptrdiff_t *local_0_start, ptrdiff_t *local_n1, ptrdiff_t *local_1_start) {
  const INT n[3] = {n0, n1, n2};
  INT lni[3], lis[3], lno[3], los[3];
  MPI_Comm cart = MPI_COMM_NULL;
  int np = 1;
  MPI_Comm_size(comm, &np);
  if (pfft_create_procmesh_1d(comm, np, &cart)) return 0;
  const INT alloc = pfft_local_size_dft_3d(n, cart, F_TRANSPOSED_OUT, lni, lis, lno, los);
  MPI_Comm_free(&cart);
  *local_n0 = lni[0];
  *local_0_start = lis[0];
  *local_n1 = lno[1];       // PFFT reports the transposed block in logical order: dimension 1 is the split one
  *local_1_start = los[1];
  return alloc;
}

void *fftw_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, void *in, void *out, MPI_Comm comm, int sign, unsigned flags) {
  const INT n[3] = {n0, n1, n2};
  MPI_Comm cart = MPI_COMM_NULL;
  int np = 1;
  MPI_Comm_size(comm, &np);
  if (pfft_create_procmesh_1d(comm, np, &cart)) return nullptr;
  pfft_plan_s *pl = pfft_plan_dft_3d(n, in, out, cart, sign, fftw_flags_to_pfft(flags));
  if (!pl) {
    MPI_Comm_free(&cart);
    return nullptr;
  }
  return new FftwCompatPlan{pl, cart};
}

void *fftw_plan_dft_3d(int n0, int n1, int n2, void *in, void *out, int sign, unsigned flags) {
  return fftw_mpi_plan_dft_3d(n0, n1, n2, in, out, MPI_COMM_SELF, sign, flags);
}

void fftw_execute(void *p) {
  if (p) pfft_execute(static_cast<FftwCompatPlan *>(p)->plan);
}

void fftw_destroy_plan(void *p) {
  if (!p) return;
  FftwCompatPlan *c = static_cast<FftwCompatPlan *>(p);
  pfft_destroy_plan(c->plan);
  MPI_Comm_free(&c->cart);
  delete c;
}
}

// ---- extensions -------------------------------------------------------------------------
namespace pfb { void set_default_transport(int t); }

extern "C" {

int pfftb200_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  ensure_device();
  return e == cudaSuccess ? 0 : 1;
}

int pfftb200_get_device(void) {
  int d = -1;
  cudaGetDevice(&d);
  return d;
}

void pfftb200_set_stream(void *s) { set_default_stream(static_cast<cudaStream_t>(s)); }
void *pfftb200_get_stream(void) { return default_stream(); }

void *pfftb200_malloc_device(size_t bytes) {
  ensure_device();
  void *p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void pfftb200_free_device(void *p) {
  if (p) cudaFree(p);
}

int pfftb200_set_transport(const char *name) {
  if (!name || !strcmp(name, "auto")) set_default_transport(TR_AUTO);
  else if (!strcmp(name, "nccl")) set_default_transport(TR_NCCL);
  else if (!strcmp(name, "p2p")) set_default_transport(TR_P2P);
  else return 1;
  return 0;
}

const char *pfftb200_get_transport(void) {
  switch (default_transport()) {
    case TR_NCCL: return "nccl";
    case TR_P2P: return "p2p";
    default: return "auto";
  }
}

size_t pfftb200_plan_describe(const void *plan, char *buf, size_t buflen) {
  if (!plan) return 0;
  const PlanBase *pl = static_cast<const PlanBase *>(plan);
  std::string j = schedule_to_json(pl->sched);
  // append per-stage kernel choice
  std::string k = ",\"kernels\":[";
  for (size_t i = 0; i < pl->use_pow2.size(); i++) {
    k += i ? "," : "";
    k += pl->kernel_kind[i] == KERNEL_POW2 ? "\"pow2\"" : (pl->kernel_kind[i] == KERNEL_REG ? "\"reg\"" : (pl->kernel_kind[i] == KERNEL_MIXED ? "\"mixed\"" : "\"generic\""));
  }
  k += "],\"kernel_names\":[";
  for (size_t i = 0; i < pl->params.size(); i++) {
    // the __global__ function the launchers pick for this stage (fft_pow2.cu: launch_stage_pow2 / launch_block_class)
    const StageParams &sp = pl->params[i];
    const char *nm = pl->kernel_kind[i] == KERNEL_MIXED ? "stage_mixed_kernel" : (pl->kernel_kind[i] == KERNEL_REG ? "stage_reg_kernel" : "stage_generic_kernel");
    if (pl->use_pow2[i]) nm = sp.ntile > 0 ? "stage_blk_kernel" : ((sp.fast && sp.istride == 1) ? "fused_pair_kernel<FUSED=false>" : "stage_pow2_kernel");
    k += i ? "," : "";
    k += std::string("\"") + nm + "\"";
  }
  k += "],\"tile_lines\":[";
  for (size_t i = 0; i < pl->params.size(); i++) k += (i ? "," : "") + std::to_string(pl->params[i].tl);
  k += "],\"fused_pair\":" + std::to_string(pl->fuse.possible ? pl->fuse.first : -1);
  k += ",\"fused_active\":" + std::to_string(pl->fuse.active ? 1 : 0);
  k += ",\"fused_ring_bytes\":" + std::to_string(pl->fuse.possible ? pl->fuse.ring_bytes : 0);
  k += ",\"fused_tile_lines\":" + std::to_string(pl->fuse.possible ? pl->fuse.a.tl : 0);
  k += ",\"transport\":\"";
  k += pl->transport == TR_NCCL ? "nccl" : "p2p";
  k += "\",\"exchange_ordering\":\"";
  k += pl->transport == TR_NCCL ? "stream (nccl)" : (transport_device_sync(pl) ? "device flags" : "host barriers");
  k += "\"}";
  j = j.substr(0, j.size() - 1) + k;
  if (buf && buflen) {
    size_t m = j.size() < buflen - 1 ? j.size() : buflen - 1;
    memcpy(buf, j.data(), m);
    buf[m] = 0;
  }
  return j.size() + 1;
}

void pfftb200_execute_async(const void *plan, void *in, void *out) {
  if (plan) plan_execute(const_cast<PlanBase *>(static_cast<const PlanBase *>(plan)), in, out, false);
}

unsigned long long pfftb200_launch_count(void) { return launch_counter(); }

void pfftb200_enable_stage_timing(const void *plan, int on) {
  if (plan) const_cast<PlanBase *>(static_cast<const PlanBase *>(plan))->stage_timing = on != 0;
}

int pfftb200_exchange_times(const void *plan, double *ms, int max_exchanges) {
  if (!plan) return 0;
  const PlanBase *pl = static_cast<const PlanBase *>(plan);
  int n = (int)pl->last_xch_ms.size();
  for (int i = 0; i < n && i < max_exchanges; i++) ms[i] = pl->last_xch_ms[i];
  return n;
}

int pfftb200_stage_times(const void *plan, double *ms, int max_stages) {
  if (!plan) return 0;
  const PlanBase *pl = static_cast<const PlanBase *>(plan);
  int n = (int)pl->last_stage_ms.size();
  for (int i = 0; i < n && i < max_stages; i++) ms[i] = pl->last_stage_ms[i];
  return n;
}

}  // extern "C"
