/* include/pfft.h -- C API of pfft_b200, source-compatible with PFFT 1.0.8-alpha.
 *
 * This is the drop-in boundary: every name, argument order and flag value below
 * is what a program written against the reference's header (api/pfft.h:69-572)
 * expects; the implementation behind it is new (sm_100a kernels + NVLink
 * transports).  Each block cites the reference interface it
 * replaces.  Double (`pfft_`) and single (`pfftf_`) precision are provided;
 * `pfftl_` (long double, api/pfft.h:524) has no GPU equivalent and is not declared.
 *
 * Memory: `pfft_alloc_*` returns CUDA managed memory (host code may dereference
 * it, kernels use it in place).  Plans also accept plain device pointers (used
 * as is) and ordinary host pointers (staged through device memory inside
 * pfft_execute).  See include/pfft_b200.h for the non-PFFT extensions.
 */
#ifndef PFFT_H
#define PFFT_H 1
#define PFFT_B200 1

#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <mpi.h>
#include <fftw3-mpi.h>

#ifdef __cplusplus
extern "C" {
#endif

/* r2r kinds and signs: same numeric values as FFTW's (api/pfft.h:45-58) */
#define PFFT_R2HC FFTW_R2HC
#define PFFT_HC2R FFTW_HC2R
#define PFFT_DHT FFTW_DHT
#define PFFT_REDFT00 FFTW_REDFT00
#define PFFT_REDFT01 FFTW_REDFT01
#define PFFT_REDFT10 FFTW_REDFT10
#define PFFT_REDFT11 FFTW_REDFT11
#define PFFT_RODFT00 FFTW_RODFT00
#define PFFT_RODFT01 FFTW_RODFT01
#define PFFT_RODFT10 FFTW_RODFT10
#define PFFT_RODFT11 FFTW_RODFT11
#define PFFT_FORWARD FFTW_FORWARD
#define PFFT_BACKWARD FFTW_BACKWARD

/* plan flags (api/pfft.h:528-545): bit values are ABI */
#define PFFT_TRANSPOSED_NONE (0U)
#define PFFT_TRANSPOSED_IN (1U << 0)
#define PFFT_TRANSPOSED_OUT (1U << 1)
#define PFFT_SHIFTED_NONE (0U)
#define PFFT_SHIFTED_IN (1U << 2)
#define PFFT_SHIFTED_OUT (1U << 3)
#define PFFT_MEASURE (0U)
#define PFFT_ESTIMATE (1U << 4)
#define PFFT_PATIENT (1U << 5)
#define PFFT_EXHAUSTIVE (1U << 6)
#define PFFT_NO_TUNE (0U)
#define PFFT_TUNE (1U << 7)
#define PFFT_PRESERVE_INPUT (1U << 8)
#define PFFT_DESTROY_INPUT (1U << 9)
#define PFFT_BUFFERED_INPLACE (1U << 10)
#define PFFT_PADDED_R2C (1U << 11)
#define PFFT_PADDED_C2R (PFFT_PADDED_R2C)

#define PFFT_DEFAULT_BLOCK FFTW_MPI_DEFAULT_BLOCK
#define PFFT_DEFAULT_BLOCKS NULL
#define PFFT_NO_GCELLS NULL

/* argument types of pfft_get_args (api/pfft.h:553-559) */
#define PFFT_INT (1U)
#define PFFT_PTRDIFF_T (2U)
#define PFFT_FLOAT (3U)
#define PFFT_DOUBLE (4U)
#define PFFT_LDOUBLE (5U)
#define PFFT_UNSIGNED (6U)
#define PFFT_SWITCH (7U)

/* ghost-cell flags (api/pfft.h:561-572) */
#define PFFT_GC_TRANSPOSED_NONE (0U)
#define PFFT_GC_TRANSPOSED (1U << 0)
#define PFFT_GC_SENDRECV (1U << 1)
#define PFFT_GC_RMA (1U << 2)
#define PFFT_GC_R2C (1U << 3)
#define PFFT_GC_PADDED (1U << 4)
#define PFFT_GC_C2R (PFFT_GC_R2C)
#define PFFT_GC_PADDED_R2C (PFFT_GC_R2C | PFFT_GC_PADDED)
#define PFFT_GC_PADDED_C2R (PFFT_GC_C2R | PFFT_GC_PADDED)

typedef fftw_complex pfft_complex;   /* api/pfft.h:515 */
typedef fftwf_complex pfftf_complex; /* api/pfft.h:516 */
typedef fftw_r2r_kind pfft_r2r_kind;
typedef fftwf_r2r_kind pfftf_r2r_kind;

#define PFFT_CONCAT(prefix, name) prefix##name
#define PFFT_MANGLE_DOUBLE(name) PFFT_CONCAT(pfft_, name)
#define PFFT_MANGLE_FLOAT(name) PFFT_CONCAT(pfftf_, name)

/* argument packs shared by many prototypes */
#define PFFT_B200_LOCAL_OUT_ ptrdiff_t *local_ni, ptrdiff_t *local_i_start, ptrdiff_t *local_no, ptrdiff_t *local_o_start
#define PFFT_B200_MANY_ int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no, ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock
#define PFFT_B200_BLOCK_ const ptrdiff_t *n, const ptrdiff_t *local_n, const ptrdiff_t *local_start

#define PFFT_B200_DECLARE_API(P, R, C, K)                                                           \
  /* opaque handles (api/pfft.h:71-74) */                                                           \
  typedef struct P(plan_s) *P(plan);                                                                \
  typedef struct P(gcplan_s) *P(gcplan);                                                            \
  typedef struct P(timer_s) *P(timer);                                                              \
  typedef struct P(gctimer_s) *P(gctimer);                                                          \
  /* library life cycle + memory (api/pfft.h:77-85; kernel/malloc.c:28-42) */                       \
  void P(init)(void);                                                                               \
  void P(cleanup)(void);                                                                            \
  void P(plan_with_nthreads)(int nthreads);                                                         \
  int P(get_nthreads)(void);                                                                        \
  void *P(malloc)(size_t n);                                                                        \
  R *P(alloc_real)(size_t n);                                                                       \
  C *P(alloc_complex)(size_t n);                                                                    \
  void P(free)(void *p);                                                                            \
  /* execution (api/pfft.h:87-92; api/api-basic.c:1044-1183) */                                     \
  void P(execute)(const P(plan) plan);                                                              \
  void P(execute_dft)(const P(plan) plan, C *in, C *out);                                           \
  void P(execute_dft_r2c)(const P(plan) plan, R *in, C *out);                                       \
  void P(execute_dft_c2r)(const P(plan) plan, C *in, R *out);                                       \
  void P(execute_r2r)(const P(plan) plan, R *in, R *out);                                           \
  void P(destroy_plan)(P(plan) plan);                                                               \
  /* test-data contract (api/pfft.h:94-149; api/api-basic.c:60-189,254-479) */                      \
  void P(init_input_complex_3d)(PFFT_B200_BLOCK_, C *data);                                         \
  void P(init_input_complex)(int rnk_n, PFFT_B200_BLOCK_, C *data);                                 \
  void P(init_input_complex_hermitian_3d)(PFFT_B200_BLOCK_, C *data);                               \
  void P(init_input_complex_hermitian)(int rnk_n, PFFT_B200_BLOCK_, C *data);                       \
  void P(init_input_real_3d)(PFFT_B200_BLOCK_, R *data);                                            \
  void P(init_input_real)(int rnk_n, PFFT_B200_BLOCK_, R *data);                                    \
  void P(clear_input_complex_3d)(PFFT_B200_BLOCK_, C *data);                                        \
  void P(clear_input_complex)(int rnk_n, PFFT_B200_BLOCK_, C *data);                                \
  void P(clear_input_complex_hermitian_3d)(PFFT_B200_BLOCK_, C *data);                              \
  void P(clear_input_complex_hermitian)(int rnk_n, PFFT_B200_BLOCK_, C *data);                      \
  void P(clear_input_real_3d)(PFFT_B200_BLOCK_, R *data);                                           \
  void P(clear_input_real)(int rnk_n, PFFT_B200_BLOCK_, R *data);                                   \
  R P(check_output_complex_3d)(PFFT_B200_BLOCK_, const C *data, MPI_Comm comm);                     \
  R P(check_output_complex)(int rnk_n, PFFT_B200_BLOCK_, const C *data, MPI_Comm comm);             \
  R P(check_output_complex_hermitian_3d)(PFFT_B200_BLOCK_, const C *data, MPI_Comm comm);           \
  R P(check_output_complex_hermitian)(int rnk_n, PFFT_B200_BLOCK_, const C *data, MPI_Comm comm);   \
  R P(check_output_real_3d)(PFFT_B200_BLOCK_, const R *data, MPI_Comm comm);                        \
  R P(check_output_real)(int rnk_n, PFFT_B200_BLOCK_, const R *data, MPI_Comm comm);                \
  /* data distribution (api/pfft.h:151-275; kernel/partrafo.c:99-315) */                            \
  ptrdiff_t P(local_size_dft_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_dft_r2c_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_dft_c2r_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_r2r_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_dft)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_dft_r2c)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_dft_c2r)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_r2r)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_many_dft)(PFFT_B200_MANY_, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_many_dft_r2c)(PFFT_B200_MANY_, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_many_dft_c2r)(PFFT_B200_MANY_, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  ptrdiff_t P(local_size_many_r2r)(PFFT_B200_MANY_, MPI_Comm comm_cart, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_dft_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_dft_r2c_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_dft_c2r_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_r2r_3d)(const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_dft)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_dft_r2c)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_dft_c2r)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_r2r)(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_many_dft)(int rnk_n, const ptrdiff_t *ni, const ptrdiff_t *no, const ptrdiff_t *iblock, const ptrdiff_t *oblock, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_many_dft_r2c)(int rnk_n, const ptrdiff_t *ni, const ptrdiff_t *no, const ptrdiff_t *iblock, const ptrdiff_t *oblock, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_many_dft_c2r)(int rnk_n, const ptrdiff_t *ni, const ptrdiff_t *no, const ptrdiff_t *iblock, const ptrdiff_t *oblock, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  void P(local_block_many_r2r)(int rnk_n, const ptrdiff_t *ni, const ptrdiff_t *no, const ptrdiff_t *iblock, const ptrdiff_t *oblock, MPI_Comm comm_cart, int pid, unsigned pfft_flags, PFFT_B200_LOCAL_OUT_); \
  /* planners (api/pfft.h:277-343; kernel/partrafo.c:317-525) */                                    \
  P(plan) P(plan_dft_3d)(const ptrdiff_t *n, C *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_dft_r2c_3d)(const ptrdiff_t *n, R *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_dft_c2r_3d)(const ptrdiff_t *n, C *in, R *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_r2r_3d)(const ptrdiff_t *n, R *in, R *out, MPI_Comm comm_cart, const K *kinds, unsigned pfft_flags); \
  P(plan) P(plan_dft)(int rnk_n, const ptrdiff_t *n, C *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_dft_r2c)(int rnk_n, const ptrdiff_t *n, R *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_dft_c2r)(int rnk_n, const ptrdiff_t *n, C *in, R *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_r2r)(int rnk_n, const ptrdiff_t *n, R *in, R *out, MPI_Comm comm_cart, const K *kinds, unsigned pfft_flags); \
  P(plan) P(plan_many_dft)(PFFT_B200_MANY_, C *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_many_dft_r2c)(PFFT_B200_MANY_, R *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_many_dft_c2r)(PFFT_B200_MANY_, C *in, R *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_many_r2r)(PFFT_B200_MANY_, R *in, R *out, MPI_Comm comm_cart, const K *kinds, unsigned pfft_flags); \
  P(plan) P(plan_many_dft_skipped)(PFFT_B200_MANY_, const int *skip_trafos, C *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_many_dft_r2c_skipped)(PFFT_B200_MANY_, const int *skip_trafos, R *in, C *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_many_dft_c2r_skipped)(PFFT_B200_MANY_, const int *skip_trafos, C *in, R *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags); \
  P(plan) P(plan_many_r2r_skipped)(PFFT_B200_MANY_, const int *skip_trafos, R *in, R *out, MPI_Comm comm_cart, const K *kinds, unsigned pfft_flags); \
  /* small helpers (api/pfft.h:345-384; util/util.c, util/getargs.c:41-66) */                      \
  ptrdiff_t P(prod_INT)(int d, const ptrdiff_t *vec);                                               \
  ptrdiff_t P(sum_INT)(int d, const ptrdiff_t *vec);                                                \
  int P(equal_INT)(int d, const ptrdiff_t *vec1, const ptrdiff_t *vec2);                            \
  void P(vcopy_INT)(int d, const ptrdiff_t *vec1, ptrdiff_t *vec2);                                 \
  void P(vadd_INT)(int d, const ptrdiff_t *vec1, const ptrdiff_t *vec2, ptrdiff_t *sum);            \
  void P(vsub_INT)(int d, const ptrdiff_t *vec1, const ptrdiff_t *vec2, ptrdiff_t *sum);            \
  void P(apr_complex_3d)(const C *data, const ptrdiff_t *local_n, const ptrdiff_t *local_start, const char *name, MPI_Comm comm); \
  void P(apr_complex_permuted_3d)(const C *data, const ptrdiff_t *local_n, const ptrdiff_t *local_start, int perm0, int perm1, int perm2, const char *name, MPI_Comm comm); \
  void P(apr_real_3d)(const R *data, const ptrdiff_t *local_n, const ptrdiff_t *local_start, const char *name, MPI_Comm comm); \
  void P(apr_real_permuted_3d)(const R *data, const ptrdiff_t *local_n, const ptrdiff_t *local_start, int perm0, int perm1, int perm2, const char *name, MPI_Comm comm); \
  void P(get_args)(int argc, char **argv, const char *name, int neededArgs, unsigned type, void *parameter); \
  /* timers (api/pfft.h:386-412; kernel/timer.c) */                                                 \
  void P(reset_timer)(P(plan) ths);                                                                 \
  P(timer) P(get_timer)(const P(plan) ths);                                                         \
  void P(print_average_timer)(const P(plan) ths, MPI_Comm comm);                                    \
  void P(print_average_timer_adv)(const P(plan) ths, MPI_Comm comm);                                \
  void P(write_average_timer)(const P(plan) ths, const char *name, MPI_Comm comm);                  \
  void P(write_average_timer_adv)(const P(plan) ths, const char *name, MPI_Comm comm);              \
  P(timer) P(copy_timer)(const P(timer) orig);                                                      \
  void P(average_timer)(P(timer) ths);                                                              \
  P(timer) P(add_timers)(const P(timer) sum1, const P(timer) sum2);                                 \
  P(timer) P(reduce_max_timer)(const P(timer) ths, MPI_Comm comm);                                  \
  double *P(convert_timer2vec)(const P(timer) ths);                                                 \
  P(timer) P(convert_vec2timer)(const double *times);                                               \
  void P(destroy_timer)(P(timer) ths);                                                              \
  /* rank-0 printing (api/pfft.h:414-421; util/util.c:175-204) */                                   \
  void P(vfprintf)(MPI_Comm comm, FILE *stream, const char *format, va_list ap);                    \
  void P(fprintf)(MPI_Comm comm, FILE *stream, const char *format, ...);                            \
  void P(printf)(MPI_Comm comm, const char *format, ...);                                           \
  /* process meshes (api/pfft.h:423-431; kernel/procmesh.c:36-64) */                                \
  int P(create_procmesh)(int rnk_n, MPI_Comm comm, const int *np, MPI_Comm *comm_cart);             \
  int P(create_procmesh_1d)(MPI_Comm comm, int np0, MPI_Comm *comm_cart_1d);                        \
  int P(create_procmesh_2d)(MPI_Comm comm, int np0, int np1, MPI_Comm *comm_cart_2d);               \
  /* ghost cells (api/pfft.h:433-502; gcell/gcells_plan.c) */                                       \
  ptrdiff_t P(local_size_gc_3d)(const ptrdiff_t *local_n, const ptrdiff_t *local_start, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, ptrdiff_t *local_ngc, ptrdiff_t *local_gc_start); \
  ptrdiff_t P(local_size_gc)(int rnk_n, const ptrdiff_t *local_n, const ptrdiff_t *local_start, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, ptrdiff_t *local_ngc, ptrdiff_t *local_gc_start); \
  ptrdiff_t P(local_size_many_gc)(int rnk_n, const ptrdiff_t *local_n, const ptrdiff_t *local_start, ptrdiff_t howmany, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, ptrdiff_t *local_ngc, ptrdiff_t *local_gc_start); \
  P(gcplan) P(plan_rgc_3d)(const ptrdiff_t *n, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, R *data, MPI_Comm comm_cart, unsigned gc_flags); \
  P(gcplan) P(plan_cgc_3d)(const ptrdiff_t *n, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, C *data, MPI_Comm comm_cart, unsigned gc_flags); \
  P(gcplan) P(plan_rgc)(int rnk_n, const ptrdiff_t *n, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, R *data, MPI_Comm comm_cart, unsigned gc_flags); \
  P(gcplan) P(plan_cgc)(int rnk_n, const ptrdiff_t *n, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, C *data, MPI_Comm comm_cart, unsigned gc_flags); \
  P(gcplan) P(plan_many_rgc)(int rnk_n, const ptrdiff_t *n, ptrdiff_t howmany, const ptrdiff_t *block, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, R *data, MPI_Comm comm_cart, unsigned gc_flags); \
  P(gcplan) P(plan_many_cgc)(int rnk_n, const ptrdiff_t *n, ptrdiff_t howmany, const ptrdiff_t *block, const ptrdiff_t *gc_below, const ptrdiff_t *gc_above, C *data, MPI_Comm comm_cart, unsigned gc_flags); \
  void P(exchange)(P(gcplan) ths);                                                                  \
  void P(reduce)(P(gcplan) ths);                                                                    \
  void P(destroy_gcplan)(P(gcplan) ths);                                                            \
  void P(reset_gctimers)(P(gcplan) ths);                                                            \
  P(gctimer) P(get_gctimer_exg)(const P(gcplan) ths);                                               \
  P(gctimer) P(get_gctimer_red)(const P(gcplan) ths);                                               \
  void P(print_average_gctimer)(const P(gcplan) ths, MPI_Comm comm);                                \
  void P(print_average_gctimer_adv)(const P(gcplan) ths, MPI_Comm comm);                            \
  void P(write_average_gctimer)(const P(gcplan) ths, const char *name, MPI_Comm comm);              \
  void P(write_average_gctimer_adv)(const P(gcplan) ths, const char *name, MPI_Comm comm);          \
  P(gctimer) P(copy_gctimer)(const P(gctimer) orig);                                                \
  void P(average_gctimer)(P(gctimer) ths);                                                          \
  P(gctimer) P(add_gctimers)(const P(gctimer) sum1, const P(gctimer) sum2);                         \
  P(gctimer) P(reduce_max_gctimer)(const P(gctimer) ths, MPI_Comm comm);                            \
  void P(convert_gctimer2vec)(const P(gctimer) ths, double *times);                                 \
  P(gctimer) P(convert_vec2gctimer)(const double *times);                                           \
  void P(destroy_gctimer)(P(gctimer) ths);

PFFT_B200_DECLARE_API(PFFT_MANGLE_DOUBLE, double, pfft_complex, pfft_r2r_kind)
PFFT_B200_DECLARE_API(PFFT_MANGLE_FLOAT, float, pfftf_complex, pfftf_r2r_kind)

#ifdef __cplusplus
}
#endif
#endif /* PFFT_H */
