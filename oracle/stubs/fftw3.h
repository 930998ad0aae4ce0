/* oracle/stubs/fftw3.h -- TEST INFRASTRUCTURE ONLY.
 * Types and constants of FFTW 3.3's public header that the reference sources
 * mention, with prototypes only for the functions they call.  FFTW itself is
 * a third-party dependency that is absent from /root/reference and from this
 * image (reference pins >= 3.3.3, pfft.pc.in:10); none of these functions is
 * ever linked -- the integer oracle stops before any of them is reached. */
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <stddef.h>
#include <stdio.h>

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

enum fftw_r2r_kind_do_not_use_me {
  FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3, FFTW_REDFT01 = 4,
  FFTW_REDFT10 = 5, FFTW_REDFT11 = 6, FFTW_RODFT00 = 7, FFTW_RODFT01 = 8,
  FFTW_RODFT10 = 9, FFTW_RODFT11 = 10
};
struct fftw_iodim64_do_not_use_me { ptrdiff_t n, is, os; };

#define FFTW_CONCAT(a, b) a##b
#define FFTW_MANGLE_DOUBLE(name) FFTW_CONCAT(fftw_, name)
#define FFTW_MANGLE_FLOAT(name) FFTW_CONCAT(fftwf_, name)
#define FFTW_MANGLE_LONG_DOUBLE(name) FFTW_CONCAT(fftwl_, name)

#if defined(_Complex_I) && defined(complex) && defined(I)
#define ORACLE_CPLX(R, C) typedef R _Complex C
#else
#define ORACLE_CPLX(R, C) typedef R C[2]
#endif

#define ORACLE_FFTW_API(X, R, C)                                               \
  ORACLE_CPLX(R, C);                                                           \
  typedef struct X(plan_s) * X(plan);                                          \
  typedef struct fftw_iodim64_do_not_use_me X(iodim64);                        \
  typedef enum fftw_r2r_kind_do_not_use_me X(r2r_kind);                        \
  void X(execute)(const X(plan) p);                                            \
  void X(execute_dft)(const X(plan) p, C *in, C *out);                         \
  void X(execute_dft_r2c)(const X(plan) p, R *in, C *out);                     \
  void X(execute_dft_c2r)(const X(plan) p, C *in, R *out);                     \
  void X(execute_r2r)(const X(plan) p, R *in, R *out);                         \
  X(plan) X(plan_guru64_dft)(int, const X(iodim64) *, int, const X(iodim64) *, \
                             C *, C *, int, unsigned);                         \
  X(plan) X(plan_guru64_dft_r2c)(int, const X(iodim64) *, int,                 \
                                 const X(iodim64) *, R *, C *, unsigned);      \
  X(plan) X(plan_guru64_dft_c2r)(int, const X(iodim64) *, int,                 \
                                 const X(iodim64) *, C *, R *, unsigned);      \
  X(plan) X(plan_guru64_r2r)(int, const X(iodim64) *, int, const X(iodim64) *, \
                             R *, R *, const X(r2r_kind) *, unsigned);         \
  void X(destroy_plan)(X(plan) p);                                             \
  void X(forget_wisdom)(void);                                                 \
  void *X(malloc)(size_t n);                                                   \
  R *X(alloc_real)(size_t n);                                                  \
  C *X(alloc_complex)(size_t n);                                               \
  void X(free)(void *p);                                                       \
  int X(init_threads)(void);                                                   \
  void X(plan_with_nthreads)(int n);                                           \
  void X(cleanup_threads)(void);

ORACLE_FFTW_API(FFTW_MANGLE_DOUBLE, double, fftw_complex)
ORACLE_FFTW_API(FFTW_MANGLE_FLOAT, float, fftwf_complex)
ORACLE_FFTW_API(FFTW_MANGLE_LONG_DOUBLE, long double, fftwl_complex)
#endif
