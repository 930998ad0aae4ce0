/* include/fftw3.h -- compatibility shim, NOT FFTW.
 *
 * The reference's public header pulls in <fftw3-mpi.h> (api/pfft.h:29) and aliases
 * its complex / r2r-kind types and a few constants to FFTW's (api/pfft.h:45-58,515-520).
 * pfft_b200 computes every transform with its own sm_100a kernels, so only those
 * type names and constant values are provided here, so that programs written
 * against PFFT keep compiling.  No fftw_* function exists in this library.
 * The numeric values are FFTW 3.3's public ABI (fftw3.h of FFTW 3.3.x).
 */
#ifndef PFFT_B200_FFTW3_COMPAT_H
#define PFFT_B200_FFTW3_COMPAT_H 1
#include <stddef.h>

#define PFFT_B200_FFTW_SHIM 1

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)

#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

typedef enum {
  FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2,
  FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5, FFTW_REDFT11 = 6,
  FFTW_RODFT00 = 7, FFTW_RODFT01 = 8, FFTW_RODFT10 = 9, FFTW_RODFT11 = 10
} pfft_b200_r2r_kind_t;
typedef pfft_b200_r2r_kind_t fftw_r2r_kind;
typedef pfft_b200_r2r_kind_t fftwf_r2r_kind;
typedef pfft_b200_r2r_kind_t fftwl_r2r_kind;

/* C99 complex when <complex.h> was included first (as the reference tests do,
 * tests/simple_check_c2c.c:1-2), else the two-element array FFTW also uses. */
#if !defined(__cplusplus) && defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex fftw_complex;
typedef float _Complex fftwf_complex;
typedef long double _Complex fftwl_complex;
#else
typedef double fftw_complex[2];
typedef float fftwf_complex[2];
typedef long double fftwl_complex[2];
#endif

/* FFTW's allocator names, used directly by a few PFFT programs
 * (e.g. /root/reference/tests/simple_check_ousam_r2r.c): same memory as pfft_malloc. */
#ifdef __cplusplus
extern "C" {
#endif
void *fftw_malloc(size_t n);
double *fftw_alloc_real(size_t n);
fftw_complex *fftw_alloc_complex(size_t n);
void fftw_free(void *p);
void *fftwf_malloc(size_t n);
float *fftwf_alloc_real(size_t n);
fftwf_complex *fftwf_alloc_complex(size_t n);
void fftwf_free(void *p);

/* The few FFTW planner/executor names PFFT programs use to compare PFFT with FFTW
 * (/root/reference/tests/bench_c2c.c:272-398, `-pfft_cmp_fftw`).  NOT FFTW: the same sm_100a stage
 * kernels behind a one-rank (fftw_plan_dft_3d) or 1-D slab (fftw_mpi_plan_dft_3d, fftw3-mpi.h) mesh. */
typedef struct pfft_b200_fftw_plan_s *fftw_plan;
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
#ifdef __cplusplus
}
#endif

#define FFTW_CONCAT(prefix, name) prefix##name
#define FFTW_MANGLE_DOUBLE(name) FFTW_CONCAT(fftw_, name)
#define FFTW_MANGLE_FLOAT(name) FFTW_CONCAT(fftwf_, name)
#define FFTW_MANGLE_LONG_DOUBLE(name) FFTW_CONCAT(fftwl_, name)

#endif
