/* oracle/stubs/fftw3-mpi.h -- TEST INFRASTRUCTURE ONLY.
 * Constants and prototypes of FFTW-MPI 3.3 named by the reference
 * (kernel/transpose.c:132,174; api/api-basic.c:217,227). Never linked. */
#ifndef ORACLE_STUB_FFTW3_MPI_H
#define ORACLE_STUB_FFTW3_MPI_H
#include <fftw3.h>
#include <mpi.h>

#define FFTW_MPI_DEFAULT_BLOCK (0)
#define FFTW_MPI_SCRAMBLED_IN (1U << 27)
#define FFTW_MPI_SCRAMBLED_OUT (1U << 28)
#define FFTW_MPI_TRANSPOSED_IN (1U << 29)
#define FFTW_MPI_TRANSPOSED_OUT (1U << 30)

#define ORACLE_FFTW_MPI_API(X, XM, R)                                          \
  void XM(init)(void);                                                         \
  void XM(cleanup)(void);                                                      \
  ptrdiff_t XM(local_size_many_transposed)(                                    \
      int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t b0,            \
      ptrdiff_t b1, MPI_Comm comm, ptrdiff_t *ln0, ptrdiff_t *l0s,             \
      ptrdiff_t *ln1, ptrdiff_t *l1s);                                         \
  X(plan) XM(plan_many_transpose)(ptrdiff_t n0, ptrdiff_t n1,                  \
                                  ptrdiff_t howmany, ptrdiff_t b0,             \
                                  ptrdiff_t b1, R *in, R *out, MPI_Comm comm,  \
                                  unsigned flags);                             \
  void XM(execute_r2r)(const X(plan) p, R *in, R *out);

#define ORACLE_XMD(n) fftw_mpi_##n
ORACLE_FFTW_MPI_API(FFTW_MANGLE_DOUBLE, ORACLE_XMD, double)
#define ORACLE_XMF(n) fftwf_mpi_##n
#define ORACLE_XML(n) fftwl_mpi_##n
ORACLE_FFTW_MPI_API(FFTW_MANGLE_FLOAT, ORACLE_XMF, float)
ORACLE_FFTW_MPI_API(FFTW_MANGLE_LONG_DOUBLE, ORACLE_XML, long double)
#endif
