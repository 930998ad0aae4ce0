/* include/fftw3-mpi.h -- compatibility shim, NOT FFTW-MPI (see fftw3.h). */
#ifndef PFFT_B200_FFTW3_MPI_COMPAT_H
#define PFFT_B200_FFTW3_MPI_COMPAT_H 1
#include <fftw3.h>
#include <mpi.h>

#define FFTW_MPI_DEFAULT_BLOCK (0)
#define FFTW_MPI_TRANSPOSED_IN (1U << 29)
#define FFTW_MPI_TRANSPOSED_OUT (1U << 30)

/* FFTW-MPI's slab interface as far as PFFT's benchmark program calls it (see fftw3.h): slabs
 * [n0/P][n1][n2], transposed output [n1/P][n0][n2] -- PFFT's own layouts on a 1-D process mesh. */
#ifdef __cplusplus
extern "C" {
#endif
ptrdiff_t fftw_mpi_local_size_3d_transposed(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm, ptrdiff_t *local_n0,
                                            ptrdiff_t *local_0_start, ptrdiff_t *local_n1, ptrdiff_t *local_1_start);
fftw_plan fftw_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in, fftw_complex *out, MPI_Comm comm,
                               int sign, unsigned flags);
#ifdef __cplusplus
}
#endif

#endif
