/* include/fftw3-mpi.h -- compatibility shim, NOT FFTW-MPI (see fftw3.h). */
#ifndef PFFT_B200_FFTW3_MPI_COMPAT_H
#define PFFT_B200_FFTW3_MPI_COMPAT_H 1
#include <fftw3.h>
#include <mpi.h>

#define FFTW_MPI_DEFAULT_BLOCK (0)
#define FFTW_MPI_TRANSPOSED_IN (1U << 29)
#define FFTW_MPI_TRANSPOSED_OUT (1U << 30)

#endif
