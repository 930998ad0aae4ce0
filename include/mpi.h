/* include/mpi.h -- "minimpi": the single-node MPI subset that ships with pfft_b200.
 *
 * Why it exists: the reference's public header includes <mpi.h> (api/pfft.h:28)
 * and every reference test program starts with MPI_Init (tests/simple_check_c2c.c:22),
 * but neither an MPI implementation nor mpirun exists in the B200 image.  One
 * rank = one process = one GPU; ranks of a job find each other through a POSIX
 * shared-memory segment created by the launcher (pfft_b200/tools/pfftrun.c), by
 * the Python host layer (pfft_b200.bootstrap) or, under torchrun, derived from
 * RANK / WORLD_SIZE / MASTER_PORT.  Only small host scalars ever travel through
 * it (sizes, timers, max-error reductions, NCCL ids, CUDA IPC handles); bulk data
 * moves GPU-to-GPU over NVLink (NCCL or peer-mapped stores).
 *
 * Covered surface: exactly what the reference's library + C tests touch on the
 * FFT path minus point-to-point/RMA (the ghost-cell halo
 * moves device memory directly).  A site with a real MPI would build the library
 * against that MPI's <mpi.h> instead; nothing outside this subset is used.
 */
#ifndef PFFT_B200_MINIMPI_H
#define PFFT_B200_MINIMPI_H 1

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MINIMPI 1
#define MPI_VERSION 2
#define MPI_SUBVERSION 2

typedef struct minimpi_comm_s *MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef ptrdiff_t MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

extern struct minimpi_comm_s minimpi_comm_world_obj;
extern struct minimpi_comm_s minimpi_comm_self_obj;
#define MPI_COMM_WORLD (&minimpi_comm_world_obj)
#define MPI_COMM_SELF (&minimpi_comm_self_obj)
#define MPI_COMM_NULL ((MPI_Comm)0)

#define MPI_SUCCESS 0
#define MPI_ERR_COMM 5
#define MPI_ERR_ARG 12
#define MPI_ERR_OTHER 15
#define MPI_UNDEFINED (-32766)
#define MPI_PROC_NULL (-2)
#define MPI_IN_PLACE ((void *)1)
#define MPI_MAX_PROCESSOR_NAME 256

/* topology kinds (MPI_Topo_test) */
#define MPI_GRAPH 1
#define MPI_CART 2

/* datatypes */
#define MPI_DATATYPE_NULL 0
#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_UNSIGNED 4
#define MPI_LONG 5
#define MPI_UNSIGNED_LONG 6
#define MPI_LONG_LONG 7
#define MPI_LONG_LONG_INT MPI_LONG_LONG
#define MPI_FLOAT 8
#define MPI_DOUBLE 9
#define MPI_LONG_DOUBLE 10
#define MPI_AINT 11

/* reduction ops */
#define MPI_OP_NULL 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_PROD 4
#define MPI_LAND 5
#define MPI_LOR 6

int MPI_Init(int *argc, char ***argv);
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Initialized(int *flag);
int MPI_Finalized(int *flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);
double MPI_Wtick(void);
int MPI_Get_processor_name(char *name, int *resultlen);

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);

int MPI_Topo_test(MPI_Comm comm, int *status);
int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int *dims, const int *periods,
                    int reorder, MPI_Comm *comm_cart);
int MPI_Cartdim_get(MPI_Comm comm, int *ndims);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int *dims, int *periods, int *coords);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords);
int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest);
int MPI_Cart_sub(MPI_Comm comm, const int *remain_dims, MPI_Comm *newcomm);
int MPI_Dims_create(int nnodes, int ndims, int *dims);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                  int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
               int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype,
                  MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op,
               int root, MPI_Comm comm);

/* ---- minimpi extensions (not MPI) ------------------------------------------------ */
/* Join job `jobname` as rank `rank` of `size` before MPI_Init (the Python host layer
 * calls this after agreeing on a job name through torch.distributed).  Returns 0. */
int minimpi_bootstrap(const char *jobname, int rank, int size);
/* world rank of member `r` of `comm` (used to address peers' GPUs). */
int minimpi_world_rank(MPI_Comm comm, int r);
/* 1 when this process was started as part of a multi-rank job */
int minimpi_is_parallel(void);

#ifdef __cplusplus
}
#endif
#endif /* PFFT_B200_MINIMPI_H */
