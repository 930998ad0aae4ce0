/* oracle/stubs/mpi.h -- TEST INFRASTRUCTURE ONLY (never shipped, never linked
 * into the product library).
 *
 * Single-process stand-in for <mpi.h>, just wide enough to compile the
 * *integer-only* part of the reference (block decomposition, local_block_*,
 * local_size_gc*, init_input_*) from the sources under /root/reference into
 * oracle/_ref/libpfft_refint.so (recipe: oracle/Makefile).  A communicator is a
 * small record {ndims, dims, size, rank}; "rank" is whatever the test driver
 * says it is, so every pid of a virtual mesh can be queried from one process.
 * Communication entry points are only declared; calling one aborts at link
 * or run time, which is what we want: the oracle never communicates. */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
#include <stddef.h>

typedef struct oracle_stub_comm {
  int ndims;      /* 0 = not Cartesian */
  int dims[8];
  int size;
  int rank;
} *MPI_Comm;

typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Win;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef struct { int src, tag, err; } MPI_Status;

extern struct oracle_stub_comm oracle_stub_world, oracle_stub_self;
#define MPI_COMM_WORLD (&oracle_stub_world)
#define MPI_COMM_SELF  (&oracle_stub_self)
#define MPI_COMM_NULL  ((MPI_Comm)0)

#define MPI_SUCCESS 0
#define MPI_CART 1
#define MPI_GRAPH 2
#define MPI_UNDEFINED (-32766)
#define MPI_FLOAT 10
#define MPI_DOUBLE 11
#define MPI_LONG_DOUBLE 12
#define MPI_INT 13
#define MPI_MAX 100
#define MPI_MIN 101
#define MPI_SUM 102
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_INFO_NULL 0
#define MPI_ORDER_C 56
#define MPI_MODE_NOPRECEDE 1
#define MPI_MODE_NOSTORE 2
#define MPI_MODE_NOPUT 4
#define MPI_MODE_NOSUCCEED 8

/* implemented in stub_mpi.c */
int MPI_Comm_size(MPI_Comm c, int *size);
int MPI_Comm_rank(MPI_Comm c, int *rank);
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out);
int MPI_Comm_free(MPI_Comm *c);
int MPI_Topo_test(MPI_Comm c, int *status);
int MPI_Cart_create(MPI_Comm c, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *out);
int MPI_Cartdim_get(MPI_Comm c, int *ndims);
int MPI_Cart_get(MPI_Comm c, int maxdims, int *dims, int *periods, int *coords);
int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int *coords);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm c);
int MPI_Finalize(void);

/* declared only (the integer oracle must never reach them) */
int MPI_Cart_sub(MPI_Comm c, const int *remain, MPI_Comm *out);
int MPI_Cart_shift(MPI_Comm c, int dir, int disp, int *src, int *dst);
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *out);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Wait(MPI_Request *rq, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *st);
int MPI_Comm_group(MPI_Comm c, MPI_Group *g);
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *out);
int MPI_Group_free(MPI_Group *g);
int MPI_Win_create(void *base, MPI_Aint size, int disp, MPI_Info info, MPI_Comm c, MPI_Win *w);
int MPI_Win_free(MPI_Win *w);
int MPI_Win_fence(int assert_, MPI_Win w);
int MPI_Win_post(MPI_Group g, int assert_, MPI_Win w);
int MPI_Win_start(MPI_Group g, int assert_, MPI_Win w);
int MPI_Win_complete(MPI_Win w);
int MPI_Win_wait(MPI_Win w);
int MPI_Get(void *o, int on, MPI_Datatype ot, int rank, MPI_Aint disp, int tn, MPI_Datatype tt, MPI_Win w);
int MPI_Accumulate(const void *o, int on, MPI_Datatype ot, int rank, MPI_Aint disp, int tn, MPI_Datatype tt, MPI_Op op, MPI_Win w);
int MPI_Type_create_subarray(int nd, const int *sizes, const int *sub, const int *starts, int order, MPI_Datatype old, MPI_Datatype *nw);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);
#endif
