/* oracle/stubs/stub_mpi.c -- TEST INFRASTRUCTURE ONLY.
 * The handful of MPI topology queries the reference's integer layer performs
 * (kernel/procmesh.c:36-64,114-129; kernel/partrafo.c:548-577), answered for a
 * virtual mesh inside one process: row-major rank <-> coords, last mesh
 * dimension fastest, as every MPI implementation does for MPI_Cart_create
 * without reordering. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct oracle_stub_comm oracle_stub_world = {0, {0}, 1, 0};
struct oracle_stub_comm oracle_stub_self = {0, {0}, 1, 0};

/* test-driver hook: pretend COMM_WORLD has `size` ranks and we are `rank` */
void oracle_stub_set_world(int size, int rank)
{
  oracle_stub_world.size = size;
  oracle_stub_world.rank = rank;
}

int MPI_Comm_size(MPI_Comm c, int *size) { *size = c->size; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { *rank = c->rank; return 0; }

int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out)
{
  *out = (MPI_Comm)malloc(sizeof(**out));
  memcpy(*out, c, sizeof(**out));
  return 0;
}

int MPI_Comm_free(MPI_Comm *c)
{
  if (*c != &oracle_stub_world && *c != &oracle_stub_self) free(*c);
  *c = MPI_COMM_NULL;
  return 0;
}

int MPI_Topo_test(MPI_Comm c, int *status)
{
  *status = c->ndims > 0 ? MPI_CART : MPI_UNDEFINED;
  return 0;
}

int MPI_Cart_create(MPI_Comm c, int ndims, const int *dims, const int *periods,
                    int reorder, MPI_Comm *out)
{
  (void)periods; (void)reorder;
  MPI_Comm n = (MPI_Comm)calloc(1, sizeof(*n));
  n->ndims = ndims;
  n->size = 1;
  for (int t = 0; t < ndims; t++) { n->dims[t] = dims[t]; n->size *= dims[t]; }
  n->rank = c->rank;
  *out = n;
  return 0;
}

int MPI_Cartdim_get(MPI_Comm c, int *ndims) { *ndims = c->ndims; return 0; }

int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int *coords)
{
  for (int t = c->ndims - 1; t >= 0; t--) {
    if (t < maxdims) coords[t] = rank % c->dims[t];
    rank /= c->dims[t];
  }
  return 0;
}

int MPI_Cart_get(MPI_Comm c, int maxdims, int *dims, int *periods, int *coords)
{
  for (int t = 0; t < maxdims && t < c->ndims; t++) { dims[t] = c->dims[t]; periods[t] = 1; }
  return MPI_Cart_coords(c, c->rank, maxdims, coords);
}

/* single process: a reduction over "all ranks" is the identity */
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  (void)op; (void)c;
  size_t w = t == MPI_FLOAT ? 4 : t == MPI_INT ? 4 : t == MPI_LONG_DOUBLE ? sizeof(long double) : 8;
  memcpy(r, s, w * (size_t)n);
  return 0;
}
double MPI_Wtime(void) { return 0.0; }
int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
int MPI_Finalize(void) { return 0; }
