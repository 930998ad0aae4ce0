/* CPU-only check of the minimpi subset + the communicator-facing integer API. */
#include <complex.h>
#include <pfft.h>
#include <stdio.h>

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  pfft_init();
  int rank, size;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  int np[2] = {2, size / 2};
  MPI_Comm cart;
  if (pfft_create_procmesh_2d(MPI_COMM_WORLD, np[0], np[1], &cart)) {
    pfft_fprintf(MPI_COMM_WORLD, stderr, "need an even number of ranks\n");
    MPI_Finalize();
    return 1;
  }
  ptrdiff_t n[3] = {29, 27, 31}, lni[3], lis[3], lno[3], los[3];
  ptrdiff_t alloc = pfft_local_size_dft_3d(n, cart, PFFT_TRANSPOSED_OUT, lni, lis, lno, los);
  double x = rank + 1.0, sum = 0, mx = 0;
  MPI_Allreduce(&x, &sum, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  MPI_Reduce(&x, &mx, 1, MPI_DOUBLE, MPI_MAX, 0, cart);
  long total_i = lni[0] * lni[1] * lni[2], total_o = lno[0] * lno[1] * lno[2], gi = 0, go = 0;
  MPI_Allreduce(&total_i, &gi, 1, MPI_LONG, MPI_SUM, cart);
  MPI_Allreduce(&total_o, &go, 1, MPI_LONG, MPI_SUM, cart);
  int coords[2], dims[2], per[2];
  MPI_Cart_get(cart, 2, dims, per, coords);
  MPI_Comm row;
  int remain[2] = {0, 1};
  MPI_Cart_sub(cart, remain, &row);
  int rr, rs;
  MPI_Comm_rank(row, &rr);
  MPI_Comm_size(row, &rs);
  int ok = (sum == size * (size + 1) / 2.0) && gi == 29 * 27 * 31 && go == 29 * 27 * 31 && rr == coords[1] && rs == dims[1];
  int allok = 0;
  MPI_Allreduce(&ok, &allok, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
  for (int it = 0; it < 2000; it++) MPI_Barrier(it % 2 ? row : cart);
  pfft_printf(MPI_COMM_WORLD, "ranks=%d sum=%g max=%g alloc=%td ok=%d\n", size, sum, mx, alloc, allok);
  printf("rank %d coords (%d,%d) ni=[%td,%td,%td] no=[%td,%td,%td] os=[%td,%td,%td]\n", rank, coords[0], coords[1],
         lni[0], lni[1], lni[2], lno[0], lno[1], lno[2], los[0], los[1], los[2]);
  MPI_Comm_free(&row);
  MPI_Comm_free(&cart);
  MPI_Finalize();
  return allok ? 0 : 1;
}
