// gcell.cu -- ghost-cell plans: pfft_plan_*gc, pfft_exchange, pfft_reduce.
//
// Semantics (reference gcell/gcells_plan.c:80-199,379-486, gcell/gcells_sendrecv.c;
// net effect as in SURVEY.md 3.4):
//   exchange: the dense local block loc_n (x tuple) becomes, in place, a dense block of shape
//             ngc = gc_below + loc_n + gc_above holding the global array at indices
//             (loc_start - gc_below ... ) taken mod n -- halos may be wider than a
//             neighbour's block and corners are filled.
//   reduce:   the adjoint: every owned element becomes the sum of itself and of all its halo
//             copies on any rank; the array shrinks back to loc_n and the tail is zeroed.
// The reference moves slabs dimension by dimension with MPI_Isend/Irecv (or MPI_Get /
// MPI_Accumulate on a window) and packs/unpacks on the host.  Here every rank publishes a
// copy of its block in a peer-mapped (CUDA IPC) staging area and ONE kernel gathers each
// output element straight from the owning rank's memory over NVLink (exchange), resp. sums
// all copies of each owned element (reduce): no packing, no per-dimension rounds, no
// diagonal messages.  Source lookups are driven by small per-dimension tables built on
// the host, so the kernels contain no integer division by run-time block sizes.
//
// Cost model: the call is in place and the array changes its shape (row pitch and plane pitch grow by the
// halo widths), so every interior element moves as well -- one read + one write of the local block is the
// floor for the re-layout, on top of the halo bytes that cross NVLink.  The implementation publishes the
// block with one device-to-device copy (peers read their halos from it, this rank reads its interior from
// it) and writes the new shape with one row-wise kernel: 2 reads + 2 writes of the block.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "gcell.h"
#include "plan.h"

namespace pfb {

namespace {

#define GC_CUDA_OK(call)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "pfft_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      abort();                                                                                  \
    }                                                                                           \
  } while (0)

struct GcDev {
  // exchange: per dim t and position i in [0, ngc_t): owner coordinate and index inside its block
  const int *own_c[3], *own_l[3];
  // extents of the block of coordinate c along dim t (for strides inside the owner's staging copy)
  const int *ext[3];
  // reduce: CSR per dim t over local index j in [0, loc_n_t): copies (coordinate, position in its ngc block)
  const int *cp_off[3], *cp_c[3], *cp_i[3];
  const int *ngc_of[3];          // ngc extent of coordinate c along dim t
  void *const *peer;             // staging area of cart rank r
  int np[3];
  int loc_n[3], ngc[3];
  int tuple;
};

// Both kernels walk ROWS of the last dimension: the owner / copy lookups of dimensions 0 and 1 are done once per
// row, the threads of a CTA run along the row (coalesced accesses, whole grid points of `tuple` reals as one vector
// when tuple is 1 or 2), and no thread executes an integer division per element.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) gc_gather_kernel(GcDev d, T *__restrict__ out) {
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  const int rows = d.ngc[0] * d.ngc[1];
  const int per_point = d.tuple / VEC;               // vectors per grid point
  const int row_vecs = d.ngc[2] * per_point;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int i0 = row / d.ngc[1], i1 = row - i0 * d.ngc[1];
    const int c0 = d.own_c[0][i0], c1 = d.own_c[1][i1];
    const long long l01 = (long long)d.own_l[0][i0] * d.ext[1][c1] + d.own_l[1][i1];
    const int rank01 = (c0 * d.np[1] + c1) * d.np[2];
    V *dst = reinterpret_cast<V *>(out) + (long long)row * row_vecs;
    for (int x = threadIdx.x; x < row_vecs; x += blockDim.x) {
      const int i2 = per_point == 1 ? x : x / per_point, h = per_point == 1 ? 0 : x - i2 * per_point;
      const int c2 = d.own_c[2][i2];
      const V *src = reinterpret_cast<const V *>(d.peer[rank01 + c2]) + (l01 * d.ext[2][c2] + d.own_l[2][i2]) * per_point + h;
      dst[x] = *src;
    }
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) gc_reduce_kernel(GcDev d, T *__restrict__ out) {
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  const int rows = d.loc_n[0] * d.loc_n[1];
  const int per_point = d.tuple / VEC;
  const int row_vecs = d.loc_n[2] * per_point;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int j0 = row / d.loc_n[1], j1 = row - j0 * d.loc_n[1];
    V *dst = reinterpret_cast<V *>(out) + (long long)row * row_vecs;
    for (int x = threadIdx.x; x < row_vecs; x += blockDim.x) {
      const int j2 = per_point == 1 ? x : x / per_point, h = per_point == 1 ? 0 : x - j2 * per_point;
      V acc;
#pragma unroll
      for (int q = 0; q < VEC; q++) acc.v[q] = 0;
      for (int a = d.cp_off[0][j0]; a < d.cp_off[0][j0 + 1]; a++) {
        const int c0 = d.cp_c[0][a], i0 = d.cp_i[0][a];
        for (int b = d.cp_off[1][j1]; b < d.cp_off[1][j1 + 1]; b++) {
          const int c1 = d.cp_c[1][b], i1 = d.cp_i[1][b];
          const long long i01 = (long long)i0 * d.ngc_of[1][c1] + i1;
          const int rank01 = (c0 * d.np[1] + c1) * d.np[2];
          for (int c = d.cp_off[2][j2]; c < d.cp_off[2][j2 + 1]; c++) {
            const int c2 = d.cp_c[2][c];
            const V sv = *(reinterpret_cast<const V *>(d.peer[rank01 + c2]) + (i01 * d.ngc_of[2][c2] + d.cp_i[2][c]) * per_point + h);
#pragma unroll
            for (int q = 0; q < VEC; q++) acc.v[q] += sv.v[q];
          }
        }
      }
      dst[x] = acc;
    }
  }
}

INT pos_mod(INT a, INT n) {
  INT r = a % n;
  return r < 0 ? r + n : r;
}

}  // namespace

// reference gcell/gcells_plan.c:51-76
INT gc_local_size(int d, const INT *ln, const INT *ls, INT howmany, const INT *gb, const INT *ga, INT *ngc, INT *gcs) {
  INT mem = howmany;
  for (int t = 0; t < d; t++) {
    const INT b = gb ? gb[t] : 0, a = ga ? ga[t] : 0;
    ngc[t] = b + ln[t] + a;
    gcs[t] = ls[t] - b;
    mem *= ngc[t];
  }
  return mem;
}

struct GcPlan {
  int prec = PREC_F64;
  MPI_Comm comm = MPI_COMM_NULL;
  int np[3] = {1, 1, 1}, coords[3] = {0, 0, 0};
  int rnk_pm = 0;
  INT n[3], blk[3], gb[3], ga[3], loc_n[3], loc_start[3], ngc[3];
  INT tuple = 1;            // reals per grid point
  void *data = nullptr;
  void *stage = nullptr;    // my published copy
  size_t stage_bytes = 0;
  std::vector<void *> opened;
  void **peer_dev = nullptr;
  std::vector<int *> dev_tables;
  GcDev dev;
  void *host_tmp = nullptr;   // device copy for plain host `data`
  cudaStream_t stream = nullptr;
  GcTimer exg, red;
};

GcPlan *gc_plan_create(int prec, int rnk_n, const INT *n_user, INT howmany, const INT *block, const INT *gb_user,
                       const INT *ga_user, void *data, MPI_Comm comm, unsigned gc_flags, bool is_complex) {
  set_error("");
  INT gb[3] = {0, 0, 0}, ga[3] = {0, 0, 0};
  bool nothing = true;
  for (int t = 0; t < rnk_n && t < 3; t++) {
    gb[t] = gb_user ? gb_user[t] : 0;
    ga[t] = ga_user ? ga_user[t] : 0;
    if (gb[t] || ga[t]) nothing = false;
  }
  for (int t = 3; t < rnk_n; t++)
    if ((gb_user && gb_user[t]) || (ga_user && ga_user[t])) nothing = false;
  if (nothing) return nullptr;                       // reference gcells_plan.c:101-111
  if (rnk_n != 3) {                                  // reference gcells_plan.c:113-117
    int r = 0;
    MPI_Comm_rank(comm, &r);
    if (r == 0) fprintf(stderr, "Error: Gcell send only available for three dimensions.\n");
    return nullptr;
  }
  ensure_device();
  GcPlan *g = new GcPlan;
  g->prec = prec;
  g->data = data;
  g->stream = default_stream();
  g->comm = assure_cart(comm);
  MPI_Cartdim_get(g->comm, &g->rnk_pm);
  int dims[8] = {1, 1, 1}, per[8], co[8] = {0, 0, 0};
  MPI_Cart_get(g->comm, 8, dims, per, co);
  const int r = g->rnk_pm;
  for (int t = 0; t < r && t < 3; t++) {
    g->np[t] = dims[t];
    g->coords[t] = co[t];
  }
  // physical sizes (reference api/api-adv.c:232-324)
  INT pn[3] = {n_user[0], n_user[1], n_user[2]};
  unsigned flags = gc_flags;
  if (is_complex) {
    flags &= ~16u;                                   // PFFT_GC_PADDED ignored for complex arrays
    if (flags & 8u) pn[2] = n_user[2] / 2 + 1;       // PFFT_GC_R2C: physical size of the half spectrum
    g->tuple = 2 * howmany;
  } else {
    if (flags & 16u) pn[2] = 2 * (n_user[2] / 2 + 1);
    g->tuple = howmany;
  }
  // PFFT_GC_TRANSPOSED rotates the distributed dims (reference gcells_plan.c:222-234)
  for (int t = 0; t < 3; t++) {
    const int s = (t < r && (flags & 1u)) ? (t + 1) % r : t;
    g->n[t] = pn[s];
    g->gb[t] = gb[s];
    g->ga[t] = ga[s];
  }
  const bool remap3d = r == 3;
  for (int t = 0; t < 3; t++) {
    if (t < r) g->blk[t] = block_default(g->n[t], (block && !remap3d) ? block[t] : 0, g->np[t]);
    else g->blk[t] = g->n[t];
  }
  if (remap3d) {
    // 3-D mesh: the blocks of the 3dto2d remap's input layout (reference api/api-adv.c:275-286)
    Problem p;
    p.rnk_n = 3;
    p.rnk_pm = 3;
    for (int t = 0; t < 3; t++) p.np[t] = g->np[t];
    const Mesh3dto2d m = mesh_3dto2d(p);
    const INT o0 = block_default(g->n[0], 0, g->np[0] * m.q0), o1 = block_default(g->n[1], 0, g->np[1] * m.q1);
    g->blk[0] = o0 * m.q0;
    g->blk[1] = o1 * m.q1;
    g->blk[2] = block_default(g->n[2], 0, m.q0 * m.q1);
  }
  for (int t = 0; t < 3; t++) {
    g->loc_n[t] = block_extent(g->n[t], g->blk[t], g->coords[t]);
    g->loc_start[t] = block_offset(g->n[t], g->blk[t], g->coords[t]);
    g->ngc[t] = g->gb[t] + g->loc_n[t] + g->ga[t];
  }
  // ---- host tables
  const size_t rb = prec == PREC_F64 ? 8 : 4;
  std::vector<int> tab;
  auto push = [&](const std::vector<int> &v) {
    int *d = nullptr;
    GC_CUDA_OK(cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(int)));
    if (!v.empty()) GC_CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
    g->dev_tables.push_back(d);
    return d;
  };
  size_t max_loc = (size_t)g->tuple, max_ngc = (size_t)g->tuple;
  {
    size_t a = g->tuple, b = g->tuple;
    for (int t = 0; t < 3; t++) {
      // staging must hold the largest block of ANY rank? no: each rank stages its own data only
      a *= (size_t)g->loc_n[t];
      b *= (size_t)g->ngc[t];
    }
    max_loc = a;
    max_ngc = b;
  }
  for (int t = 0; t < 3; t++) {
    const INT n = g->n[t], blk = g->blk[t];
    std::vector<int> oc(g->ngc[t]), ol(g->ngc[t]), ext(g->np[t]), ngcof(g->np[t]);
    for (INT i = 0; i < g->ngc[t]; i++) {
      const INT gidx = pos_mod(g->loc_start[t] - g->gb[t] + i, n);
      const INT c = gidx / blk;
      oc[i] = (int)c;
      ol[i] = (int)(gidx - c * blk);
    }
    for (int c = 0; c < g->np[t]; c++) {
      ext[c] = (int)block_extent(n, blk, c);
      ngcof[c] = (int)(g->gb[t] + ext[c] + g->ga[t]);
    }
    // copies of my element j: positions i' on coordinate c' with (start_c' - gb + i') mod n == start_me + j
    std::vector<int> off(g->loc_n[t] + 1, 0), cc, ci;
    for (INT j = 0; j < g->loc_n[t]; j++) {
      const INT gidx = g->loc_start[t] + j;
      for (int c = 0; c < g->np[t]; c++) {
        const INT start_c = block_offset(n, blk, c);
        INT i0 = pos_mod(gidx - (start_c - g->gb[t]), n);
        for (INT i = i0; i < ngcof[c]; i += n) {
          cc.push_back(c);
          ci.push_back((int)i);
        }
      }
      off[j + 1] = (int)cc.size();
    }
    g->dev.own_c[t] = push(oc);
    g->dev.own_l[t] = push(ol);
    g->dev.ext[t] = push(ext);
    g->dev.ngc_of[t] = push(ngcof);
    g->dev.cp_off[t] = push(off);
    g->dev.cp_c[t] = push(cc);
    g->dev.cp_i[t] = push(ci);
    g->dev.np[t] = g->np[t];
    g->dev.loc_n[t] = (int)g->loc_n[t];
    g->dev.ngc[t] = (int)g->ngc[t];
  }
  g->dev.tuple = (int)g->tuple;
  // ---- staging area, published to every rank of the mesh
  g->stage_bytes = std::max<size_t>(std::max(max_loc, max_ngc) * rb, 256);
  GC_CUDA_OK(cudaMalloc(&g->stage, g->stage_bytes));
  int np_all = 1, me = 0;
  MPI_Comm_size(g->comm, &np_all);
  MPI_Comm_rank(g->comm, &me);
  cudaIpcMemHandle_t mine, *all = new cudaIpcMemHandle_t[np_all];
  memset(&mine, 0, sizeof mine);
  int ok = 1;
  if (np_all > 1 && cudaIpcGetMemHandle(&mine, g->stage) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  MPI_Allgather(&mine, (int)sizeof mine, MPI_BYTE, all, (int)sizeof mine, MPI_BYTE, g->comm);
  std::vector<void *> peers(np_all, nullptr);
  peers[me] = g->stage;
  for (int rk = 0; rk < np_all && ok; rk++) {
    if (rk == me) continue;
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[rk], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    peers[rk] = p;
    g->opened.push_back(p);
  }
  delete[] all;
  int all_ok = 0;
  MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, g->comm);
  if (!all_ok) {
    set_error("ghost cells: CUDA IPC mapping of the staging areas failed");
    gc_plan_destroy(g);
    return nullptr;
  }
  GC_CUDA_OK(cudaMalloc(&g->peer_dev, np_all * sizeof(void *)));
  GC_CUDA_OK(cudaMemcpy(g->peer_dev, peers.data(), np_all * sizeof(void *), cudaMemcpyHostToDevice));
  g->dev.peer = g->peer_dev;
  return g;
}

static bool gc_device_accessible(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

template <typename T>
static void gc_run(GcPlan *g, bool reduce) {
  GcTimer &tm = reduce ? g->red : g->exg;
  const double t0 = MPI_Wtime();
  const size_t rb = sizeof(T);
  size_t loc_elems = (size_t)g->tuple, ngc_elems = (size_t)g->tuple;
  for (int t = 0; t < 3; t++) {
    loc_elems *= (size_t)g->loc_n[t];
    ngc_elems *= (size_t)g->ngc[t];
  }
  cudaStream_t st = g->stream;
  T *data = static_cast<T *>(g->data);
  const bool host = !gc_device_accessible(g->data);
  if (host) {
    if (!g->host_tmp) GC_CUDA_OK(cudaMalloc(&g->host_tmp, std::max<size_t>(ngc_elems * rb, 256)));
    data = static_cast<T *>(g->host_tmp);
    GC_CUDA_OK(cudaMemcpyAsync(data, g->data, (reduce ? ngc_elems : loc_elems) * rb, cudaMemcpyHostToDevice, st));
  }
  // 1. publish my block (exchange: the owned block; reduce: the block with halos)
  const size_t pub = (reduce ? ngc_elems : loc_elems) * rb;
  if (pub) GC_CUDA_OK(cudaMemcpyAsync(g->stage, data, pub, cudaMemcpyDeviceToDevice, st));
  GC_CUDA_OK(cudaStreamSynchronize(st));
  const double t1 = MPI_Wtime();
  MPI_Barrier(g->comm);
  // 2. one kernel pulls everything it needs from the owners
  const size_t out_points = (reduce ? loc_elems : ngc_elems) / (size_t)g->tuple;
  if (out_points) {
    const long long rows = reduce ? (long long)g->loc_n[0] * g->loc_n[1] : (long long)g->ngc[0] * g->ngc[1];
    const int row_len = (int)((reduce ? g->loc_n[2] : g->ngc[2]) * g->tuple);
    const int threads = row_len >= 256 ? 256 : (row_len >= 128 ? 128 : 64);
    const long long blocks = std::min<long long>(rows, 148 * 16);
    // whole grid points as one vector when a point is 1 or 2 reals (real / complex arrays without tuples)
    const bool vec2 = g->tuple % 2 == 0 && (reinterpret_cast<uintptr_t>(data) % (2 * sizeof(T)) == 0);
    if (reduce) {
      if (vec2) gc_reduce_kernel<T, 2><<<(unsigned)blocks, threads, 0, st>>>(g->dev, data);
      else gc_reduce_kernel<T, 1><<<(unsigned)blocks, threads, 0, st>>>(g->dev, data);
    } else {
      if (vec2) gc_gather_kernel<T, 2><<<(unsigned)blocks, threads, 0, st>>>(g->dev, data);
      else gc_gather_kernel<T, 1><<<(unsigned)blocks, threads, 0, st>>>(g->dev, data);
    }
    launch_counter()++;
    GC_CUDA_OK(cudaGetLastError());
  }
  if (reduce && ngc_elems > loc_elems)   // the reference zeroes the tail of the shrunken array (gcells_plan.c:483-485)
    GC_CUDA_OK(cudaMemsetAsync(data + loc_elems, 0, (ngc_elems - loc_elems) * rb, st));
  if (host) GC_CUDA_OK(cudaMemcpyAsync(g->data, data, ngc_elems * rb, cudaMemcpyDeviceToHost, st));
  GC_CUDA_OK(cudaStreamSynchronize(st));
  MPI_Barrier(g->comm);   // nobody republishes while a peer may still be reading
  const double t2 = MPI_Wtime();
  tm.iter++;
  tm.whole += t2 - t0;
  tm.pad_zeros += t1 - t0;
  tm.exchange += t2 - t1;
}

void gc_exchange(GcPlan *g) {
  if (!g) return;   // reference api/api-basic.c:560
  if (g->prec == PREC_F64) gc_run<double>(g, false);
  else gc_run<float>(g, false);
}

void gc_reduce(GcPlan *g) {
  if (!g) return;
  if (g->prec == PREC_F64) gc_run<double>(g, true);
  else gc_run<float>(g, true);
}

void gc_plan_destroy(GcPlan *g) {
  if (!g) return;
  cudaStreamSynchronize(g->stream);
  for (void *p : g->opened) cudaIpcCloseMemHandle(p);
  for (int *t : g->dev_tables) cudaFree(t);
  if (g->peer_dev) cudaFree(g->peer_dev);
  if (g->stage) cudaFree(g->stage);
  if (g->host_tmp) cudaFree(g->host_tmp);
  if (g->comm != MPI_COMM_NULL) MPI_Comm_free(&g->comm);
  delete g;
}

void gc_reset_timers(GcPlan *g) {
  if (g) g->exg = g->red = GcTimer();
}
GcTimer *gc_get_timer(GcPlan *g, int which) { return g ? new GcTimer(which ? g->red : g->exg) : nullptr; }

void gc_print_timers(GcPlan *g, MPI_Comm comm, FILE *f, bool adv) {
  if (!g) return;
  int r = 0, np = 1;
  MPI_Comm_rank(comm, &r);
  MPI_Comm_size(comm, &np);
  const char *names[2] = {"pfft_gc_exg", "pfft_gc_red"};
  const GcTimer *src[2] = {&g->exg, &g->red};
  for (int k = 0; k < 2; k++) {
    GcTimer *m = gctimer_reduce_max(src[k], comm);
    gctimer_average(m);
    if (r == 0) {
      fprintf(f, "%s_iter(%d) = %d;  %s(%d) = %.3e;\n", names[k], np, src[k]->iter, names[k], np, m->whole);
      if (adv) fprintf(f, "%s_pad_zeros(%d) = %.3e;  %s_exchange(%d) = %.3e;\n", names[k], np, m->pad_zeros, names[k], np, m->exchange);
      fflush(f);
    }
    delete m;
  }
}

void gc_write_timers(GcPlan *g, const char *name, MPI_Comm comm, bool adv) {
  int r = 0;
  MPI_Comm_rank(comm, &r);
  FILE *f = r == 0 ? fopen(name, "w") : nullptr;
  gc_print_timers(g, comm, f ? f : stdout, adv);
  if (f) fclose(f);
}

GcTimer *gctimer_copy(const GcTimer *t) { return t ? new GcTimer(*t) : nullptr; }
void gctimer_average(GcTimer *t) {
  if (!t || t->iter <= 0) return;
  t->whole /= t->iter;
  t->pad_zeros /= t->iter;
  t->exchange /= t->iter;
  t->iter = 1;
}
GcTimer *gctimer_add(const GcTimer *a, const GcTimer *b) {
  GcTimer *r = new GcTimer(*a);
  r->iter += b->iter;
  r->whole += b->whole;
  r->pad_zeros += b->pad_zeros;
  r->exchange += b->exchange;
  return r;
}
GcTimer *gctimer_reduce_max(const GcTimer *t, MPI_Comm comm) {
  double v[4], m[4];
  gctimer_to_vec(t, v);
  MPI_Allreduce(v, m, 4, MPI_DOUBLE, MPI_MAX, comm);
  return gctimer_from_vec(m);
}
void gctimer_to_vec(const GcTimer *t, double *v) {
  v[0] = t->iter;
  v[1] = t->whole;
  v[2] = t->pad_zeros;
  v[3] = t->exchange;
}
GcTimer *gctimer_from_vec(const double *v) {
  GcTimer *t = new GcTimer;
  t->iter = (int)v[0];
  t->whole = v[1];
  t->pad_zeros = v[2];
  t->exchange = v[3];
  return t;
}

}  // namespace pfb
