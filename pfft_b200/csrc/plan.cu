// plan.cu -- plan life cycle and the execute loop (replaces the reference's
// plan_partrafo / execute_full, kernel/partrafo.c:317-525, api/api-basic.c:1044-1107).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>

#include "plan.h"

namespace pfb {

namespace {
cudaStream_t g_stream = nullptr;
int g_transport = -1;
bool g_device_ready = false;
}  // namespace

void set_error(const std::string &msg) { last_error_ref() = msg; }

cudaStream_t default_stream() { return g_stream; }
void set_default_stream(cudaStream_t s) { g_stream = s; }

int default_transport() {
  if (g_transport < 0) {
    g_transport = TR_AUTO;
    const char *e = getenv("PFFT_B200_TRANSPORT");
    if (e && !strcmp(e, "nccl")) g_transport = TR_NCCL;
    if (e && !strcmp(e, "p2p")) g_transport = TR_P2P;
  }
  return g_transport;
}
void set_default_transport(int t) { g_transport = t; }

// One rank = one GPU.  Respect the caller's current device unless told otherwise
// (PFFT_B200_DEVICE, or PFFT_B200_AUTODEVICE=1 set by pfftrun: rank % device count).
void ensure_device() {
  if (g_device_ready) return;
  g_device_ready = true;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "pfft_b200: no CUDA device available -- this library has no CPU fallback\n");
    abort();
  }
  const char *d = getenv("PFFT_B200_DEVICE");
  if (d) {
    cudaSetDevice(atoi(d) % ndev);
  } else if (getenv("PFFT_B200_AUTODEVICE")) {
    int rank = 0, inited = 0;
    MPI_Initialized(&inited);
    if (inited) MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    cudaSetDevice(rank % ndev);
  }
  cudaFree(0);
}

MPI_Comm assure_cart(MPI_Comm comm) {
  MPI_Comm cart = MPI_COMM_NULL;
  int status = MPI_UNDEFINED;
  MPI_Topo_test(comm, &status);
  if (status == MPI_CART) {
    MPI_Comm_dup(comm, &cart);
  } else {
    int np = 1, per = 1;
    MPI_Comm_size(comm, &np);
    MPI_Cart_create(comm, 1, &np, &per, 1, &cart);
  }
  return cart;
}

#define CUDA_OK(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "pfft_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      abort();                                                                                 \
    }                                                                                          \
  } while (0)

static bool all_ranks_ok(MPI_Comm comm, bool ok) {
  int mine = ok ? 1 : 0, all = 0;
  MPI_Allreduce(&mine, &all, 1, MPI_INT, MPI_MIN, comm);
  return all != 0;
}

// Planning-time device allocations: a failure (out of memory) makes the planner return NULL with
// pfftb200_last_error set, like the reference's planners (kernel/partrafo.c:337-373), instead of aborting.
static bool try_malloc(void **p, size_t bytes, std::string *err) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaSuccess) return true;
  cudaGetLastError();
  *p = nullptr;
  if (err && err->empty()) *err = "cudaMalloc(" + std::to_string(bytes) + " bytes) failed: " + cudaGetErrorString(e);
  return false;
}

static void *upload_device(const void *host, size_t bytes, void *ctx) {
  PlanBase *pl = static_cast<PlanBase *>(ctx);
  void *dev = nullptr;
  CUDA_OK(cudaMalloc(&dev, bytes));      // (twiddle tables: kilobytes)
  CUDA_OK(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice));
  pl->tables.push_back(dev);
  return dev;
}

// table exp(-2 pi i m / (8 D)) of the DCT/DST pre- and post-twiddles (kernels.h)
template <typename T>
void *make_r2r_table(int D, UploadFn upload, void *ctx) {
  const int M = 8 * D;
  std::vector<double> roots(2 * (size_t)M);
  make_twiddles_f64(M, roots.data());
  std::vector<T> host(roots.begin(), roots.end());
  return upload(host.data(), host.size() * sizeof(T), ctx);
}
template void *make_r2r_table<float>(int, UploadFn, void *);
template void *make_r2r_table<double>(int, UploadFn, void *);

// Kernel family of every stage and its parameters:
//   pow2   power-of-two complex lines, 64..4096, held in registers (fft_pow2.cu; the headline path)
//   reg    real lines of length 2 * {2^k, 3 * 2^k} as packed half-length transforms and complex lines of
//          length 3 * 2^k, held in registers (fft_reg.cu)
//   mixed  everything else: any length, odd real lines, DCT/DST, copy stages (fft_mixed.cu)
//   generic  the round-1 any-length kernel, kept behind PFFT_B200_GENERIC=1 for A/B comparisons
template <typename T>
static bool build_stage_params(PlanBase *pl, std::string *err) {
  struct TableSet {
    void *dev;
    size_t off2, off3;
  };
  std::map<int, TableSet> table_of;
  static const bool old_generic = [] {
    const char *e = getenv("PFFT_B200_GENERIC");
    return e && atoi(e) != 0;
  }();
  const Schedule &s = pl->sched;
  for (size_t i = 0; i < s.stages.size(); i++) {
    const Stage &g = s.stages[i];
    StageParams sp;
    if (!stage_params_basic(g, sp, err)) return false;
    const int L = sp.L;
    int kind = stage_kernel_family<T>(g, L);       // KERNEL_POW2 / KERNEL_REG / KERNEL_MIXED
    const bool fast = kind == KERNEL_POW2;
    if (kind == KERNEL_MIXED && old_generic) kind = KERNEL_GENERIC;
    if (g.op == OP_R2R) {
      auto jt = table_of.find(-8 * sp.r2r_D);
      if (jt == table_of.end()) jt = table_of.emplace(-8 * sp.r2r_D, TableSet{make_r2r_table<T>(sp.r2r_D, upload_device, pl), 0, 0}).first;
      sp.tw_r2r = jt->second.dev;
    }
    if (kind == KERNEL_REG) {
      if (!reg_prepare<T>(g, sp, upload_device, pl, err)) return false;
      stage_params_tiles(g, sp, sp.tl);
    } else if (kind == KERNEL_MIXED) {
      if (!mixed_prepare<T>(g, sp, upload_device, pl, err)) return false;
      stage_params_tiles(g, sp, sp.tl);
    } else {
      const int tl = fast ? pow2_pick_tile<T>(g, L) : generic_pick_tile<T>(g, L);
      if (tl <= 0) {
        *err = "transform length " + std::to_string(L) + " does not fit the shared-memory kernels";
        return false;
      }
      stage_params_tiles(g, sp, tl);
      sp.nfac = factorize_generic(L, sp.fac);
      if (g.op != OP_COPY) {
        auto it = table_of.find(2 * L + (fast ? 1 : 0));
        if (it == table_of.end()) {
          // unit roots in fp64 (< 1 ulp), then the layouts the kernels want, rounded once to T
          std::vector<double> roots(2 * (size_t)L);
          make_twiddles_f64(L, roots.data());
          std::vector<double> all(roots);
          size_t off2 = 0, off3 = 0;
          if (fast) {
            std::vector<double> pp;
            pow2_twiddle_tables(L, roots.data(), &pp, &off2, &off3);
            off2 += L;
            off3 += L;
            all.insert(all.end(), pp.begin(), pp.end());
          }
          std::vector<T> host(all.begin(), all.end());
          void *dev = upload_device(host.data(), host.size() * sizeof(T), pl);
          it = table_of.emplace(2 * L + (fast ? 1 : 0), TableSet{dev, off2, off3}).first;
        }
        const TableSet &ts = it->second;
        sp.twiddle = ts.dev;
        if (fast) {
          sp.tw2 = static_cast<const T *>(ts.dev) + 2 * ts.off2;
          sp.tw3 = static_cast<const T *>(ts.dev) + 2 * ts.off3;
          pow2_prepare<T>(g, sp);
        }
      }
    }
    if (g.ntile > 0 && !(fast && sp.fast)) {
      *err = "internal: micro-blocked stage is not eligible for the register-resident kernel";
      return false;
    }
    pl->params.push_back(sp);
    pl->use_pow2.push_back(fast ? 1 : 0);
    pl->kernel_kind.push_back(kind);
  }
  return true;
}

// ---- plane-fused last pair -------------------------------------------------------------
namespace {
struct DimSE {
  long long stride, extent;
};
// canonical form of an address set given as a sum of strided dimensions
std::vector<DimSE> canonical(std::vector<DimSE> v) {
  std::vector<DimSE> w;
  for (auto &d : v)
    if (d.extent > 1) w.push_back(d);
  std::sort(w.begin(), w.end(), [](const DimSE &x, const DimSE &y) { return x.stride < y.stride; });
  std::vector<DimSE> out;
  for (auto &d : w) {
    if (!out.empty() && out.back().stride * out.back().extent == d.stride) out.back().extent *= d.extent;
    else out.push_back(d);
  }
  return out;
}
bool same_set(const std::vector<DimSE> &a, const std::vector<DimSE> &b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++)
    if (a[i].stride != b[i].stride || a[i].extent != b[i].extent) return false;
  return true;
}
// Make the plane (stride `ps` on the given side, `planes` of them) batch dimension 0 of sp,
// splitting it out of a merged batch dimension if necessary.
bool plane_first(StageParams &sp, bool out_side, long long ps, long long planes) {
  int nb = sp.nbatch;
  long long ext[kMaxBatch + 1], is[kMaxBatch + 1], os[kMaxBatch + 1];
  int tile = sp.tile_dim;
  for (int k = 0; k < nb; k++) { ext[k] = sp.bext[k]; is[k] = sp.bis[k]; os[k] = sp.bos[k]; }
  int kp = -1;
  for (int k = 0; k < nb; k++) {
    const long long st = out_side ? os[k] : is[k];
    if (st == ps && ext[k] == planes) { kp = k; break; }
  }
  if (kp < 0) {
    for (int k = 0; k < nb && kp < 0; k++) {
      const long long st = out_side ? os[k] : is[k];
      if (st <= 0 || st >= ps || ps % st) continue;
      const long long q = ps / st;
      if (ext[k] % q || ext[k] / q != planes || nb + 1 > kMaxBatch) continue;
      // split dimension k into (planes, q): insert the inner part behind it
      for (int j = nb; j > k + 1; j--) { ext[j] = ext[j - 1]; is[j] = is[j - 1]; os[j] = os[j - 1]; }
      ext[k + 1] = q; is[k + 1] = is[k]; os[k + 1] = os[k];
      ext[k] = planes; is[k] *= q; os[k] *= q;
      if (tile == k) tile = k + 1;
      else if (tile > k) tile++;
      nb++;
      kp = k;
    }
  }
  if (kp < 0 || kp == tile) return false;
  // rotate kp to the front
  const long long e0 = ext[kp], i0 = is[kp], o0 = os[kp];
  for (int j = kp; j > 0; j--) { ext[j] = ext[j - 1]; is[j] = is[j - 1]; os[j] = os[j - 1]; }
  ext[0] = e0; is[0] = i0; os[0] = o0;
  if (tile >= 0 && tile < kp) tile++;
  sp.nbatch = nb;
  sp.tile_dim = tile;
  for (int k = 0; k < nb; k++) { sp.bext[k] = ext[k]; sp.bis[k] = is[k]; sp.bos[k] = os[k]; }
  return true;
}
}  // namespace

template <typename T>
static void setup_fused_pair(PlanBase *pl) {
  FusedPair &f = pl->fuse;
  f.possible = false;
  // Opt-in (PFFT_B200_FUSE=1): measured on B200 at 1024^3 fp64 the pair costs 14.2-14.6 ms against
  // 7.2 + 5.5 ms for the two stages run separately -- the first stage's cross-line exchange with
  // two-line tiles and the ticket/flag traffic cost more than the halved HBM traffic returns.
  const char *env = getenv("PFFT_B200_FUSE");
  if (!env || atoi(env) == 0) return;
  const Schedule &s = pl->sched;
  const int nst = (int)s.stages.size();
  if (nst < 2) return;
  const int i = nst - 2;
  const Stage &ga = s.stages[i], &gb = s.stages[i + 1];
  if (!pl->use_pow2[i] || !pl->use_pow2[i + 1]) return;
  StageParams a = pl->params[i], b = pl->params[i + 1];
  if (!a.fast || !b.fast || a.L != b.L || a.sign != b.sign) return;
  if (a.ntile > 0 || b.ntile > 0) return;   // micro-blocked chains have their own kernel
  if (sizeof(T) != 8) return;   // 8-byte cp.async is L1-cached (.ca): ring data written by other SMs could be stale
  if (ga.exchange >= 0 && s.exchanges[ga.exchange].nparts > 1) return;
  if (a.noseg != 1 || b.noseg != 1 || a.istride != 1 || b.istride != 1) return;
  if (a.ntiles <= 0 || b.ntiles <= 0 || a.nbatch < 1 || b.nbatch < 1) return;
  // plane = outermost dimension of the intermediate array
  int ka = 0;
  for (int k = 1; k < a.nbatch; k++)
    if (a.bos[k] > a.bos[ka]) ka = k;
  const long long ps = a.bos[ka];
  long long planes = a.bext[ka];
  // everything else A writes / B reads must stay inside one plane slot
  {
    long long span = (long long)(a.L - 1) * a.ostride + 1;
    for (int k = 0; k < a.nbatch; k++)
      if (k != ka) span += (a.bext[k] - 1) * a.bos[k];
    if (span > ps) return;
  }
  if (!plane_first(a, true, ps, planes)) return;
  if (!plane_first(b, false, ps, planes)) return;
  {
    long long span = (long long)(b.L - 1) * b.istride + 1;
    for (int k = 1; k < b.nbatch; k++) span += (b.bext[k] - 1) * b.bis[k];
    if (span > ps) return;
  }
  const char *er = getenv("PFFT_B200_RING"), *ed = getenv("PFFT_B200_DELTA");
  const int delta = ed ? std::max(1, atoi(ed)) : 2;
  const int ring = er ? std::max(delta + 1, atoi(er)) : delta + 2;
  const size_t plane_bytes = (size_t)ps * 2 * sizeof(T);
  if (planes < 2 * ring || plane_bytes * ring > ((size_t)96 << 20)) return;
  // tiles: short runs are fine on the ring side (it lives in L2), so small independent CTAs
  const int tl = fused_pick_tile<T>(a.L);
  auto retile = [&](StageParams &sp) {
    sp.tl = tl;
    if (sp.tile_dim < 0) return false;
    if (sp.bext[sp.tile_dim] < tl) return false;
    sp.tiles_along = (sp.bext[sp.tile_dim] + tl - 1) / tl;
    long long others = 1;
    for (int k = 0; k < sp.nbatch; k++)
      if (k != sp.tile_dim) others *= sp.bext[k];
    sp.ntiles = others * sp.tiles_along;
    return sp.ntiles < (1ll << 30);
  };
  if (!retile(a) || !retile(b)) return;
  a.ring_out = ring;
  a.ring_in = 0;
  b.ring_in = ring;
  b.ring_out = 0;
  f.fp.planes = (int)planes;
  f.fp.t1 = (int)(a.ntiles / planes);
  f.fp.t2 = (int)(b.ntiles / planes);
  f.fp.delta = delta;
  f.fp.ring = ring;
  f.fp.done = nullptr;
  // same addresses for plane p on the pair's input and output side?
  {
    std::vector<DimSE> in_set, out_set;
    in_set.push_back({a.istride, a.L});
    for (int k = 1; k < a.nbatch; k++) in_set.push_back({a.bis[k], a.bext[k]});
    out_set.push_back({b.ostride, b.L});
    for (int k = 1; k < b.nbatch; k++) out_set.push_back({b.bos[k], b.bext[k]});
    f.inplace_ok = a.bis[0] == b.bos[0] && same_set(canonical(in_set), canonical(out_set));
  }
  f.a = a;
  f.b = b;
  f.first = i;
  f.ring_bytes = plane_bytes * ring;
  f.possible = true;
  (void)gb;
}

PlanBase *plan_create(int prec, const Problem &p, void *in, void *out, MPI_Comm comm) {
  ensure_device();
  PlanBase *pl = new PlanBase;
  pl->prec = prec;
  pl->prob = p;
  pl->planned_in = in;
  pl->planned_out = out;
  pl->stream = default_stream();
  pl->comm_cart = assure_cart(comm);
  int pid = 0;
  MPI_Comm_rank(pl->comm_cart, &pid);
  std::string err;
  bool ok = build_schedule(p, pid, &pl->sched);
  if (!ok) err = pl->sched.error;
  if (ok) ok = prec == PREC_F64 ? build_stage_params<double>(pl, &err) : build_stage_params<float>(pl, &err);
  if (ok) { if (prec == PREC_F64) setup_fused_pair<double>(pl); else setup_fused_pair<float>(pl); }
  if (!all_ranks_ok(pl->comm_cart, ok)) {
    set_error(err.empty() ? "planning failed on another rank" : err);
    plan_destroy(pl);
    return nullptr;
  }
  const Schedule &s = pl->sched;
  // one communicator per exchange group (rows / columns of the mesh, remap sub-groups):
  // colour = lowest member rank, key = position inside the group
  for (int g = 0; g < s.ngroups; g++) {
    int color = s.groups[g].members[0];
    for (int q = 1; q < s.groups[g].size; q++) color = std::min(color, s.groups[g].members[q]);
    MPI_Comm_split(pl->comm_cart, color, s.groups[g].me, &pl->comm_1d[g]);
  }
  const size_t rb = pl->elem_real_bytes();
  auto bytes_of = [&](const INT *ln, bool real) {
    size_t m = (size_t)p.howmany * rb * (real ? 1 : 2);
    for (int t = 0; t < p.rnk_n; t++) m *= (size_t)ln[t];
    return m;
  };
  pl->user_in_bytes = bytes_of(s.ls.lni, s.stages.front().in_real);
  pl->user_out_bytes = bytes_of(s.ls.lno, s.stages.back().out_real);
  pl->scratch_bytes = (size_t)s.scratch_elems * rb;
  if (pl->scratch_bytes == 0) pl->scratch_bytes = 256;
  bool any_exchange = false;
  for (auto &x : s.exchanges)
    if (x.nparts > 1) any_exchange = true;
  int tr = default_transport();
  if (tr == TR_AUTO) tr = TR_P2P;
  pl->transport = tr;
  const size_t nst = s.stages.size();
  // footprint (bytes) of what sits between stage i and stage i+1
  pl->boundary_bytes.assign(nst > 0 ? nst - 1 : 0, 0);
  pl->boundary_remote.assign(nst > 0 ? nst - 1 : 0, 0);
  for (size_t i = 0; i + 1 < nst; i++) {
    const Stage &g = s.stages[i];
    const size_t es = rb * (g.out_real ? 1 : 2);
    INT elems = 0;
    for (int q = 0; q < g.noseg; q++) elems += g.oseg_cnt[q];
    if (g.exchange >= 0) {
      const Exchange &x = s.exchanges[g.exchange];
      elems = std::max<INT>(elems, x.recv_cnt * x.nparts);
      pl->boundary_remote[i] = x.nparts > 1;
    }
    pl->boundary_bytes[i] = (size_t)elems * es;
  }
  // plans with real exchanges own their receive areas from the start (peers map them);
  // purely local plans borrow user buffers where they can and allocate scratch lazily
  if (any_exchange) {
    std::string aerr;
    bool aok = true;
    if (nst >= 2) aok = try_malloc(&pl->scratch[0], pl->scratch_bytes, &aerr) && aok;
    if (nst >= 3) aok = try_malloc(&pl->scratch[1], pl->scratch_bytes, &aerr) && aok;
    if (tr == TR_NCCL) aok = try_malloc(&pl->scratch[2], pl->scratch_bytes, &aerr) && aok;
    pl->scratch_cap[0] = pl->scratch[0] ? pl->scratch_bytes : 0;
    pl->scratch_cap[1] = pl->scratch[1] ? pl->scratch_bytes : 0;
    if (!all_ranks_ok(pl->comm_cart, aok)) {
      set_error(aerr.empty() ? "receive areas could not be allocated on another rank" : aerr);
      plan_destroy(pl);
      return nullptr;
    }
  }
  if (any_exchange) {
    std::string terr;
    bool tok = transport_setup(pl, &terr);
    if (!all_ranks_ok(pl->comm_cart, tok)) {
      set_error(terr.empty() ? "transport setup failed on another rank" : terr);
      plan_destroy(pl);
      return nullptr;
    }
  }
  if (pl->fuse.possible) {
    CUDA_OK(cudaMalloc(&pl->fuse.ring, pl->fuse.ring_bytes));
    void *cnt = nullptr;
    CUDA_OK(cudaMalloc(&cnt, sizeof(unsigned) * (2 * (size_t)pl->fuse.fp.planes + 1)));
    pl->fuse.fp.done = static_cast<unsigned *>(cnt);
  }
  pl->events.resize(2 * nst);
  for (auto &e : pl->events) CUDA_OK(cudaEventCreate(&e));
  pl->timer.shape(p.rnk_pm);
  pl->last_stage_ms.assign(nst, 0.0);
  pl->last_xch_ms.assign(s.exchanges.size(), 0.0);
  return pl;
}

void plan_destroy(PlanBase *pl) {
  if (!pl) return;
  cudaStreamSynchronize(pl->stream);   // (the null stream too: kernels of this plan may still run)
  transport_teardown(pl);
  for (auto &e : pl->events) cudaEventDestroy(e);
  if (pl->fuse.ring) cudaFree(pl->fuse.ring);
  if (pl->fuse.fp.done) cudaFree(pl->fuse.fp.done);
  for (void *t : pl->tables) cudaFree(t);
  if (pl->mixed_ws) cudaFree(pl->mixed_ws);
  for (int k = 0; k < 3; k++)
    if (pl->scratch[k]) cudaFree(pl->scratch[k]);
  for (int t = 0; t < kMaxGroups; t++)
    if (pl->comm_1d[t] != MPI_COMM_NULL) MPI_Comm_free(&pl->comm_1d[t]);
  if (pl->comm_cart != MPI_COMM_NULL) MPI_Comm_free(&pl->comm_cart);
  delete pl;
}

// Is `p` usable by a kernel as is (device or managed memory)?
static bool device_accessible(const void *p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---- host-pointer executes ---------------------------------------------------------------------
// Arrays the device cannot address (plain or pinned host memory) are staged through two process-wide
// device areas.  The copies run on two copy streams of their own, in chunks, so that consecutive
// executes keep BOTH PCIe directions busy: the device-to-host copy of one transform's output overlaps the
// host-to-device copy of the next transform's input (a 3-D FFT needs all of its input before any output
// is final, so inside ONE transform the two copies cannot overlap).  Ordering:
//   * read-after-write through host memory: every D2H chunk records an event; an H2D chunk whose host
//     range overlaps a pending D2H waits for the covering chunk's event -- backward(forward(x)) streams;
//   * the input staging area is reused: H2D waits for the previous staged transform's last kernel;
//   * the output staging area is reused: a staged transform's kernels wait for the previous D2H.
// pfft_execute then synchronises (the reference call blocks); pfftb200_execute_async does not.
namespace {
constexpr size_t kCopyChunk = (size_t)128 << 20;
struct HostPending {
  const char *base;
  size_t bytes;
  std::vector<cudaEvent_t> done;   // one per kCopyChunk
};
std::vector<HostPending> g_pending;
std::vector<cudaEvent_t> g_event_pool;
cudaStream_t g_h2d = nullptr, g_d2h = nullptr;
cudaEvent_t g_compute_done = nullptr, g_d2h_done = nullptr;
bool g_have_compute_done = false, g_have_d2h_done = false;

cudaEvent_t take_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return e;
}
void copy_streams() {
  if (g_h2d) return;
  CUDA_OK(cudaStreamCreateWithFlags(&g_h2d, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&g_d2h, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&g_compute_done, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&g_d2h_done, cudaEventDisableTiming));
}
// forget device-to-host copies that have landed
void purge_pending() {
  for (size_t i = 0; i < g_pending.size();) {
    if (cudaEventQuery(g_pending[i].done.back()) == cudaSuccess) {
      for (cudaEvent_t e : g_pending[i].done) g_event_pool.push_back(e);
      g_pending.erase(g_pending.begin() + i);
    } else {
      cudaGetLastError();
      i++;
    }
  }
}
}  // namespace

// Process-wide staging areas for host-pointer executes (0: input, 1: output); grow only.
static void *staging_area(int which, size_t bytes, cudaStream_t st) {
  static void *area[2] = {nullptr, nullptr};
  static size_t cap[2] = {0, 0};
  if (cap[which] < bytes) {
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaDeviceSynchronize());
    if (area[which]) cudaFree(area[which]);
    CUDA_OK(cudaMalloc(&area[which], bytes));
    cap[which] = bytes;
  }
  return area[which];
}

static void staged_h2d(void *dev, const void *host, size_t bytes, cudaStream_t st) {
  copy_streams();
  purge_pending();
  if (g_have_compute_done) CUDA_OK(cudaStreamWaitEvent(g_h2d, g_compute_done, 0));
  const char *h = static_cast<const char *>(host);
  for (size_t off = 0; off < bytes; off += kCopyChunk) {
    const size_t len = std::min(kCopyChunk, bytes - off);
    for (const HostPending &p : g_pending) {
      if (h + off + len <= p.base || p.base + p.bytes <= h + off) continue;
      // the D2H stream runs its chunks in order: waiting for the last overlapping one covers them all
      const size_t last_byte = std::min<size_t>((size_t)(h + off + len - p.base), p.bytes) - 1;
      CUDA_OK(cudaStreamWaitEvent(g_h2d, p.done[last_byte / kCopyChunk], 0));
    }
    CUDA_OK(cudaMemcpyAsync(static_cast<char *>(dev) + off, h + off, len, cudaMemcpyHostToDevice, g_h2d));
  }
  cudaEvent_t landed = take_event();
  CUDA_OK(cudaEventRecord(landed, g_h2d));
  CUDA_OK(cudaStreamWaitEvent(st, landed, 0));
  g_event_pool.push_back(landed);      // (a recorded event may be re-recorded once its waits are enqueued)
}

static void staged_d2h(void *host, const void *dev, size_t bytes) {
  CUDA_OK(cudaStreamWaitEvent(g_d2h, g_compute_done, 0));   // (recorded by the caller behind the last kernel)
  HostPending p;
  p.base = static_cast<const char *>(host);
  p.bytes = bytes;
  for (size_t off = 0; off < bytes; off += kCopyChunk) {
    const size_t len = std::min(kCopyChunk, bytes - off);
    CUDA_OK(cudaMemcpyAsync(static_cast<char *>(host) + off, static_cast<const char *>(dev) + off, len, cudaMemcpyDeviceToHost, g_d2h));
    cudaEvent_t e = take_event();
    CUDA_OK(cudaEventRecord(e, g_d2h));
    p.done.push_back(e);
  }
  if (p.done.empty()) return;
  CUDA_OK(cudaEventRecord(g_d2h_done, g_d2h));
  g_have_d2h_done = true;
  g_pending.push_back(p);
}

// Choose, for every boundary between two stages, the buffer that holds it.  A stage never
// writes the buffer it reads; the last intermediate is never the output array; boundaries
// that an exchange fills are plan-owned scratch (peers map them); user arrays are used only
// when they are large enough, and the input array only when it may be destroyed
// (PFFT_DESTROY_INPUT, in-place, or staged host input) -- the default PFFT_PRESERVE_INPUT
// semantics of the reference (kernel/partrafo.c:963-975) therefore holds by construction.
// Among the feasible assignments the one with the least scratch memory wins.
static void assign_buffers(PlanBase *pl, bool inplace, bool destroyable) {
  const int key = (inplace ? 1 : 0) | (destroyable ? 2 : 0);
  if (pl->assign_key == key && !pl->assign.empty()) return;
  const size_t nb_all = pl->boundary_bytes.size();
  bool any_remote = false;
  for (size_t i = 0; i < nb_all; i++) any_remote |= pl->boundary_remote[i] != 0;
  // the fused pair hands its intermediate through the ring; its input array must not be the
  // array it writes unless both use the same addresses plane by plane
  FusedPair &fu = pl->fuse;
  fu.active = fu.possible && nb_all >= 1;
  if (fu.active && nb_all == 1 && inplace && !fu.inplace_ok) fu.active = false;
  const size_t nb = fu.active ? nb_all - 1 : nb_all;
  if (any_remote) {
    // peers store into these areas: every rank must pick the same ones, whatever its block sizes
    pl->assign.resize(nb_all);
    for (size_t i = 0; i < nb_all; i++) pl->assign[i] = (i % 2) ? BUF_B : BUF_A;
    if (fu.active) pl->assign[nb_all - 1] = BUF_RING;
    pl->assign_key = key;
    return;
  }
  std::vector<int> best, cur(nb, 0);
  int best_cost = 1 << 30;
  const int cand[4] = {BUF_USER_IN, BUF_USER_OUT, BUF_A, BUF_B};
  std::function<void(size_t, int)> rec = [&](size_t i, int used_mask) {
    const int cost = ((used_mask >> BUF_A) & 1) + ((used_mask >> BUF_B) & 1);
    if (cost >= best_cost) return;
    if (i == nb) {
      best = cur;
      best_cost = cost;
      return;
    }
    const int prev = i == 0 ? BUF_USER_IN : cur[i - 1];
    for (int c : cand) {
      int id = c;
      if (inplace && id == BUF_USER_OUT) continue;                  // same array as BUF_USER_IN
      if (id == prev) continue;
      if (inplace && prev == BUF_USER_IN && id == BUF_USER_IN) continue;
      if (i + 1 == nb && (id == BUF_USER_OUT || (inplace && id == BUF_USER_IN)) && !(fu.active && fu.inplace_ok)) continue;
      if (pl->boundary_remote[i] && id < BUF_A) continue;
      if (id == BUF_USER_IN && (!destroyable || pl->boundary_bytes[i] > pl->user_in_bytes)) continue;
      if (id == BUF_USER_OUT && pl->boundary_bytes[i] > pl->user_out_bytes) continue;
      if (id == BUF_B && !((used_mask >> BUF_A) & 1)) continue;     // symmetry: take A before B
      cur[i] = id;
      rec(i + 1, used_mask | (1 << id));
    }
  };
  rec(0, 0);
  if (best.size() != nb) {   // cannot happen: A/B alternation is always feasible
    best.assign(nb, BUF_A);
    for (size_t i = 0; i < nb; i++) best[i] = (i % 2) ? BUF_B : BUF_A;
  }
  if (fu.active) best.push_back(BUF_RING);
  pl->assign = best;
  pl->assign_key = key;
  // scratch on demand (plans with exchanges allocated theirs up front)
  for (int k = 0; k < 2; k++) {
    size_t need = 0;
    for (size_t i = 0; i < nb; i++)
      if (best[i] == BUF_A + k) need = std::max(need, pl->boundary_bytes[i]);   // (the ring is plan-owned)
    if (need > pl->scratch_cap[k]) {
      CUDA_OK(cudaStreamSynchronize(pl->stream));
      if (pl->scratch[k]) cudaFree(pl->scratch[k]);
      CUDA_OK(cudaMalloc(&pl->scratch[k], need));
      pl->scratch_cap[k] = need;
    }
  }
}

void plan_execute(PlanBase *pl, void *in, void *out, bool blocking) {
  const Schedule &s = pl->sched;
  cudaStream_t st = pl->stream;
  if (!in) in = pl->planned_in;
  if (!out) out = pl->planned_out;
  // host pointers are staged through device memory (counted in the end-to-end numbers);
  // the staging areas are process-wide and shared by all plans
  void *dev_in = in, *dev_out = out;
  bool copy_back = false, staged_in = false;
  if (pl->user_in_bytes && !device_accessible(in)) {
    dev_in = staging_area(0, pl->user_in_bytes, st);
    staged_h2d(dev_in, in, pl->user_in_bytes, st);
    staged_in = true;
  }
  if (pl->user_out_bytes && !device_accessible(out)) {
    dev_out = staging_area(1, pl->user_out_bytes, st);
    if (g_have_d2h_done) CUDA_OK(cudaStreamWaitEvent(st, g_d2h_done, 0));   // the area may still be on its way to the host
    copy_back = true;
  }
  const size_t nst = s.stages.size();
  const size_t rb = pl->elem_real_bytes();
  // where each intermediate lives: user buffers when allowed and large enough, else scratch
  const bool destroyable = staged_in || (pl->prob.flags & F_DESTROY_INPUT) || dev_in == dev_out;
  assign_buffers(pl, dev_in == dev_out, destroyable);
  auto buffer_ptr = [&](int id) -> void * {
    switch (id) {
      case BUF_USER_IN: return dev_in;
      case BUF_USER_OUT: return dev_out;
      case BUF_RING: return pl->fuse.ring;
      default: return pl->scratch[id - BUF_A];
    }
  };
  double xch_host[16] = {0};
  const bool dsync = transport_device_sync(pl);
  if (dsync) transport_begin_execute(pl);
  for (size_t i = 0; i < nst; i++) {
    const Stage &g = s.stages[i];
    StageParams sp = pl->params[i];
    sp.in = i == 0 ? dev_in : buffer_ptr(pl->assign[i - 1]);
    const bool last = i + 1 == nst;
    const bool xch = !last && g.exchange >= 0 && s.exchanges[g.exchange].nparts > 1;
    if (last) {
      sp.out[0] = dev_out;
    } else if (xch) {
      transport_stage_outputs(pl, (int)i, sp.out);
    } else {
      char *base = static_cast<char *>(buffer_ptr(pl->assign[i]));
      const size_t es = rb * (g.out_real ? 1 : 2);
      for (int q = 0; q < g.noseg; q++) sp.out[q] = base + (size_t)g.oseg_off[q] * es;
    }
    if (pl->fuse.active && (int)i == pl->fuse.first) {
      // the last two stages as one plane-fused launch
      StageParams a = pl->fuse.a, b = pl->fuse.b;
      a.in = sp.in;
      a.out[0] = pl->fuse.ring;
      b.in = pl->fuse.ring;
      b.out[0] = dev_out;
      if (pl->stage_timing) cudaEventRecord(pl->events[2 * i], st);
      cudaError_t e = pl->prec == PREC_F64 ? launch_fused_pow2<double>(a, b, pl->fuse.fp, st)
                                            : launch_fused_pow2<float>(a, b, pl->fuse.fp, st);
      CUDA_OK(e);
      if (pl->stage_timing) {
        cudaEventRecord(pl->events[2 * i + 1], st);
        cudaEventRecord(pl->events[2 * i + 2], st);
        cudaEventRecord(pl->events[2 * i + 3], st);
      }
      break;
    }
    double t0 = 0;
    if (dsync) {
      transport_wait_stage(pl, (int)i);
    } else if (xch) {
      t0 = MPI_Wtime();
      transport_before_stage(pl, (int)i);
      xch_host[g.exchange % 16] += MPI_Wtime() - t0;
    }
    if (pl->stage_timing) cudaEventRecord(pl->events[2 * i], st);
    if (sp.ntiles > 0) {
      cudaError_t e;
      const int kk = pl->kernel_kind[i];
      if (pl->prec == PREC_F64)
        e = kk == KERNEL_POW2 ? launch_stage_pow2<double>(sp, st)
            : kk == KERNEL_REG ? launch_stage_reg<double>(sp, st)
            : kk == KERNEL_MIXED ? launch_stage_mixed<double>(sp, &pl->mixed_ws, &pl->mixed_ws_bytes, st)
                                 : launch_stage_generic<double>(sp, st);
      else
        e = kk == KERNEL_POW2 ? launch_stage_pow2<float>(sp, st)
            : kk == KERNEL_REG ? launch_stage_reg<float>(sp, st)
            : kk == KERNEL_MIXED ? launch_stage_mixed<float>(sp, &pl->mixed_ws, &pl->mixed_ws_bytes, st)
                                 : launch_stage_generic<float>(sp, st);
      CUDA_OK(e);
    }
    if (pl->stage_timing) cudaEventRecord(pl->events[2 * i + 1], st);
    if (dsync) {
      transport_signal_stage(pl, (int)i);
    } else if (xch) {
      t0 = MPI_Wtime();
      transport_after_stage(pl, (int)i);
      xch_host[g.exchange % 16] += MPI_Wtime() - t0;
    }
  }
  if (staged_in || copy_back) {
    // the staging areas are free for the next host-pointer execute once these kernels are done
    copy_streams();
    CUDA_OK(cudaEventRecord(g_compute_done, st));
    g_have_compute_done = true;
  }
  if (copy_back) staged_d2h(out, dev_out, pl->user_out_bytes);
  if (blocking) {
    CUDA_OK(cudaStreamSynchronize(st));
    if (copy_back) CUDA_OK(cudaStreamSynchronize(g_d2h));
    if (pl->stage_timing) {
      TimerData &tm = pl->timer;
      tm.iter++;
      float whole = 0;
      cudaEventElapsedTime(&whole, pl->events.front(), pl->events.back());
      tm.whole += whole * 1e-3;
      for (size_t i = 0; i < nst; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pl->events[2 * i], pl->events[2 * i + 1]);
        pl->last_stage_ms[i] = ms;
        // the reference's slots (kernel/timer.c:297-319): trafo[] per transform stage, remap_3dto2d[] for the
        // stages of the 3-D mesh remap; itwiddle / otwiddle stay 0 (the modulations are fused into the stages)
        const int slot = s.stages[i].timer_slot;
        if (slot >= 0) tm.trafo[std::min<size_t>((size_t)slot, tm.trafo.size() - 1)] += ms * 1e-3;
        else tm.remap_3dto2d[slot == -1 ? 0 : 1] += ms * 1e-3;
      }
      size_t xslot = 0;
      for (size_t i = 0; i < nst; i++) {
        const int x = s.stages[i].exchange;
        if (x < 0 || (size_t)x >= s.exchanges.size()) continue;
        pl->last_xch_ms[x] = xch_host[x % 16] * 1e3;
        if (s.stages[i].timer_slot < 0) tm.remap_3dto2d[s.stages[i].timer_slot == -1 ? 0 : 1] += xch_host[x % 16];
        else if (!tm.remap.empty()) tm.remap[std::min<size_t>(xslot++, tm.remap.size() - 1)] += xch_host[x % 16];
      }
    }
  }
}

}  // namespace pfb
