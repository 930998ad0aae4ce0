// plan.cu -- plan life cycle and the execute loop (replaces the reference's
// plan_partrafo / execute_full, kernel/partrafo.c:317-525, api/api-basic.c:1044-1107).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
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

template <typename T>
static bool build_stage_params(PlanBase *pl, std::string *err) {
  struct TableSet {
    void *dev;
    size_t off2, off3;
    bool has_pow2;
  };
  std::map<int, TableSet> table_of;
  const Schedule &s = pl->sched;
  for (size_t i = 0; i < s.stages.size(); i++) {
    const Stage &g = s.stages[i];
    StageParams sp;
    memset(&sp, 0, sizeof sp);
    if (g.op == OP_R2R) {
      *err = "r2r transforms are not implemented yet";
      return false;
    }
    if (g.n > (1 << 24)) {
      *err = "transform length too large";
      return false;
    }
    const int L = (int)g.n;
    sp.op = g.op;
    sp.sign = g.sign;
    sp.r2r_kind = g.r2r_kind;
    sp.n = (int)g.n;
    sp.L = L;
    sp.nin = (int)g.nin; sp.zin = (int)g.zin; sp.nout = (int)g.nout; sp.zout = (int)g.zout;
    sp.istride = g.istride; sp.iseg_stride = g.iseg_stride; sp.iblk = (int)g.iblk;
    sp.ostride = g.ostride; sp.oblk = (int)g.oblk; sp.noseg = g.noseg;
    sp.nbatch = g.nbatch;
    long long lines = 1;
    for (int k = 0; k < g.nbatch; k++) {
      sp.bext[k] = g.batch[k].extent;
      sp.bis[k] = g.batch[k].istride;
      sp.bos[k] = g.batch[k].ostride;
      lines *= g.batch[k].extent;
    }
    if (g.in_elems == 0 && g.nin > 0) lines = 0;   // some batch extent of size 1 was dropped but another is 0
    sp.tile_dim = g.tile_dim;
    sp.in_real = g.in_real; sp.out_real = g.out_real; sp.conj_in = g.conj_in; sp.conj_out = g.conj_out;
    sp.mod_in = {g.mod_in.on, (int)g.mod_in.start, (int)g.mod_in.half, g.mod_in.extra};
    sp.mod_out = {g.mod_out.on, (int)g.mod_out.start, (int)g.mod_out.half, g.mod_out.extra};
    const bool fast = pow2_supported<T>(g, L);
    int tl = fast ? pow2_pick_tile<T>(g, L) : generic_pick_tile<T>(g, L);
    if (tl <= 0) {
      *err = "transform length " + std::to_string(L) + " does not fit the shared-memory kernels yet";
      return false;
    }
    sp.tl = tl;
    if (g.tile_dim >= 0) {
      long long others = 1;
      for (int k = 0; k < g.nbatch; k++)
        if (k != g.tile_dim) others *= g.batch[k].extent;
      sp.tiles_along = (g.batch[g.tile_dim].extent + tl - 1) / tl;
      sp.ntiles = others * sp.tiles_along;
    } else {
      sp.tiles_along = 1;
      sp.ntiles = lines;
    }
    if (g.nout == 0 || lines == 0) sp.ntiles = 0;
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
        void *dev = nullptr;
        CUDA_OK(cudaMalloc(&dev, host.size() * sizeof(T)));
        CUDA_OK(cudaMemcpy(dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
        pl->tables.push_back(dev);
        it = table_of.emplace(2 * L + (fast ? 1 : 0), TableSet{dev, off2, off3, fast}).first;
      }
      const TableSet &ts = it->second;
      sp.twiddle = ts.dev;
      if (fast) {
        sp.tw2 = static_cast<const T *>(ts.dev) + 2 * ts.off2;
        sp.tw3 = static_cast<const T *>(ts.dev) + 2 * ts.off3;
        pow2_prepare<T>(g, sp);
      }
    }
    pl->params.push_back(sp);
    pl->use_pow2.push_back(fast ? 1 : 0);
  }
  return true;
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
  if (!all_ranks_ok(pl->comm_cart, ok)) {
    set_error(err.empty() ? "planning failed on another rank" : err);
    plan_destroy(pl);
    return nullptr;
  }
  const Schedule &s = pl->sched;
  const int r = s.rnk_pm_eff;
  for (int t = 0; t < r; t++) {
    int remain[kMaxMesh] = {0, 0, 0};
    remain[t] = 1;
    MPI_Cart_sub(pl->comm_cart, remain, &pl->comm_1d[t]);
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
  if (nst >= 2) CUDA_OK(cudaMalloc(&pl->scratch[0], pl->scratch_bytes));
  if (nst >= 3) CUDA_OK(cudaMalloc(&pl->scratch[1], pl->scratch_bytes));
  if (any_exchange && tr == TR_NCCL) CUDA_OK(cudaMalloc(&pl->scratch[2], pl->scratch_bytes));
  if (any_exchange) {
    std::string terr;
    bool tok = transport_setup(pl, &terr);
    if (!all_ranks_ok(pl->comm_cart, tok)) {
      set_error(terr.empty() ? "transport setup failed on another rank" : terr);
      plan_destroy(pl);
      return nullptr;
    }
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
  if (pl->stream || true) cudaStreamSynchronize(pl->stream);
  transport_teardown(pl);
  for (auto &e : pl->events) cudaEventDestroy(e);
  for (void *t : pl->tables) cudaFree(t);
  for (int k = 0; k < 3; k++)
    if (pl->scratch[k]) cudaFree(pl->scratch[k]);
  if (pl->stage_in) cudaFree(pl->stage_in);
  if (pl->stage_out) cudaFree(pl->stage_out);
  for (int t = 0; t < kMaxMesh; t++)
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

void plan_execute(PlanBase *pl, void *in, void *out, bool blocking) {
  const Schedule &s = pl->sched;
  cudaStream_t st = pl->stream;
  if (!in) in = pl->planned_in;
  if (!out) out = pl->planned_out;
  // host pointers are staged through device memory (counted in the end-to-end numbers)
  void *dev_in = in, *dev_out = out;
  bool copy_back = false;
  if (pl->user_in_bytes && !device_accessible(in)) {
    if (pl->stage_in_bytes < pl->user_in_bytes) {
      if (pl->stage_in) cudaFree(pl->stage_in);
      CUDA_OK(cudaMalloc(&pl->stage_in, pl->user_in_bytes));
      pl->stage_in_bytes = pl->user_in_bytes;
    }
    CUDA_OK(cudaMemcpyAsync(pl->stage_in, in, pl->user_in_bytes, cudaMemcpyHostToDevice, st));
    dev_in = pl->stage_in;
  }
  if (pl->user_out_bytes && !device_accessible(out)) {
    if (pl->stage_out_bytes < pl->user_out_bytes) {
      if (pl->stage_out) cudaFree(pl->stage_out);
      CUDA_OK(cudaMalloc(&pl->stage_out, pl->user_out_bytes));
      pl->stage_out_bytes = pl->user_out_bytes;
    }
    dev_out = pl->stage_out;
    copy_back = true;
  }
  const size_t nst = s.stages.size();
  const size_t rb = pl->elem_real_bytes();
  double xch_host[16] = {0};
  for (size_t i = 0; i < nst; i++) {
    const Stage &g = s.stages[i];
    StageParams sp = pl->params[i];
    sp.in = g.in_buf == BUF_USER_IN ? dev_in : pl->scratch[g.in_buf - BUF_A];
    const bool last = i + 1 == nst;
    const bool xch = !last && g.exchange >= 0 && s.exchanges[g.exchange].nparts > 1;
    if (last) {
      sp.out[0] = dev_out;
    } else if (xch) {
      transport_stage_outputs(pl, (int)i, sp.out);
    } else {
      char *base = static_cast<char *>(pl->scratch[g.out_buf - BUF_A]);
      const size_t es = rb * (g.out_real ? 1 : 2);
      for (int q = 0; q < g.noseg; q++) sp.out[q] = base + (size_t)g.oseg_off[q] * es;
    }
    double t0 = 0;
    if (xch) {
      t0 = MPI_Wtime();
      transport_before_stage(pl, (int)i);
      xch_host[g.exchange % 16] += MPI_Wtime() - t0;
    }
    if (pl->stage_timing) cudaEventRecord(pl->events[2 * i], st);
    if (sp.ntiles > 0) {
      cudaError_t e;
      if (pl->prec == PREC_F64) e = pl->use_pow2[i] ? launch_stage_pow2<double>(sp, st) : launch_stage_generic<double>(sp, st);
      else e = pl->use_pow2[i] ? launch_stage_pow2<float>(sp, st) : launch_stage_generic<float>(sp, st);
      CUDA_OK(e);
    }
    if (pl->stage_timing) cudaEventRecord(pl->events[2 * i + 1], st);
    if (xch) {
      t0 = MPI_Wtime();
      transport_after_stage(pl, (int)i);
      xch_host[g.exchange % 16] += MPI_Wtime() - t0;
    }
  }
  if (copy_back) CUDA_OK(cudaMemcpyAsync(out, pl->stage_out, pl->user_out_bytes, cudaMemcpyDeviceToHost, st));
  if (blocking) {
    CUDA_OK(cudaStreamSynchronize(st));
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
        tm.trafo[std::min<size_t>(i, tm.trafo.size() - 1)] += ms * 1e-3;
      }
      for (size_t x = 0; x < s.exchanges.size(); x++) {
        pl->last_xch_ms[x] = xch_host[x % 16] * 1e3;
        if (!tm.remap.empty()) tm.remap[std::min<size_t>(x, tm.remap.size() - 1)] += xch_host[x % 16];
      }
    }
  }
}

}  // namespace pfb
