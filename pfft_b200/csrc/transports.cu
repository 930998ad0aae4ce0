// transports.cu -- how the chunks a stage produces reach the ranks that consume them.
// Replaces the reference's FFTW-MPI transpose plans (kernel/transpose.c:140-181,224-316,
// MPI_Alltoall(v) underneath) with one of two NVLink transports:
//
//  p2p  : every rank maps its peers' receive areas (CUDA IPC) at plan time; the stage
//         kernel's epilogue stores each output element straight into the consuming
//         rank's memory -- pack + send + unpack are the store itself.  Ordering is a
//         barrier over the mesh dimension's communicator before ("receive areas are
//         free") and after ("all chunks have landed") the stage.  Also works between
//         ranks that share one GPU, which is how multi-rank plans are tested on a
//         single-GPU box.  When every rank of the plan has a GPU of its own the ordering
//         is done ON THE DEVICE instead (progress counters in peer-mapped memory, a
//         one-CTA signal kernel behind every stage and a one-CTA wait kernel in front of
//         the stages that need it): no host synchronisation, the whole transform is
//         stream-ordered and pfftb200_execute_async really is asynchronous.
//  nccl : the stage writes per-destination chunks into a send area and a grouped
//         ncclSend/ncclRecv all-to-all(v) moves them, stream-ordered, no host sync.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <vector>

#include "plan.h"

namespace pfb {

struct PeerMap {
  // [buf 0/1][cart rank] -> pointer valid in this process (own pointer for self)
  std::vector<void *> ptr[2];
  std::vector<void *> opened;   // to close
  // cart rank of member q of mesh dimension m
  std::vector<int> member_rank[kMaxGroups];
  // device-side ordering (see the head of this file): flags[0] = number of executes started,
  // flags[kFlagBase + r] = progress of cart rank r as r itself wrote it: 64 * execute + stages completed
  bool dev_sync = false;
  unsigned *flags = nullptr;
  unsigned **peer_flags_dev = nullptr;   // device array [np]: rank r's flags (null: not a peer, or self)
  int me = 0, np = 1;
};
constexpr int kFlagBase = 16, kFlagWords = 256;
constexpr size_t kFlagBytes = (size_t)2 << 20;   // an allocation of its own (IPC handles map whole allocations)

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) return api;
#define LOAD(field, sym) *(void **)(&api.field) = dlsym(api.handle, sym)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  api.ok = api.GetUniqueId && api.CommInitRank && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
  return api;
}

// NCCL communicators are expensive to build: cache them per set of world ranks.
std::map<std::vector<int>, ncclComm_t> &nccl_cache() {
  static std::map<std::vector<int>, ncclComm_t> c;
  return c;
}

// ---- device-side ordering kernels -----------------------------------------------------------
struct WaitList {
  int n;
  int rank[2 * kMaxSeg];
  int back[2 * kMaxSeg];        // 1: the target lies in the previous execute
  unsigned code[2 * kMaxSeg];   // stages completed
  unsigned long long timeout_ns;   // give up (loudly) after this long: PFFT_B200_WAIT_TIMEOUT_S, default 600 s
};
__global__ void xch_begin_kernel(unsigned *flags) { flags[0] += 1u; }
__global__ void xch_signal_kernel(const unsigned *flags, unsigned *const *peer, int me, int np, unsigned code) {
  const int r = threadIdx.x;
  if (r >= np || !peer[r]) return;
  const unsigned v = flags[0] * 64u + code;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer[r] + kFlagBase + me), "r"(v) : "memory");
}
__global__ void xch_wait_kernel(const unsigned *flags, const __grid_constant__ WaitList w) {
  const int i = threadIdx.x;
  if (i >= w.n) return;
  const unsigned target = (flags[0] - (unsigned)w.back[i]) * 64u + w.code[i];
  const unsigned *p = flags + kFlagBase + w.rank[i];
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (unsigned spins = 0;; spins++) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if ((int)(v - target) >= 0) break;
    __nanosleep(200);
    if ((spins & 0xfff) == 0xfff) {
      // a peer that never arrives (crashed rank, mismatched sequence of executes) must not hang the GPU for good --
      // and must not be ignored either: the kernel traps, the next CUDA call of this rank reports the error
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > w.timeout_ns) {
        printf("pfft_b200: gave up waiting for rank %d (progress %u, wanted %u)\n", w.rank[i], v, target);
        __trap();
      }
    }
  }
}

void fill_members(PlanBase *pl) {
  PeerMap *pm = pl->peers;
  const Schedule &s = pl->sched;
  for (int g = 0; g < s.ngroups; g++) pm->member_rank[g].assign(s.groups[g].members, s.groups[g].members + s.groups[g].size);
}

}  // namespace

bool transport_setup(PlanBase *pl, std::string *err) {
  pl->peers = new PeerMap;
  fill_members(pl);
  int np = 1, me = 0;
  MPI_Comm_size(pl->comm_cart, &np);
  MPI_Comm_rank(pl->comm_cart, &me);
  bool ok = true;
  if (pl->transport == TR_P2P) {
    struct Handles {
      cudaIpcMemHandle_t h[3];   // receive areas A, B and the ordering flags
      int has[3];
      char bus_id[32];           // which GPU this rank drives
    } mine, *all;
    memset(&mine, 0, sizeof mine);
    PeerMap *pm = pl->peers;
    pm->me = me;
    pm->np = np;
    {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetPCIBusId(mine.bus_id, (int)sizeof mine.bus_id, dev);
      // progress of every rank starts as "execute 0 complete"
      std::vector<unsigned> init(kFlagWords, 0u);
      for (int r = 0; r < kFlagWords - kFlagBase; r++) init[kFlagBase + r] = 63u;
      if (np <= kFlagWords - kFlagBase && cudaMalloc(&pm->flags, kFlagBytes) == cudaSuccess) {
        cudaMemcpy(pm->flags, init.data(), kFlagWords * sizeof(unsigned), cudaMemcpyHostToDevice);
        mine.has[2] = cudaIpcGetMemHandle(&mine.h[2], pm->flags) == cudaSuccess ? 1 : 0;
      }
      cudaGetLastError();
    }
    for (int b = 0; b < 2; b++) {
      mine.has[b] = pl->scratch[b] != nullptr;
      if (mine.has[b]) {
        cudaError_t e = cudaIpcGetMemHandle(&mine.h[b], pl->scratch[b]);
        if (e != cudaSuccess) {
          *err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
          cudaGetLastError();
          mine.has[b] = -1;
          ok = false;
        }
      }
    }
    all = new Handles[np];
    MPI_Allgather(&mine, (int)sizeof(Handles), MPI_BYTE, all, (int)sizeof(Handles), MPI_BYTE, pl->comm_cart);
    // which cart ranks do I ever write to?
    std::vector<char> is_peer(np, 0);
    for (auto &x : pl->sched.exchanges)
      if (x.nparts > 1)
        for (int q = 0; q < x.nparts; q++) is_peer[pl->peers->member_rank[x.mesh_dim][q]] = 1;
    for (int b = 0; b < 2; b++) {
      pl->peers->ptr[b].assign(np, nullptr);
      pl->peers->ptr[b][me] = pl->scratch[b];
      for (int rk = 0; rk < np && ok; rk++) {
        if (rk == me || !is_peer[rk] || all[rk].has[b] <= 0) continue;
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[rk].h[b], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          *err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
          cudaGetLastError();
          ok = false;
          break;
        }
        pl->peers->ptr[b][rk] = p;
        pl->peers->opened.push_back(p);
      }
    }
    // device-side ordering needs a GPU per rank (a waiting kernel must never keep a peer's kernels off
    // the GPU they share) and few enough stages for the progress code
    const char *env = getenv("PFFT_B200_P2P_SYNC");
    bool dsync = ok && !(env && !strcmp(env, "host")) && np <= 64 && pl->sched.stages.size() < 60;
    for (int rk = 0; rk < np && dsync; rk++) {
      if (all[rk].has[2] <= 0) dsync = false;
      for (int r2 = 0; r2 < rk && dsync; r2++)
        if (!strncmp(all[rk].bus_id, all[r2].bus_id, sizeof mine.bus_id)) dsync = false;
    }
    if (dsync) {
      std::vector<unsigned *> pf(np, nullptr);
      for (int rk = 0; rk < np && dsync; rk++) {
        if (rk == me || !is_peer[rk]) continue;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[rk].h[2], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          dsync = false;
          break;
        }
        pf[rk] = static_cast<unsigned *>(p);
        pm->opened.push_back(p);
      }
      if (dsync && cudaMalloc(&pm->peer_flags_dev, np * sizeof(unsigned *)) == cudaSuccess)
        cudaMemcpy(pm->peer_flags_dev, pf.data(), np * sizeof(unsigned *), cudaMemcpyHostToDevice);
      else
        dsync = false;
    }
    // all ranks or none
    int mine_ok = dsync ? 1 : 0, all_ok = 0;
    MPI_Allreduce(&mine_ok, &all_ok, 1, MPI_INT, MPI_MIN, pl->comm_cart);
    pm->dev_sync = all_ok != 0;
    delete[] all;
  } else {
    NcclApi &api = nccl();
    if (!api.ok) {
      *err = "libnccl.so.2 could not be loaded";
      return false;
    }
    std::vector<int> world(np);
    for (int rk = 0; rk < np; rk++) world[rk] = minimpi_world_rank(pl->comm_cart, rk);
    auto it = nccl_cache().find(world);
    if (it == nccl_cache().end()) {
      ncclUniqueId id;
      memset(&id, 0, sizeof id);
      if (me == 0) api.GetUniqueId(&id);
      MPI_Bcast(&id, (int)sizeof id, MPI_BYTE, 0, pl->comm_cart);
      ncclComm_t c = nullptr;
      ncclResult_t r = api.CommInitRank(&c, np, id, me);
      if (r != ncclSuccess) {
        *err = std::string("ncclCommInitRank: ") + (api.GetErrorString ? api.GetErrorString(r) : "error");
        return false;
      }
      it = nccl_cache().emplace(world, c).first;
    }
    pl->nccl_comm = it->second;
  }
  return ok;
}

void transport_teardown(PlanBase *pl) {
  if (!pl->peers) return;
  if (pl->transport == TR_P2P) {
    // peers may still be storing into my receive areas or flags
    cudaStreamSynchronize(pl->stream);
    MPI_Barrier(pl->comm_cart);
  }
  for (void *p : pl->peers->opened) cudaIpcCloseMemHandle(p);
  if (pl->peers->flags) cudaFree(pl->peers->flags);
  if (pl->peers->peer_flags_dev) cudaFree(pl->peers->peer_flags_dev);
  delete pl->peers;
  pl->peers = nullptr;
}

void transport_stage_outputs(PlanBase *pl, int i, void **out) {
  const Stage &g = pl->sched.stages[i];
  const Exchange &x = pl->sched.exchanges[g.exchange];
  const size_t es = pl->elem_real_bytes() * (g.out_real ? 1 : 2);
  const int b = pl->assign[i] - BUF_A;
  for (int q = 0; q < x.nparts; q++) {
    if (q == x.me) {
      out[q] = static_cast<char *>(pl->scratch[b]) + (size_t)x.me * x.recv_cnt * es;
    } else if (pl->transport == TR_P2P) {
      char *base = static_cast<char *>(pl->peers->ptr[b][pl->peers->member_rank[x.mesh_dim][q]]);
      out[q] = base ? base + (size_t)x.me * x.peer_recv_cnt[q] * es : nullptr;
    } else {
      out[q] = static_cast<char *>(pl->scratch[2]) + (size_t)g.oseg_off[q] * es;
    }
  }
}

bool transport_device_sync(const PlanBase *pl) { return pl->peers && pl->transport == TR_P2P && pl->peers->dev_sync; }

void transport_begin_execute(PlanBase *pl) {
  xch_begin_kernel<<<1, 1, 0, pl->stream>>>(pl->peers->flags);
}

// Device-side ordering in front of stage i: the rule is exchange_waits (core.h / planner.cpp), a pure function of
// the schedule that the CPU tests model-check (tests/test_exchange_ordering.py); here it becomes one wait kernel.
void transport_wait_stage(PlanBase *pl, int i) {
  PeerMap *pm = pl->peers;
  const std::vector<ExchangeWait> waits = exchange_waits(pl->sched, pl->assign, i);
  static const unsigned long long timeout_ns = [] {
    const char *e = getenv("PFFT_B200_WAIT_TIMEOUT_S");
    const double sec = e ? atof(e) : 600.0;
    return (unsigned long long)((sec > 0 ? sec : 600.0) * 1e9);
  }();
  WaitList w;
  w.n = 0;
  w.timeout_ns = timeout_ns;
  for (const ExchangeWait &e : waits) {
    if (w.n >= 2 * kMaxSeg) break;
    w.rank[w.n] = e.rank;
    w.back[w.n] = e.back;
    w.code[w.n] = (unsigned)e.code;
    w.n++;
  }
  if (w.n > 0) xch_wait_kernel<<<1, 2 * kMaxSeg, 0, pl->stream>>>(pm->flags, w);
}

void transport_signal_stage(PlanBase *pl, int i) {
  PeerMap *pm = pl->peers;
  xch_signal_kernel<<<1, 64, 0, pl->stream>>>(pm->flags, pm->peer_flags_dev, pm->me, pm->np, (unsigned)(i + 1));
}

void transport_before_stage(PlanBase *pl, int i) {
  if (pl->transport != TR_P2P) return;
  const Stage &g = pl->sched.stages[i];
  const Exchange &x = pl->sched.exchanges[g.exchange];
  // peers must have finished reading the areas this stage is about to overwrite
  cudaStreamSynchronize(pl->stream);
  MPI_Barrier(pl->comm_1d[x.mesh_dim]);
}

void transport_after_stage(PlanBase *pl, int i) {
  const Stage &g = pl->sched.stages[i];
  const Exchange &x = pl->sched.exchanges[g.exchange];
  if (pl->transport == TR_P2P) {
    cudaStreamSynchronize(pl->stream);
    MPI_Barrier(pl->comm_1d[x.mesh_dim]);
    return;
  }
  NcclApi &api = nccl();
  ncclComm_t comm = static_cast<ncclComm_t>(pl->nccl_comm);
  const size_t es = pl->elem_real_bytes() * (g.out_real ? 1 : 2);
  const int b = pl->assign[i] - BUF_A;
  char *recv_base = static_cast<char *>(pl->scratch[b]);
  char *send_base = static_cast<char *>(pl->scratch[2]);
  api.GroupStart();
  for (int q = 0; q < x.nparts; q++) {
    if (q == x.me) continue;
    const int peer = pl->peers->member_rank[x.mesh_dim][q];
    if (x.send_cnt[q] > 0)
      api.Send(send_base + (size_t)g.oseg_off[q] * es, (size_t)x.send_cnt[q] * es, ncclUint8, peer, comm, pl->stream);
    if (x.recv_cnt > 0)
      api.Recv(recv_base + (size_t)q * x.recv_cnt * es, (size_t)x.recv_cnt * es, ncclUint8, peer, comm, pl->stream);
  }
  api.GroupEnd();
}

}  // namespace pfb
