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
//         single-GPU box.
//  nccl : the stage writes per-destination chunks into a send area and a grouped
//         ncclSend/ncclRecv all-to-all(v) moves them, stream-ordered, no host sync.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
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
};

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
      cudaIpcMemHandle_t h[2];
      int has[2];
    } mine, *all;
    memset(&mine, 0, sizeof mine);
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
  for (void *p : pl->peers->opened) cudaIpcCloseMemHandle(p);
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
