// plan.h -- the plan object behind pfft_plan / pfftf_plan and its execution engine.
#pragma once
#include <cuda_runtime.h>
#include <mpi.h>

#include <string>
#include <vector>

#include "core.h"
#include "kernels.h"

namespace pfb {

enum Precision : int { PREC_F64 = 0, PREC_F32 = 1 };
enum TransportKind : int { TR_AUTO = 0, TR_NCCL = 1, TR_P2P = 2 };
enum KernelKind : int { KERNEL_GENERIC = 0, KERNEL_POW2 = 1, KERNEL_MIXED = 2, KERNEL_REG = 3 };

// Same fields as the reference's timer (kernel/ipfft.h:208-219) so that
// pfft_convert_timer2vec keeps its documented layout (kernel/timer.c:297-319).
struct TimerData {
  int rnk_pm = 0, rnk_trafo = 0, rnk_remap = 0;
  int iter = 0;
  double whole = 0;
  std::vector<double> trafo, remap;
  double remap_3dto2d[2] = {0, 0};
  double itwiddle = 0, otwiddle = 0;
  void shape(int rnk_pm_) {
    rnk_pm = rnk_pm_;
    rnk_trafo = 2 * rnk_pm_ + 2;
    rnk_remap = 2 * rnk_pm_;
    trafo.assign(rnk_trafo, 0.0);
    remap.assign(rnk_remap, 0.0);
  }
};

struct PeerMap;   // transports.cu

// Last two stages of a rank run as one plane-fused kernel (fft_pow2.cu: fused_pair_kernel)
struct FusedPair {
  bool possible = false;     // kernel-side conditions hold
  bool active = false;       // used by the current buffer assignment
  bool inplace_ok = false;   // plane p of the pair's input and of its output occupy the same addresses
  int first = -1;            // index of the pair's first stage
  StageParams a, b;          // stage parameters with the plane as batch dimension 0
  FusePlanes fp = {};
  void *ring = nullptr;      // intermediate: fp.ring plane slots
  size_t ring_bytes = 0;
};

struct PlanBase {
  int prec = PREC_F64;
  Problem prob;
  Schedule sched;
  MPI_Comm comm_cart = MPI_COMM_NULL;
  MPI_Comm comm_1d[kMaxGroups] = {MPI_COMM_NULL, MPI_COMM_NULL, MPI_COMM_NULL, MPI_COMM_NULL, MPI_COMM_NULL};   // one per exchange group
  void *planned_in = nullptr, *planned_out = nullptr;
  // device scratch: [0]=A, [1]=B (stage ping-pong / receive areas), [2]=W (NCCL send area)
  void *scratch[3] = {nullptr, nullptr, nullptr};
  size_t scratch_bytes = 0;
  size_t scratch_cap[3] = {0, 0, 0};
  size_t user_in_bytes = 0, user_out_bytes = 0;
  // buffer holding what stage i hands to stage i+1 (BufId), chosen per kind of execute
  std::vector<size_t> boundary_bytes;
  std::vector<char> boundary_remote;
  std::vector<int> assign;
  int assign_key = -1;
  std::vector<StageParams> params;
  std::vector<int> use_pow2;
  std::vector<int> kernel_kind;   // KernelKind per stage
  void *mixed_ws = nullptr;       // global workspace of the any-length kernel (lines beyond shared memory)
  size_t mixed_ws_bytes = 0;
  std::vector<void *> tables;     // device twiddle tables owned by the plan
  cudaStream_t stream = nullptr;
  int transport = TR_P2P;
  PeerMap *peers = nullptr;
  void *nccl_comm = nullptr;
  TimerData timer;
  std::vector<cudaEvent_t> events;   // 2 per stage
  bool stage_timing = true;
  std::vector<double> last_stage_ms;
  std::vector<double> last_xch_ms;
  FusedPair fuse;
  size_t elem_real_bytes() const { return prec == PREC_F64 ? 8 : 4; }
};

PlanBase *plan_create(int prec, const Problem &p, void *in, void *out, MPI_Comm comm);
void plan_execute(PlanBase *pl, void *in, void *out, bool blocking);
void plan_destroy(PlanBase *pl);

// process-wide settings (pfft_b200.h)
cudaStream_t default_stream();
void set_default_stream(cudaStream_t s);
int default_transport();
void ensure_device();
void set_error(const std::string &msg);
std::string &last_error_ref();

// communicator helpers shared with the ghost-cell module
MPI_Comm assure_cart(MPI_Comm comm);      // dup of a Cartesian comm, or a fresh 1-D mesh (reference kernel/procmesh.c:114-129)
void fill_problem(Problem *p, int kind, int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
                  const INT *iblock, const INT *oblock, int rnk_pm, const int *np, int sign, const int *kinds,
                  const int *skip, unsigned flags);

// transports.cu
struct ExchangeCtx;
bool transport_setup(PlanBase *pl, std::string *err);
void transport_teardown(PlanBase *pl);
// resolve the output base pointers of a stage followed by exchange `x` (before launching it)
void transport_stage_outputs(PlanBase *pl, int stage_idx, void **out_ptrs);
// device-side ordering of the p2p transport (every rank on a GPU of its own): no host synchronisation
bool transport_device_sync(const PlanBase *pl);
void transport_begin_execute(PlanBase *pl);
void transport_wait_stage(PlanBase *pl, int stage_idx);
void transport_signal_stage(PlanBase *pl, int stage_idx);
void transport_before_stage(PlanBase *pl, int stage_idx);
void transport_after_stage(PlanBase *pl, int stage_idx);

}  // namespace pfb
