// core.h -- host-side data model of pfft_b200: problem description, block
// decomposition, and the stage list ("schedule") a plan executes.  Pure integer
// code; no MPI, no CUDA.  The same planner serves a real rank and a "virtual" rank
// (tests enumerate every rank of a mesh in one process and simulate the schedule).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace pfb {

using INT = ptrdiff_t;
constexpr int kMaxDims = 8;    // rnk_n
constexpr int kMaxMesh = 3;    // rnk_pm
constexpr int kMaxBatch = 4;   // batch dims a stage kernel indexes (after merging)
constexpr int kMaxSeg = 32;    // ranks along one mesh dimension
constexpr int kMaxTile = 64;   // lines of an explicit tile (micro-blocked layouts)
constexpr int kMaxGroups = 5;  // exchange groups of a schedule: mesh dimensions (+ 2 for the 3-D -> 2-D mesh remap)

// public flag bits (include/pfft.h; reference api/pfft.h:528-545)
enum : unsigned {
  F_TRANSPOSED_IN = 1u << 0, F_TRANSPOSED_OUT = 1u << 1, F_SHIFTED_IN = 1u << 2, F_SHIFTED_OUT = 1u << 3,
  F_PRESERVE_INPUT = 1u << 8, F_DESTROY_INPUT = 1u << 9, F_PADDED_R2C = 1u << 11
};

enum class Kind : int { C2C = 0, R2C = 1, C2R = 2, R2R = 3 };

// ---- 1-D block decomposition (reference kernel/block.c:54-95) ----------------------
INT block_count(INT n, INT blk);
INT block_default(INT n, INT user_blk, int nprocs);   // user block or ceil(n/P)
INT block_extent(INT n, INT blk, int which);          // 0 beyond the last block
INT block_offset(INT n, INT blk, int which);          // 0 (not n) beyond the last block

struct Problem {
  int rnk_n = 0;
  INT n[kMaxDims] = {0}, ni[kMaxDims] = {0}, no[kMaxDims] = {0};
  INT howmany = 1;
  bool has_iblock = false, has_oblock = false;
  INT iblock[kMaxDims] = {0}, oblock[kMaxDims] = {0};
  int rnk_pm = 0;              // mesh rank as the user built it (1..3)
  int np[kMaxMesh] = {1, 1, 1};
  Kind kind = Kind::C2C;
  int sign = -1;
  int r2r_kinds[kMaxDims] = {0};
  bool has_skip = false;
  int skip[kMaxMesh + 1] = {0};
  unsigned flags = 0;
};

// what pfft_local_block_* / pfft_local_size_* report (reference kernel/partrafo.c:99-315)
struct LocalSizes {
  INT lni[kMaxDims], lis[kMaxDims], lno[kMaxDims], los[kMaxDims];
};

// Effective 2-D mesh behind a 3-D mesh for 3-D data (reference kernel/procmesh.c:191-391)
struct Mesh3dto2d {
  bool active = false;
  int q0 = 1, q1 = 1;
};
Mesh3dto2d mesh_3dto2d(const Problem &p);
void mesh_coords(int rnk_pm, const int *np, int pid, int *coords);   // row-major, last fastest

// returns false when the reference planner would return NULL (kernel/partrafo.c:337-373)
bool problem_is_legal(const Problem &p, std::string *why);
void local_block(const Problem &p, int pid, LocalSizes *out);
// elements (complex for c2c/r2c/c2r, real for r2r) the caller must allocate per array
INT alloc_local(const Problem &p, int pid);

// ---- schedule ---------------------------------------------------------------------
enum StageOp : int { OP_COPY = 0, OP_C2C = 1, OP_R2C = 2, OP_C2R = 3, OP_R2R = 4 };

struct BatchDim {
  INT extent, istride, ostride;
  int dim;   // logical dimension id (for diagnostics; -1 = tuple)
};

// ±1 modulation fused into a stage (reference api/api-basic.c:1186-1285):
// factor(idx) = (g < half) ? ((g odd) ? -1 : 1) * extra : 1  with g = idx + start
struct SignMod {
  int on = 0;
  INT start = 0, half = 0;
  int extra = 1;
};

// One pass over the local array: (optionally) transform along one dimension and
// re-lay the data out for what follows (next local stage, an exchange, or the user).
// Element units: `in_real`/`out_real` say whether strides count reals or complex numbers.
struct Stage {
  int op = OP_COPY;
  int sign = -1;
  int r2r_kind = 0;
  int dim = -1;               // logical dimension transformed / re-laid out
  INT n = 1;                  // logical transform length (embed target)
  INT nin = 1, zin = 0;       // elements present on input along `dim`, and their offset inside the length-n line
  INT nout = 1, zout = 0;     // outputs kept: logical indices [zout, zout+nout)
  // input addressing along `dim`: j -> seg = j / iblk; seg*iseg_stride + (j - seg*iblk)*istride
  INT istride = 1, iblk = 1, iseg_stride = 0;
  // output addressing along `dim`: k -> seg = k / oblk; obase[seg] + (k - seg*oblk)*ostride
  INT ostride = 1, oblk = 1;
  int noseg = 1;
  INT oseg_off[kMaxSeg] = {0};   // element offset of chunk `seg` inside the destination's receive area (peer-local)
  INT oseg_cnt[kMaxSeg] = {0};   // elements of chunk `seg` (what an exchange moves)
  int nbatch = 0;
  BatchDim batch[kMaxBatch];
  int tile_dim = -1;          // index into batch[]: lines of one tile differ along it
  // Micro-blocked layouts (planner.cpp: plan_microblocks).  Along `dim`, inside a chunk, index x sits at
  //   (x / iblk2) * iblk2_stride + (x % iblk2) * istride      (same with o... on the output side);
  // 1 = plain.  With ntile > 0 the tile is explicit: its lines are the combinations of the low parts
  // of the micro-blocked batch dimensions, line l at tile_ioff[l] / tile_ooff[l] from the tile's base,
  // and batch[] enumerates whole tiles (tile_dim = -1).
  INT iblk2 = 1, iblk2_stride = 0;
  INT oblk2 = 1, oblk2_stride = 0;
  int ntile = 0;
  INT tile_ioff[kMaxTile] = {0}, tile_ooff[kMaxTile] = {0};
  // Bank swizzle baked into the blocked layout: inside block b (index along the gathered dimension) the
  // lines are permuted, so that a consumer reading `iblk2` words out of each of several consecutive
  // blocks (dense in shared memory) hits different banks.  Consumer: line l of the tile sits at
  // tile_ioff[l] ^ ((b & iswz_mask) * iblk2), b = (index inside the segment) / iblk2.  Producer: in a tile
  // whose batch coordinate oswz_batch is c, line l goes to tile_ooff[l] ^ ((c & oswz_mask) << oswz_shift).
  int iswz_mask = 0;
  int oswz_mask = 0, oswz_shift = 0, oswz_batch = -1;
  bool in_real = false, out_real = false;
  bool conj_in = false, conj_out = false;
  SignMod mod_in, mod_out;
  // plumbing
  int in_buf = 0, out_buf = 0;   // buffer ids, see BufId
  int exchange = -1;          // index into Schedule::exchanges applied to this stage's output, or -1
  INT in_elems = 0, out_elems = 0;   // footprint in elements (for buffer sizing / roofline bytes)
  // where the stage's time is booked in the reference's timer layout (kernel/timer.c:297-319):
  // >= 0: trafo[slot]; -1 / -2: remap_3dto2d[0] / [1] (stages of the 3-D-mesh remap)
  int timer_slot = 0;
};

enum BufId : int { BUF_USER_IN = 0, BUF_USER_OUT = 1, BUF_A = 2, BUF_B = 3, BUF_RING = 4 };

// All-to-all over one mesh dimension. Chunk `p` of the producing stage goes to member
// `p` of that mesh dimension's communicator and lands at recv_off[me] elements inside
// its receive buffer (uniform chunk size per receiver).
struct Exchange {
  int mesh_dim = 0;
  int nparts = 1;
  int me = 0;
  INT send_cnt[kMaxSeg] = {0};    // elements I send to member p
  INT recv_cnt = 0;               // elements of every chunk I receive
  INT peer_recv_cnt[kMaxSeg] = {0};   // chunk size at member p (offset of my chunk there = me * that)
  bool elem_real = false;
};

struct Schedule {
  Problem prob;
  int pid = 0;
  int rnk_pm_eff = 0;           // mesh rank the schedule runs on (2 for a remapped 3-D mesh)
  int np_eff[kMaxMesh] = {1, 1, 1};
  int coords_eff[kMaxMesh] = {0, 0, 0};
  // Exchange groups (Exchange::mesh_dim indexes them): the rows / columns of the mesh the
  // schedule runs on and, for 3-D data on a 3-D mesh, the two sub-groups of the remap
  // (reference kernel/procmesh.c:191-391).  members = ranks of the user's Cartesian communicator.
  int ngroups = 0;
  struct Group {
    int size = 1, me = 0;
    int members[kMaxSeg] = {0};
  } groups[kMaxGroups];
  LocalSizes ls;
  std::vector<Stage> stages;
  std::vector<Exchange> exchanges;
  INT scratch_elems = 0;        // complex (or real for r2r) elements per scratch buffer
  std::string error;            // non-empty: unsupported configuration
};

// Device-side ordering of the p2p transport (transports.cu) as a pure function of the schedule: what stage `stage`
// must wait for before it may run.  Every rank publishes its progress (64 * execute + stages completed) to its peers;
//  (a) a stage whose input was stored by the members of an exchange group waits until all of them have completed
//      the stage before;
//  (b) a stage that stores into its peers' receive area b waits until every peer has completed the stage that
//      consumed what area b held before -- stage j + 1, j the previous boundary kept in b (in this execute, else the
//      last one of the previous execute: back = 1).
// `assign[i]` = buffer id holding the boundary between stage i and i + 1 (any ids; equal ids = same area).
struct ExchangeWait {
  int rank;   // rank of the user's Cartesian communicator to wait for
  int back;   // 1: the target lies in the previous execute
  int code;   // stages that rank must have completed
};
std::vector<ExchangeWait> exchange_waits(const Schedule &s, const std::vector<int> &assign, int stage);
// does stage i hand its output to an exchange with more than one member?
bool boundary_is_remote(const Schedule &s, int i);

// Build the schedule of rank `pid`. Returns false (with sched->error) when the
// configuration is legal for the reference but not yet supported here.
bool build_schedule(const Problem &p, int pid, Schedule *sched);
// Thread geometry of the register-resident power-of-two kernels, shared by the planner (tile shapes
// of micro-blocked layouts) and fft_pow2.cu: points per thread, threads per line, lines per 512-thread tile.
int pow2_points_per_thread(int L);
inline int pow2_threads_per_line(int L) { return L / pow2_points_per_thread(L); }
std::string schedule_to_json(const Schedule &s);

}  // namespace pfb
