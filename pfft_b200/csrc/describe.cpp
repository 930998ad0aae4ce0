// describe.cpp -- MPI-free, CUDA-free entry points over the integer layer and the planner
// (include/pfft_b200.h): used by tests to enumerate every rank of a mesh in one process.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "core.h"
#include "kernels.h"
#include "pfft_b200.h"

namespace pfb {
std::string &last_error_ref() {
  static thread_local std::string e;
  return e;
}

void fill_problem(Problem *p, int kind, int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
                  const INT *iblock, const INT *oblock, int rnk_pm, const int *np, int sign, const int *kinds,
                  const int *skip, unsigned flags) {
  p->rnk_n = rnk_n;
  for (int t = 0; t < rnk_n && t < kMaxDims; t++) {
    p->n[t] = n ? n[t] : (ni ? ni[t] : 0);
    p->ni[t] = ni ? ni[t] : p->n[t];
    p->no[t] = no ? no[t] : p->n[t];
    if (kinds) p->r2r_kinds[t] = kinds[t];
  }
  p->howmany = howmany;
  p->rnk_pm = rnk_pm;
  for (int t = 0; t < rnk_pm && t < kMaxMesh; t++) p->np[t] = np[t];
  p->has_iblock = iblock != nullptr;
  p->has_oblock = oblock != nullptr;
  for (int t = 0; t < rnk_pm && t < kMaxMesh; t++) {
    if (iblock) p->iblock[t] = iblock[t];
    if (oblock) p->oblock[t] = oblock[t];
  }
  p->kind = (Kind)kind;
  p->sign = sign;
  p->has_skip = skip != nullptr;
  if (skip)
    for (int t = 0; t <= rnk_pm && t <= kMaxMesh; t++) p->skip[t] = skip[t];
  p->flags = flags;
}
}  // namespace pfb

namespace pfb { void set_error(const std::string &msg); }
using namespace pfb;

namespace {
struct HostTables {
  std::vector<void *> ptrs;
  ~HostTables() {
    for (void *p : ptrs) free(p);
  }
};
void *upload_host(const void *host, size_t bytes, void *ctx) {
  void *p = malloc(bytes ? bytes : 1);
  memcpy(p, host, bytes);
  static_cast<HostTables *>(ctx)->ptrs.push_back(p);
  return p;
}
template <typename T>
int emulate_stage(const Stage &g, const void *in, void *const *outs) {
  StageParams sp;
  std::string err;
  HostTables tabs;
  if (!stage_params_basic(g, sp, &err)) { set_error(err); return 1; }
  if (g.ntile > 0 || g.iblk2 > 1 || g.oblk2 > 1) { set_error("micro-blocked stages belong to the register-resident kernel"); return 2; }
  if (g.op == OP_R2R) sp.tw_r2r = make_r2r_table<T>(sp.r2r_D, upload_host, &tabs);
  if (!mixed_prepare<T>(g, sp, upload_host, &tabs, &err)) { set_error(err); return 1; }
  stage_params_tiles(g, sp, sp.tl);
  sp.in = in;
  for (int q = 0; q < g.noseg; q++) sp.out[q] = outs[q];
  sp.mx.gws = 0;
  emulate_stage_mixed<T>(sp);
  return 0;
}
// which kernel would run stage g, and with what tile geometry (host planning only, no device)
template <typename T>
std::string describe_stage_kernel(const Stage &g) {
  StageParams sp;
  std::string err;
  HostTables tabs;
  if (!stage_params_basic(g, sp, &err)) return "{\"error\":\"" + err + "\"}";
  const int fam = stage_kernel_family<T>(g, sp.L);
  const size_t csize = 2 * sizeof(T);
  std::string j = "{";
  if (fam == 1) {
    const int tl = pow2_pick_tile<T>(g, sp.L);
    const int threads = tl * pow2_threads_per_line(sp.L);
    j += "\"family\":\"pow2\",\"tl\":" + std::to_string(tl) + ",\"threads\":" + std::to_string(threads) +
         ",\"micro_blocked\":" + std::to_string(g.ntile > 0 ? 1 : 0);
  } else if (fam == 3) {
    if (!reg_prepare<T>(g, sp, upload_host, &tabs, &err)) return "{\"error\":\"" + err + "\"}";
    const RegParams &rg = sp.rg;
    const int TL = rg.NL / rg.E;
    const size_t smem = ((size_t)sp.tl * (rg.pitch + rg.spitch) + rg.table_elems) * csize;
    j += "\"family\":\"reg\",\"tl\":" + std::to_string(sp.tl) + ",\"threads\":" + std::to_string(sp.tl * TL) +
         ",\"class_threads\":" + std::to_string(rg.maxt) + ",\"points_per_thread\":" + std::to_string(rg.E) +
         ",\"complex_points\":" + std::to_string(rg.NL) + ",\"kind\":" + std::to_string(rg.kind) +
         ",\"smem_bytes\":" + std::to_string(smem) + ",\"simple_in\":" + std::to_string(rg.simple_in) +
         ",\"simple_out\":" + std::to_string(rg.simple_out) + ",\"third_zero\":" + std::to_string(rg.third_zero);
  } else {
    if (g.op == OP_R2R) sp.tw_r2r = make_r2r_table<T>(sp.r2r_D, upload_host, &tabs);
    if (!mixed_prepare<T>(g, sp, upload_host, &tabs, &err)) return "{\"error\":\"" + err + "\"}";
    j += "\"family\":\"mixed\",\"tl\":" + std::to_string(sp.tl) + ",\"complex_points\":" + std::to_string(sp.mx.L) +
         ",\"transform_length\":" + std::to_string(sp.mx.Lc) + ",\"bluestein\":" + std::to_string(sp.mx.bluestein) +
         ",\"half_real\":" + std::to_string(sp.mx.half_real) + ",\"global_workspace\":" + std::to_string(sp.mx.gws) +
         ",\"smem_bytes\":" + std::to_string((size_t)sp.tl * sp.mx.pitch * csize) + ",\"passes\":" + std::to_string(sp.mx.npass);
  }
  return j + "}";
}
}  // namespace

extern "C" {

// TEST SUPPORT / introspection without a device: the kernel family and tile geometry the planner would pick for
// every stage of rank `pid` (JSON list).  prec 0: fp64, 1: fp32.
size_t pfftb200_describe_kernels(int prec, int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no,
                                 ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm,
                                 const int *np, int pid, int sign, const int *kinds, const int *skip_trafos,
                                 unsigned pfft_flags, char *buf, size_t buflen) {
  Problem p;
  fill_problem(&p, kind, rnk_n, n, ni, no, howmany, iblock, oblock, rnk_pm, np, sign, kinds, skip_trafos, pfft_flags);
  Schedule s;
  std::string j = "[";
  if (!build_schedule(p, pid, &s)) {
    j = "{\"error\":\"" + s.error + "\"}";
  } else {
    for (size_t i = 0; i < s.stages.size(); i++)
      j += (i ? "," : "") + (prec == 0 ? describe_stage_kernel<double>(s.stages[i]) : describe_stage_kernel<float>(s.stages[i]));
    j += "]";
  }
  if (buf && buflen) {
    size_t k = j.size() < buflen - 1 ? j.size() : buflen - 1;
    memcpy(buf, j.data(), k);
    buf[k] = 0;
  }
  return j.size() + 1;
}

// TEST SUPPORT / introspection without a device: the device-side ordering of the p2p transport for rank `pid` --
// per boundary whether an exchange fills it, which receive area (0 / 1, alternating) holds it and which ranks store
// into it; per stage the progress values it waits for (exchange_waits, core.h).  tests/test_exchange_ordering.py
// model-checks the protocol with these lists.
size_t pfftb200_describe_exchange_ordering(int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no,
                                           ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm,
                                           const int *np, int pid, int sign, const int *kinds, const int *skip_trafos,
                                           unsigned pfft_flags, char *buf, size_t buflen) {
  Problem p;
  fill_problem(&p, kind, rnk_n, n, ni, no, howmany, iblock, oblock, rnk_pm, np, sign, kinds, skip_trafos, pfft_flags);
  Schedule s;
  std::string j;
  if (!build_schedule(p, pid, &s)) {
    j = "{\"error\":\"" + s.error + "\"}";
  } else {
    const int nst = (int)s.stages.size(), nb = nst > 0 ? nst - 1 : 0;
    std::vector<int> assign(nb);
    for (int i = 0; i < nb; i++) assign[i] = i % 2;       // (plan.cu: assign_buffers with remote boundaries)
    j = "{\"pid\":" + std::to_string(pid) + ",\"nstages\":" + std::to_string(nst) + ",\"remote\":[";
    for (int i = 0; i < nb; i++) j += (i ? "," : "") + std::to_string(boundary_is_remote(s, i) ? 1 : 0);
    j += "],\"buffer\":[";
    for (int i = 0; i < nb; i++) j += (i ? "," : "") + std::to_string(assign[i]);
    j += "],\"writers\":[";
    for (int i = 0; i < nb; i++) {
      j += i ? ",[" : "[";
      if (boundary_is_remote(s, i)) {
        const Exchange &x = s.exchanges[s.stages[i].exchange];
        for (int q = 0; q < x.nparts; q++) j += (q ? "," : "") + std::to_string(s.groups[x.mesh_dim].members[q]);
      } else {
        j += std::to_string(pid);
      }
      j += "]";
    }
    j += "],\"waits\":[";
    for (int i = 0; i < nst; i++) {
      j += i ? ",[" : "[";
      const std::vector<ExchangeWait> w = exchange_waits(s, assign, i);
      for (size_t k = 0; k < w.size(); k++)
        j += std::string(k ? "," : "") + "[" + std::to_string(w[k].rank) + "," + std::to_string(w[k].back) + "," + std::to_string(w[k].code) + "]";
      j += "]";
    }
    j += "]}";
  }
  if (buf && buflen) {
    size_t k = j.size() < buflen - 1 ? j.size() : buflen - 1;
    memcpy(buf, j.data(), k);
    buf[k] = 0;
  }
  return j.size() + 1;
}

const char *pfftb200_version(void) { return "pfft_b200 0.1 (sm_100a)"; }

const char *pfftb200_last_error(void) { return last_error_ref().c_str(); }


// TEST SUPPORT: run stage `stage` of rank `pid`'s schedule through the BODY OF THE CUDA KERNEL
// (fft_mixed.h, the code stage_mixed_kernel executes) on the CPU, threads emulated one after the other
// between barriers.  `in`: the stage's input buffer, `outs`: one buffer per output chunk (host memory).
// prec 0: fp64, 1: fp32.  Returns 0, or non-zero with pfftb200_last_error() set.  Not a CPU fallback:
// nothing in the library calls it.
int pfftb200_emulate_stage(int prec, int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no,
                           ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm, const int *np,
                           int pid, int sign, const int *kinds, const int *skip_trafos, unsigned pfft_flags, int stage,
                           const void *in, void *const *outs) {
  Problem p;
  fill_problem(&p, kind, rnk_n, n, ni, no, howmany, iblock, oblock, rnk_pm, np, sign, kinds, skip_trafos, pfft_flags);
  Schedule s;
  if (!build_schedule(p, pid, &s)) { set_error(s.error); return 1; }
  if (stage < 0 || stage >= (int)s.stages.size()) { set_error("no such stage"); return 1; }
  return prec == 0 ? emulate_stage<double>(s.stages[stage], in, outs) : emulate_stage<float>(s.stages[stage], in, outs);
}

size_t pfftb200_describe_schedule(int kind, int rnk_n, const ptrdiff_t *n, const ptrdiff_t *ni, const ptrdiff_t *no,
                                  ptrdiff_t howmany, const ptrdiff_t *iblock, const ptrdiff_t *oblock, int rnk_pm,
                                  const int *np, int pid, int sign, const int *kinds, const int *skip_trafos,
                                  unsigned pfft_flags, char *buf, size_t buflen) {
  Problem p;
  fill_problem(&p, kind, rnk_n, n, ni, no, howmany, iblock, oblock, rnk_pm, np, sign, kinds, skip_trafos, pfft_flags);
  Schedule s;
  build_schedule(p, pid, &s);
  std::string j = schedule_to_json(s);
  if (buf && buflen) {
    size_t k = j.size() < buflen - 1 ? j.size() : buflen - 1;
    memcpy(buf, j.data(), k);
    buf[k] = 0;
  }
  return j.size() + 1;
}

void pfftb200_local_block(int kind, int rnk_n, const ptrdiff_t *ni, const ptrdiff_t *no, const ptrdiff_t *iblock,
                          const ptrdiff_t *oblock, int rnk_pm, const int *np, int pid, unsigned pfft_flags,
                          ptrdiff_t *local_ni, ptrdiff_t *local_i_start, ptrdiff_t *local_no,
                          ptrdiff_t *local_o_start) {
  Problem p;
  fill_problem(&p, kind, rnk_n, ni, ni, no, 1, iblock, oblock, rnk_pm, np, -1, nullptr, nullptr, pfft_flags);
  LocalSizes ls;
  local_block(p, pid, &ls);
  for (int t = 0; t < rnk_n; t++) {
    local_ni[t] = ls.lni[t];
    local_i_start[t] = ls.lis[t];
    local_no[t] = ls.lno[t];
    local_o_start[t] = ls.los[t];
  }
}

}  // extern "C"
