// fft_reg.cu -- host side of the register-resident kernel for real lines and lengths 3 * 2^k
// (fft_reg_kernel.h; instantiated per kind and precision in fft_reg_k*.cu): which stages it takes, tile and
// shared-memory geometry, twiddle tables.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "kernels.h"

namespace pfb {

cudaError_t launch_reg_k0_f64(StageParams &sp, cudaStream_t stream);
cudaError_t launch_reg_k1_f64(StageParams &sp, cudaStream_t stream);
cudaError_t launch_reg_k2_f64(StageParams &sp, cudaStream_t stream);
cudaError_t launch_reg_k0_f32(StageParams &sp, cudaStream_t stream);
cudaError_t launch_reg_k1_f32(StageParams &sp, cudaStream_t stream);
cudaError_t launch_reg_k2_f32(StageParams &sp, cudaStream_t stream);

namespace {

constexpr size_t kSmemLimit = 227 * 1024;

// complex line length -> (sub-transform length, points per thread, Q); the table of launch_reg_kind (fft_reg_kernel.h)
bool reg_geometry(int NL, int kind, int *nsub, int *E, int *q) {
  switch (NL) {
    case 64: case 128: case 256: case 512: case 1024: case 2048:
      if (kind == 0) return false;      // power-of-two complex lines belong to fft_pow2.cu
      *nsub = NL;
      *q = 1;
      break;
    case 192: case 384: case 768: case 1536:
      *nsub = NL / 3;
      *q = 3;
      break;
    default:
      return false;
  }
  *E = *nsub <= 512 ? 8 : 16;
  return true;
}

// threads per CTA the kernel is compiled for (launch_reg_one)
template <typename T>
int reg_class(int E, int Q) {
  return E == 8 ? ((sizeof(T) == 4 && Q == 1) ? 1024 : 768) : 512;
}

// per-pass twiddle tables [r-1][k] of the sub-transform (same layout as pow2_twiddle_tables, for a given E)
void reg_pass_tables(int L, int E, const double *roots, std::vector<double> *table) {
  int rem = L / E;
  const int R1 = E, R2 = rem >= E ? E : rem;
  rem /= R2;
  const int R3 = rem;
  auto emit = [&](int NS, int R) {
    const int tstep = L / (NS * R);
    for (int r = 1; r < R; r++)
      for (int k = 0; k < NS; k++) {
        const long long idx = (long long)k * r * tstep;
        table->push_back(roots[2 * idx]);
        table->push_back(roots[2 * idx + 1]);
      }
  };
  emit(R1, R2);
  if (R3 > 1) emit(R1 * R2, R3);
}

int kind_of(const Stage &g) { return g.op == OP_R2C ? 1 : (g.op == OP_C2R ? 2 : 0); }

}  // namespace

template <typename T>
bool reg_supported(const Stage &g, int L) {
  static const bool off = [] {
    const char *e = getenv("PFFT_B200_NO_REG");
    return e && atoi(e) != 0;
  }();
  if (off || g.ntile > 0 || g.noseg > kMaxSeg) return false;
  long long lines = 1;
  for (int k = 0; k < g.nbatch; k++) {
    if (g.batch[k].extent >= (1ll << 31)) return false;
    lines *= g.batch[k].extent;
  }
  if (lines >= (1ll << 31)) return false;
  auto all_even = [&](bool in_side) {
    for (int k = 0; k < g.nbatch; k++)
      if ((in_side ? g.batch[k].istride : g.batch[k].ostride) % 2) return false;
    return true;
  };
  int NL = L;
  const int kind = kind_of(g);
  if (g.op == OP_C2C) {
    if (g.in_real || g.out_real) return false;
  } else if (g.op == OP_R2C) {
    // packed pairs of reals: contiguous rows starting at even offsets, even window
    if (!g.in_real || g.out_real || g.n % 2 || L != g.n) return false;
    if (g.istride != 1 || g.iseg_stride != 0 || g.nin % 2 || g.zin % 2 || !all_even(true)) return false;
    NL = (int)(g.n / 2);
    if (g.zout < 0 || g.zout + g.nout > NL + 1) return false;
  } else if (g.op == OP_C2R) {
    if (g.in_real || !g.out_real || g.n % 2 || L != g.n) return false;
    if (g.ostride != 1 || g.noseg != 1 || g.nout % 2 || g.zout % 2 || !all_even(false)) return false;
    NL = (int)(g.n / 2);
    if (g.zin < 0 || g.zin + g.nin > NL + 1) return false;
  } else {
    return false;
  }
  int nsub, E, q;
  return reg_geometry(NL, kind, &nsub, &E, &q);
}

template <typename T>
bool reg_prepare(const Stage &g, StageParams &sp, UploadFn upload, void *ctx, std::string *err) {
  RegParams &rg = sp.rg;
  memset(&rg, 0, sizeof rg);
  rg.kind = kind_of(g);
  rg.NL = rg.kind ? sp.n / 2 : sp.L;
  int nsub = 0, Q = 1, E = 8;
  if (!reg_geometry(rg.NL, rg.kind, &nsub, &E, &Q)) {
    *err = "internal: length not served by the register-resident kernel";
    return false;
  }
  const int TL = Q * nsub / E;
  const size_t csize = 2 * sizeof(T);
  rg.E = E;
  rg.maxt = reg_class<T>(E, Q);
  // ---- tile: as many lines as the thread class holds -- one persistent CTA per SM with a staging buffer of its own,
  // all of its warps needed to cover latency; a strided side then moves tl elements per point in one run
  static const int forced = [] {
    const char *e = getenv("PFFT_B200_REG_TL");
    return e ? atoi(e) : 0;
  }();
  int tl = std::max(1, rg.maxt / TL);
  if (forced > 0) tl = std::min(forced, tl);
  if (g.tile_dim < 0) tl = 1;
  else if ((INT)tl > g.batch[g.tile_dim].extent) tl = (int)std::max<INT>(1, g.batch[g.tile_dim].extent);
  // ---- buffer line: the Q sub-line regions resp. the padded natural-order line (+ the Nyquist bin of c2r input)
  const int RS = Q == 1 ? nsub + (nsub >> 4) : ((nsub + (nsub >> 4) + 7) / 8 * 8 + 3);
  using std::max;
  const int P1 = E, rem1 = nsub / E, P2 = rem1 >= E ? E : rem1, P3 = rem1 / P2;
  const int twn = (P2 - 1) * P1 + (P3 > 1 ? (P3 - 1) * P1 * P2 : 0) + (Q > 1 ? rg.NL : 0) + (rg.kind ? rg.NL + 1 : 0);
  for (;;) {
    const int base = (max(Q * RS, rg.NL + (rg.NL >> 4) + 1) + 7) / 8 * 8;
    rg.pitch = base + (tl >= 8 ? 1 : (tl >= 4 ? 2 : (tl >= 2 ? 4 : 0)));   // lines of a tile start in different banks (tile-minor accesses)
    rg.spitch = (rg.NL + 2) | 1;                    // dense staging lines (+ the Nyquist bin), odd pitch
    const size_t need = ((size_t)tl * (rg.pitch + rg.spitch) + twn) * csize;
    if (need <= kSmemLimit) break;
    if (tl == 1) {
      *err = "internal: line does not fit the register-resident kernel";
      return false;
    }
    tl = tl > 2 ? tl - (tl > 8 ? 4 : (tl > 4 ? 2 : 1)) : 1;
  }
  sp.tl = tl;
  static const int line_bars = [] {
    const char *e = getenv("PFFT_B200_REG_LINEBAR");
    return e ? atoi(e) : 1;
  }();
  sp.line_bars = line_bars;
  // simple sides: one chunk, window starting at index 0 (the kernel checks the count), no modulation on the store
  const bool one_in_chunk = g.iseg_stride == 0 || g.iblk >= g.nin;
  rg.simple_in = (one_in_chunk && g.zin == 0) ? 1 : 0;
  rg.simple_out = (g.noseg == 1 && g.zout == 0 && !g.mod_out.on && !g.conj_out) ? 1 : 0;
  {
    const long long present = rg.kind == 1 ? (g.zin + g.nin) / 2 : g.zin + g.nin;      // packed / complex points up to the end of the window
    rg.third_zero = (Q == 3 && rg.kind != 2 && present <= 2ll * nsub) ? 1 : 0;
  }
  rg.diblk = make_fastdiv((unsigned)std::max(1, sp.iblk));
  rg.doblk = make_fastdiv((unsigned)std::max(1, sp.oblk));
  // ---- tables: [pass 2 | pass 3 of the sub-transform | w_NL^m | w_n^k], rounded once from fp64 roots
  std::vector<double> all;
  {
    std::vector<double> roots(2 * (size_t)nsub), pp;
    make_twiddles_f64(nsub, roots.data());
    reg_pass_tables(nsub, E, roots.data(), &pp);
    all.insert(all.end(), pp.begin(), pp.end());
  }
  if (Q > 1) {
    std::vector<double> roots(2 * (size_t)rg.NL);
    make_twiddles_f64(rg.NL, roots.data());
    all.insert(all.end(), roots.begin(), roots.end());
  }
  if (rg.kind) {
    std::vector<double> roots(4 * (size_t)rg.NL);
    make_twiddles_f64(2 * rg.NL, roots.data());
    all.insert(all.end(), roots.begin(), roots.begin() + 2 * (size_t)(rg.NL + 1));
  }
  if ((int)(all.size() / 2) != twn) {
    *err = "internal: twiddle table size mismatch";
    return false;
  }
  std::vector<T> host(all.begin(), all.end());
  rg.table_elems = twn;
  rg.tables = upload(host.data(), host.size() * sizeof(T), ctx);
  return true;
}

template <typename T>
cudaError_t launch_stage_reg(StageParams &sp, cudaStream_t stream) {
  if (sp.ntiles <= 0) return cudaSuccess;
  if (sp.ntiles >= (1ll << 31)) return cudaErrorInvalidValue;
  sp.rg.dalong = make_fastdiv((unsigned)std::max<long long>(1, sp.tiles_along));
  for (int k = 0; k < kMaxBatch; k++) sp.rg.dbext[k] = make_fastdiv((unsigned)std::max<long long>(1, k < sp.nbatch ? sp.bext[k] : 1));
  const bool f64 = sizeof(T) == 8;
  switch (sp.rg.kind) {
    case 0: return f64 ? launch_reg_k0_f64(sp, stream) : launch_reg_k0_f32(sp, stream);
    case 1: return f64 ? launch_reg_k1_f64(sp, stream) : launch_reg_k1_f32(sp, stream);
    default: return f64 ? launch_reg_k2_f64(sp, stream) : launch_reg_k2_f32(sp, stream);
  }
}

template bool reg_supported<float>(const Stage &, int);
template bool reg_supported<double>(const Stage &, int);
template bool reg_prepare<float>(const Stage &, StageParams &, UploadFn, void *, std::string *);
template bool reg_prepare<double>(const Stage &, StageParams &, UploadFn, void *, std::string *);
template cudaError_t launch_stage_reg<float>(StageParams &, cudaStream_t);
template cudaError_t launch_stage_reg<double>(StageParams &, cudaStream_t);

}  // namespace pfb
