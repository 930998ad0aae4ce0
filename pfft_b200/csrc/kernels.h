// kernels.h -- device-side contract of one schedule stage (see core.h: Stage).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "core.h"

namespace pfb {

constexpr int kMaxFactors = 24;

struct SignModDev {
  int on, start, half, extra;
};

// x / d for a run-time constant d (round-up method, exact for 32-bit x)
struct FastDiv {
  unsigned m, s1, s2;
};
inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  unsigned l = 0;
  while ((1ull << l) < d) l++;
  f.m = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.s1 = l < 1 ? l : 1;
  f.s2 = l > 0 ? l - 1 : 0;
  return f;
}

// any-length stage kernel (fft_mixed.h / fft_mixed.cu)
constexpr int kMaxMixedPass = 24;
struct MixedPass {
  int R, LR, Ns, tstep;      // radix, butterflies per line (Lc / R), product of earlier radices, twiddle table step
  FastDiv dLR, dNs;
};
struct MixedParams {
  int L;                     // complex points of a line: n, n/2 for even-length real lines, 2D for DCT/DST
  int Lc;                    // length of the transforms actually run: L, or the Bluestein length M >= 2L - 1
  int half_real;             // 0: complex line, 1: r2c on n/2 packed points, 2: c2r on n/2 packed points
  int bluestein, swap, zero_fill;
  int in_pairs, out_pairs;   // packed real lines moved two reals at a time (contiguous, even offsets)
  int pitch;                 // complex elements per buffer line (padded)
  int tl_shift;              // log2(tl)
  int gws;                   // buffers live in the global workspace (line does not fit shared memory)
  int npass;
  MixedPass pass[kMaxMixedPass];
  FastDiv dnin, dnout, diblk, doblk, dL, dLc, dnin2, dnout2;
  const int *rev;            // Lc entries: natural index -> digit-reversed position (null: identity)
  const void *tw;            // Lc entries exp(-2 pi i k / Lc)
  const void *tw_half;       // n/2 + 1 entries exp(-2 pi i k / n) (packed real lines)
  const void *chirp;         // Bluestein: L entries exp(-i pi j^2 / L)
  const void *bhat;          // Bluestein: Lc entries FFT(conj chirp, wrapped) / Lc
  void *ws;                  // global workspace
  long long ws_stride;       // complex elements per CTA
};

// register-resident kernel for real lines and lengths 3 * 2^k (fft_reg_kernel.h)
struct RegParams {
  int NL;                    // complex points per line: n, or n / 2 for real lines
  int kind;                  // 0 complex, 1 r2c, 2 c2r
  int pitch;                 // complex elements per buffer line
  int spitch;                // pitch of the staging buffer's dense lines (inputs are fetched into a buffer of their own)
  int simple_in, simple_out; // one chunk, window starting at 0, no modulation / conjugation on that side's addressing
  int E, maxt;               // points per thread and thread class the stage was planned for
  int third_zero;            // Q = 3: the last third of every input line is zero padding
  FastDiv diblk, doblk;
  FastDiv dalong, dbext[kMaxBatch];   // tile walk: tiles along the tile dimension, batch extents (filled at launch)
  const void *tables;        // [pass-2 | pass-3 twiddles of the sub-transform | w_NL^m (Q > 1) | w_n^k, k <= n/2 (real lines)]
  int table_elems;           // complex entries of `tables` (host bookkeeping: shared-memory footprint)
};

// Plain-old-data copy of a Stage plus pointers; passed to kernels by value.
struct StageParams {
  const void *in;
  void *out[kMaxSeg];        // base pointer of output chunk `seg`
  int op, sign, r2r_kind;
  int n;                     // logical transform length
  int L;                     // complex FFT length executed on a line
  int nin, zin, nout, zout;
  long long istride, iseg_stride;
  int iblk;
  long long ostride;
  int oblk, noseg;
  int nbatch;
  long long bext[kMaxBatch], bis[kMaxBatch], bos[kMaxBatch];
  int tile_dim;              // -1: single line
  int tl;                    // lines per tile
  long long tiles_along;     // tiles along tile_dim
  long long ntiles;
  int in_real, out_real, conj_in, conj_out;
  SignModDev mod_in, mod_out;
  const void *twiddle;       // L entries exp(-2*pi*i*k/L) in the stage's precision
  // r2r (DCT/DST): the line is transformed as a zero-padded complex DFT of length L = 2D with a
  // twiddle before and after: Y_k = 2 F(sum_j w_j x_j exp(-i pi (j + a)(k + b) / D)), F = Re | -Im
  const void *tw_r2r;        // 8D entries exp(-2*pi*i*m/(8D))
  int r2r_a2, r2r_b2, r2r_D; // 2a, 2b, D
  int r2r_sine, r2r_half0, r2r_halfn;   // F = -Im; w_0 = 1/2; w_{n-1} = 1/2
  int nfac;
  int fac[kMaxFactors];
  // power-of-two fast path (fft_pow2.cu)
  int fast;                  // table-driven addressing applies
  int line_bars;             // per-line named barriers between passes (else CTA-wide)
  int ring_in, ring_out;     // fused pair: batch dimension 0 (the plane) wraps around a ring of this many slots
  const void *tw2, *tw3;     // per-pass twiddle tables [r-1][k]
  // micro-blocked layouts / explicit tiles (core.h: Stage::iblk2 ...): handled by stage_blk_kernel
  int iblk2, oblk2, ntile;
  int iswz_mask, oswz_mask, oswz_shift, oswz_batch;
  int half_cta;              // tile-minor exchange inside half-CTAs of tl/2 lines (stages without remote chunks)
  long long iblk2_stride, oblk2_stride;
  void *outp[16];            // out[out_seg[e]] + out_off[e], resolved at launch
  long long tile_ioff[kMaxTile], tile_ooff[kMaxTile];
  long long in_off[16];      // element offset of line index e*THREADS on input
  long long out_off[16];     // ... on output, inside chunk out_seg[e]
  int out_seg[16];
  MixedParams mx;
  RegParams rg;
};

// Plane-fused pair of stages (fft_pow2.cu: fused_pair_kernel).
struct FusePlanes {
  int planes;                // extent of the plane dimension (batch index 0 of both stages)
  int t1, t2;                // tiles per plane of the first / second stage
  int delta;                 // the second stage runs `delta` planes behind the first
  int ring;                  // plane slots of the intermediate ring buffer
  unsigned target1, target2; // counter value that means "plane complete"
  unsigned div_m, div_s1, div_s2;   // fast division by t1 + t2
  unsigned *done;            // [2][planes] completion counters
};

// host-side helpers (fft_tables.cpp)
bool stage_params_basic(const Stage &g, StageParams &sp, std::string *err);
void stage_params_tiles(const Stage &g, StageParams &sp, int tl);
int factorize_generic(int L, int *fac);
// parameters of an r2r kind (FFTW's enum values) on a line of n reals; false: unsupported kind
bool r2r_params(int kind, int n, int *a2, int *b2, int *D, int *sine, int *half0, int *halfn);          // radices for the generic shared-memory kernel
void make_twiddles_f64(int L, double *re_im);    // 2*L doubles, accurate to < 1 ulp
void make_twiddles_f32(int L, float *re_im);

// generic (any length) stage kernel; returns cudaErrorInvalidValue if the line does not fit
template <typename T>
cudaError_t launch_stage_generic(StageParams &sp, cudaStream_t stream);
template <typename T>
int generic_pick_tile(const Stage &g, int L);    // lines per tile (0: does not fit in shared memory)

// any-length stage kernel (fft_mixed.cu).  `upload` copies a host table to where the kernel will read it
// (device memory for the real thing, host memory for the CPU emulation of the kernel body).
using UploadFn = void *(*)(const void *host, size_t bytes, void *ctx);
template <typename T>
bool mixed_prepare(const Stage &g, StageParams &sp, UploadFn upload, void *ctx, std::string *err);
template <typename T>
cudaError_t launch_stage_mixed(StageParams &sp, void **ws, size_t *ws_bytes, cudaStream_t stream);
template <typename T>
void emulate_stage_mixed(StageParams &sp);      // the kernel body on the CPU (tests only)
template <typename T>
void *make_r2r_table(int D, UploadFn upload, void *ctx);   // plan.cu

// fast path: power-of-two complex lines held in registers (fft_pow2.cu)
template <typename T>
bool pow2_supported(const Stage &g, int L);
template <typename T>
cudaError_t launch_stage_pow2(StageParams &sp, cudaStream_t stream);
template <typename T>
int pow2_pick_tile(const Stage &g, int L);
template <typename T>
void pow2_prepare(const Stage &g, StageParams &sp);
template <typename T>
int fused_pick_tile(int L);
template <typename T>
cudaError_t launch_fused_pow2(StageParams &a, StageParams &b, FusePlanes &fp, cudaStream_t stream);
void pow2_twiddle_tables(int L, const double *roots, std::vector<double> *table, size_t *off2, size_t *off3);

// real lines (packed half-length transforms) and complex lines of length 3 * 2^k held in registers (fft_reg.cu)
template <typename T>
bool reg_supported(const Stage &g, int L);
template <typename T>
bool reg_prepare(const Stage &g, StageParams &sp, UploadFn upload, void *ctx, std::string *err);
template <typename T>
cudaError_t launch_stage_reg(StageParams &sp, cudaStream_t stream);

unsigned long long &launch_counter();

// Kernel family of a stage (plan.cu: KernelKind values): 1 pow2 (fft_pow2.cu), 3 reg (fft_reg.cu), 2 mixed (fft_mixed.cu)
template <typename T>
inline int stage_kernel_family(const Stage &g, int L) {
  if (pow2_supported<T>(g, L)) return 1;
  if (reg_supported<T>(g, L)) return 3;
  return 2;
}

}  // namespace pfb
