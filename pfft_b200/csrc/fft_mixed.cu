// fft_mixed.cu -- any-length stage kernel for sm_100a: host-side planning (radix schedule, Bluestein
// tables, tile and buffer geometry), the CUDA kernel around the shared body in fft_mixed.h, and the CPU
// emulation of that body for the planner tests.  Replaces, per stage, the reference's embed loop + FFTW
// plan + FFTW copy plan + truncate loop (kernel/outrafo.c:154-168, kernel/sertrafo.c:489-554,1073-1245).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "fft_mixed.h"

namespace pfb {

namespace {

constexpr int kThreads = 256;      // per CTA when two CTAs share an SM; one resident CTA runs 2 * kThreads
constexpr size_t kSmemMax = 227 * 1024;

// radices with a codelet, large ones first (the first pass needs no twiddles: give it the most work)
bool radix_schedule(int L, std::vector<int> *out) {
  out->clear();
  int m = L;
  std::vector<int> f;
  int twos = 0;
  while (m % 2 == 0) { m /= 2; twos++; }
  while (twos >= 4) { f.push_back(16); twos -= 4; }
  if (twos == 3) f.push_back(8);
  if (twos == 2) f.push_back(4);
  if (twos == 1) {
    // a lone factor 2: merge with a 16 into 8 * 4 where possible (two balanced passes instead of 16 * 2)
    auto it = std::find(f.begin(), f.end(), 16);
    if (it != f.end()) { *it = 8; f.push_back(4); }
    else f.push_back(2);
  }
  for (int p : {13, 11, 7, 5, 3})
    while (m % p == 0) { f.push_back(p); m /= p; }
  if (m != 1) return false;
  std::sort(f.begin(), f.end(), [](int a, int b) { return a > b; });
  if ((int)f.size() > kMaxMixedPass) return false;
  // The first pass touches positions R i + r (stride-R butterflies): conflict-free in shared memory for radix 16
  // (17-element padded pitch) and for odd radices, 2- to 4-way conflicts for 2, 4, 8 -- so a 16 goes first if there
  // is one, else the largest odd radix; later passes have strides that are multiples of 16 or odd.
  auto lead = std::find(f.begin(), f.end(), 16);
  if (lead == f.end()) lead = std::find_if(f.begin(), f.end(), [](int r) { return r % 2 == 1; });
  if (lead != f.end() && lead != f.begin()) std::rotate(f.begin(), lead, lead + 1);
  *out = f;
  return true;
}

void unit_roots_ld(long long L, std::vector<long double> *re, std::vector<long double> *im) {
  // exp(-2 pi i k / L) through the fp64 table maker's octant folding would lose the long double digits;
  // plain cosl/sinl of the reduced angle are accurate to ~1e-19 here, ample for tables rounded to fp64
  const long double two_pi = 6.283185307179586476925286766559005768L;
  re->resize(L);
  im->resize(L);
  for (long long k = 0; k < L; k++) {
    const long double a = two_pi * (long double)k / (long double)L;
    (*re)[k] = cosl(a);
    (*im)[k] = -sinl(a);
  }
}

// in-place radix-2 FFT in long double (host, planning time only)
void fft_pow2_ld(std::vector<long double> &re, std::vector<long double> &im) {
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  std::vector<long double> wr, wi;
  unit_roots_ld((long long)n, &wr, &wi);
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t step = n / len;
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; k++) {
        const long double ur = re[i + k], ui = im[i + k];
        const long double xr = re[i + k + len / 2], xi = im[i + k + len / 2];
        const long double vr = xr * wr[k * step] - xi * wi[k * step], vi = xr * wi[k * step] + xi * wr[k * step];
        re[i + k] = ur + vr; im[i + k] = ui + vi;
        re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
      }
  }
}

template <typename T>
void *upload_complex(const std::vector<long double> &re, const std::vector<long double> &im, UploadFn upload, void *ctx) {
  std::vector<T> h(2 * re.size());
  for (size_t i = 0; i < re.size(); i++) { h[2 * i] = (T)re[i]; h[2 * i + 1] = (T)im[i]; }
  return upload(h.data(), h.size() * sizeof(T), ctx);
}

int floor_pow2(long long x) {
  int p = 1;
  while ((long long)p * 2 <= x) p *= 2;
  return p;
}

struct DevExec {
  template <class F>
  __host__ __device__ __forceinline__ void run(F &&f) {
#if defined(__CUDA_ARCH__)
    f((int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
#endif
  }
};

struct HostExec {
  int nthr;
  template <class F>
  __host__ __device__ void run(F &&f) {
    for (int t = 0; t < nthr; t++) f(t, nthr);
  }
};

template <typename T, bool GWS, int MAXT, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB) stage_mixed_kernel(const __grid_constant__ StageParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *buf = GWS ? reinterpret_cast<cx<T> *>(sp.mx.ws) + (size_t)blockIdx.x * (size_t)sp.mx.ws_stride
                   : reinterpret_cast<cx<T> *>(smem_raw);
  DevExec ex;
  for (long long tile = blockIdx.x; tile < sp.ntiles; tile += gridDim.x) {
    long long ibase, obase, t_is, t_os;
    int tvalid;
    mixed_locate(sp, tile, &ibase, &obase, &tvalid, &t_is, &t_os);
    mixed_tile<T>(sp, buf, ibase, obase, tvalid, t_is, t_os, ex);
  }
}

}  // namespace

// Decide how the stage runs in the any-length kernel and build its tables.
template <typename T>
bool mixed_prepare(const Stage &g, StageParams &sp, UploadFn upload, void *ctx, std::string *err) {
  MixedParams &mx = sp.mx;
  memset(&mx, 0, sizeof mx);
  const int n = sp.n;
  const size_t csize = 2 * sizeof(T);
  static const bool no_half = [] {
    const char *e = getenv("PFFT_B200_NO_HALFREAL");
    return e && atoi(e) != 0;
  }();
  // even-length real lines are packed as n / 2 complex points (reference: FFTW's r2c / c2r plans, kernel/sertrafo.c:517-530)
  mx.half_real = 0;
  if (!no_half && n % 2 == 0 && n >= 4) {
    if (g.op == OP_R2C && g.in_real && !g.out_real) mx.half_real = 1;
    if (g.op == OP_C2R && !g.in_real && g.out_real) mx.half_real = 2;
  }
  mx.L = mx.half_real ? n / 2 : sp.L;           // (sp.L = 2D for DCT/DST lines)
  mx.swap = (g.op == OP_C2C && g.sign > 0) || (g.op == OP_C2R && !mx.half_real);
  std::vector<int> radices;
  mx.bluestein = 0;
  mx.Lc = mx.L;
  if (g.op == OP_COPY) {
    radices.clear();
  } else if (!radix_schedule(mx.L, &radices)) {
    // a prime factor without a codelet: Bluestein, two power-of-two transforms of length M >= 2L - 1
    mx.bluestein = 1;
    int M = 16;
    while (M < 2 * mx.L - 1) M *= 2;
    mx.Lc = M;
    radix_schedule(M, &radices);
  }
  mx.npass = (int)radices.size();
  int Ns = 1;
  for (int i = 0; i < mx.npass; i++) {
    MixedPass &p = mx.pass[i];
    p.R = radices[i];
    p.LR = mx.Lc / p.R;
    p.Ns = Ns;
    p.tstep = mx.Lc / (Ns * p.R);
    p.dLR = make_fastdiv((unsigned)p.LR);
    p.dNs = make_fastdiv((unsigned)p.Ns);
    Ns *= p.R;
  }
  // buffer line: padded transform length
  mx.pitch = mx.Lc + (mx.Lc >> 4) + 2;
  // what the load leaves untouched must be cleared: zero padding of pruned inputs, the Bluestein tail, DCT/DST padding
  const int line_in = mx.half_real == 2 ? n / 2 + 1 : (mx.half_real == 1 ? n : (g.op == OP_C2R ? n / 2 + 1 : n));
  mx.zero_fill = (sp.nin != line_in || sp.zin != 0 || mx.bluestein || g.op == OP_R2R) ? 1 : 0;
  // ---- tile: lines per CTA (a power of two) against the shared-memory budget
  const size_t per_line = (size_t)mx.pitch * csize;
  const bool strided = g.istride != 1 || g.ostride != 1;
  const int want = strided ? (int)std::max<size_t>(1, 128 / csize) : 1;   // 128-byte runs on a strided side
  const long long fit1 = (long long)((kSmemMax - 1024) / per_line), fit2 = (long long)((kSmemMax / 2 - 1024) / per_line);
  int tl;
  mx.gws = 0;
  // two resident CTAs overlap each other's phases; short lines get more of them per tile (>= 4096 points) so that
  // the per-phase barriers are amortised
  int amort = 1;
  while (amort * mx.Lc < 4096 && amort < 64) amort *= 2;
  if (fit2 >= want) tl = std::min(floor_pow2(fit2), std::max(std::max(want, 4), amort));
  else if (fit1 >= 1) tl = std::min(floor_pow2(fit1), want);
  else { tl = 1; mx.gws = 1; }
  static const int forced = [] {
    const char *e = getenv("PFFT_B200_MIXED_TL");
    return e ? atoi(e) : 0;
  }();
  if (forced > 0 && !mx.gws && (long long)forced <= fit1) tl = floor_pow2(forced);
  if (g.tile_dim < 0) tl = 1;
  else if ((INT)tl > g.batch[g.tile_dim].extent) tl = floor_pow2(std::max<INT>(1, g.batch[g.tile_dim].extent));
  sp.tl = tl;
  mx.tl_shift = 0;
  while ((1 << mx.tl_shift) < tl) mx.tl_shift++;
  if ((long long)tl * std::max(sp.nin, std::max(sp.nout, mx.Lc)) >= (1ll << 31)) {
    *err = "transform length too large for the stage kernel";
    return false;
  }
  mx.dnin = make_fastdiv((unsigned)std::max(1, sp.nin));
  mx.dnout = make_fastdiv((unsigned)std::max(1, sp.nout));
  mx.diblk = make_fastdiv((unsigned)std::max(1, sp.iblk));
  mx.doblk = make_fastdiv((unsigned)std::max(1, sp.oblk));
  mx.dnin2 = make_fastdiv((unsigned)std::max(1, sp.nin / 2));
  mx.dnout2 = make_fastdiv((unsigned)std::max(1, sp.nout / 2));
  mx.dL = make_fastdiv((unsigned)std::max(1, mx.L));
  // packed real lines whose rows are contiguous and start at even offsets move two reals per access
  auto all_even = [&](bool in_side) {
    for (int k = 0; k < g.nbatch; k++)
      if ((in_side ? g.batch[k].istride : g.batch[k].ostride) % 2) return false;
    return true;
  };
  mx.in_pairs = (mx.half_real == 1 && g.istride == 1 && g.iseg_stride == 0 && sp.nin % 2 == 0 && sp.zin % 2 == 0 && all_even(true)) ? 1 : 0;
  mx.out_pairs = (mx.half_real == 2 && g.ostride == 1 && g.noseg == 1 && sp.nout % 2 == 0 && sp.zout % 2 == 0 && all_even(false)) ? 1 : 0;
  mx.dLc = make_fastdiv((unsigned)std::max(1, mx.Lc));
  // ---- tables
  if (mx.npass > 1) {
    // digit reversal for the in-place decimation-in-time passes: n = r_P + R_P (r_{P-1} + R_{P-1} (...)) sits at
    // sum_p r_p Ns_p (Ns_p = product of the radices of the passes before p)
    std::vector<int> rev(mx.Lc);
    for (int i = 0; i < mx.Lc; i++) {
      int m = i, pos = 0;
      for (int q = mx.npass - 1; q >= 0; q--) {
        pos += (m % mx.pass[q].R) * mx.pass[q].Ns;
        m /= mx.pass[q].R;
      }
      rev[i] = pos;
    }
    mx.rev = static_cast<const int *>(upload(rev.data(), rev.size() * sizeof(int), ctx));
  }
  std::vector<long double> re, im;
  if (mx.npass > 0) {
    unit_roots_ld(mx.Lc, &re, &im);
    mx.tw = upload_complex<T>(re, im, upload, ctx);
  }
  if (mx.half_real) {
    unit_roots_ld(n, &re, &im);
    re.resize(n / 2 + 1);
    im.resize(n / 2 + 1);
    mx.tw_half = upload_complex<T>(re, im, upload, ctx);
  }
  if (mx.bluestein) {
    const long long L = mx.L, M = mx.Lc;
    // chirp c[j] = exp(-i pi j^2 / L) = exp(-2 pi i (j^2 mod 2L) / 2L)
    std::vector<long double> r2, i2;
    unit_roots_ld(2 * L, &r2, &i2);
    std::vector<long double> cr(L), ci(L);
    for (long long j = 0; j < L; j++) {
      const long long q = (j * j) % (2 * L);
      cr[j] = r2[q];
      ci[j] = i2[q];
    }
    mx.chirp = upload_complex<T>(cr, ci, upload, ctx);
    // b[m] = conj c[|m|] wrapped into length M, spectrum scaled by 1 / M (the inverse transform's normalisation)
    std::vector<long double> br(M, 0.0L), bi(M, 0.0L);
    for (long long m = 0; m < L; m++) {
      br[m] = cr[m];
      bi[m] = -ci[m];
      if (m) { br[M - m] = cr[m]; bi[M - m] = -ci[m]; }
    }
    fft_pow2_ld(br, bi);
    for (long long m = 0; m < M; m++) { br[m] /= (long double)M; bi[m] /= (long double)M; }
    mx.bhat = upload_complex<T>(br, bi, upload, ctx);
  }
  return true;
}

template <typename T>
cudaError_t launch_stage_mixed(StageParams &sp, void **ws, size_t *ws_bytes, cudaStream_t stream) {
  if (sp.ntiles <= 0) return cudaSuccess;
  const size_t tile_bytes = (size_t)sp.tl * sp.mx.pitch * 2 * sizeof(T);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sp.mx.gws) {
    // lines beyond shared memory: ping-pong buffers in a global workspace (L2-resident for moderate lengths)
    long long grid = std::min<long long>(sp.ntiles, 2ll * sms);
    const size_t cap = (size_t)2 << 30;
    if ((size_t)grid * tile_bytes > cap) grid = std::max<long long>(1, (long long)(cap / tile_bytes));
    const size_t need = (size_t)grid * tile_bytes;
    if (*ws_bytes < need) {
      cudaError_t e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) return e;
      if (*ws) cudaFree(*ws);
      *ws = nullptr;
      *ws_bytes = 0;
      e = cudaMalloc(ws, need);
      if (e != cudaSuccess) return e;
      *ws_bytes = need;
    }
    sp.mx.ws = *ws;
    sp.mx.ws_stride = (long long)(tile_bytes / (2 * sizeof(T)));
    stage_mixed_kernel<T, true, kThreads><<<(unsigned)grid, kThreads, 0, stream>>>(sp);
    launch_counter()++;
    return cudaGetLastError();
  }
  // two resident CTAs of 256 threads need <= 128 registers per thread (the fp64 radix-13/16 butterflies then spill a
  // little); PFFT_B200_MIXED_MINB=1 lifts the cap (one resident CTA with all the registers it wants)
  static const int minb = [] {
    const char *e = getenv("PFFT_B200_MIXED_MINB");
    return e ? atoi(e) : 2;
  }();
  using K = void (*)(StageParams);
  K kerns[3] = {stage_mixed_kernel<T, false, kThreads, 1>, stage_mixed_kernel<T, false, kThreads, 2>,
                stage_mixed_kernel<T, false, 2 * kThreads, 1>};
  static bool attr_set = false;
  if (!attr_set) {
    for (K k : kerns) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
      if (e != cudaSuccess) return e;
    }
    attr_set = true;
  }
  if (tile_bytes > kSmemMax) return cudaErrorInvalidValue;
  const bool two_fit = 2 * (tile_bytes + 1024) <= kSmemMax;
  // one resident CTA: 512 threads keep 16 warps per SM (fp32 only: the fp64 butterflies need more than 128 registers)
  const int which = (two_fit && minb >= 2) ? 1 : ((!two_fit && sizeof(T) == 4) ? 2 : 0);
  const int threads = which == 2 ? 2 * kThreads : kThreads;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kerns[which], threads, tile_bytes);
  if (per_sm < 1) per_sm = 1;
  const long long grid = std::min<long long>(sp.ntiles, (long long)sms * per_sm);
  kerns[which]<<<(unsigned)grid, threads, tile_bytes, stream>>>(sp);
  launch_counter()++;
  return cudaGetLastError();
}

// The kernel body on the CPU: same functions, threads run one after the other between barriers.
template <typename T>
void emulate_stage_mixed(StageParams &sp) {
  if (sp.ntiles <= 0) return;
  std::vector<cx<T>> buf((size_t)sp.tl * sp.mx.pitch);
  HostExec ex{kThreads};
  for (long long tile = 0; tile < sp.ntiles; tile++) {
    long long ibase, obase, t_is, t_os;
    int tvalid;
    mixed_locate(sp, tile, &ibase, &obase, &tvalid, &t_is, &t_os);
    mixed_tile<T>(sp, buf.data(), ibase, obase, tvalid, t_is, t_os, ex);
  }
}

template bool mixed_prepare<float>(const Stage &, StageParams &, UploadFn, void *, std::string *);
template bool mixed_prepare<double>(const Stage &, StageParams &, UploadFn, void *, std::string *);
template cudaError_t launch_stage_mixed<float>(StageParams &, void **, size_t *, cudaStream_t);
template cudaError_t launch_stage_mixed<double>(StageParams &, void **, size_t *, cudaStream_t);
template void emulate_stage_mixed<float>(StageParams &);
template void emulate_stage_mixed<double>(StageParams &);

}  // namespace pfb
