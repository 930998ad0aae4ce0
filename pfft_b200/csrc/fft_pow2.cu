// fft_pow2.cu -- register-resident power-of-two stage kernel for sm_100a (the hot
// kernel of the headline benchmark: three of these per 1024^3 transform).
//
// One CTA owns a tile of `tl` lines.  A line of N complex points is spread over
// N/E threads, E points per thread, in the "strided" distribution
//     thread t holds indices  t, t + N/E, t + 2N/E, ...
// which the Stockham autosort formulation preserves: every pass reads exactly
// that set and the last pass leaves its outputs in it again.  Consequences:
//   * the first pass loads straight from global memory into registers and the last
//     pass stores straight from registers, both coalesced -- either along the line
//     (thread index = line-major) or, for the strided/transposed side of a stage,
//     along the tile (thread index = tile-minor, so a warp touches `tl` neighbouring
//     lines = tl*16 contiguous bytes per point);
//   * shared memory is used only for the exchanges between passes (2 for N = 1024),
//     and the change between the load mapping and the store mapping of threads
//     happens for free inside the last exchange: no separate transpose, pack or
//     unpack pass ever touches HBM.
// Zero padding (ni < n), truncation (no < n), chunked sources/destinations of the
// exchanges, the +-1 index-shift modulation and conjugation are all applied in the
// load / store of this one kernel.  Backward transforms swap re/im on load and store
// around the same forward butterflies.
//
// Reference counterpart: FFTW guru64 plan + FFTW rank-0 copy plan per stage
// (kernel/sertrafo.c:489-554,604-646), i.e. two passes over memory; here it is one.
#include <cuda_runtime.h>
#include <stdlib.h>

#include <vector>

#include "kernels.h"

namespace pfb {

namespace {

template <typename T> struct C2 { using type = void; };
template <> struct C2<double> { using type = double2; };
template <> struct C2<float> { using type = float2; };

template <typename T>
struct alignas(2 * sizeof(T)) cx {
  T x, y;
};

template <typename T>
__device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
  cx<T> r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}

// ---- forward DFTs of length 2, 4, 8, 16 on registers ---------------------------------
// dftR leaves output q in slot perm(q); slot_of<R>(q) gives that slot.
template <typename T>
__device__ __forceinline__ void dft2(cx<T> &a, cx<T> &b) {
  cx<T> t = a;
  a.x = t.x + b.x; a.y = t.y + b.y;
  b.x = t.x - b.x; b.y = t.y - b.y;
}

// natural-order 4-point forward DFT: (a,b,c,d) -> (X0,X1,X2,X3)
template <typename T>
__device__ __forceinline__ void dft4(cx<T> &a, cx<T> &b, cx<T> &c, cx<T> &d) {
  cx<T> s0{a.x + c.x, a.y + c.y}, s1{a.x - c.x, a.y - c.y};
  cx<T> s2{b.x + d.x, b.y + d.y}, s3{b.x - d.x, b.y - d.y};
  a.x = s0.x + s2.x; a.y = s0.y + s2.y;          // X0
  c.x = s0.x - s2.x; c.y = s0.y - s2.y;          // X2
  b.x = s1.x + s3.y; b.y = s1.y - s3.x;          // X1 = s1 - i*s3
  d.x = s1.x - s3.y; d.y = s1.y + s3.x;          // X3 = s1 + i*s3
}

template <typename T, int R>
struct Dft;

template <typename T>
struct Dft<T, 2> {
  __device__ __forceinline__ static void run(cx<T> *v) { dft2(v[0], v[1]); }
};
template <typename T>
struct Dft<T, 4> {
  __device__ __forceinline__ static void run(cx<T> *v) { dft4(v[0], v[1], v[2], v[3]); }
};
template <typename T>
struct Dft<T, 8> {
  // r = 2*r1 + r0 (r0 in 0..1, r1 in 0..3), q = q0 + 4*q1 (q0 in 0..3, q1 in 0..1)
  __device__ __forceinline__ static void run(cx<T> *v) {
    // step 1: for each r0, 4-point DFT over r1 of v[2*r1 + r0]  -> slot r0 + 2*q0
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    // step 2: twiddle A_{r0=1}[q0] by w8^{q0}
    const T h = (T)0.70710678118654752440;
    { cx<T> t = v[3]; v[3].x = (t.x + t.y) * h; v[3].y = (t.y - t.x) * h; }     // w8^1 = (1 - i)/sqrt2
    { cx<T> t = v[5]; v[5].x = t.y; v[5].y = -t.x; }                              // w8^2 = -i
    { cx<T> t = v[7]; v[7].x = (t.y - t.x) * h; v[7].y = -(t.x + t.y) * h; }      // w8^3 = (-1 - i)/sqrt2
    // step 3: for each q0, 2-point DFT over r0 of slots (2*q0, 2*q0+1) -> V[q0 + 4*q1] in slot 2*q0 + q1
    dft2(v[0], v[1]);
    dft2(v[2], v[3]);
    dft2(v[4], v[5]);
    dft2(v[6], v[7]);
  }
};
template <typename T>
struct Dft<T, 16> {
  __device__ __forceinline__ static void run(cx<T> *v) {
    // step 1: for each r0 in 0..3, DFT4 over r1 of v[4*r1 + r0] -> slot r0 + 4*q0
    dft4(v[0], v[4], v[8], v[12]);
    dft4(v[1], v[5], v[9], v[13]);
    dft4(v[2], v[6], v[10], v[14]);
    dft4(v[3], v[7], v[11], v[15]);
    // step 2: slot r0 + 4*q0 *= w16^{r0*q0}
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;   // cos, sin(pi/8)
    const T h = (T)0.70710678118654752440;
    auto mul = [](cx<T> &z, T wr, T wi) { cx<T> t = z; z.x = t.x * wr - t.y * wi; z.y = t.x * wi + t.y * wr; };
    // q0 = 1: r0 = 1,2,3 -> w16^1, w16^2, w16^3
    mul(v[5], c1, -s1);
    { cx<T> t = v[6]; v[6].x = (t.x + t.y) * h; v[6].y = (t.y - t.x) * h; }
    mul(v[7], s1, -c1);
    // q0 = 2: w16^2, w16^4, w16^6
    { cx<T> t = v[9]; v[9].x = (t.x + t.y) * h; v[9].y = (t.y - t.x) * h; }
    { cx<T> t = v[10]; v[10].x = t.y; v[10].y = -t.x; }
    { cx<T> t = v[11]; v[11].x = (t.y - t.x) * h; v[11].y = -(t.x + t.y) * h; }
    // q0 = 3: w16^3, w16^6, w16^9
    mul(v[13], s1, -c1);
    { cx<T> t = v[14]; v[14].x = (t.y - t.x) * h; v[14].y = -(t.x + t.y) * h; }
    mul(v[15], -c1, s1);
    // step 3: for each q0, DFT4 over r0 of slots 4*q0 .. 4*q0+3 -> V[q0 + 4*q1] in slot 4*q0 + q1
    dft4(v[0], v[1], v[2], v[3]);
    dft4(v[4], v[5], v[6], v[7]);
    dft4(v[8], v[9], v[10], v[11]);
    dft4(v[12], v[13], v[14], v[15]);
  }
};
// where output q of Dft<T, R>::run ends up: 8-point V[q0 + 4*q1] in slot 2*q0 + q1,
// 16-point V[q0 + 4*q1] in slot 4*q0 + q1
template <int R>
__device__ constexpr int slot_of(int q) {
  return R == 8 ? (2 * (q % 4) + q / 4) : (R == 16 ? (4 * (q % 4) + q / 4) : q);
}

__device__ __forceinline__ int sign_mod_dev(const SignModDev &m, int idx) {
  const int g = idx + m.start;
  if (g >= m.half) return 1;
  return ((g & 1) ? -1 : 1) * m.extra;
}

constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x / 2); }

// padded shared-memory index: one extra element every 16 keeps the radix-16 scatter conflict-free
__device__ __forceinline__ int phys(int idx) { return idx + (idx >> 4); }

template <int N>
constexpr int line_pitch(int skew) { return N + (N >> 4) + skew; }

// One Stockham pass on the E register-resident points of a line.
//  R : radix, NS : product of the radices of earlier passes, LAST : outputs stay in registers.
//  twp : this pass' twiddles laid out [r-1][k] (k < NS), so that neighbouring threads read
//        neighbouring entries: w_N^{k * r * N/(NS*R)}
template <typename T, int N, int E, int R, int NS, bool LAST>
__device__ __forceinline__ void pass(cx<T> *x, int t, const cx<T> *__restrict__ twp, cx<T> *line_smem) {
  constexpr int THREADS = N / E;
  constexpr int B = E / R;              // butterflies per thread
#pragma unroll
  for (int b = 0; b < B; b++) {
    const int j = t + b * THREADS;
    const int k = j & (NS - 1);
    cx<T> v[R];
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = x[b + r * B];
    if (NS > 1) {
      // twiddles come from shared memory, four at a time (keeps register pressure down)
#pragma unroll
      for (int r0 = 1; r0 < R; r0 += 4) {
        cx<T> w[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (r0 + i < R) w[i] = twp[(r0 + i - 1) * NS + k];
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (r0 + i < R) v[r0 + i] = cmul(v[r0 + i], w[i]);
      }
    }
    Dft<T, R>::run(v);
    if (LAST) {
#pragma unroll
      for (int q = 0; q < R; q++) x[b + q * B] = v[slot_of<R>(q)];
    } else {
      const int base = ((j - k) * R) + k;   // (j / NS) * NS * R + k
#pragma unroll
      for (int q = 0; q < R; q++) line_smem[phys(base + q * NS)] = v[slot_of<R>(q)];
    }
  }
}

template <int N, int E>
struct Passes {
  static constexpr int R1 = E;
  static constexpr int REM1 = N / E;
  static constexpr int R2 = REM1 >= E ? E : REM1;
  static constexpr int REM2 = REM1 / R2;
  static constexpr int R3 = REM2 >= E ? E : REM2;
  static constexpr int REM3 = REM2 / R3;
  static_assert(REM3 == 1, "at most three passes");
  static constexpr int NPASS = R3 > 1 ? 3 : 2;
};

// streaming 16-byte / 8-byte global load that does not allocate in L1 (the twiddle tables live there)
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_stream(const float2 *p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

// cp.async of BYTES (8 or 16) global -> shared; `ok == false` writes zeros instead (src-size 0)
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(unsigned dst_smem, const void *src, bool ok) {
  const int n = ok ? BYTES : 0;
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
}

// FAST: no pruning, no index-shift modulation, chunk boundaries aligned with the thread
// distribution -> every address is  base(thread, tile) + table[e]  with the table in the
// constant bank.  Otherwise the general (integer-division) addressing is used.
template <typename T, int N, int E, int MAXT, bool FAST>
__global__ void __launch_bounds__(MAXT, (MAXT <= 256 ? 2 : 1)) stage_pow2_kernel(const __grid_constant__ StageParams sp) {
  using P = Passes<N, E>;
  using V = typename C2<T>::type;
  constexpr int THREADS = N / E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *smem = reinterpret_cast<cx<T> *>(smem_raw);
  const int tl = sp.tl;
  const int skew = tl <= 8 ? 8 / tl : 0;
  const int pitch = N + (N >> 4) + skew;
  const int tid = threadIdx.x;
  // per-pass twiddle tables live in shared memory behind the exchange buffer (persistent CTA:
  // loaded once): pass 2 has (R2-1)*R1 entries, pass 3 has (R3-1)*R1*R2
  constexpr int TW2 = (P::R2 - 1) * P::R1;
  constexpr int TW3 = P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0;
  cx<T> *tw_s = smem + (size_t)tl * pitch;
  {
    const cx<T> *g2 = reinterpret_cast<const cx<T> *>(sp.tw2);
    for (int i = tid; i < TW2 + TW3; i += blockDim.x) tw_s[i] = g2[i];   // the two tables are contiguous
  }
  const cx<T> *tw2 = tw_s;
  const cx<T> *tw3 = tw_s + TW2;
  __syncthreads();
  // load mapping: along the line when the input is line-contiguous, else tile-minor
  const bool in_line_major = sp.istride == 1;
  const bool out_line_major = sp.ostride == 1;
  const int t_in = in_line_major ? tid % THREADS : tid / tl;
  const int tt_in = in_line_major ? tid / THREADS : tid % tl;
  const int t_out = out_line_major ? tid % THREADS : tid / tl;
  const int tt_out = out_line_major ? tid / THREADS : tid % tl;
  const bool backward = sp.sign > 0;

  // tile bookkeeping in 32 bits (the host refuses the fast kernels beyond 2^31 tiles)
  const unsigned ntiles = (unsigned)sp.ntiles;
  auto locate = [&](unsigned tile, long long &ibase, long long &obase, int &tvalid) {
    unsigned rest = tile;
    ibase = 0;
    obase = 0;
    tvalid = 1;
    if (sp.tile_dim >= 0) {
      const unsigned along = (unsigned)sp.tiles_along;
      const unsigned chunk = rest % along;
      rest /= along;
      const long long first = (long long)chunk * tl;
      const long long left = sp.bext[sp.tile_dim] - first;
      tvalid = left < tl ? (int)left : tl;
      ibase = first * sp.bis[sp.tile_dim];
      obase = first * sp.bos[sp.tile_dim];
    }
#pragma unroll
    for (int k = kMaxBatch - 1; k >= 0; k--) {
      if (k >= sp.nbatch || k == sp.tile_dim) continue;
      const unsigned ext = (unsigned)sp.bext[k];
      const unsigned c = rest % ext;
      rest /= ext;
      ibase += (long long)c * sp.bis[k];
      obase += (long long)c * sp.bos[k];
    }
  };
  const long long t_is = sp.tile_dim >= 0 ? sp.bis[sp.tile_dim] : 0;
  const long long t_os = sp.tile_dim >= 0 ? sp.bos[sp.tile_dim] : 0;
  cx<T> *const my_line_in = smem + tt_in * pitch;
  const unsigned my_line_in_s = (unsigned)__cvta_generic_to_shared(my_line_in);

  // Asynchronous fetch (cp.async, L2 only) of a tile's inputs into the exchange buffer: every
  // thread fetches exactly the E points it will consume, so only cp.async.wait_all is needed
  // before reading them back.  Missing points (padding, lines beyond a ragged tile) are zero-filled.
  auto prefetch = [&](long long ibase, int tvalid) {
    const bool live = tt_in < tvalid;
    if (FAST) {
      const cx<T> *in = reinterpret_cast<const cx<T> *>(sp.in) + ibase + (long long)tt_in * t_is + (long long)t_in * sp.istride;
#pragma unroll
      for (int e = 0; e < E; e++)
        cp_async_zfill<sizeof(cx<T>)>(my_line_in_s + (unsigned)(phys(t_in + e * THREADS) * sizeof(cx<T>)),
                                      live ? (const void *)(in + sp.in_off[e]) : sp.in, live);
    } else {
      const cx<T> *in = reinterpret_cast<const cx<T> *>(sp.in) + ibase + (long long)tt_in * t_is;
      const bool seg_in = sp.iseg_stride != 0;
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int j = t_in + e * THREADS - sp.zin;   // position in the input line
        const bool ok = live && j >= 0 && j < sp.nin;
        long long off = 0;
        if (ok) {
          if (seg_in) {
            const int seg = j / sp.iblk;
            off = (long long)seg * sp.iseg_stride + (long long)(j - seg * sp.iblk) * sp.istride;
          } else {
            off = (long long)j * sp.istride;
          }
        }
        cp_async_zfill<sizeof(cx<T>)>(my_line_in_s + (unsigned)(phys(t_in + e * THREADS) * sizeof(cx<T>)),
                                      ok ? (const void *)(in + off) : sp.in, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  long long ibase, obase;
  int tvalid;
  unsigned tile = blockIdx.x;
  if (tile < ntiles) {
    locate(tile, ibase, obase, tvalid);
    prefetch(ibase, tvalid);
  }
  for (; tile < ntiles; tile += gridDim.x) {
    // ---- my E points: wait for the asynchronous fetch, pick them up from shared memory
    cx<T> x[E];
    asm volatile("cp.async.wait_all;" ::: "memory");
    {
      const bool cj = sp.conj_in != 0;
#pragma unroll
      for (int e = 0; e < E; e++) {
        const cx<T> r = my_line_in[phys(t_in + e * THREADS)];
        T re = r.x, im = cj ? -r.y : r.y;
        if (!FAST && sp.mod_in.on) {
          if (sign_mod_dev(sp.mod_in, t_in + e * THREADS - sp.zin) < 0) { re = -re; im = -im; }
        }
        x[e].x = backward ? im : re;
        x[e].y = backward ? re : im;
      }
    }
    __syncthreads();   // all inputs are in registers: the buffer may be overwritten by pass 1

    // ---- passes
    cx<T> *line_w = my_line_in;    // exchange buffers are addressed per line
    pass<T, N, E, P::R1, 1, false>(x, t_in, nullptr, line_w);
    __syncthreads();
    if (P::NPASS == 3) {
      {
        const cx<T> *line_r = my_line_in;
#pragma unroll
        for (int e = 0; e < E; e++) x[e] = line_r[phys(t_in + e * THREADS)];
      }
      __syncthreads();
      pass<T, N, E, P::R2, P::R1, false>(x, t_in, tw2, line_w);
      __syncthreads();
    }
    {
      const cx<T> *line_r = smem + tt_out * pitch;
#pragma unroll
      for (int e = 0; e < E; e++) x[e] = line_r[phys(t_out + e * THREADS)];
    }
    __syncthreads();   // exchange buffer is free again: start fetching the next tile behind the last pass
    const long long obase_cur = obase;
    const int tvalid_cur = tvalid;
    {
      const unsigned next = tile + gridDim.x;
      if (next < ntiles) {
        locate(next, ibase, obase, tvalid);
        prefetch(ibase, tvalid);
      }
    }
    if (P::NPASS == 3) pass<T, N, E, (P::R3 > 1 ? P::R3 : 2), P::R1 * P::R2, true>(x, t_out, tw3, nullptr);
    else pass<T, N, E, P::R2, P::R1, true>(x, t_out, tw2, nullptr);

    // ---- store the kept outputs
    if (tt_out < tvalid_cur) {
      const bool cj = sp.conj_out != 0;
      if (FAST) {
        const long long thread_off = obase_cur + (long long)tt_out * t_os + (long long)t_out * sp.ostride;
#pragma unroll
        for (int e = 0; e < E; e++) {
          V raw;
          raw.x = backward ? x[e].y : x[e].x;
          raw.y = backward ? x[e].x : x[e].y;
          if (cj) raw.y = -raw.y;
          cx<T> *out = reinterpret_cast<cx<T> *>(sp.out[sp.out_seg[e]]) + thread_off + sp.out_off[e];
          *reinterpret_cast<V *>(out) = raw;
        }
      } else {
        const bool seg_out = sp.noseg > 1;
#pragma unroll
        for (int e = 0; e < E; e++) {
          const int kk = t_out + e * THREADS - sp.zout;
          if (kk < 0 || kk >= sp.nout) continue;
          V raw;
          raw.x = backward ? x[e].y : x[e].x;
          raw.y = backward ? x[e].x : x[e].y;
          if (sp.mod_out.on && sign_mod_dev(sp.mod_out, kk) < 0) { raw.x = -raw.x; raw.y = -raw.y; }
          if (cj) raw.y = -raw.y;
          int seg = 0, kl = kk;
          if (seg_out) {
            seg = kk / sp.oblk;
            kl = kk - seg * sp.oblk;
          }
          cx<T> *out = reinterpret_cast<cx<T> *>(sp.out[seg]) + obase_cur + (long long)tt_out * t_os + (long long)kl * sp.ostride;
          *reinterpret_cast<V *>(out) = raw;
        }
      }
    }
  }
}

// (N, E) pairs compiled in; E = points per thread.  Two block-size classes per pair:
// <= 256 threads (two CTAs per SM) and <= 512 threads (one CTA per SM, twice the lines per
// tile, i.e. twice as long contiguous runs on a strided side).
template <typename T, int N, int E, int MAXT>
cudaError_t launch_block_class(StageParams &sp, cudaStream_t stream) {
  constexpr int THREADS = N / E;
  const int tl = sp.tl;
  const int skew = tl <= 8 ? 8 / tl : 0;
  using P = Passes<N, E>;
  constexpr int TWN = (P::R2 - 1) * P::R1 + (P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0);
  const size_t smem = ((size_t)tl * (N + (N >> 4) + skew) + TWN) * 2 * sizeof(T);
  auto fast = stage_pow2_kernel<T, N, E, MAXT, true>;
  auto slow = stage_pow2_kernel<T, N, E, MAXT, false>;
  auto kern = sp.fast ? fast : slow;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(slow, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tl * THREADS, smem);
  if (per_sm < 1) per_sm = 1;
  const long long grid = std::min<long long>(sp.ntiles, (long long)sms * per_sm);
  kern<<<(unsigned)grid, tl * THREADS, smem, stream>>>(sp);
  launch_counter()++;
  return cudaGetLastError();
}

template <typename T, int N, int E>
cudaError_t launch_variant(StageParams &sp, cudaStream_t stream) {
  constexpr int THREADS = N / E;
  if (sp.tl * THREADS <= 256) return launch_block_class<T, N, E, (THREADS > 256 ? THREADS : 256)>(sp, stream);
  return launch_block_class<T, N, E, (THREADS > 512 ? THREADS : 512)>(sp, stream);
}

constexpr int points_per_thread(int n) {
  // three passes at most: 64 -> 8*8, 128 -> 8*8*2, 256 -> 16*16, 512 -> 8*8*8, 1024 -> 16*16*4, ...
  return (n == 64 || n == 128 || n == 512) ? 8 : 16;
}

}  // namespace

template <typename T>
bool pow2_supported(const Stage &g, int L) {
  if (g.op != OP_C2C) return false;
  if (L < 64 || L > 4096 || (L & (L - 1))) return false;
  if (g.in_real || g.out_real) return false;
  if (g.noseg > kMaxSeg) return false;
  long long lines = 1;
  for (int k = 0; k < g.nbatch; k++) {
    if (g.batch[k].extent >= (1ll << 31)) return false;
    lines *= g.batch[k].extent;
  }
  if (lines >= (1ll << 31)) return false;
  return true;
}

template <typename T>
int pow2_pick_tile(const Stage &g, int L) {
  // Lines per tile.  Default: fill a 256-thread block (fp64) / 512-thread block (fp32): two
  // resident CTAs per SM in fp64, and a strided side then moves tl * sizeof(complex) = 64
  // contiguous bytes per point at N = 1024.  PFFT_B200_TL overrides (experiments).
  const int threads = L / points_per_thread(L);
  int block = sizeof(T) == 8 ? 256 : 512;
  static const int forced = [] {
    const char *e = getenv("PFFT_B200_TL");
    return e ? atoi(e) : 0;
  }();
  const bool strided = g.istride != 1 || g.ostride != 1;
  // measured on B200 (1024^3 fp64): 128-byte runs on the strided side beat two resident CTAs
  // (8.4 vs 10.5 ms per pass; 6.8 vs 10.5 ms when half of the stores cross NVLink)
  if (strided && sizeof(T) == 8) block = 512;
  int tl = block / threads;
  if (forced > 0 && strided) tl = forced;
  while (tl * threads > 512 && tl > 1) tl /= 2;
  if (tl < 1) tl = 1;
  if (g.tile_dim < 0) tl = 1;
  else if ((INT)tl > g.batch[g.tile_dim].extent) {
    tl = 1;
    while (tl * 2 <= g.batch[g.tile_dim].extent) tl *= 2;
  }
  return tl;
}

template <typename T>
cudaError_t launch_stage_pow2(StageParams &sp, cudaStream_t stream) {
  if (sp.ntiles <= 0) return cudaSuccess;
  switch (sp.L) {
    case 64: return launch_variant<T, 64, 8>(sp, stream);
    case 128: return launch_variant<T, 128, 8>(sp, stream);
    case 256: return launch_variant<T, 256, 16>(sp, stream);
    case 512: return launch_variant<T, 512, 8>(sp, stream);
    case 1024: return launch_variant<T, 1024, 16>(sp, stream);
    case 2048: return launch_variant<T, 2048, 16>(sp, stream);
    case 4096: return launch_variant<T, 4096, 16>(sp, stream);
    default: return cudaErrorInvalidValue;
  }
}

// Decide whether the table-driven addressing applies and fill the tables.
template <typename T>
void pow2_prepare(const Stage &g, StageParams &sp) {
  const int L = sp.L;
  const int E = points_per_thread(L);
  const int threads = L / E;
  bool fast = g.nin == L && g.zin == 0 && g.nout == L && g.zout == 0 && !g.mod_in.on && !g.mod_out.on;
  const bool seg_in = g.iseg_stride != 0 && g.iblk < g.nin;
  const bool seg_out = g.noseg > 1;
  if (seg_in && g.iblk % threads != 0) fast = false;
  if (seg_out && g.oblk % threads != 0) fast = false;
  sp.fast = fast ? 1 : 0;
  if (!fast) return;
  for (int e = 0; e < E; e++) {
    const long long idx = (long long)e * threads;
    if (seg_in) {
      const long long seg = idx / g.iblk;
      sp.in_off[e] = seg * g.iseg_stride + (idx - seg * g.iblk) * g.istride;
    } else {
      sp.in_off[e] = idx * g.istride;
    }
    if (seg_out) {
      const long long seg = idx / g.oblk;
      sp.out_seg[e] = (int)seg;
      sp.out_off[e] = (idx - seg * g.oblk) * g.ostride;
    } else {
      sp.out_seg[e] = 0;
      sp.out_off[e] = idx * g.ostride;
    }
  }
}

// Per-pass twiddle tables [r-1][k]: returns element offsets of the pass-2 / pass-3 tables
// inside `table` (complex entries, forward sign).
void pow2_twiddle_tables(int L, const double *roots /* 2*L */, std::vector<double> *table, size_t *off2, size_t *off3) {
  const int E = points_per_thread(L);
  int rem = L / E;
  const int R1 = E;
  const int R2 = rem >= E ? E : rem;
  rem /= R2;
  const int R3 = rem;
  table->clear();
  auto emit = [&](int NS, int R) {
    const int tstep = L / (NS * R);
    for (int r = 1; r < R; r++)
      for (int k = 0; k < NS; k++) {
        const long long idx = (long long)k * r * tstep;
        table->push_back(roots[2 * idx]);
        table->push_back(roots[2 * idx + 1]);
      }
  };
  *off2 = 0;
  emit(R1, R2);
  *off3 = table->size() / 2;
  if (R3 > 1) emit(R1 * R2, R3);
}

template bool pow2_supported<float>(const Stage &, int);
template bool pow2_supported<double>(const Stage &, int);
template int pow2_pick_tile<float>(const Stage &, int);
template int pow2_pick_tile<double>(const Stage &, int);
template cudaError_t launch_stage_pow2<float>(StageParams &, cudaStream_t);
template cudaError_t launch_stage_pow2<double>(StageParams &, cudaStream_t);
template void pow2_prepare<float>(const Stage &, StageParams &);
template void pow2_prepare<double>(const Stage &, StageParams &);

}  // namespace pfb
