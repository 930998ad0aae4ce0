// fft_regs.h -- device primitives of the register-resident FFT kernels (fft_pow2.cu, fft_reg.cu):
// complex arithmetic, radix-2/4/8/16 butterflies on registers, the Stockham pass with its padded
// shared-memory exchange, the radix plan of a line length, and the cp.async helpers.
// A line of N points is spread over N / E threads, E points per thread, in the "strided" distribution
//     thread t holds indices  t, t + N/E, t + 2N/E, ...
// which every pass preserves (see fft_pow2.cu).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

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

// phys(i + e * STEP) given i (and pi = phys(i)): for steps that are multiples of 16 the padding
// advances by STEP / 16 per step, so the address is pi + a compile-time constant
template <int STEP>
__device__ __forceinline__ int phys_at(int i, int pi, int e) {
  if (STEP % 16 == 0) return pi + e * (STEP + STEP / 16);
  return phys(i + e * STEP);
}

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
      if (NS % 16 == 0) {
        const int pb = phys(base);
#pragma unroll
        for (int q = 0; q < R; q++) line_smem[pb + q * (NS + NS / 16)] = v[slot_of<R>(q)];
      } else if (NS == 1 && R == 16) {
        const int pb = 17 * j;                // phys(16 j + q) = 16 j + q + j for q < 16
#pragma unroll
        for (int q = 0; q < R; q++) line_smem[pb + q] = v[slot_of<R>(q)];
      } else {
#pragma unroll
        for (int q = 0; q < R; q++) line_smem[phys(base + q * NS)] = v[slot_of<R>(q)];
      }
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
  else if (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
}

template <int BYTES>
__device__ __forceinline__ void cp_async_plain(unsigned dst_smem, const void *src) {
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}

}  // namespace

}  // namespace pfb
