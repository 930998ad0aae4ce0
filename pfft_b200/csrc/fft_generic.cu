// fft_generic.cu -- any-length stage kernel (sm_100a).
//
// One CTA processes a tile of `tl` lines of the transformed dimension: it gathers
// them from global memory (straight out of the user's array or out of the chunks
// an exchange delivered), zero-pads / sign-modulates / conjugates on the fly,
// runs a mixed-radix Stockham FFT between two shared-memory buffers, and scatters
// the kept outputs to their final place (next stage's layout, or per-destination
// chunk of the next exchange, or the user's output array).  This replaces, per
// stage, the reference's embed loop + FFTW plan + FFTW copy plan + truncate loop
// (kernel/outrafo.c:154-168, kernel/sertrafo.c:1073-1245, kernel/ousample.c:391-506)
// with one read and one write of the array.
//
// The butterflies here are "one output per thread": out[q] = sum_r in[r] * w^(r*e),
// which handles any radix (29, 31, ...) with exact table twiddles.  Power-of-two
// complex lines take the register-resident fast path in fft_pow2.cu instead.
#include <cuda_runtime.h>

#include "kernels.h"

namespace pfb {

template <typename T> struct Cplx;
template <> struct Cplx<double> { using type = double2; };
template <> struct Cplx<float> { using type = float2; };

template <typename T>
__device__ __forceinline__ typename Cplx<T>::type mk(T a, T b) {
  typename Cplx<T>::type r;
  r.x = a;
  r.y = b;
  return r;
}

__device__ __forceinline__ int sign_mod(const SignModDev &m, int idx) {
  // reference api/api-basic.c:1213-1234: (-1)^g below the half, with g = idx + start
  const int g = idx + m.start;
  if (g >= m.half) return 1;
  return ((g & 1) ? -1 : 1) * m.extra;
}

template <typename T>
__global__ void __launch_bounds__(256) stage_generic_kernel(const __grid_constant__ StageParams sp) {
  using C = typename Cplx<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = sp.L;
  const int tl = sp.tl;
  C *buf0 = reinterpret_cast<C *>(smem_raw);
  C *buf1 = buf0 + (size_t)tl * L;
  const C *tw = reinterpret_cast<const C *>(sp.twiddle);
  const int tid = threadIdx.x, nthr = blockDim.x;

  for (long long tile = blockIdx.x; tile < sp.ntiles; tile += gridDim.x) {
    // ---- locate the tile
    long long rest = tile;
    long long ibase = 0, obase = 0;
    int tvalid = 1;
    long long t_is = 0, t_os = 0;
    if (sp.tile_dim >= 0) {
      const long long chunk = rest % sp.tiles_along;
      rest /= sp.tiles_along;
      const long long first = chunk * tl;
      const long long left = sp.bext[sp.tile_dim] - first;
      tvalid = left < tl ? (int)left : tl;
      t_is = sp.bis[sp.tile_dim];
      t_os = sp.bos[sp.tile_dim];
      ibase = first * t_is;
      obase = first * t_os;
    }
    for (int k = sp.nbatch - 1; k >= 0; k--) {
      if (k == sp.tile_dim) continue;
      const long long c = rest % sp.bext[k];
      rest /= sp.bext[k];
      ibase += c * sp.bis[k];
      obase += c * sp.bos[k];
    }

    // ---- load (zero-padded) lines into buf0
    const int lines_elems = tvalid * L;
    for (int e = tid; e < lines_elems; e += nthr) buf0[e] = mk<T>(0, 0);
    __syncthreads();
    {
      const int total = tvalid * sp.nin;
      const bool contiguous = sp.istride == 1;
      for (int e = tid; e < total; e += nthr) {
        int tt, j;
        if (contiguous) { tt = e / sp.nin; j = e - tt * sp.nin; }
        else { j = e / tvalid; tt = e - j * tvalid; }
        const int seg = j / sp.iblk;
        const long long off = ibase + (long long)tt * t_is + (long long)seg * sp.iseg_stride +
                              (long long)(j - seg * sp.iblk) * sp.istride;
        C v;
        if (sp.in_real) {
          v.x = reinterpret_cast<const T *>(sp.in)[off];
          v.y = 0;
        } else {
          v = reinterpret_cast<const C *>(sp.in)[off];
        }
        if (sp.conj_in) v.y = -v.y;
        if (sp.mod_in.on && sign_mod(sp.mod_in, j) < 0) { v.x = -v.x; v.y = -v.y; }
        if (sp.op == OP_R2R) {
          // w_j x_j exp(-i pi b jj / D), jj = position inside the logical line of n reals
          const int jj = j + sp.zin;
          T xr = v.x;
          if ((jj == 0 && sp.r2r_half0) || (jj == sp.n - 1 && sp.r2r_halfn)) xr *= (T)0.5;
          const C w = reinterpret_cast<const C *>(sp.tw_r2r)[(int)(((long long)2 * sp.r2r_b2 * jj) % (8ll * sp.r2r_D))];
          v.x = xr * w.x;
          v.y = xr * w.y;
        }
        buf0[tt * L + j + sp.zin] = v;
      }
    }
    __syncthreads();
    if (sp.op == OP_C2R) {
      // Hermitian completion of the half spectrum: X[n-k] = conj(X[k]); imaginary parts of
      // the DC / Nyquist bins only reach the (discarded) imaginary output.
      const int half = (L - 1) / 2;
      for (int e = tid; e < tvalid * half; e += nthr) {
        const int tt = e / half, k = e - tt * half + 1;
        C v = buf0[tt * L + k];
        v.y = -v.y;
        buf0[tt * L + L - k] = v;
      }
      __syncthreads();
    }

    // ---- mixed-radix Stockham passes, one output per thread
    C *src = buf0, *dst = buf1;
    if (sp.op != OP_COPY) {
      const bool backward = sp.op != OP_R2R && sp.sign > 0;
      int Ns = 1;
      for (int f = 0; f < sp.nfac; f++) {
        const int R = sp.fac[f];
        if (R == 1) continue;
        const int LR = L / R;           // butterflies per line
        const int tstep = L / (Ns * R);  // table step of this pass' twiddle
        for (int e = tid; e < lines_elems; e += nthr) {
          const int tt = e / L;
          const int o = e - tt * L;     // output slot: q-th output of butterfly jj
          const int jj = o % LR, q = o / LR;
          const int k = jj % Ns;
          // exponent step: twiddle w^(r*k*tstep) times butterfly root W_R^(q*r) = w^(r*q*LR)
          int estep = k * tstep + q * LR;
          estep %= L;
          const C *line = src + tt * L + jj;
          T ar = 0, ai = 0;
          int ex = 0;
          for (int r = 0; r < R; r++) {
            const C x = line[r * LR];
            C w = tw[ex];
            if (backward) w.y = -w.y;
            ar += x.x * w.x - x.y * w.y;
            ai += x.x * w.y + x.y * w.x;
            ex += estep;
            if (ex >= L) ex -= L;
          }
          const int base = (jj / Ns) * Ns * R + k;
          dst[tt * L + base + q * Ns] = mk<T>(ar, ai);
        }
        __syncthreads();
        C *t = src; src = dst; dst = t;
        Ns *= R;
      }
    }

    // ---- store the kept outputs
    {
      const int total = tvalid * sp.nout;
      const bool contiguous = sp.ostride == 1;
      for (int e = tid; e < total; e += nthr) {
        int tt, kk;
        if (contiguous) { tt = e / sp.nout; kk = e - tt * sp.nout; }
        else { kk = e / tvalid; tt = e - kk * tvalid; }
        C v = src[tt * L + kk + sp.zout];
        if (sp.op == OP_R2R) {
          // 2 F(exp(-i pi a (k + b) / D) X_k)
          const int k = kk + sp.zout;
          const C w = reinterpret_cast<const C *>(sp.tw_r2r)[(int)(((long long)sp.r2r_a2 * (2 * k + sp.r2r_b2)) % (8ll * sp.r2r_D))];
          const T re = v.x * w.x - v.y * w.y, im = v.x * w.y + v.y * w.x;
          v.x = sp.r2r_sine ? (T)-2 * im : (T)2 * re;
          v.y = 0;
        }
        if (sp.mod_out.on && sign_mod(sp.mod_out, kk) < 0) { v.x = -v.x; v.y = -v.y; }
        if (sp.conj_out) v.y = -v.y;
        const int seg = kk / sp.oblk;
        const long long off = obase + (long long)tt * t_os + (long long)(kk - seg * sp.oblk) * sp.ostride;
        if (sp.out_real) reinterpret_cast<T *>(sp.out[seg])[off] = v.x;
        else reinterpret_cast<C *>(sp.out[seg])[off] = v;
      }
    }
    __syncthreads();
  }
}

template <typename T>
int generic_pick_tile(const Stage &g, int L) {
  const size_t csize = 2 * sizeof(T);
  const size_t budget = 200 * 1024;
  size_t per_line = 2 * (size_t)L * csize;
  if (per_line > budget) return 0;
  int tl = (int)(96 * 1024 / per_line);
  if (tl < 1) tl = 1;
  // strided sides want at least 128 contiguous bytes per access
  const int want = (int)(128 / csize);
  const bool strided = g.istride != 1 || g.ostride != 1;
  if (strided && tl < want) tl = (int)std::min<size_t>(want, budget / per_line);
  if (tl > 16) tl = 16;
  if (g.tile_dim < 0) tl = 1;
  else if ((INT)tl > g.batch[g.tile_dim].extent) tl = (int)g.batch[g.tile_dim].extent;
  if (tl < 1) tl = 1;
  return tl;
}

template <typename T>
cudaError_t launch_stage_generic(StageParams &sp, cudaStream_t stream) {
  if (sp.ntiles <= 0) return cudaSuccess;
  const size_t smem = 2 * (size_t)sp.tl * sp.L * 2 * sizeof(T);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stage_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024)));
  long long grid = std::min<long long>(sp.ntiles, (long long)sms * per_sm);
  stage_generic_kernel<T><<<(unsigned)grid, 256, smem, stream>>>(sp);
  launch_counter()++;
  return cudaGetLastError();
}

template int generic_pick_tile<float>(const Stage &, int);
template int generic_pick_tile<double>(const Stage &, int);
template cudaError_t launch_stage_generic<float>(StageParams &, cudaStream_t);
template cudaError_t launch_stage_generic<double>(StageParams &, cudaStream_t);

}  // namespace pfb
