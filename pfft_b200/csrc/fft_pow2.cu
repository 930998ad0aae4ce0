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

#include <algorithm>
#include <type_traits>
#include <vector>

#include "fft_regs.h"
#include "kernels.h"

namespace pfb {

namespace {

// FAST: no pruning, no index-shift modulation, chunk boundaries aligned with the thread
// distribution -> every address is  base(thread, tile) + table[e]  with the table in the
// constant bank.  Otherwise the general (integer-division) addressing is used.
// DIR: 1 forward, 2 backward (the re/im swap is then pure register naming), 0 decided at run time.
// INLM: the input is known to be line-contiguous (load mapping fixed at compile time).
template <typename T, int N, int E, int MAXT, bool FAST, int DIR, bool INLM>
__global__ void __launch_bounds__(MAXT, (MAXT <= 256 ? 2 : 1)) stage_pow2_kernel(const __grid_constant__ StageParams sp) {
  using P = Passes<N, E>;
  using V = typename C2<T>::type;
  constexpr int THREADS = N / E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *smem = reinterpret_cast<cx<T> *>(smem_raw);
  const int tl = sp.tl;
  const int skew = tl >= 8 ? 1 : 8 / tl;
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
  const bool in_line_major = INLM ? true : sp.istride == 1;
  const bool out_line_major = sp.ostride == 1;
  const int t_in = in_line_major ? tid % THREADS : tid / tl;
  const int tt_in = in_line_major ? tid / THREADS : tid % tl;
  const int t_out = out_line_major ? tid % THREADS : tid / tl;
  const int tt_out = out_line_major ? tid / THREADS : tid % tl;
  const bool backward = DIR == 0 ? sp.sign > 0 : DIR == 2;
  const bool c2r_line = !FAST && sp.op == OP_C2R;   // Hermitian half spectrum in, real line out
  // Barriers between passes: while a line is handled by whole warps of its own (line-major
  // mapping), only those warps need to meet -- named barrier 1 + line -- so the lines of a tile
  // drift apart and their load / butterfly / exchange phases overlap.  Where the store mapping is
  // tile-minor (strided side), the last exchange mixes lines and needs the CTA-wide barrier.
  // (tiles of more than 15 lines -- 16-line fp32 tiles -- share a named barrier between `bg` neighbouring lines)
  int bg = 1;
  while (tl / bg > 15) bg *= 2;
  const bool lbar_in = sp.line_bars && in_line_major && (THREADS % 32 == 0) && tl % bg == 0;
  const bool lbar_out = lbar_in && out_line_major && sp.line_bars != 2;   // 2: experiment (CTA-wide around the last exchange)
  const int bar_id = 1 + tt_in / bg, bar_cnt = bg * THREADS;
  auto sync_lines = [&](bool per_line) {
    if (per_line) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_cnt) : "memory");
    else __syncthreads();
  };

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
      // (lines beyond a ragged tile are never stored: whatever their buffers hold is fine)
      const cx<T> *in = reinterpret_cast<const cx<T> *>(sp.in) + ibase + (long long)tt_in * t_is + (long long)t_in * sp.istride;
      if (live) {
#pragma unroll
        for (int e = 0; e < E; e++)
          cp_async_plain<sizeof(cx<T>)>(my_line_in_s + (unsigned)(phys(t_in + e * THREADS) * sizeof(cx<T>)), in + sp.in_off[e]);
      }
    } else {
      const long long line0 = ibase + (long long)tt_in * t_is;
      const bool seg_in = sp.iseg_stride != 0;
#pragma unroll
      for (int e = 0; e < E; e++) {
        int j = t_in + e * THREADS - sp.zin;   // position in the input line
        // c2r: the upper half of the spectrum is the mirrored, conjugated lower half (conjugation on pick-up)
        if (c2r_line && j > N / 2) j = N - j;
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
        const unsigned dst = my_line_in_s + (unsigned)(phys(t_in + e * THREADS) * sizeof(cx<T>));
        if (sp.in_real)      // r2c: real parts only (strides count reals); the imaginary half of the slot is ignored
          cp_async_zfill<sizeof(T)>(dst, ok ? (const void *)(reinterpret_cast<const T *>(sp.in) + line0 + off) : sp.in, ok);
        else
          cp_async_zfill<sizeof(cx<T>)>(dst, ok ? (const void *)(reinterpret_cast<const cx<T> *>(sp.in) + line0 + off) : sp.in, ok);
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
      const bool cj = !FAST && sp.conj_in != 0;   // (the host keeps conjugating stages off the FAST path)
#pragma unroll
      for (int e = 0; e < E; e++) {
        const cx<T> r = my_line_in[phys(t_in + e * THREADS)];
        T re = r.x, im = cj ? -r.y : r.y;
        if (!FAST && sp.in_real) im = 0;
        if (c2r_line && t_in + e * THREADS - sp.zin > N / 2) im = -im;
        if (!FAST && sp.mod_in.on) {
          // the modulation belongs to the element as stored (reference api/api-basic.c twiddle_input acts on the
          // half spectrum before it is completed): a mirrored c2r bin takes the factor of its source index
          int jm = t_in + e * THREADS - sp.zin;
          if (c2r_line && jm > N / 2) jm = N - jm;
          if (sign_mod_dev(sp.mod_in, jm) < 0) { re = -re; im = -im; }
        }
        x[e].x = backward ? im : re;
        x[e].y = backward ? re : im;
      }
    }
    sync_lines(lbar_in);   // all inputs are in registers: the buffer may be overwritten by pass 1

    // ---- passes
    cx<T> *line_w = my_line_in;    // exchange buffers are addressed per line
    pass<T, N, E, P::R1, 1, false>(x, t_in, nullptr, line_w);
    if (P::NPASS == 3) {
      sync_lines(lbar_in);
      {
        const cx<T> *line_r = my_line_in;
#pragma unroll
        for (int e = 0; e < E; e++) x[e] = line_r[phys(t_in + e * THREADS)];
      }
      sync_lines(lbar_in);
      pass<T, N, E, P::R2, P::R1, false>(x, t_in, tw2, line_w);
    }
    sync_lines(lbar_out);
    {
      const cx<T> *line_r = smem + tt_out * pitch;
#pragma unroll
      for (int e = 0; e < E; e++) x[e] = line_r[phys(t_out + e * THREADS)];
    }
    sync_lines(lbar_out);   // exchange buffer is free again: start fetching the next tile behind the last pass
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
      const bool cj = !FAST && sp.conj_out != 0;
      if (FAST) {
        const long long thread_off = obase_cur + (long long)tt_out * t_os + (long long)t_out * sp.ostride;
#pragma unroll
        for (int e = 0; e < E; e++) {
          V raw;
          raw.x = backward ? x[e].y : x[e].x;
          raw.y = backward ? x[e].x : x[e].y;
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
          const long long off = obase_cur + (long long)tt_out * t_os + (long long)kl * sp.ostride;
          if (sp.out_real) reinterpret_cast<T *>(sp.out[seg])[off] = raw.x;      // c2r: strides count reals
          else *reinterpret_cast<V *>(reinterpret_cast<cx<T> *>(sp.out[seg]) + off) = raw;
        }
      }
    }
  }
}

// ---- plane-fused pair ------------------------------------------------------------------
// Two consecutive stages A, B of one rank (no exchange in between) share a batch dimension
// that neither transforms: the "plane" dimension.  Stage B's lines of plane p need exactly
// stage A's lines of plane p, so the pair is run as ONE persistent kernel that walks the
// planes in order -- tiles of A(p) followed by tiles of B(p - delta) -- and hands the
// intermediate through a small ring of plane slots that never leaves the 126 MB L2:
// A's outputs are written with an evict-last policy, B reads them back from L2, and only A's
// reads and B's writes (both contiguous lines, evict-first) touch HBM.  A pair therefore costs
// one read + one write of the local array instead of two of each.  Ordering is done with
// per-plane completion counters in global memory (release increments per warp, acquire polls);
// because every CTA takes its items in increasing order and an item only depends on items far
// behind it, the minimum unfinished item can always run: no deadlock as long as all CTAs are
// resident (grid = SMs x occupancy).
// Reference counterpart: two FFTW plans + copy plans with the whole array between them
// (kernel/partrafo-transposed.c:224-285); nothing like it exists there.
// The counters are polled with relaxed loads: everything read behind a successful poll is
// fetched with cp.async.cg, i.e. straight from L2, after the branch that depends on the poll.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill_hint(unsigned dst_smem, const void *src, bool ok, unsigned long long pol) {
  const int n = ok ? BYTES : 0;
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst_smem), "l"(src), "r"(n), "l"(pol) : "memory");
  else
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2, %3;" ::"r"(dst_smem), "l"(src), "r"(n), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(double2 *p, double2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(float2 *p, float2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
// x / d for the runtime constant d behind (m, s1, s2) (round-up method, exact for 32-bit x)
__device__ __forceinline__ unsigned fast_div(unsigned x, unsigned m, unsigned s1, unsigned s2) {
  const unsigned t = __umulhi(m, x);
  return (t + ((x - t) >> s1)) >> s2;
}

// FUSED == false: the same body runs ONE stage (spA) over a static tile sequence -- this is the
// fast path of every power-of-two stage with contiguous input lines.
template <typename T, int N, int E, int MAXT, bool BWD, bool FUSED>
__global__ void __launch_bounds__(MAXT, (MAXT <= 128 ? 4 : (MAXT <= 256 ? 2 : 1)))
fused_pair_kernel(const __grid_constant__ StageParams spA, const __grid_constant__ StageParams spB,
                  const __grid_constant__ FusePlanes fp) {
  using P = Passes<N, E>;
  using V = typename C2<T>::type;
  constexpr int THREADS = N / E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *smem = reinterpret_cast<cx<T> *>(smem_raw);
  const int tl = spA.tl;
  const int skew = tl >= 8 ? 1 : 8 / tl;
  const int pitch = N + (N >> 4) + skew;
  const int tid = threadIdx.x;
  constexpr int TW2 = (P::R2 - 1) * P::R1;
  constexpr int TW3 = P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0;
  cx<T> *tw_s = smem + (size_t)tl * pitch;
  {
    const cx<T> *g2 = reinterpret_cast<const cx<T> *>(spA.tw2);
    for (int i = tid; i < TW2 + TW3; i += blockDim.x) tw_s[i] = g2[i];
  }
  const cx<T> *tw2 = tw_s;
  const cx<T> *tw3 = tw_s + TW2;
  __syncthreads();
  unsigned long long pol_first, pol_last;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));

  // both stages read contiguous lines: load / butterfly mapping is line-major throughout
  const int t_in = tid % THREADS, tt_in = tid / THREADS;
  constexpr bool backward = BWD;
  int bg = 1;                       // lines sharing one named barrier (at most 15 groups)
  while (tl / bg > 15) bg *= 2;
  const bool per_line = (THREADS % 32 == 0) && tl % bg == 0;
  const int bar_id = 1 + tt_in / bg, bar_cnt = bg * THREADS;
  cx<T> *const my_line_in = smem + tt_in * pitch;
  const unsigned my_line_in_s = (unsigned)__cvta_generic_to_shared(my_line_in);
  const int pt_in = phys(t_in);
  auto sync_lines = [&](bool pl) {
    if (pl) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_cnt) : "memory");
    else __syncthreads();
  };

  const unsigned per_round = (unsigned)(fp.t1 + fp.t2);
  const unsigned nitems = FUSED ? per_round * (unsigned)fp.planes : (unsigned)spA.ntiles;
  const unsigned n_head = (unsigned)fp.delta * (unsigned)fp.t1;                  // rounds with first-stage tiles only
  const unsigned n_mid = (unsigned)(fp.planes - fp.delta) * per_round;
  // item -> (stage, plane, tile inside the plane).  Items are handed out in this order by a
  // ticket counter, so all CTAs advance together and what an item depends on (two rounds
  // back) is practically always complete.
  auto decode = [&](unsigned item, int &which, unsigned &w, int &plane) {
    if (!FUSED) {
      which = 0;
      plane = 0;
      w = item;
    } else if (item < n_head) {
      which = 0;
      plane = (int)(item / (unsigned)fp.t1);
      w = item - (unsigned)plane * (unsigned)fp.t1;
    } else if (item < n_head + n_mid) {
      const unsigned j = item - n_head;
      const unsigned rnd = fast_div(j, fp.div_m, fp.div_s1, fp.div_s2);
      w = j - rnd * per_round;
      if (w < (unsigned)fp.t1) {
        which = 0;
        plane = (int)rnd + fp.delta;
      } else {
        which = 1;
        w -= (unsigned)fp.t1;
        plane = (int)rnd;
      }
    } else {
      const unsigned j = item - n_head - n_mid;
      which = 1;
      const unsigned q = j / (unsigned)fp.t2;
      plane = fp.planes - fp.delta + (int)q;
      w = j - q * (unsigned)fp.t2;
    }
  };
  auto flag_of = [&](int which, int plane) -> const unsigned * {
    if (which) return fp.done + plane;
    return fp.done + fp.planes + (plane >= fp.ring ? plane - fp.ring : 0);
  };
  auto flag_ok = [&](int which, int plane, unsigned v) -> bool {
    if (!FUSED) return true;
    if (which) return v >= fp.target1;
    return plane < fp.ring || v >= fp.target2;
  };
  auto locate = [&](const StageParams &sp, int plane, unsigned w, long long &ibase, long long &obase, int &tvalid) {
    unsigned rest = w;
    ibase = 0;
    obase = 0;
    tvalid = 1;
    if (sp.tile_dim >= 0) {
      unsigned chunk = rest;
      if (!FUSED || sp.nbatch > 2) {
        const unsigned along = (unsigned)sp.tiles_along;
        const unsigned qa = fast_div(rest, sp.rg.dalong.m, sp.rg.dalong.s1, sp.rg.dalong.s2);   // (host-made magic numbers)
        chunk = rest - qa * along;
        rest = qa;
      }
      const long long first = (long long)chunk * tl;
      const long long left = sp.bext[sp.tile_dim] - first;
      tvalid = left < tl ? (int)left : tl;
      ibase = first * sp.bis[sp.tile_dim];
      obase = first * sp.bos[sp.tile_dim];
    }
    if (!FUSED || sp.nbatch > 2) {
#pragma unroll
      for (int k = kMaxBatch - 1; k >= (FUSED ? 1 : 0); k--) {
        if (k >= sp.nbatch || k == sp.tile_dim) continue;
        const unsigned ext = (unsigned)sp.bext[k];
        const unsigned qe = fast_div(rest, sp.rg.dbext[k].m, sp.rg.dbext[k].s1, sp.rg.dbext[k].s2);
        const unsigned c = rest - qe * ext;
        rest = qe;
        ibase += (long long)c * sp.bis[k];
        obase += (long long)c * sp.bos[k];
      }
    }
    if (FUSED) {
      // batch dimension 0 is the plane: the ring side wraps around
      const unsigned pi = sp.ring_in ? (unsigned)plane % (unsigned)sp.ring_in : (unsigned)plane;
      const unsigned po = sp.ring_out ? (unsigned)plane % (unsigned)sp.ring_out : (unsigned)plane;
      ibase += (long long)pi * sp.bis[0];
      obase += (long long)po * sp.bos[0];
    }
  };
  auto prefetch = [&](const StageParams &sp, int which, long long ibase, int tvalid) {
    const long long t_is = sp.tile_dim >= 0 ? sp.bis[sp.tile_dim] : 0;
    const cx<T> *in = reinterpret_cast<const cx<T> *>(sp.in) + ibase + (long long)tt_in * t_is + (long long)t_in * sp.istride;
    const unsigned long long pol = which ? pol_last : pol_first;
    if (tt_in < tvalid) {
#pragma unroll
      for (int e = 0; e < E; e++) {
        const unsigned dst = my_line_in_s + (unsigned)(phys_at<THREADS>(t_in, pt_in, e) * sizeof(cx<T>));
        if (FUSED) cp_async_zfill_hint<sizeof(cx<T>)>(dst, in + sp.in_off[e], true, pol);
        else cp_async_plain<sizeof(cx<T>)>(dst, in + sp.in_off[e]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // completion of a tile is published one tile late, when its stores have long drained and the
  // release costs nothing; `sig` < 0: nothing pending
  int sig = -1;
  auto publish = [&]() {
    if (FUSED && sig >= 0) {
      __syncwarp();
      if ((tid & 31) == 0) red_release_add(fp.done + sig, 1u);
      sig = -1;
    }
  };

  __shared__ unsigned next_item_s[2];
  int parity = 0;
  unsigned *const ticket = fp.done + 2 * fp.planes;
  int which = 0, plane = 0;
  unsigned w = 0;
  unsigned item = blockIdx.x;   // the first gridDim.x items are taken statically, tickets continue from there
  long long ibase = 0, obase = 0;
  int tvalid = 0;
  if (item < nitems) {
    decode(item, which, w, plane);
    if (FUSED)
      while (!flag_ok(which, plane, ld_relaxed_u32(flag_of(which, plane)))) __nanosleep(100);
    locate(which ? spB : spA, plane, w, ibase, obase, tvalid);
    prefetch(which ? spB : spA, which, ibase, tvalid);
  }
  while (item < nitems) {
    const StageParams &sp = (FUSED && which) ? spB : spA;
    const bool out_line_major = sp.ostride == 1;
    const int t_out = out_line_major ? t_in : tid / tl;
    const int tt_out = out_line_major ? tt_in : tid % tl;
    const bool lbar_out = per_line && out_line_major;
    // claim the next item now; the ticket travels while this tile is computed
    unsigned claimed = 0;
    if (FUSED && tid == 0) claimed = atomicAdd(ticket, 1u) + gridDim.x;
    cx<T> x[E];
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int e = 0; e < E; e++) {
      const cx<T> r = my_line_in[phys_at<THREADS>(t_in, pt_in, e)];
      x[e].x = backward ? r.y : r.x;
      x[e].y = backward ? r.x : r.y;
    }
    sync_lines(per_line);
    pass<T, N, E, P::R1, 1, false>(x, t_in, nullptr, my_line_in);
    if (P::NPASS == 3) {
      sync_lines(per_line);
#pragma unroll
      for (int e = 0; e < E; e++) x[e] = my_line_in[phys_at<THREADS>(t_in, pt_in, e)];
      sync_lines(per_line);
      pass<T, N, E, P::R2, P::R1, false>(x, t_in, tw2, my_line_in);
    }
    sync_lines(lbar_out);
    {
      const cx<T> *line_r = smem + tt_out * pitch;
      const int pt_out = phys(t_out);
#pragma unroll
      for (int e = 0; e < E; e++) x[e] = line_r[phys_at<THREADS>(t_out, pt_out, e)];
    }
    unsigned next;
    if (FUSED) {
      if (tid == 0) next_item_s[parity] = claimed;
      __syncthreads();   // exchange buffers are free again, and everybody learns the next item
      next = next_item_s[parity];
      parity ^= 1;
    } else {
      sync_lines(lbar_out);
      next = item + gridDim.x;
    }
    publish();   // the previous tile of this CTA
    const long long obase_cur = obase;
    const int tvalid_cur = tvalid;
    int nwhich = 0, nplane = 0;
    unsigned nw = 0;
    bool fetched = false;
    if (next < nitems) {
      decode(next, nwhich, nw, nplane);
      if (!FUSED || flag_ok(nwhich, nplane, ld_relaxed_u32(flag_of(nwhich, nplane)))) {
        locate((FUSED && nwhich) ? spB : spA, nplane, nw, ibase, obase, tvalid);
        prefetch((FUSED && nwhich) ? spB : spA, nwhich, ibase, tvalid);
        fetched = true;
      }
    }
    if (P::NPASS == 3) pass<T, N, E, (P::R3 > 1 ? P::R3 : 2), P::R1 * P::R2, true>(x, t_out, tw3, nullptr);
    else pass<T, N, E, P::R2, P::R1, true>(x, t_out, tw2, nullptr);
    if (tt_out < tvalid_cur) {
      const long long t_os = sp.tile_dim >= 0 ? sp.bos[sp.tile_dim] : 0;
      const long long thread_off = obase_cur + (long long)tt_out * t_os + (long long)t_out * sp.ostride;
      const unsigned long long pol = which ? pol_first : pol_last;
#pragma unroll
      for (int e = 0; e < E; e++) {
        V raw;
        raw.x = backward ? x[e].y : x[e].x;
        raw.y = backward ? x[e].x : x[e].y;
        cx<T> *out = reinterpret_cast<cx<T> *>(sp.out[sp.out_seg[e]]) + thread_off + sp.out_off[e];
        if (FUSED) st_hint(reinterpret_cast<V *>(out), raw, pol);
        else *reinterpret_cast<V *>(out) = raw;
      }
    }
    if (FUSED) sig = which * fp.planes + plane;
    if (FUSED && next < nitems && !fetched) {
      // must wait: say first that this tile is done (whoever we wait for may be waiting for us)
      publish();
      while (!flag_ok(nwhich, nplane, ld_relaxed_u32(flag_of(nwhich, nplane)))) __nanosleep(100);
      locate(nwhich ? spB : spA, nplane, nw, ibase, obase, tvalid);
      prefetch(nwhich ? spB : spA, nwhich, ibase, tvalid);
    }
    item = next;
    which = nwhich;
    plane = nplane;
    w = nw;
  }
  publish();
}

// magic numbers of the divisions in the kernels' tile walk (tiles along the tile dimension, batch extents)
void fill_walk_divisors(StageParams &sp) {
  sp.rg.dalong = make_fastdiv((unsigned)std::max<long long>(1, sp.tiles_along));
  for (int k = 0; k < kMaxBatch; k++) sp.rg.dbext[k] = make_fastdiv((unsigned)std::max<long long>(1, k < sp.nbatch ? sp.bext[k] : 1));
}

template <typename T, int N, int E, int MAXT>
cudaError_t launch_fused_class(StageParams &a, StageParams &b, FusePlanes &fp, cudaStream_t stream) {
  fill_walk_divisors(a);
  fill_walk_divisors(b);
  constexpr int THREADS = N / E;
  using P = Passes<N, E>;
  constexpr int TWN = (P::R2 - 1) * P::R1 + (P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0);
  const int tl = a.tl;
  const int skew = tl >= 8 ? 1 : 8 / tl;
  const size_t smem = ((size_t)tl * (N + (N >> 4) + skew) + TWN) * 2 * sizeof(T);
  auto kfwd = fused_pair_kernel<T, N, E, MAXT, false, true>;
  auto kbwd = fused_pair_kernel<T, N, E, MAXT, true, true>;
  auto kern = a.sign > 0 ? kbwd : kfwd;
  static bool attr_set = false;
  if (!attr_set) {
    // (the kernel also has a few bytes of static shared memory)
    cudaError_t e = cudaFuncSetAttribute(kfwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kbwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tl * THREADS, smem);
  if (per_sm < 1) return cudaErrorInvalidValue;
  const int warps = tl * THREADS / 32;
  fp.target1 = (unsigned)fp.t1 * (unsigned)warps;
  fp.target2 = (unsigned)fp.t2 * (unsigned)warps;
  {
    // magic numbers of x / (t1 + t2)
    const unsigned d = (unsigned)(fp.t1 + fp.t2);
    unsigned l = 0;
    while ((1ull << l) < d) l++;
    fp.div_m = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    fp.div_s1 = l < 1 ? l : 1;
    fp.div_s2 = l > 0 ? l - 1 : 0;
  }
  const long long nitems = (long long)(fp.t1 + fp.t2) * fp.planes;
  if (nitems >= (1ll << 31)) return cudaErrorInvalidValue;
  // all CTAs must be co-resident (they wait for one another)
  const long long grid = std::min<long long>(nitems, (long long)sms * per_sm);
  // completion counters [2][planes] and the ticket counter behind them
  cudaError_t e = cudaMemsetAsync(fp.done, 0, sizeof(unsigned) * (2 * (size_t)fp.planes + 1), stream);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)grid, tl * THREADS, smem, stream>>>(a, b, fp);
  launch_counter()++;
  return cudaGetLastError();
}

template <typename T, int N, int E>
cudaError_t launch_fused_variant(StageParams &a, StageParams &b, FusePlanes &fp, cudaStream_t stream) {
  constexpr int THREADS = N / E;
  const int block = a.tl * THREADS;
  if (block <= 128) return launch_fused_class<T, N, E, (THREADS > 128 ? THREADS : 128)>(a, b, fp, stream);
  if (block <= 256) return launch_fused_class<T, N, E, (THREADS > 256 ? THREADS : 256)>(a, b, fp, stream);
  return launch_fused_class<T, N, E, (THREADS > 512 ? THREADS : 512)>(a, b, fp, stream);
}


// ---- micro-blocked stages ------------------------------------------------------------------
// Stage of a chain of power-of-two transforms whose intermediate arrays are stored as dense
// micro-blocks (planner.cpp: plan_microblocks).  A tile is `tl` whole lines given by an offset
// table; on a blocked side the lines of a tile interleave inside the blocks, so
//   * loads from a blocked input use the tile-minor thread mapping (lane -> line first, then
//     point): a warp fetches whole blocks, 512 / 256 contiguous bytes at 1024^3 fp64;
//   * the butterflies always run line-major (a line belongs to whole warps: per-line named
//     barriers between the passes);
//   * stores to a blocked output use the tile-minor mapping again -- the change of mapping
//     happens inside the last exchange through shared memory, as in the plain kernels.
// User-facing sides (first stage's input, last stage's output) are contiguous lines and keep the
// line-major mapping.  Every address is  tile base + thread base + table[e].
//
// One 512-thread CTA per SM owns the whole register file (16 points x 2 x fp64 per thread), so
// nothing but this CTA can cover its memory latency.  Hence the tile is DOUBLE-BUFFERED: the
// inputs of tile i+1 stream (cp.async, completion counted by an mbarrier) into a staging buffer
// of their own during ALL of tile i's passes, and the exchanges between the passes go through a
// second buffer.  Both fit because the exchange buffer holds 8-byte words only: fp64 lines are
// exchanged in two rounds (real parts, then imaginary parts), fp32 lines as whole complex words.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
// arrival that fires when all cp.async of this thread issued so far have landed (counted in the init count)
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// butterflies of one pass, outputs left in registers: output q of butterfly b in x[b + q * B]
template <typename T, int N, int E, int R, int NS>
__device__ __forceinline__ void pass_regs(cx<T> *x, int t, const cx<T> *__restrict__ twp) {
  pass<T, N, E, R, NS, true>(x, t, twp, nullptr);
}
// where pass (R, NS) puts output q of butterfly b: index into a padded line of 8-byte words
template <int N, int E, int R, int NS>
__device__ __forceinline__ int scatter_index(int t, int b, int q) {
  constexpr int THREADS = N / E;
  const int j = t + b * THREADS;
  const int k = j & (NS - 1);
  const int base = ((j - k) * R) + k;
  if (NS % 16 == 0) return phys(base) + q * (NS + NS / 16);
  if (NS == 1 && R == 16) return 17 * j + q;
  return phys(base + q * NS);
}
// PART 0: real parts, 1: imaginary parts (XE = T); 2: whole complex words (XE = cx<T>)
template <typename T, int N, int E, int R, int NS, int PART, typename XE>
__device__ __forceinline__ void scatter(const cx<T> *x, int t, XE *line) {
  constexpr int B = E / R;
#pragma unroll
  for (int b = 0; b < B; b++)
#pragma unroll
    for (int q = 0; q < R; q++) {
      const int idx = scatter_index<N, E, R, NS>(t, b, q);
      if constexpr (PART == 0) line[idx] = x[b + q * B].x;
      else if constexpr (PART == 1) line[idx] = x[b + q * B].y;
      else line[idx] = x[b + q * B];
    }
}

// TMA bulk copy global -> shared of `bytes` (multiple of 16, both addresses 16-byte aligned); completion is
// counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load(unsigned dst_smem, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}

// Geometry of the staging buffer, shared by the kernel and its launcher.  The tile's input is a few
// contiguous pieces of global memory -- its lines (line-major input) or, for a micro-blocked input, the
// blocks coming from one source rank (tl * iblk words: blocks are contiguous along the gathered
// dimension) -- copied densely, piece after piece, in chunks of <= 4 KB by the first `nchunks` threads.
struct StageGeom {
  int blocked, nchunks, chunk_elems, piece_elems;
};
__host__ __device__ inline StageGeom stage_geom(int N, int tl, int iblk2, long long iblk, long long iseg_stride, int elem_bytes) {
  StageGeom g;
  g.blocked = iblk2 > 1;
  g.piece_elems = g.blocked ? tl * (int)(iseg_stride ? iblk : N) : N;
  g.chunk_elems = 4096 / elem_bytes;
  if (g.chunk_elems > g.piece_elems) g.chunk_elems = g.piece_elems;
  g.nchunks = tl * N / g.chunk_elems;
  return g;
}

template <typename T, int N, int E, int MAXT, bool BWD>
__global__ void __launch_bounds__(MAXT, 1) stage_blk_kernel(const __grid_constant__ StageParams sp) {
  using P = Passes<N, E>;
  using V = typename C2<T>::type;
  constexpr int THREADS = N / E;
  constexpr bool SPLIT = sizeof(T) == 8;                 // exchange real and imaginary parts separately
  using XE = typename std::conditional<SPLIT, T, cx<T>>::type;   // 8-byte exchange word
  static_assert(sizeof(XE) == 8, "exchange words are 8 bytes");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tl = sp.tl;
  const StageGeom sg = stage_geom(N, tl, sp.iblk2, sp.iblk, sp.iseg_stride, (int)sizeof(cx<T>));
  // exchange buffer: lines of a tile must fall into different banks for the tile-minor gathers
  // (8 lines x 2 x 8-byte words per half-warp: pitch = 2 mod 16)
  // Half-CTA mode (stages whose chunks are all local): the change to the tile-minor store mapping is
  // done inside groups of tl/2 lines with 256-thread named barriers, so the two halves of the CTA drift
  // apart like the lines do; a warp then stores tl/2 lines x 8 points = 64-byte pieces, merged in L2.
  const bool grouped = sp.half_cta != 0;
  const int pitch_x = N + (N >> 4) + (grouped ? 4 : 2);
  cx<T> *stage = reinterpret_cast<cx<T> *>(smem_raw);
  XE *xbuf = reinterpret_cast<XE *>(stage + (size_t)tl * N);
  constexpr int TW2 = (P::R2 - 1) * P::R1;
  constexpr int TW3 = P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0;
  cx<T> *tw_s = reinterpret_cast<cx<T> *>(xbuf + (size_t)tl * pitch_x);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(tw_s + TW2 + TW3);
  const int tid = threadIdx.x;
  {
    const cx<T> *g2 = reinterpret_cast<const cx<T> *>(sp.tw2);
    for (int i = tid; i < TW2 + TW3; i += blockDim.x) tw_s[i] = g2[i];   // (tables are contiguous, [r-1][k])
  }
  const unsigned bar_full = (unsigned)__cvta_generic_to_shared(bars);
  const unsigned bar_empty = bar_full + 8;
  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_empty, blockDim.x);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const cx<T> *tw2 = tw_s;
  const cx<T> *tw3 = tw_s + TW2;
  __syncthreads();

  const bool out_lm = sp.oblk2 == 1;                             // line-major mapping on the output side
  const int t_c = tid % THREADS, tt_c = tid / THREADS;           // butterflies
  const int lg = tl >> 1, grp = tid >> 8, u = tid & 255;         // half-CTA mode: lines per group, my group
  const int t_s = out_lm ? t_c : (grouped ? u / lg : tid / tl);
  const int tt_s = out_lm ? tt_c : (grouped ? grp * lg + u % lg : tid % tl);
  const bool per_line = (THREADS % 32 == 0) && tl <= 15;
  const bool lbar_out = per_line && out_lm;
  auto sync_lines = [&](bool pl) {
    if (pl) asm volatile("bar.sync %0, %1;" ::"r"(1 + tt_c), "n"(THREADS) : "memory");
    else if (grouped) asm volatile("bar.sync %0, 256;" ::"r"(9 + grp) : "memory");
    else __syncthreads();
  };
  // my E input points inside the staging buffer: first + e * step
  const cx<T> *my_first;
  int my_step;
  if (sg.blocked) {
    // block j / tG of the tile, my line's words inside it (lines are XOR-permuted from block to block:
    // a quarter-warp reads 8 / tG consecutive blocks and must hit different banks)
    const int b = t_c / sp.iblk2, bk = tl * sp.iblk2;
    my_first = stage + b * bk + ((int)sp.tile_ioff[tt_c] ^ ((b & sp.iswz_mask) * sp.iblk2)) + (t_c % sp.iblk2);
    my_step = (THREADS / sp.iblk2) * bk;
  } else {
    my_first = stage + tt_c * N + t_c;
    my_step = THREADS;
  }
  XE *const my_x = xbuf + tt_c * pitch_x;
  const XE *const out_x = xbuf + tt_s * pitch_x;
  const int pt_c = phys(t_c), pt_s = phys(t_s);
  // the piece of the tile this thread fetches (TMA bulk copy), if any
  const bool loader = tid < sg.nchunks;
  const unsigned chunk_bytes = (unsigned)(sg.chunk_elems * sizeof(cx<T>));
  const unsigned tile_bytes = chunk_bytes * (unsigned)sg.nchunks;
  const cx<T> *chunk_src = reinterpret_cast<const cx<T> *>(sp.in);
  unsigned chunk_dst = (unsigned)__cvta_generic_to_shared(stage);
  if (loader) {
    const int first = tid * sg.chunk_elems, q = first / sg.piece_elems, within = first % sg.piece_elems;
    chunk_src += (sg.blocked ? (long long)q * sp.iseg_stride : sp.tile_ioff[q]) + within;
    chunk_dst += (unsigned)(first * sizeof(cx<T>));
  }
  const long long out_line = sp.tile_ooff[tt_s];
  const long long out_point = out_lm ? (long long)t_s * sp.ostride
                                     : (long long)(t_s / sp.oblk2) * sp.oblk2_stride + (long long)(t_s % sp.oblk2) * sp.ostride;
  const unsigned ntiles = (unsigned)sp.ntiles;
  // obase includes this thread's line: its place inside the output blocks depends on the tile (bank swizzle)
  auto locate = [&](unsigned tile, long long &ibase, long long &obase) {
    unsigned rest = tile;
    ibase = 0;
    obase = 0;
    unsigned cswz = 0;
#pragma unroll
    for (int k = kMaxBatch - 1; k >= 0; k--) {
      if (k >= sp.nbatch) continue;
      const unsigned ext = (unsigned)sp.bext[k];
      const unsigned qe = fast_div(rest, sp.rg.dbext[k].m, sp.rg.dbext[k].s1, sp.rg.dbext[k].s2);   // (host-made magic numbers)
      const unsigned c = rest - qe * ext;
      rest = qe;
      ibase += (long long)c * sp.bis[k];
      obase += (long long)c * sp.bos[k];
      if (k == sp.oswz_batch) cswz = c;
    }
    obase += out_line ^ (long long)((cswz & (unsigned)sp.oswz_mask) << sp.oswz_shift);
  };
  auto prefetch = [&](long long ibase) {
    if (tid == 0) mbar_arrive_expect_tx(bar_full, tile_bytes);
    if (loader) bulk_load(chunk_dst, chunk_src + ibase, chunk_bytes, bar_full);
  };

  long long ibase = 0, obase = 0;
  unsigned tile = blockIdx.x;
  if (tile < ntiles) {
    locate(tile, ibase, obase);
    prefetch(ibase);
  }
  unsigned phase = 0;
  for (; tile < ntiles; tile += gridDim.x, phase ^= 1) {
    cx<T> x[E];
    mbar_wait(bar_full, phase);
#pragma unroll
    for (int e = 0; e < E; e++) {
      const cx<T> r = my_first[e * my_step];
      x[e].x = BWD ? r.y : r.x;
      x[e].y = BWD ? r.x : r.y;
    }
    mbar_arrive(bar_empty);          // my part of the staging buffer may be overwritten
    pass_regs<T, N, E, P::R1, 1>(x, t_c, nullptr);
    const long long obase_cur = obase;
    const unsigned next = tile + gridDim.x;
    if (next < ntiles) locate(next, ibase, obase);
    // tile-minor output: the previous tile's last gather read every line of xbuf; everybody is past it once
    // everybody has emptied the staging buffer of this tile
    if (!lbar_out && !grouped) mbar_wait(bar_empty, phase);
#define PFB_EXCHANGE(R_, NS_, SRC, T_, PT_, PL)                                             \
  if constexpr (SPLIT) {                                                                    \
    T re_[E];                                                                               \
    scatter<T, N, E, R_, NS_, 0>(x, t_c, my_x);                                             \
    sync_lines(PL);                                                                         \
    _Pragma("unroll") for (int e = 0; e < E; e++) re_[e] = (SRC)[phys_at<THREADS>(T_, PT_, e)];   \
    sync_lines(PL);                                                                         \
    scatter<T, N, E, R_, NS_, 1>(x, t_c, my_x);                                             \
    sync_lines(PL);                                                                         \
    _Pragma("unroll") for (int e = 0; e < E; e++) {                                         \
      x[e].y = (SRC)[phys_at<THREADS>(T_, PT_, e)];                                         \
      x[e].x = re_[e];                                                                      \
    }                                                                                       \
  } else {                                                                                  \
    scatter<T, N, E, R_, NS_, 2>(x, t_c, my_x);                                             \
    sync_lines(PL);                                                                         \
    _Pragma("unroll") for (int e = 0; e < E; e++) x[e] = (SRC)[phys_at<THREADS>(T_, PT_, e)];     \
  }
    if (P::NPASS == 3) {
      PFB_EXCHANGE(P::R1, 1, my_x, t_c, pt_c, per_line)
      sync_lines(per_line);          // the line's words are free again
      pass_regs<T, N, E, P::R2, P::R1>(x, t_c, tw2);
    }
    // fetch the next tile (its staging buffer must have been emptied by everybody; that was long ago)
    if (next < ntiles) {
      if ((lbar_out || grouped) && (loader || tid == 0)) mbar_wait(bar_empty, phase);
      prefetch(ibase);
    }
    if (P::NPASS == 3) {
      PFB_EXCHANGE(P::R2, P::R1, out_x, t_s, pt_s, lbar_out)
      if (lbar_out) sync_lines(true);
      else if (grouped) sync_lines(false);     // my group's words are free for the next tile
      pass_regs<T, N, E, (P::R3 > 1 ? P::R3 : 2), P::R1 * P::R2>(x, t_s, tw3);
    } else {
      PFB_EXCHANGE(P::R1, 1, out_x, t_s, pt_s, lbar_out)
      if (lbar_out) sync_lines(true);
      else if (grouped) sync_lines(false);
      pass_regs<T, N, E, P::R2, P::R1>(x, t_s, tw2);
    }
#undef PFB_EXCHANGE
    const long long thread_off = obase_cur + out_point;
#pragma unroll
    for (int e = 0; e < E; e++) {
      V raw;
      raw.x = BWD ? x[e].y : x[e].x;
      raw.y = BWD ? x[e].x : x[e].y;
      *reinterpret_cast<V *>(reinterpret_cast<cx<T> *>(sp.outp[e]) + thread_off) = raw;
    }
  }
}

template <typename T, int N, int E, int MAXT>
cudaError_t launch_blk_class(StageParams &sp, cudaStream_t stream) {
  constexpr int THREADS = N / E;
  using P = Passes<N, E>;
  constexpr int TWN = (P::R2 - 1) * P::R1 + (P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0);
  const int tl = sp.tl;
  fill_walk_divisors(sp);
  const StageGeom sg = stage_geom(N, tl, sp.iblk2, sp.iblk, sp.iseg_stride, 2 * (int)sizeof(T));
  // half-CTA tile-minor exchange: opt-in (PFFT_B200_HALFCTA=1), 512-thread tiles of 8 lines, blocked output,
  // every chunk local (64-byte stores over NVLink reach 477 GB/s against 707 GB/s for 128 bytes and more)
  static const int half_env = [] {
    const char *e = getenv("PFFT_B200_HALFCTA");
    return e ? atoi(e) : 0;
  }();
  sp.half_cta = (half_env && tl == 8 && THREADS == 64 && sp.oblk2 > 1 && sp.noseg == 1) ? 1 : 0;
  const size_t pitch_x = N + (N >> 4) + (sp.half_cta ? 4 : 2);
  const size_t smem = (size_t)tl * N * 2 * sizeof(T) + tl * pitch_x * 8 + (size_t)TWN * 2 * sizeof(T) + 16;
  // contiguous pieces only: lines of a line-major input, block runs of a micro-blocked one
  if (sg.blocked ? sp.iblk2_stride != (long long)tl * sp.iblk2 : (sp.istride != 1 || sp.iseg_stride != 0)) return cudaErrorInvalidValue;
  if ((tl * N) % sg.chunk_elems || sg.piece_elems % sg.chunk_elems) return cudaErrorInvalidValue;
  for (int e = 0; e < E; e++) sp.outp[e] = static_cast<char *>(sp.out[sp.out_seg[e]]) + sp.out_off[e] * (long long)(2 * sizeof(T));
  auto kf = stage_blk_kernel<T, N, E, MAXT, false>;
  auto kb = stage_blk_kernel<T, N, E, MAXT, true>;
  auto kern = sp.sign > 0 ? kb : kf;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (smem > 227 * 1024 || tl * THREADS > MAXT || sg.nchunks > tl * THREADS) return cudaErrorInvalidValue;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tl * THREADS, smem);
  if (per_sm < 1) return cudaErrorInvalidValue;
  const long long grid = std::min<long long>(sp.ntiles, (long long)sms * per_sm);
  kern<<<(unsigned)grid, tl * THREADS, smem, stream>>>(sp);
  launch_counter()++;
  return cudaGetLastError();
}

// (N, E) pairs compiled in; E = points per thread.  Two block-size classes per pair:
// <= 256 threads (two CTAs per SM) and <= 512 threads (one CTA per SM, twice the lines per
// tile, i.e. twice as long contiguous runs on a strided side).
template <typename T, int N, int E, int MAXT>
cudaError_t launch_block_class(StageParams &sp, cudaStream_t stream) {
  constexpr int THREADS = N / E;
  const int tl = sp.tl;
  const int skew = tl >= 8 ? 1 : 8 / tl;
  using P = Passes<N, E>;
  constexpr int TWN = (P::R2 - 1) * P::R1 + (P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0);
  const size_t smem = ((size_t)tl * (N + (N >> 4) + skew) + TWN) * 2 * sizeof(T);
  fill_walk_divisors(sp);
  if (sp.fast && sp.istride == 1) {
    // contiguous input lines: the lean body shared with the plane-fused pair
    auto kf = fused_pair_kernel<T, N, E, MAXT, false, false>;
    auto kb = fused_pair_kernel<T, N, E, MAXT, true, false>;
    auto k1 = sp.sign > 0 ? kb : kf;
    static bool attr1_set = false;
    if (!attr1_set) {
      cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e != cudaSuccess) return e;
      attr1_set = true;
    }
    int dev1 = 0, sms1 = 148, per_sm1 = 1;
    cudaGetDevice(&dev1);
    cudaDeviceGetAttribute(&sms1, cudaDevAttrMultiProcessorCount, dev1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, k1, tl * THREADS, smem);
    if (per_sm1 < 1) per_sm1 = 1;
    const long long grid1 = std::min<long long>(sp.ntiles, (long long)sms1 * per_sm1);
    FusePlanes none = {};
    k1<<<(unsigned)grid1, tl * THREADS, smem, stream>>>(sp, sp, none);
    launch_counter()++;
    return cudaGetLastError();
  }
  using K = void (*)(StageParams);
  // [0] general (run-time direction), [1] / [2] FAST forward / backward (strided input lines here)
  K kerns[3] = {stage_pow2_kernel<T, N, E, MAXT, false, 0, false>, stage_pow2_kernel<T, N, E, MAXT, true, 1, false>,
                stage_pow2_kernel<T, N, E, MAXT, true, 2, false>};
  K kern = kerns[sp.fast ? 1 + (sp.sign > 0 ? 1 : 0) : 0];
  static bool attr_set = false;
  if (!attr_set) {
    for (K k : kerns) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
    }
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
  // fp32: 16-line tiles of 1024 threads on 64 registers (128-byte runs on a strided side)
  if constexpr (sizeof(T) == 4) {
    if (sp.tl * THREADS > 512) return launch_block_class<T, N, E, 1024>(sp, stream);
  }
  return launch_block_class<T, N, E, (THREADS > 512 ? THREADS : 512)>(sp, stream);
}

int points_per_thread(int n) { return pow2_points_per_thread(n); }   // core.h: shared with the planner

}  // namespace

template <typename T>
bool pow2_supported(const Stage &g, int L) {
  if (L < 64 || L > 4096 || (L & (L - 1))) return false;
  // Real lines: the any-length kernel runs them as n/2 packed complex points (fft_mixed.h).  The path below
  // (full-length complex transform: r2c with imaginary parts zero, c2r with the upper half of the spectrum
  // completed by Hermitian symmetry on load) is kept behind PFFT_B200_REAL_POW2=1 for comparison.
  static const bool real_pow2 = [] {
    const char *e = getenv("PFFT_B200_REAL_POW2");
    return e && atoi(e) != 0;
  }();
  if ((g.op == OP_R2C || g.op == OP_C2R) && !real_pow2) return false;
  if (g.op == OP_R2C) {
    if (!g.in_real || g.out_real) return false;
  } else if (g.op == OP_C2R) {
    if (g.in_real || !g.out_real || g.zin != 0 || g.nin != L / 2 + 1) return false;
  } else if (g.op != OP_C2C || g.in_real || g.out_real) {
    return false;
  }
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
  if (g.ntile > 0) return g.ntile;   // explicit tile of a micro-blocked chain (the planner's choice)
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
  // fp32: 16 lines of 64 threads = 128-byte runs (PFFT_B200_F32_BLOCK=512 restores the 8-line tiles)
  static const int f32_block = [] {
    const char *e = getenv("PFFT_B200_F32_BLOCK");
    return e ? atoi(e) : 1024;
  }();
  const int max_block = sizeof(T) == 4 ? f32_block : 512;
  if (strided && sizeof(T) == 4 && threads >= 32) block = max_block;
  int tl = block / threads;
  if (forced > 0 && strided) tl = forced;
  static const int forced_c = [] {
    const char *e = getenv("PFFT_B200_TLC");
    return e ? atoi(e) : 0;
  }();
  if (forced_c > 0 && !strided) tl = forced_c;
  while (tl * threads > max_block && tl > 1) tl /= 2;
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
  if (sp.ntile > 0) {
    // micro-blocked chain: tiles of 512 threads
    switch (sp.L) {
      case 64: return launch_blk_class<T, 64, 8, 512>(sp, stream);
      case 128: return launch_blk_class<T, 128, 8, 512>(sp, stream);
      case 256: return launch_blk_class<T, 256, 16, 512>(sp, stream);
      case 512: return launch_blk_class<T, 512, 8, 512>(sp, stream);
      case 1024: return launch_blk_class<T, 1024, 16, 512>(sp, stream);
      case 2048: return launch_blk_class<T, 2048, 16, 512>(sp, stream);
      case 4096: return launch_blk_class<T, 4096, 16, 512>(sp, stream);
      default: return cudaErrorInvalidValue;
    }
  }
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

// Threads of a block must be whole warps and at least 128 for the fused pair; lines per tile
// are chosen for 32-byte runs on the ring side (L2 sectors), not for HBM.
template <typename T>
int fused_pick_tile(int L) {
  const int threads = L / points_per_thread(L);
  int tl = (int)(32 / (2 * sizeof(T)));
  static const int forced = [] {
    const char *e = getenv("PFFT_B200_FTL");
    return e ? atoi(e) : 0;
  }();
  if (forced > 0) tl = forced;
  while (tl * threads < 128) tl *= 2;
  while (tl * threads > 512 && tl > 1) tl /= 2;
  return tl;
}

template <typename T>
cudaError_t launch_fused_pow2(StageParams &a, StageParams &b, FusePlanes &fp, cudaStream_t stream) {
  if (a.L != b.L || a.tl != b.tl) return cudaErrorInvalidValue;
  switch (a.L) {
    case 64: return launch_fused_variant<T, 64, 8>(a, b, fp, stream);
    case 128: return launch_fused_variant<T, 128, 8>(a, b, fp, stream);
    case 256: return launch_fused_variant<T, 256, 16>(a, b, fp, stream);
    case 512: return launch_fused_variant<T, 512, 8>(a, b, fp, stream);
    case 1024: return launch_fused_variant<T, 1024, 16>(a, b, fp, stream);
    case 2048: return launch_fused_variant<T, 2048, 16>(a, b, fp, stream);
    case 4096: return launch_fused_variant<T, 4096, 16>(a, b, fp, stream);
    default: return cudaErrorInvalidValue;
  }
}

// Decide whether the table-driven addressing applies and fill the tables.
template <typename T>
void pow2_prepare(const Stage &g, StageParams &sp) {
  const int L = sp.L;
  const int E = points_per_thread(L);
  const int threads = L / E;
  bool fast = g.op == OP_C2C && g.nin == L && g.zin == 0 && g.nout == L && g.zout == 0 && !g.mod_in.on && !g.mod_out.on && !g.conj_in && !g.conj_out;
  const bool seg_in = g.iseg_stride != 0 && g.iblk < g.nin;
  const bool seg_out = g.noseg > 1;
  if (seg_in && g.iblk % threads != 0) fast = false;
  if (seg_out && g.oblk % threads != 0) fast = false;
  if (threads % g.iblk2 != 0 || threads % g.oblk2 != 0) fast = false;
  sp.fast = fast ? 1 : 0;
  static const int line_bars = [] {
    const char *e = getenv("PFFT_B200_LINEBAR");
    return e ? atoi(e) : 1;
  }();
  sp.line_bars = line_bars;   // (tiles of more than 15 lines share barriers between neighbouring lines)
  if (!fast) return;
  for (int e = 0; e < E; e++) {
    const long long idx = (long long)e * threads;
    // (idx is a multiple of the micro-block sizes: only the high part moves)
    auto along = [](long long x, long long blk2, long long blk2_stride, long long stride) {
      return blk2 > 1 ? (x / blk2) * blk2_stride : x * stride;
    };
    if (seg_in) {
      const long long seg = idx / g.iblk;
      sp.in_off[e] = seg * g.iseg_stride + along(idx - seg * g.iblk, g.iblk2, g.iblk2_stride, g.istride);
    } else {
      sp.in_off[e] = along(idx, g.iblk2, g.iblk2_stride, g.istride);
    }
    if (seg_out) {
      const long long seg = idx / g.oblk;
      sp.out_seg[e] = (int)seg;
      sp.out_off[e] = along(idx - seg * g.oblk, g.oblk2, g.oblk2_stride, g.ostride);
    } else {
      sp.out_seg[e] = 0;
      sp.out_off[e] = along(idx, g.oblk2, g.oblk2_stride, g.ostride);
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
template int fused_pick_tile<float>(int);
template int fused_pick_tile<double>(int);
template cudaError_t launch_fused_pow2<float>(StageParams &, StageParams &, FusePlanes &, cudaStream_t);
template cudaError_t launch_fused_pow2<double>(StageParams &, StageParams &, FusePlanes &, cudaStream_t);
template void pow2_prepare<float>(const Stage &, StageParams &);
template void pow2_prepare<double>(const Stage &, StageParams &);

}  // namespace pfb
