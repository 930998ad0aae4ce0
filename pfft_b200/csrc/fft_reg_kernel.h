// fft_reg_kernel.h -- register-resident stage kernel for
//   * real lines of even length n = 2M: r2c / c2r run as ONE packed complex transform of M points with the
//     Hermitian post- / pre-processing fused into the last / first shared-memory exchange (the reference plans
//     true real transforms: kernel/sertrafo.c:517-530, fftw_plan_guru64_dft_r2c / _c2r);
//   * complex lines of length 3 * 2^k (768 = 1.5 x 512, the oversampled sizes of kernel/ousample.c).
// A line of NL = Q * NSUB complex points (Q = 1 or 3) belongs to TL = Q * NSUB / E threads.  Thread r = Q t + q
// runs sub-transform q (points with output index = q mod Q) of length NSUB on E registers:
//   fetch     cp.async puts the tile's lines, dense, into a staging buffer of their own WHILE the tile before is
//             being transformed (one persistent CTA per SM: nothing else can cover its memory latency)
//   prologue  y_q[j] = w_NL^(j q) sum_r' x[j + r' NSUB] w_Q^(r' q)   (radix-Q decimation in frequency on pick-up
//             from the staging buffer, together with zero padding, +-1 modulation, conjugation and the re/im
//             swap of backward transforms; c2r first turns the half spectrum into the packed spectrum in place)
//   passes    the Stockham passes of fft_regs.h inside the sub-line's own region of the exchange buffer
//   output    X[Q k' + q] sits in thread (t, q) as element k' = t + e NSUB/E, i.e. output index r + e TL:
//             the line's natural strided distribution over its TL threads -- contiguous lines are stored
//             straight from registers, strided ones (and r2c, whose post-processing needs Z[k] and Z[M-k]) after
//             one more exchange (tile-minor store mapping: a warp writes tl neighbouring lines per point).
// Addressing: a "simple" side (one chunk, window starting at index 0, no modulation) is a per-thread pointer plus
// compile-time multiples of one step; anything else (shifted windows, gathered chunks, per-destination segments,
// modulations) takes the general path.  One read and one write of the array per stage.
#pragma once
#include "fft_regs.h"

namespace pfb {

namespace {

__device__ __forceinline__ unsigned rg_div(unsigned x, const FastDiv &f) {
  const unsigned t = __umulhi(f.m, x);
  return (t + ((x - t) >> f.s1)) >> f.s2;
}

// region stride of a sub-line inside the exchange buffer: the Q sub-lines of one line are touched by neighbouring
// lanes at the same offset, so their regions start 3 (mod 8) elements apart
template <int NSUB, int Q>
struct RegGeom {
  static constexpr int RS = Q == 1 ? NSUB + (NSUB >> 4) : ((NSUB + (NSUB >> 4) + 7) / 8 * 8 + 3);
};

// KIND 0: complex line (BWD: backward, re/im swapped around the forward butterflies), 1: r2c (n = 2 NL reals in,
// NL + 1 complex out), 2: c2r (NL + 1 complex in, n reals out).
// MAXT: threads per CTA the kernel is compiled for (512: 128 registers per thread, 768: 85, 1024: 64).
template <typename T, int NSUB, int E, int Q, int KIND, bool BWD, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) stage_reg_kernel(const __grid_constant__ StageParams sp) {
  using P = Passes<NSUB, E>;
  using V = typename C2<T>::type;
  constexpr int TS = NSUB / E;          // threads per sub-line
  constexpr int TL = Q * TS;            // threads per line
  constexpr int NL = Q * NSUB;          // complex points per line
  constexpr int RS = RegGeom<NSUB, Q>::RS;
  constexpr int TW2 = (P::R2 - 1) * P::R1;
  constexpr int TW3 = P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0;
  constexpr int TWQ = Q > 1 ? NL : 0;
  constexpr int TWH = KIND != 0 ? NL + 1 : 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tl = sp.tl;
  const int pitch = sp.rg.pitch, spitch = sp.rg.spitch;
  cx<T> *const xbuf = reinterpret_cast<cx<T> *>(smem_raw);          // exchanges between passes (padded index)
  cx<T> *const sbuf = xbuf + (size_t)tl * pitch;                    // staging: dense lines, odd pitch
  cx<T> *const tw_s = sbuf + (size_t)tl * spitch;
  const int tid = threadIdx.x;
  {
    const cx<T> *g = reinterpret_cast<const cx<T> *>(sp.rg.tables);          // [tw2 | tw3 | twq | twh], contiguous
    for (int i = tid; i < TW2 + TW3 + TWQ + TWH; i += blockDim.x) tw_s[i] = g[i];
  }
  const cx<T> *const tw2 = tw_s;
  const cx<T> *const tw3 = tw_s + TW2;
  const cx<T> *const twq = tw_s + TW2 + TW3;
  const cx<T> *const twh = tw_s + TW2 + TW3 + TWQ;
  __syncthreads();

  // ---- thread mappings
  const int tt_c = tid / TL, r_c = tid % TL;             // butterflies: line-major
  const int t = r_c / Q, q = r_c % Q;
  const bool in_lm = KIND == 1 || sp.istride == 1;       // packed real input lines are contiguous by construction
  const bool out_lm = KIND == 2 || sp.ostride == 1;
  const int t_in = in_lm ? r_c : tid / tl, tt_in = in_lm ? tt_c : tid % tl;
  const int t_out = out_lm ? r_c : tid / tl, tt_out = out_lm ? tt_c : tid % tl;
  // Barriers between passes: only the threads of a line have to meet.  Lines are grouped so that a group is made
  // of whole warps (1 line of 64 threads, 2 of 48, 4 of 24 ...) and at most 15 groups exist; every group has a
  // named barrier of its own, so the groups of a tile drift apart and their shared-memory and arithmetic phases
  // overlap.  Exchanges that mix the lines of a tile (tile-minor mappings) use the CTA-wide barrier.
  constexpr int BG0 = TL % 32 == 0 ? 1 : (2 * TL % 32 == 0 ? 2 : (4 * TL % 32 == 0 ? 4 : (8 * TL % 32 == 0 ? 8 : 16)));
  int bg = BG0;
  while (tl / bg > 15) bg *= 2;
  const bool per_line = sp.line_bars && bg < tl && tl % bg == 0;
  const bool lbar_in = per_line && in_lm, lbar_out = per_line && out_lm;
  const int bar_id = 1 + tt_c / bg, bar_cnt = bg * TL;
  auto sync_lines = [&](bool pl) {
    if (pl) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_cnt) : "memory");
    else __syncthreads();
  };
  cx<T> *const my_reg = xbuf + tt_c * pitch + q * RS;    // exchange region of my sub-line
  const int pt = phys(t);
  cx<T> *const my_sline = sbuf + tt_c * spitch;          // input line I transform
  const unsigned in_line_s = (unsigned)__cvta_generic_to_shared(sbuf + tt_in * spitch + t_in);

  const unsigned ntiles = (unsigned)sp.ntiles;
  auto locate = [&](unsigned tile, long long &ibase, long long &obase, int &tvalid) {
    unsigned rest = tile;
    ibase = 0;
    obase = 0;
    tvalid = 1;
    if (sp.tile_dim >= 0) {
      // (divisions by run-time extents through host-made magic numbers: a few instructions instead of ~25 each)
      const unsigned along = (unsigned)sp.tiles_along;
      const unsigned qa = rg_div(rest, sp.rg.dalong);
      const unsigned chunk = rest - qa * along;
      rest = qa;
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
      const unsigned qe = rg_div(rest, sp.rg.dbext[k]);
      const unsigned c = rest - qe * ext;
      rest = qe;
      ibase += (long long)c * sp.bis[k];
      obase += (long long)c * sp.bos[k];
    }
  };
  const long long t_is = sp.tile_dim >= 0 ? sp.bis[sp.tile_dim] : 0;
  const long long t_os = sp.tile_dim >= 0 ? sp.bos[sp.tile_dim] : 0;
  const bool seg_in = sp.iseg_stride != 0;
  const bool seg_out = sp.noseg > 1;
  const bool simple_in = sp.rg.simple_in != 0, simple_out = sp.rg.simple_out != 0;
  // simple sides: element e of this thread sits e * step behind the thread's first element
  const long long in_first = (KIND == 1 ? 2ll : sp.istride) * t_in, in_step = (KIND == 1 ? 2ll : sp.istride) * TL;
  const long long out_first = (KIND == 2 ? 2ll : sp.ostride) * t_out, out_step = (KIND == 2 ? 2ll : sp.ostride) * TL;

  // ---- asynchronous fetch of a tile's inputs (cp.async; zero-filled outside the input window)
  auto fetch_general = [&](int e_idx, long long line0, bool live) {
    const int idx = t_in + e_idx * TL;
    const unsigned dst = in_line_s + (unsigned)(e_idx * TL * sizeof(cx<T>));
    if (KIND == 1) {
      // packed point idx = reals 2 idx, 2 idx + 1 of the zero-padded line; zin and nin are even
      const int jr = 2 * idx - sp.zin;
      const bool ok = live && jr >= 0 && jr < sp.nin;
      cp_async_zfill<sizeof(cx<T>)>(dst, ok ? (const void *)(reinterpret_cast<const T *>(sp.in) + line0 + jr) : sp.in, ok);
    } else {
      const int j = idx - sp.zin;
      const bool ok = live && j >= 0 && j < sp.nin;
      long long off = 0;
      if (ok) {
        if (seg_in) {
          const int seg = (int)rg_div((unsigned)j, sp.rg.diblk);
          off = (long long)seg * sp.iseg_stride + (long long)(j - seg * sp.iblk) * sp.istride;
        } else {
          off = (long long)j * sp.istride;
        }
      }
      cp_async_zfill<sizeof(cx<T>)>(dst, ok ? (const void *)(reinterpret_cast<const cx<T> *>(sp.in) + line0 + off) : sp.in, ok);
    }
  };
  auto prefetch = [&](long long ibase, int tvalid) {
    const bool live = tt_in < tvalid;
    const long long line0 = ibase + (long long)tt_in * t_is;
    if (simple_in) {
      // one chunk, window starting at 0: element e is present iff its index lies below the input count
      const int lim = live ? (KIND == 1 ? sp.nin >> 1 : sp.nin) : 0;
      if (KIND == 1) {
        const T *src = reinterpret_cast<const T *>(sp.in) + line0 + in_first;
#pragma unroll
        for (int e = 0; e < E; e++) {
          const bool ok = t_in + e * TL < lim;
          cp_async_zfill<sizeof(cx<T>)>(in_line_s + (unsigned)(e * TL * sizeof(cx<T>)), ok ? (const void *)(src + e * in_step) : sp.in, ok);
        }
      } else {
        const cx<T> *src = reinterpret_cast<const cx<T> *>(sp.in) + line0 + in_first;
#pragma unroll
        for (int e = 0; e < E; e++) {
          const bool ok = t_in + e * TL < lim;
          cp_async_zfill<sizeof(cx<T>)>(in_line_s + (unsigned)(e * TL * sizeof(cx<T>)), ok ? (const void *)(src + e * in_step) : sp.in, ok);
        }
        if (KIND == 2 && t_in == 0) {
          const bool ok = NL < lim;
          cp_async_zfill<sizeof(cx<T>)>(in_line_s + (unsigned)(NL * sizeof(cx<T>)), ok ? (const void *)(src + E * in_step) : sp.in, ok);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; e++) fetch_general(e, line0, live);
      if (KIND == 2 && t_in == 0) fetch_general(E, line0, live);      // the Nyquist bin
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const bool third_zero = Q == 3 && sp.rg.third_zero != 0;
  const bool plain_pick = KIND == 2 || !(sp.mod_in.on || (KIND == 0 && sp.conj_in));
  // element `idx` of my input line as the butterflies want it
  auto pick = [&](int idx) -> cx<T> {
    cx<T> v = my_sline[idx];
    if (!plain_pick) {
      if (KIND == 0) {
        if (sp.conj_in) v.y = -v.y;
        if (sp.mod_in.on && sign_mod_dev(sp.mod_in, idx - sp.zin) < 0) { v.x = -v.x; v.y = -v.y; }
      } else if (KIND == 1) {
        if (sign_mod_dev(sp.mod_in, 2 * idx - sp.zin) < 0) v.x = -v.x;
        if (sign_mod_dev(sp.mod_in, 2 * idx + 1 - sp.zin) < 0) v.y = -v.y;
      }
    }
    if (BWD) { const T s = v.x; v.x = v.y; v.y = s; }
    return v;
  };

  long long ibase, obase;
  int tvalid;
  unsigned tile = blockIdx.x;
  if (tile < ntiles) {
    locate(tile, ibase, obase, tvalid);
    prefetch(ibase, tvalid);
  }
  for (; tile < ntiles; tile += gridDim.x) {
    cx<T> x[E];
    asm volatile("cp.async.wait_all;" ::: "memory");
    sync_lines(lbar_in);                       // the whole line has landed (every thread fetched a part of it)
    if (KIND == 2) {
      // c2r pre-processing, in place, one thread per pair (k, M - k), M = NL:
      //   Zf[k] = (X[k] + conj X[M-k]) + i conj(w^k) (X[k] - conj X[M-k]),  w = exp(-2 pi i / n)
      // so that x[2j] + i x[2j+1] = sum_k Zf[k] exp(+2 pi i j k / M); the backward sum runs through the forward
      // butterflies on swapped parts (stored swapped here, swapped back on the way out).
#pragma unroll
      for (int e = 0; e <= E / 2; e++) {
        const int k = r_c + e * TL;
        if (k > NL / 2) break;
        cx<T> a = my_sline[k], b = my_sline[NL - k];
        if (sp.conj_in) { a.y = -a.y; b.y = -b.y; }
        if (sp.mod_in.on) {
          if (sign_mod_dev(sp.mod_in, k - sp.zin) < 0) { a.x = -a.x; a.y = -a.y; }
          if (sign_mod_dev(sp.mod_in, NL - k - sp.zin) < 0) { b.x = -b.x; b.y = -b.y; }
        }
        if (k == 0) a.y = b.y = (T)0;          // DC and Nyquist bins of a real line are real
        const cx<T> Ee{a.x + b.x, a.y - b.y}, D{a.x - b.x, a.y + b.y};
        const cx<T> w = twh[k];
        const cx<T> O{D.x * w.x + D.y * w.y, D.y * w.x - D.x * w.y};
        my_sline[k] = cx<T>{Ee.y + O.x, Ee.x - O.y};
        if (k != 0 && 2 * k != NL) my_sline[NL - k] = cx<T>{O.x - Ee.y, Ee.x + O.y};
      }
      sync_lines(per_line);                    // (written and read by the line's own threads)
    }
    // ---- prologue: my E points of sub-transform q
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int j = t + e * TS;
      if (Q == 1) {
        x[e] = pick(j);
      } else {
        // (pruned input whose last third is known to be zero padding -- 512 of 768: not even read)
        const cx<T> a0 = pick(j), a1 = pick(j + NSUB), a2 = third_zero ? cx<T>{(T)0, (T)0} : pick(j + 2 * NSUB);
        const cx<T> s{a1.x + a2.x, a1.y + a2.y}, d{a1.x - a2.x, a1.y - a2.y};
        if (q == 0) {
          x[e] = cx<T>{a0.x + s.x, a0.y + s.y};
        } else {
          const T h = q == 1 ? (T)0.86602540378443864676 : (T)-0.86602540378443864676;
          const cx<T> y{a0.x - (T)0.5 * s.x + h * d.y, a0.y - (T)0.5 * s.y - h * d.x};
          x[e] = cmul(y, twq[j * q]);
        }
      }
    }
    sync_lines(lbar_in);                       // inputs are in registers: the staging buffer is free
    const long long obase_cur = obase;
    const int tvalid_cur = tvalid;
    {
      const unsigned next = tile + gridDim.x;
      if (next < ntiles) {
        locate(next, ibase, obase, tvalid);
        prefetch(ibase, tvalid);
      }
    }
    // ---- passes of the sub-transform
    pass<T, NSUB, E, P::R1, 1, false>(x, t, nullptr, my_reg);
    if (P::NPASS == 3) {
      sync_lines(per_line);
#pragma unroll
      for (int e = 0; e < E; e++) x[e] = my_reg[phys_at<TS>(t, pt, e)];
      sync_lines(per_line);
      pass<T, NSUB, E, P::R2, P::R1, false>(x, t, tw2, my_reg);
    }
    sync_lines(per_line);
#pragma unroll
    for (int e = 0; e < E; e++) x[e] = my_reg[phys_at<TS>(t, pt, e)];
    const bool regs_out = KIND == 2 || (KIND == 0 && out_lm);      // stored straight from registers
    if (!regs_out) sync_lines(per_line);       // the line's regions are free for the last exchange
    if (P::NPASS == 3) pass<T, NSUB, E, (P::R3 > 1 ? P::R3 : 2), P::R1 * P::R2, true>(x, t, tw3, nullptr);
    else pass<T, NSUB, E, P::R2, P::R1, true>(x, t, tw2, nullptr);
    // x[e] = output r_c + e TL of my line

    if (KIND == 2) {
      // ---- packed real pairs, contiguous line: reals 2j, 2j+1 of point j = r_c + e TL (parts swapped back)
      if (tt_c < tvalid_cur) {
        T *out = reinterpret_cast<T *>(sp.out[0]) + obase_cur + (long long)tt_c * t_os;
        if (simple_out) {
          out += out_first;
#pragma unroll
          for (int e = 0; e < E; e++) {
            if (2 * (r_c + e * TL) >= sp.nout) continue;
            V raw;
            raw.x = x[e].y;
            raw.y = x[e].x;
            *reinterpret_cast<V *>(out + e * out_step) = raw;
          }
        } else {
#pragma unroll
          for (int e = 0; e < E; e++) {
            const int jr = 2 * (r_c + e * TL) - sp.zout;
            if (jr < 0 || jr >= sp.nout) continue;
            V raw;
            raw.x = x[e].y;
            raw.y = x[e].x;
            if (sp.mod_out.on) {
              if (sign_mod_dev(sp.mod_out, jr) < 0) raw.x = -raw.x;
              if (sign_mod_dev(sp.mod_out, jr + 1) < 0) raw.y = -raw.y;
            }
            *reinterpret_cast<V *>(out + jr) = raw;
          }
        }
      }
      continue;
    }
    auto store = [&](int k, cx<T> v, long long line0) {
      const int kk = k - sp.zout;
      if (kk < 0 || kk >= sp.nout) return;
      V raw;
      raw.x = BWD ? v.y : v.x;
      raw.y = BWD ? v.x : v.y;
      if (sp.mod_out.on && sign_mod_dev(sp.mod_out, kk) < 0) { raw.x = -raw.x; raw.y = -raw.y; }
      if (sp.conj_out) raw.y = -raw.y;
      int seg = 0, kl = kk;
      if (seg_out) {
        seg = (int)rg_div((unsigned)kk, sp.rg.doblk);
        kl = kk - seg * sp.oblk;
      }
      *reinterpret_cast<V *>(reinterpret_cast<cx<T> *>(sp.out[seg]) + line0 + (long long)kl * sp.ostride) = raw;
    };
    auto store_all = [&](int t_s, int tt_s) {
      if (tt_s >= tvalid_cur) return;
      const long long line0 = obase_cur + (long long)tt_s * t_os;
      if (simple_out) {
        cx<T> *out = reinterpret_cast<cx<T> *>(sp.out[0]) + line0 + out_first;
#pragma unroll
        for (int e = 0; e < E; e++) {
          if (t_s + e * TL >= sp.nout) continue;
          V raw;
          raw.x = BWD ? x[e].y : x[e].x;
          raw.y = BWD ? x[e].x : x[e].y;
          *reinterpret_cast<V *>(out + e * out_step) = raw;
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) store(t_s + e * TL, x[e], line0);
      }
    };
    if (regs_out) {
      store_all(r_c, tt_c);
      continue;
    }
    // ---- last exchange: natural order, dense, into the line (plus Z[0] once more behind it, so that the mirror
    // index M - k needs no special case), picked up in the store mapping
    {
      cx<T> *line_w = xbuf + tt_c * pitch + r_c;
#pragma unroll
      for (int e = 0; e < E; e++) line_w[e * TL] = x[e];
      if (KIND == 1 && r_c == 0) line_w[NL] = x[0];
    }
    sync_lines(lbar_out);
    const cx<T> *line_r = xbuf + tt_out * pitch;
    if (KIND == 0) {
#pragma unroll
      for (int e = 0; e < E; e++) x[e] = line_r[t_out + e * TL];
      sync_lines(lbar_out);                    // everybody has read the lines: the next tile may scatter into them
      store_all(t_out, tt_out);
      continue;
    }
    // ---- r2c post-processing by pairs (k, M - k), M = NL, k = t_out + e TL < M / 2:
    //   E = (Z[k] + conj Z[M-k]) / 2,  P = w^k (Z[k] - conj Z[M-k]) / (2i):   X[k] = E + P,  X[M-k] = conj(E - P)
    // (k = 0 yields the DC and the Nyquist bin; the middle bin X[M/2] = conj Z[M/2] is one extra element)
    cx<T> mid{(T)0, (T)0};
    {
      const cx<T> *za = line_r + t_out, *zb = line_r + (NL - t_out), *wk = twh + t_out;
#pragma unroll
      for (int e = 0; e < E / 2; e++) {
        const cx<T> a = za[e * TL], b = zb[-e * TL], w = wk[e * TL];
        const cx<T> Ee{(T)0.5 * (a.x + b.x), (T)0.5 * (a.y - b.y)}, D{(T)0.5 * (a.x - b.x), (T)0.5 * (a.y + b.y)};
        const cx<T> Pp{D.y * w.x + D.x * w.y, D.y * w.y - D.x * w.x};       // (D / i) w
        x[e] = cx<T>{Ee.x + Pp.x, Ee.y + Pp.y};
        x[E / 2 + e] = cx<T>{Ee.x - Pp.x, Pp.y - Ee.y};
      }
      if (t_out == 0) {
        const cx<T> m = line_r[NL / 2];
        mid = cx<T>{m.x, -m.y};
      }
    }
    sync_lines(lbar_out);                      // everybody has read the lines: the next tile may scatter into them
    if (tt_out < tvalid_cur) {
      const long long line0 = obase_cur + (long long)tt_out * t_os;
      if (simple_out) {
        cx<T> *const base = reinterpret_cast<cx<T> *>(sp.out[0]) + line0;
        cx<T> *lo = base + out_first, *hi = base + (long long)NL * sp.ostride - out_first;
#pragma unroll
        for (int e = 0; e < E / 2; e++) {
          const int k = t_out + e * TL;
          V raw;
          if (k < sp.nout) {
            raw.x = x[e].x;
            raw.y = x[e].y;
            *reinterpret_cast<V *>(lo + e * out_step) = raw;
          }
          if (NL - k < sp.nout) {
            raw.x = x[E / 2 + e].x;
            raw.y = x[E / 2 + e].y;
            *reinterpret_cast<V *>(hi - e * out_step) = raw;
          }
        }
        if (t_out == 0 && NL / 2 < sp.nout) {
          V raw;
          raw.x = mid.x;
          raw.y = mid.y;
          *reinterpret_cast<V *>(base + (long long)(NL / 2) * sp.ostride) = raw;
        }
      } else {
#pragma unroll
        for (int e = 0; e < E / 2; e++) {
          store(t_out + e * TL, x[e], line0);
          store(NL - t_out - e * TL, x[E / 2 + e], line0);
        }
        if (t_out == 0) store(NL / 2, mid, line0);
      }
    }
  }
}

template <typename T, int NSUB, int E, int Q, int KIND, bool BWD, int MAXT>
cudaError_t launch_reg_variant(StageParams &sp, cudaStream_t stream) {
  using P = Passes<NSUB, E>;
  constexpr int TL = Q * NSUB / E;
  constexpr int NL = Q * NSUB;
  constexpr int TWN = (P::R2 - 1) * P::R1 + (P::NPASS == 3 ? (P::R3 - 1) * P::R1 * P::R2 : 0) + (Q > 1 ? NL : 0) + (KIND != 0 ? NL + 1 : 0);
  auto kern = stage_reg_kernel<T, NSUB, E, Q, KIND, BWD, MAXT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int threads = sp.tl * TL;
  const size_t smem = ((size_t)sp.tl * (sp.rg.pitch + sp.rg.spitch) + TWN) * 2 * sizeof(T);
  if (threads > MAXT || smem > 227 * 1024 || sp.rg.E != E || sp.rg.maxt != MAXT) return cudaErrorInvalidValue;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (per_sm < 1) return cudaErrorInvalidValue;
  const long long grid = std::min<long long>(sp.ntiles, (long long)sms * per_sm);
  kern<<<(unsigned)grid, threads, smem, stream>>>(sp);
  launch_counter()++;
  return cudaGetLastError();
}

// Thread classes (host twin: reg_class in fft_reg.cu): fp32 lines of 8 points per thread run 1024-thread tiles on
// 64 registers (768 / 85 with the radix-3 prologue), fp64 ones 768-thread tiles on 85; 16 points per thread: 512 / 128.
template <typename T, int NSUB, int E, int Q, int KIND>
cudaError_t launch_reg_one(StageParams &sp, cudaStream_t stream) {
  constexpr int MAXT = E == 8 ? ((sizeof(T) == 4 && Q == 1) ? 1024 : 768) : 512;
  if constexpr (KIND == 0) {
    if (sp.sign > 0) return launch_reg_variant<T, NSUB, E, Q, KIND, true, MAXT>(sp, stream);
  }
  return launch_reg_variant<T, NSUB, E, Q, KIND, false, MAXT>(sp, stream);
}

// (NSUB, E, Q) by complex line length NL (host twin: reg_geometry in fft_reg.cu)
template <typename T, int KIND>
cudaError_t launch_reg_kind(StageParams &sp, cudaStream_t stream) {
  switch (sp.rg.NL) {
    case 64: if constexpr (KIND != 0) return launch_reg_one<T, 64, 8, 1, KIND>(sp, stream); break;
    case 128: if constexpr (KIND != 0) return launch_reg_one<T, 128, 8, 1, KIND>(sp, stream); break;
    case 256: if constexpr (KIND != 0) return launch_reg_one<T, 256, 8, 1, KIND>(sp, stream); break;
    case 512: if constexpr (KIND != 0) return launch_reg_one<T, 512, 8, 1, KIND>(sp, stream); break;
    case 1024: if constexpr (KIND != 0) return launch_reg_one<T, 1024, 16, 1, KIND>(sp, stream); break;
    case 2048: if constexpr (KIND != 0) return launch_reg_one<T, 2048, 16, 1, KIND>(sp, stream); break;
    case 192: return launch_reg_one<T, 64, 8, 3, KIND>(sp, stream);
    case 384: return launch_reg_one<T, 128, 8, 3, KIND>(sp, stream);
    case 768: return launch_reg_one<T, 256, 8, 3, KIND>(sp, stream);
    case 1536: return launch_reg_one<T, 512, 8, 3, KIND>(sp, stream);
    default: break;
  }
  return cudaErrorInvalidValue;
}

}  // namespace

}  // namespace pfb
