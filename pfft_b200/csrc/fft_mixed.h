// fft_mixed.h -- body of the any-length stage kernel (stage_mixed_kernel, fft_mixed.cu), written as
// host + device code: the CUDA kernel runs it with one thread per `tid`, the CPU emulation used by the
// planner tests (pfftb200_emulate_stage, tests/test_kernel_emulation.py) runs the very same functions phase by
// phase over all `tid`.  Only the barrier differs.
//
// One tile = `tl` lines of the transformed dimension held in ONE buffer (shared memory, or a per-CTA
// global workspace for lines that do not fit).  Per tile:
//   load    gather the lines from global memory (user array or the chunks an exchange delivered) into
//           digit-reversed positions, zero-pad (ni -> n), +-1 modulation, conjugation, real -> complex;
//           even-length real lines are packed as n/2 complex points (r2c) resp. built from the Hermitian
//           half spectrum (c2r)
//   core    in-place mixed-radix passes (decimation in time) with register codelets of radix 2, 3, 4, 5, 7, 8,
//           11, 13, 16 (codelets.h), one barrier per pass; lengths with a larger prime factor go through
//           Bluestein's algorithm (two power-of-two transforms of length M >= 2L - 1, the second one
//           decimation in frequency, and three pointwise products)
//   store   the kept outputs (n -> no) with the Hermitian post-processing of r2c, the DCT/DST
//           twiddles, modulation and conjugation, straight to their final place (next stage's layout,
//           per-destination chunks of the next exchange, or the user's array)
// i.e. one read and one write of the array per stage where the reference runs an embed loop, an FFTW plan,
// an FFTW copy plan and a truncate loop (kernel/outrafo.c:154-168, kernel/sertrafo.c:1073-1245,
// kernel/ousample.c:391-506).  Real transforms follow kernel/sertrafo.c:517-530 (r2c/c2r of the last dimension).
#pragma once
#include <type_traits>

#include "codelets.h"
#include "kernels.h"

namespace pfb {

PFB_HD unsigned fd_div(unsigned x, const FastDiv &f) {
#if defined(__CUDA_ARCH__)
  const unsigned t = __umulhi(f.m, x);
#else
  const unsigned t = (unsigned)(((unsigned long long)f.m * x) >> 32);
#endif
  return (t + ((x - t) >> f.s1)) >> f.s2;
}

// padded index inside a buffer line: one extra element every 16 keeps power-of-two strides off one bank
PFB_HD int mx_phys(int i) { return i + (i >> 4); }

template <typename T>
PFB_HD cx<T> mx_ldg(const cx<T> *p) {
#if defined(__CUDA_ARCH__)
  using V = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
  const V v = __ldg(reinterpret_cast<const V *>(p));
  return cx<T>{v.x, v.y};
#else
  return *p;
#endif
}

PFB_HD int mx_sign_mod(const SignModDev &m, int idx) {
  // reference api/api-basic.c:1213-1234: (-1)^g below the half, with g = idx + start
  const int g = idx + m.start;
  if (g >= m.half) return 1;
  return ((g & 1) ? -1 : 1) * m.extra;
}

// streaming global loads: the input is read exactly once -- keep it out of L1, which holds the twiddle and
// digit-reversal tables
template <typename T>
PFB_HD cx<T> mx_ld_stream(const cx<T> *p) {
#if defined(__CUDA_ARCH__)
  cx<T> r;
  if (sizeof(T) == 8)
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(*reinterpret_cast<double *>(&r.x)), "=d"(*reinterpret_cast<double *>(&r.y)) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(*reinterpret_cast<float *>(&r.x)), "=f"(*reinterpret_cast<float *>(&r.y)) : "l"(p));
  return r;
#else
  return *p;
#endif
}
template <typename T>
PFB_HD T mx_ld_stream(const T *p) {
#if defined(__CUDA_ARCH__)
  T r;
  if (sizeof(T) == 8) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(*reinterpret_cast<double *>(&r)) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(*reinterpret_cast<float *>(&r)) : "l"(p));
  return r;
#else
  return *p;
#endif
}

// v[r] *= w^(r e1) for r = 1..R-1: only the powers w^(2^b e1) come from the table (log2 R loads instead of
// R - 1 per butterfly); every other factor is a product of at most 3 of them (a few ulp), formed on the fly so
// that no second register array is alive next to the butterfly's operands
template <typename T, int R>
PFB_HD void mx_apply_twiddles(const cx<T> *tw, unsigned e1, cx<T> *v) {
  constexpr int NB = R > 8 ? 4 : (R > 4 ? 3 : (R > 2 ? 2 : 1));
  cx<T> wp[NB];
#pragma unroll
  for (int b = 0; b < NB; b++)
    if ((1 << b) < R) wp[b] = mx_ldg(tw + ((unsigned)(1 << b)) * e1);
#pragma unroll
  for (int r = 1; r < R; r++) {
    cx<T> w;
    bool have = false;
#pragma unroll
    for (int b = 0; b < NB; b++) {
      if (!((r >> b) & 1)) continue;
      w = have ? cxmul(w, wp[b]) : wp[b];
      have = true;
    }
    v[r] = cxmul(v[r], w);
  }
}

// One in-place radix-R pass over all lines of the tile.  Decimation in time (DIF == false): inputs sit in
// digit-reversed order before the first pass and the last pass leaves natural order; pass p works on blocks
// of Ns * R points (Ns = product of the earlier radices): butterfly (blk, k) reads positions
// blk * Ns * R + k + r * Ns, multiplies input r by w^(r k tstep), and writes the R outputs back to the SAME
// positions -- no second buffer and one barrier per pass.  Decimation in frequency (DIF == true) is the
// transposed pass (twiddles behind the butterfly); running the passes in reverse order takes natural order
// to digit-reversed order (second transform of the Bluestein convolution).
template <typename T, int R, bool DIF>
PFB_HD void mixed_pass_r(const MixedPass &ps, cx<T> *buf, const cx<T> *tw, int pitch, int tvalid, int tid, int nthr) {
  const unsigned LR = (unsigned)ps.LR, Ns = (unsigned)ps.Ns;
  const unsigned total = (unsigned)tvalid * LR;
  for (unsigned i = (unsigned)tid; i < total; i += (unsigned)nthr) {
    const unsigned tt = fd_div(i, ps.dLR);
    const unsigned jj = i - tt * LR;
    const unsigned blk = fd_div(jj, ps.dNs);
    const unsigned k = jj - blk * Ns;
    cx<T> v[R];
    cx<T> *line = buf + (size_t)tt * pitch;
    const unsigned base = blk * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = line[mx_phys((int)(base + r * Ns))];
    if (Ns > 1 && !DIF) mx_apply_twiddles<T, R>(tw, k * (unsigned)ps.tstep, v);
    dft_r<T, R>(v);
    if (Ns > 1 && DIF) mx_apply_twiddles<T, R>(tw, k * (unsigned)ps.tstep, v);
#pragma unroll
    for (int q = 0; q < R; q++) line[mx_phys((int)(base + q * Ns))] = v[q];
  }
}

template <typename T, bool DIF>
PFB_HD void mixed_pass(const MixedPass &ps, cx<T> *buf, const cx<T> *tw, int pitch, int tvalid, int tid, int nthr) {
  switch (ps.R) {
    case 2: mixed_pass_r<T, 2, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    case 3: mixed_pass_r<T, 3, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    case 4: mixed_pass_r<T, 4, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    case 5: mixed_pass_r<T, 5, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    case 7: mixed_pass_r<T, 7, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    case 8: mixed_pass_r<T, 8, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    case 11: mixed_pass_r<T, 11, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    case 13: mixed_pass_r<T, 13, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
    default: mixed_pass_r<T, 16, DIF>(ps, buf, tw, pitch, tvalid, tid, nthr); break;
  }
}

// (line, position) of flat work item e: lines fastest when the global side is strided (a warp then touches
// tl neighbouring lines = tl * sizeof(element) contiguous bytes per position), positions fastest when it is contiguous
struct MxItem {
  int tt, j;
  bool ok;
};
PFB_HD MxItem mx_item(unsigned e, bool contiguous, const FastDiv &dcount, int count, int tl_shift, int tvalid) {
  MxItem it;
  if (contiguous) {
    it.tt = (int)fd_div(e, dcount);
    it.j = (int)(e - (unsigned)it.tt * (unsigned)count);
    it.ok = true;
  } else {
    it.tt = (int)(e & ((1u << tl_shift) - 1));
    it.j = (int)(e >> tl_shift);
    it.ok = it.tt < tvalid;
  }
  return it;
}

// One tile.  buf: tl * pitch complex elements.  `ex.run(f)` runs f(tid, nthr) for every thread of the
// CTA and then a CTA-wide barrier.
template <typename T, typename Exec>
PFB_HD void mixed_tile(const StageParams &sp, cx<T> *buf, long long ibase, long long obase, int tvalid, long long t_is,
                       long long t_os, Exec &ex) {
  const MixedParams &mx = sp.mx;
  const int pitch = mx.pitch, tl = sp.tl, Lc = mx.Lc, L = mx.L;
  cx<T> *const B0 = buf;
  const cx<T> *const tw = reinterpret_cast<const cx<T> *>(mx.tw);
  const cx<T> *const twh = reinterpret_cast<const cx<T> *>(mx.tw_half);
  const int *const rev = mx.rev;            // natural index -> position the first (decimation-in-time) pass wants
  const bool swap_in = mx.swap != 0;        // backward transform through the forward butterflies: swap re/im on the way in and out
  const int half = mx.half_real;            // 0: complex line, 1: r2c on n/2 packed points, 2: c2r on n/2 packed points
  const int M = L;                          // (half-length real transforms: L = n/2 complex points)
  auto pos_of = [&](int idx) -> int {
#if defined(__CUDA_ARCH__)
    return mx_phys(rev ? __ldg(rev + idx) : idx);
#else
    return mx_phys(rev ? rev[idx] : idx);
#endif
  };

  // ---- 0. clear what the load does not overwrite
  if (mx.zero_fill) {
    ex.run([&](int tid, int nthr) {
      const cx<T> z{(T)0, (T)0};
      for (unsigned e = (unsigned)tid; e < (unsigned)tvalid * (unsigned)pitch; e += (unsigned)nthr) B0[e] = z;
    });
  }

  // ---- 1. load: several independent global loads in flight per thread, then the shared-memory writes
  ex.run([&](int tid, int nthr) {
    constexpr int U = sizeof(T) == 8 ? 4 : 8;
    const bool contiguous = sp.istride == 1;
    const bool pairs = mx.in_pairs != 0;       // packed r2c line, contiguous and aligned: two reals per load
    // c2r on packed points walks k = 0..M-1 and fetches X[k] and X[M-k] itself
    const int count = half == 2 ? M : (pairs ? sp.nin >> 1 : sp.nin);
    const FastDiv &dcount = half == 2 ? mx.dL : (pairs ? mx.dnin2 : mx.dnin);
    const unsigned total = (unsigned)(contiguous ? tvalid : tl) * (unsigned)count;
    const bool seg_in = sp.iseg_stride != 0;
    auto offset_of = [&](int tt, int j) -> long long {
      long long off = ibase + (long long)tt * t_is;
      if (seg_in) {
        const int seg = (int)fd_div((unsigned)j, mx.diblk);
        off += (long long)seg * sp.iseg_stride + (long long)(j - seg * sp.iblk) * sp.istride;
      } else {
        off += (long long)j * sp.istride;
      }
      return off;
    };
    for (unsigned e0 = (unsigned)tid; e0 < total; e0 += (unsigned)nthr * U) {
      cx<T> val[U], val2[U];
      unsigned okmask = 0;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned e = e0 + (unsigned)u * (unsigned)nthr;
        if (e >= total) continue;
        const MxItem it = mx_item(e, contiguous, dcount, count, mx.tl_shift, tvalid);
        if (!it.ok) continue;
        okmask |= 1u << u;
        const int j = it.j, tt = it.tt;
        if (half == 2) {
          // spectrum entries at positions k and M - k of the zero-padded half spectrum (element j sits at j + zin)
          const int ja = j - sp.zin, jb = M - j - sp.zin;
          val[u] = cx<T>{(T)0, (T)0};
          val2[u] = cx<T>{(T)0, (T)0};
          if (ja >= 0 && ja < sp.nin) val[u] = mx_ld_stream(reinterpret_cast<const cx<T> *>(sp.in) + offset_of(tt, ja));
          if (jb >= 0 && jb < sp.nin) val2[u] = mx_ld_stream(reinterpret_cast<const cx<T> *>(sp.in) + offset_of(tt, jb));
        } else if (pairs) {
          val[u] = mx_ld_stream(reinterpret_cast<const cx<T> *>(sp.in) + (((ibase + (long long)tt * t_is) >> 1) + j));   // line base counts reals (even)
        } else if (sp.in_real) {
          val[u].x = mx_ld_stream(reinterpret_cast<const T *>(sp.in) + offset_of(tt, j));
        } else {
          val[u] = mx_ld_stream(reinterpret_cast<const cx<T> *>(sp.in) + offset_of(tt, j));
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (!((okmask >> u) & 1)) continue;
        // (the item is recomputed rather than kept: an array of structs would live in local memory)
        const MxItem it = mx_item(e0 + (unsigned)u * (unsigned)nthr, contiguous, dcount, count, mx.tl_shift, tvalid);
        const int j = it.j, tt = it.tt;
        cx<T> *line = B0 + (size_t)tt * pitch;
        if (half == 2) {
          // c2r, n = 2M: x[2j] + i x[2j+1] = sum_k Zf[k] exp(+2 pi i j k / M) with
          //   Zf[k] = (X[k] + conj X[M-k]) + i exp(+2 pi i k / n) (X[k] - conj X[M-k]),  k = 0..M-1
          // (unnormalised, like FFTW's c2r); the backward sum runs through the forward butterflies on swapped parts.
          const int k = j;
          cx<T> a = val[u], b = val2[u];
          const int ja = k - sp.zin, jb = M - k - sp.zin;
          if (sp.conj_in) { a.y = -a.y; b.y = -b.y; }
          if (sp.mod_in.on) {
            if (mx_sign_mod(sp.mod_in, ja) < 0) { a.x = -a.x; a.y = -a.y; }
            if (mx_sign_mod(sp.mod_in, jb) < 0) { b.x = -b.x; b.y = -b.y; }
          }
          if (k == 0) a.y = b.y = (T)0;      // DC and Nyquist bins of a real line are real (FFTW ignores their imaginary parts)
          const cx<T> E{a.x + b.x, a.y - b.y}, D{a.x - b.x, a.y + b.y};
          const cx<T> w = mx_ldg(twh + k);                       // exp(-2 pi i k / n); we need its conjugate
          const cx<T> O{D.x * w.x + D.y * w.y, D.y * w.x - D.x * w.y};
          const cx<T> Z{E.x - O.y, E.y + O.x};
          line[pos_of(k)] = cx<T>{Z.y, Z.x};
          continue;
        }
        if (pairs) {
          cx<T> v = val[u];
          if (sp.mod_in.on) {
            if (mx_sign_mod(sp.mod_in, 2 * j) < 0) v.x = -v.x;
            if (mx_sign_mod(sp.mod_in, 2 * j + 1) < 0) v.y = -v.y;
          }
          line[pos_of(j + (sp.zin >> 1))] = v;
          continue;
        }
        const int p = j + sp.zin;          // position inside the zero-padded line
        if (sp.in_real) {
          T xr = val[u].x;
          if (sp.mod_in.on && mx_sign_mod(sp.mod_in, j) < 0) xr = -xr;
          if (half == 1) {
            // two consecutive reals = one complex point
            reinterpret_cast<T *>(line + pos_of(p >> 1))[p & 1] = xr;
          } else if (sp.op == OP_R2R) {
            // w_j x_j exp(-i pi b jj / D), jj = position inside the logical line of n reals
            if ((p == 0 && sp.r2r_half0) || (p == sp.n - 1 && sp.r2r_halfn)) xr *= (T)0.5;
            const cx<T> w = mx_ldg(reinterpret_cast<const cx<T> *>(sp.tw_r2r) + (int)(((long long)2 * sp.r2r_b2 * p) % (8ll * sp.r2r_D)));
            line[pos_of(p)] = cx<T>{xr * w.x, xr * w.y};
          } else {
            line[pos_of(p)] = swap_in ? cx<T>{(T)0, xr} : cx<T>{xr, (T)0};
          }
        } else {
          cx<T> v = val[u];
          if (sp.conj_in) v.y = -v.y;
          if (sp.mod_in.on && mx_sign_mod(sp.mod_in, j) < 0) { v.x = -v.x; v.y = -v.y; }
          if (swap_in) { const T t = v.x; v.x = v.y; v.y = t; }
          line[pos_of(p)] = v;
          if (sp.op == OP_C2R && p >= 1 && 2 * p < L) {
            // odd n: complete the spectrum, X[n-k] = conj X[k] (the value is already swapped: the conjugate mirrors as (-x, y))
            line[pos_of(L - p)] = swap_in ? cx<T>{-v.x, v.y} : cx<T>{v.x, -v.y};
          }
        }
      }
    }
  });

  // ---- 2. core transform of length L (directly, or as a Bluestein convolution of length Lc)
  if (mx.bluestein) {
    ex.run([&](int tid, int nthr) {
      const cx<T> *ch = reinterpret_cast<const cx<T> *>(mx.chirp);
      const unsigned total = (unsigned)tvalid * (unsigned)L;
      for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
        const int tt = (int)fd_div(e, mx.dL);
        const int j = (int)(e - (unsigned)tt * (unsigned)L);
        cx<T> *p = B0 + (size_t)tt * pitch + pos_of(j);
        *p = cxmul(*p, mx_ldg(ch + j));
      }
    });
  }
  for (int ps = 0; ps < mx.npass; ps++)
    ex.run([&](int tid, int nthr) { mixed_pass<T, false>(mx.pass[ps], B0, tw, pitch, tvalid, tid, nthr); });
  if (mx.bluestein) {
    // spectrum of the convolution (natural order), conjugated: the inverse transform is conj(FFT(conj(.)));
    // it runs decimation-in-frequency and leaves its result in digit-reversed order
    ex.run([&](int tid, int nthr) {
      const cx<T> *bh = reinterpret_cast<const cx<T> *>(mx.bhat);
      const unsigned total = (unsigned)tvalid * (unsigned)Lc;
      for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
        const int tt = (int)fd_div(e, mx.dLc);
        const int m = (int)(e - (unsigned)tt * (unsigned)Lc);
        cx<T> *p = B0 + (size_t)tt * pitch + mx_phys(m);
        const cx<T> y = cxmul(*p, mx_ldg(bh + m));
        *p = cx<T>{y.x, -y.y};
      }
    });
    for (int ps = mx.npass - 1; ps >= 0; ps--)
      ex.run([&](int tid, int nthr) { mixed_pass<T, true>(mx.pass[ps], B0, tw, pitch, tvalid, tid, nthr); });
  }
  // transform output k of line `line`
  auto fetch = [&](const cx<T> *line, int k) -> cx<T> {
    if (!mx.bluestein) return line[mx_phys(k)];
    const cx<T> r = line[pos_of(k)];
    return cxmul(cx<T>{r.x, -r.y}, mx_ldg(reinterpret_cast<const cx<T> *>(mx.chirp) + k));
  };

  // ---- 3. store the kept outputs
  ex.run([&](int tid, int nthr) {
    const bool contiguous = sp.ostride == 1;
    const bool seg_out = sp.noseg > 1;
    if (mx.out_pairs) {
      // packed c2r line, contiguous and aligned: one 2-real store per packed point
      const int count = sp.nout >> 1;
      const unsigned total = (unsigned)tvalid * (unsigned)count;
      for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
        const int tt = (int)fd_div(e, mx.dnout2);
        const int c = (int)(e - (unsigned)tt * (unsigned)count);
        const cx<T> r = fetch(B0 + (size_t)tt * pitch, c + (sp.zout >> 1));
        cx<T> v{r.y, r.x};
        if (sp.mod_out.on) {
          if (mx_sign_mod(sp.mod_out, 2 * c) < 0) v.x = -v.x;
          if (mx_sign_mod(sp.mod_out, 2 * c + 1) < 0) v.y = -v.y;
        }
        const long long off = obase + (long long)tt * t_os;       // reals; even by construction
        reinterpret_cast<cx<T> *>(sp.out[0])[(off >> 1) + c] = v;
      }
      return;
    }
    const unsigned total = (unsigned)(contiguous ? tvalid : tl) * (unsigned)sp.nout;
    for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
      const MxItem it = mx_item(e, contiguous, mx.dnout, sp.nout, mx.tl_shift, tvalid);
      if (!it.ok) continue;
      const int kk = it.j, tt = it.tt;
      const int k = kk + sp.zout;        // logical output index
      const cx<T> *line = B0 + (size_t)tt * pitch;
      cx<T> v;
      if (half == 1) {
        // r2c, n = 2M: X[k] = E[k] + exp(-2 pi i k / n) O[k],  E = (Z[k] + conj Z[M-k]) / 2,  O = (Z[k] - conj Z[M-k]) / (2i)
        const cx<T> a = fetch(line, k == M ? 0 : k), b = fetch(line, k == 0 ? 0 : M - k);
        const cx<T> E{(T)0.5 * (a.x + b.x), (T)0.5 * (a.y - b.y)}, D{(T)0.5 * (a.x - b.x), (T)0.5 * (a.y + b.y)};
        const cx<T> O{D.y, -D.x};
        const cx<T> w = mx_ldg(twh + k);
        v = cx<T>{E.x + O.x * w.x - O.y * w.y, E.y + O.x * w.y + O.y * w.x};
      } else if (half == 2) {
        // real output k sits in part k & 1 of packed point k >> 1 (parts swapped: backward through forward butterflies)
        const cx<T> r = fetch(line, k >> 1);
        v = cx<T>{(k & 1) ? r.x : r.y, (T)0};
      } else {
        v = fetch(line, k);
        if (swap_in) { const T t = v.x; v.x = v.y; v.y = t; }
      }
      if (sp.op == OP_R2R) {
        // 2 F(exp(-i pi a (k + b) / D) X_k)
        const cx<T> w = mx_ldg(reinterpret_cast<const cx<T> *>(sp.tw_r2r) + (int)(((long long)sp.r2r_a2 * (2 * k + sp.r2r_b2)) % (8ll * sp.r2r_D)));
        const T re = v.x * w.x - v.y * w.y, im = v.x * w.y + v.y * w.x;
        v.x = sp.r2r_sine ? (T)-2 * im : (T)2 * re;
        v.y = 0;
      }
      if (sp.mod_out.on && mx_sign_mod(sp.mod_out, kk) < 0) { v.x = -v.x; v.y = -v.y; }
      if (sp.conj_out) v.y = -v.y;
      int seg = 0, kl = kk;
      if (seg_out) {
        seg = (int)fd_div((unsigned)kk, mx.doblk);
        kl = kk - seg * sp.oblk;
      }
      const long long off = obase + (long long)tt * t_os + (long long)kl * sp.ostride;
      if (sp.out_real) reinterpret_cast<T *>(sp.out[seg])[off] = v.x;
      else reinterpret_cast<cx<T> *>(sp.out[seg])[off] = v;
    }
  });
}

// tile index -> bases (same batch walk as the other stage kernels)
PFB_HD void mixed_locate(const StageParams &sp, long long tile, long long *ibase, long long *obase, int *tvalid, long long *t_is,
                         long long *t_os) {
  long long rest = tile, ib = 0, ob = 0;
  *tvalid = 1;
  *t_is = *t_os = 0;
  if (sp.tile_dim >= 0) {
    const long long chunk = rest % sp.tiles_along;
    rest /= sp.tiles_along;
    const long long first = chunk * sp.tl;
    const long long left = sp.bext[sp.tile_dim] - first;
    *tvalid = left < sp.tl ? (int)left : sp.tl;
    *t_is = sp.bis[sp.tile_dim];
    *t_os = sp.bos[sp.tile_dim];
    ib = first * *t_is;
    ob = first * *t_os;
  }
  for (int k = sp.nbatch - 1; k >= 0; k--) {
    if (k == sp.tile_dim) continue;
    const long long c = rest % sp.bext[k];
    rest /= sp.bext[k];
    ib += c * sp.bis[k];
    ob += c * sp.bos[k];
  }
  *ibase = ib;
  *obase = ob;
}

}  // namespace pfb
