// fft_mixed.h -- body of the any-length stage kernel (stage_mixed_kernel, fft_mixed.cu), written as
// host + device code: the CUDA kernel runs it with one thread per `tid`, the CPU emulation used by the
// planner tests (pfftb200_emulate_stage, tests/test_kernel_emulation.py) runs the very same functions phase by
// phase over all `tid`.  Only the barrier differs.
//
// One tile = `tl` lines of the transformed dimension held in two ping-pong buffers (shared memory, or a
// per-CTA global workspace for lines that do not fit).  Per tile:
//   load    gather the lines from global memory (user array or the chunks an exchange delivered),
//           zero-pad (ni -> n), +-1 modulation, conjugation, real -> complex; even-length real lines are
//           packed as n/2 complex points (r2c) resp. built from the Hermitian half spectrum (c2r)
//   core    mixed-radix Stockham passes with register codelets of radix 2, 3, 4, 5, 7, 8, 11, 13, 16
//           (codelets.h); lengths with a larger prime factor go through Bluestein's algorithm (two
//           power-of-two transforms of length M >= 2L - 1 and three pointwise products)
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

template <typename T, int R>
PFB_HD void mixed_pass_r(const MixedPass &ps, const cx<T> *src, cx<T> *dst, const cx<T> *tw, int pitch, int tvalid, int tid,
                         int nthr) {
  const unsigned LR = (unsigned)ps.LR, Ns = (unsigned)ps.Ns;
  const unsigned total = (unsigned)tvalid * LR;
  for (unsigned i = (unsigned)tid; i < total; i += (unsigned)nthr) {
    const unsigned tt = fd_div(i, ps.dLR);
    const unsigned jj = i - tt * LR;
    const unsigned blk = fd_div(jj, ps.dNs);
    const unsigned k = jj - blk * Ns;
    cx<T> v[R];
    const cx<T> *line = src + (size_t)tt * pitch;
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = line[mx_phys((int)(jj + r * LR))];
    if (Ns > 1) {
      const unsigned e1 = k * (unsigned)ps.tstep;
#pragma unroll
      for (int r = 1; r < R; r++) v[r] = cxmul(v[r], mx_ldg(tw + r * e1));
    }
    dft_r<T, R>(v);
    cx<T> *o = dst + (size_t)tt * pitch;
    const unsigned base = blk * Ns * R + k;
#pragma unroll
    for (int q = 0; q < R; q++) o[mx_phys((int)(base + q * Ns))] = v[q];
  }
}

template <typename T>
PFB_HD void mixed_pass(const MixedPass &ps, const cx<T> *src, cx<T> *dst, const cx<T> *tw, int pitch, int tvalid, int tid,
                       int nthr) {
  switch (ps.R) {
    case 2: mixed_pass_r<T, 2>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    case 3: mixed_pass_r<T, 3>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    case 4: mixed_pass_r<T, 4>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    case 5: mixed_pass_r<T, 5>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    case 7: mixed_pass_r<T, 7>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    case 8: mixed_pass_r<T, 8>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    case 11: mixed_pass_r<T, 11>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    case 13: mixed_pass_r<T, 13>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
    default: mixed_pass_r<T, 16>(ps, src, dst, tw, pitch, tvalid, tid, nthr); break;
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

// One tile.  buf: 2 * tl * pitch complex elements.  `ex.run(f)` runs f(tid, nthr) for every thread of the
// CTA and then a CTA-wide barrier.
template <typename T, typename Exec>
PFB_HD void mixed_tile(const StageParams &sp, cx<T> *buf, long long ibase, long long obase, int tvalid, long long t_is,
                       long long t_os, Exec &ex) {
  const MixedParams &mx = sp.mx;
  const int pitch = mx.pitch, tl = sp.tl, Lc = mx.Lc, L = mx.L;
  cx<T> *const B0 = buf;
  cx<T> *const B1 = buf + (size_t)tl * pitch;
  const cx<T> *const tw = reinterpret_cast<const cx<T> *>(mx.tw);
  const cx<T> *const twh = reinterpret_cast<const cx<T> *>(mx.tw_half);
  const bool swap_in = mx.swap != 0;        // backward transform through the forward butterflies: swap re/im on the way in and out
  const int half = mx.half_real;            // 0: complex line, 1: r2c on n/2 packed points, 2: c2r on n/2 packed points
  const int M = L;                          // (half-length real transforms: L = n/2 complex points)

  // ---- 0. clear what the load does not overwrite
  if (mx.zero_fill) {
    ex.run([&](int tid, int nthr) {
      const cx<T> z{(T)0, (T)0};
      for (unsigned e = (unsigned)tid; e < (unsigned)tvalid * (unsigned)pitch; e += (unsigned)nthr) {
        B0[e] = z;
        if (half == 2) B1[e] = z;      // c2r on packed points: the spectrum line lives in B1
      }
    });
  }

  // ---- 1. load: kLoadBatch independent global loads in flight per thread, then the shared-memory writes
  ex.run([&](int tid, int nthr) {
    constexpr int U = 8;
    const bool contiguous = sp.istride == 1;
    const bool pairs = mx.in_pairs != 0;       // packed r2c line, contiguous and aligned: two reals per load
    const int count = pairs ? sp.nin >> 1 : sp.nin;
    const FastDiv &dcount = pairs ? mx.dnin2 : mx.dnin;
    const unsigned total = (unsigned)(contiguous ? tvalid : tl) * (unsigned)count;
    const bool seg_in = sp.iseg_stride != 0;
    for (unsigned e0 = (unsigned)tid; e0 < total; e0 += (unsigned)nthr * U) {
      MxItem it[U];
      cx<T> val[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned e = e0 + (unsigned)u * (unsigned)nthr;
        it[u].ok = false;
        if (e >= total) continue;
        it[u] = mx_item(e, contiguous, dcount, count, mx.tl_shift, tvalid);
        if (!it[u].ok) continue;
        const int j = it[u].j;
        long long off = ibase + (long long)it[u].tt * t_is;
        if (seg_in) {
          const int seg = (int)fd_div((unsigned)j, mx.diblk);
          off += (long long)seg * sp.iseg_stride + (long long)(j - seg * sp.iblk) * sp.istride;
        } else {
          off += (long long)j * sp.istride;
        }
        if (pairs) val[u] = reinterpret_cast<const cx<T> *>(sp.in)[((off - j) >> 1) + j];   // line base counts reals (even)
        else if (sp.in_real) val[u].x = reinterpret_cast<const T *>(sp.in)[off];
        else val[u] = reinterpret_cast<const cx<T> *>(sp.in)[off];
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (!it[u].ok) continue;
        const int j = it[u].j, tt = it[u].tt;
        if (pairs) {
          cx<T> v = val[u];
          if (sp.mod_in.on) {
            if (mx_sign_mod(sp.mod_in, 2 * j) < 0) v.x = -v.x;
            if (mx_sign_mod(sp.mod_in, 2 * j + 1) < 0) v.y = -v.y;
          }
          B0[(size_t)tt * pitch + mx_phys(j + (sp.zin >> 1))] = v;
          continue;
        }
        const int p = j + sp.zin;          // position inside the zero-padded line
        if (sp.in_real) {
          T xr = val[u].x;
          if (sp.mod_in.on && mx_sign_mod(sp.mod_in, j) < 0) xr = -xr;
          if (half == 1) {
            // two consecutive reals = one complex point
            reinterpret_cast<T *>(B0 + (size_t)tt * pitch + mx_phys(p >> 1))[p & 1] = xr;
          } else if (sp.op == OP_R2R) {
            // w_j x_j exp(-i pi b jj / D), jj = position inside the logical line of n reals
            if ((p == 0 && sp.r2r_half0) || (p == sp.n - 1 && sp.r2r_halfn)) xr *= (T)0.5;
            const cx<T> w = mx_ldg(reinterpret_cast<const cx<T> *>(sp.tw_r2r) + (int)(((long long)2 * sp.r2r_b2 * p) % (8ll * sp.r2r_D)));
            B0[(size_t)tt * pitch + mx_phys(p)] = cx<T>{xr * w.x, xr * w.y};
          } else {
            B0[(size_t)tt * pitch + mx_phys(p)] = swap_in ? cx<T>{(T)0, xr} : cx<T>{xr, (T)0};
          }
        } else {
          cx<T> v = val[u];
          if (sp.conj_in) v.y = -v.y;
          if (sp.mod_in.on && mx_sign_mod(sp.mod_in, j) < 0) { v.x = -v.x; v.y = -v.y; }
          if (half == 2) {
            B1[(size_t)tt * pitch + p] = v;      // Hermitian half spectrum X[0..M], unpadded indexing
          } else {
            if (swap_in) { const T t = v.x; v.x = v.y; v.y = t; }
            B0[(size_t)tt * pitch + mx_phys(p)] = v;
          }
        }
      }
    }
  });

  // ---- 1b. real-line preparation
  if (half == 2) {
    // c2r, n = 2M: x[2j] + i x[2j+1] = sum_k Zf[k] exp(+2 pi i j k / M) with
    //   Zf[k] = (X[k] + conj X[M-k]) + i exp(+2 pi i k / n) (X[k] - conj X[M-k]),  k = 0..M-1
    // (unnormalised, like FFTW's c2r); the backward sum runs through the forward butterflies on swapped parts.
    ex.run([&](int tid, int nthr) {
      const unsigned total = (unsigned)tvalid * (unsigned)M;
      for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
        const int tt = (int)fd_div(e, mx.dL);
        const int k = (int)(e - (unsigned)tt * (unsigned)M);
        const cx<T> *X = B1 + (size_t)tt * pitch;
        cx<T> a = X[k], b = X[M - k];
        if (k == 0) a.y = b.y = (T)0;      // DC and Nyquist bins of a real line are real (FFTW ignores their imaginary parts)
        const cx<T> E{a.x + b.x, a.y - b.y}, D{a.x - b.x, a.y + b.y};
        const cx<T> w = mx_ldg(twh + k);                       // exp(-2 pi i k / n); we need its conjugate
        const cx<T> O{D.x * w.x + D.y * w.y, D.y * w.x - D.x * w.y};
        const cx<T> Z{E.x - O.y, E.y + O.x};
        B0[(size_t)tt * pitch + mx_phys(k)] = cx<T>{Z.y, Z.x};  // swapped: backward through forward butterflies
      }
    });
  } else if (sp.op == OP_C2R) {
    // odd n: complete the spectrum, X[n-k] = conj X[k] (imaginary parts of the DC bin only reach the discarded
    // imaginary output); values are already swapped, so the conjugate mirrors as (-x, y)
    ex.run([&](int tid, int nthr) {
      const int hl = (L - 1) / 2;
      const unsigned total = (unsigned)tvalid * (unsigned)hl;
      for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
        const int tt = (int)(e / (unsigned)hl);
        const int k = (int)(e - (unsigned)tt * (unsigned)hl) + 1;
        cx<T> *X = B0 + (size_t)tt * pitch;
        const cx<T> v = X[mx_phys(k)];
        X[mx_phys(L - k)] = swap_in ? cx<T>{-v.x, v.y} : cx<T>{v.x, -v.y};
      }
    });
  }

  // ---- 2. core transform of length L (directly, or as a Bluestein convolution of length Lc)
  int cur = 0;
  if (mx.bluestein) {
    ex.run([&](int tid, int nthr) {
      const cx<T> *ch = reinterpret_cast<const cx<T> *>(mx.chirp);
      const unsigned total = (unsigned)tvalid * (unsigned)L;
      for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
        const int tt = (int)fd_div(e, mx.dL);
        const int j = (int)(e - (unsigned)tt * (unsigned)L);
        cx<T> *p = B0 + (size_t)tt * pitch + mx_phys(j);
        *p = cxmul(*p, mx_ldg(ch + j));
      }
    });
  }
  for (int rep = 0; rep < (mx.bluestein ? 2 : 1); rep++) {
    for (int ps = 0; ps < mx.npass; ps++) {
      const cx<T> *src = cur ? B1 : B0;
      cx<T> *dst = cur ? B0 : B1;
      ex.run([&](int tid, int nthr) { mixed_pass<T>(mx.pass[ps], src, dst, tw, pitch, tvalid, tid, nthr); });
      cur ^= 1;
    }
    if (mx.bluestein) {
      cx<T> *S = cur ? B1 : B0;
      if (rep == 0) {
        // spectrum of the convolution, conjugated: the inverse transform is conj(FFT(conj(.)))
        ex.run([&](int tid, int nthr) {
          const cx<T> *bh = reinterpret_cast<const cx<T> *>(mx.bhat);
          const unsigned total = (unsigned)tvalid * (unsigned)Lc;
          for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
            const int tt = (int)fd_div(e, mx.dLc);
            const int m = (int)(e - (unsigned)tt * (unsigned)Lc);
            cx<T> *p = S + (size_t)tt * pitch + mx_phys(m);
            const cx<T> y = cxmul(*p, mx_ldg(bh + m));
            *p = cx<T>{y.x, -y.y};
          }
        });
      } else {
        ex.run([&](int tid, int nthr) {
          const cx<T> *ch = reinterpret_cast<const cx<T> *>(mx.chirp);
          const unsigned total = (unsigned)tvalid * (unsigned)L;
          for (unsigned e = (unsigned)tid; e < total; e += (unsigned)nthr) {
            const int tt = (int)fd_div(e, mx.dL);
            const int k = (int)(e - (unsigned)tt * (unsigned)L);
            cx<T> *p = S + (size_t)tt * pitch + mx_phys(k);
            const cx<T> r{p->x, -p->y};
            *p = cxmul(r, mx_ldg(ch + k));
          }
        });
      }
    }
  }
  const cx<T> *const S = cur ? B1 : B0;

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
        const cx<T> r = S[(size_t)tt * pitch + mx_phys(c + (sp.zout >> 1))];
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
      const cx<T> *line = S + (size_t)tt * pitch;
      cx<T> v;
      if (half == 1) {
        // r2c, n = 2M: X[k] = E[k] + exp(-2 pi i k / n) O[k],  E = (Z[k] + conj Z[M-k]) / 2,  O = (Z[k] - conj Z[M-k]) / (2i)
        const cx<T> a = line[mx_phys(k == M ? 0 : k)], b = line[mx_phys(k == 0 ? 0 : M - k)];
        const cx<T> E{(T)0.5 * (a.x + b.x), (T)0.5 * (a.y - b.y)}, D{(T)0.5 * (a.x - b.x), (T)0.5 * (a.y + b.y)};
        const cx<T> O{D.y, -D.x};
        const cx<T> w = mx_ldg(twh + k);
        v = cx<T>{E.x + O.x * w.x - O.y * w.y, E.y + O.x * w.y + O.y * w.x};
      } else if (half == 2) {
        // real output k sits in part k & 1 of packed point k >> 1 (parts swapped: backward through forward butterflies)
        const cx<T> r = line[mx_phys(k >> 1)];
        v = cx<T>{(k & 1) ? r.x : r.y, (T)0};
      } else {
        v = line[mx_phys(k)];
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
