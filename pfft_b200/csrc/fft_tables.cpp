// fft_tables.cpp -- host-side tables for the FFT kernels: radix factorisation and
// twiddle factors.  Twiddles are evaluated in long double after reducing the angle
// to the first octant, so every fp64 entry is correctly rounded to well below 1 ulp;
// the kernels never call sincos (fp64 accuracy target: 1e-12 relative L2 over a
// 1024^3 transform, BASELINE.json north_star).
#include <math.h>
#include <string.h>

#include "kernels.h"

namespace pfb {

int factorize_generic(int L, int *fac) {
  int nf = 0;
  int m = L;
  while (m % 4 == 0) { fac[nf++] = 4; m /= 4; }
  for (int p = 2; p <= 7 && m > 1; p++)
    while (m % p == 0) { fac[nf++] = p; m /= p; }
  for (int p = 11; (long long)p * p <= m; p += 2)
    while (m % p == 0) { fac[nf++] = p; m /= p; }
  if (m > 1) fac[nf++] = m;
  if (nf == 0) fac[nf++] = 1;
  return nf;
}

// exp(-2*pi*i*k/L) with the angle folded into [0, pi/4]
static void unit_root(long long k, long long L, long double *c, long double *s) {
  const long double two_pi = 6.283185307179586476925286766559005768L;
  k %= L;
  if (k < 0) k += L;
  // work in eighths of a turn: position = 8k/L
  long long k8 = 8 * k;
  int oct = (int)(k8 / L);                 // 0..7
  long long rem8 = k8 - (long long)oct * L;   // 0 <= rem8 < L, angle within octant = rem8/(8L) turns
  long double a;
  bool flip = oct & 1;                      // odd octant: measure from the octant's upper edge
  if (flip) a = two_pi * (long double)(L - rem8) / (8.0L * (long double)L);
  else a = two_pi * (long double)rem8 / (8.0L * (long double)L);
  long double ca = cosl(a), sa = sinl(a);
  if (flip) { long double t = ca; ca = sa; sa = t; }   // angle within quadrant = pi/2 - a'
  // now (ca, sa) = (cos, sin) of the angle within the quadrant q = oct/2
  long double cq, sq;
  switch (oct / 2) {
    case 0: cq = ca; sq = sa; break;
    case 1: cq = -sa; sq = ca; break;
    case 2: cq = -ca; sq = -sa; break;
    default: cq = sa; sq = -ca; break;
  }
  *c = cq;
  *s = -sq;   // forward sign: exp(-i theta)
}

void make_twiddles_f64(int L, double *re_im) {
  for (int k = 0; k < L; k++) {
    long double c, s;
    unit_root(k, L, &c, &s);
    re_im[2 * k] = (double)c;
    re_im[2 * k + 1] = (double)s;
  }
}

void make_twiddles_f32(int L, float *re_im) {
  for (int k = 0; k < L; k++) {
    long double c, s;
    unit_root(k, L, &c, &s);
    re_im[2 * k] = (float)c;
    re_im[2 * k + 1] = (float)s;
  }
}

// Fields of a stage every kernel reads (plain copies of the planner's Stage; lengths, windows, strides,
// batch walk, tile table).  Tile geometry (tl, ntiles) is filled in by stage_params_tiles once the kernel
// family has chosen its lines per tile.
bool stage_params_basic(const Stage &g, StageParams &sp, std::string *err) {
  memset(&sp, 0, sizeof sp);
  if (g.n > (1 << 22)) {
    *err = "transform length too large";
    return false;
  }
  int L = (int)g.n;
  if (g.op == OP_R2R) {
    // DCT/DST line of n reals = zero-padded complex DFT of length 2D between two twiddles (kernels.h)
    if (!r2r_params(g.r2r_kind, (int)g.n, &sp.r2r_a2, &sp.r2r_b2, &sp.r2r_D, &sp.r2r_sine, &sp.r2r_half0, &sp.r2r_halfn)) {
      *err = "r2r kind " + std::to_string(g.r2r_kind) + " on length " + std::to_string(g.n) +
             " is not supported (DCT/DST kinds REDFT00..RODFT11 only)";
      return false;
    }
    L = 2 * sp.r2r_D;
  }
  sp.op = g.op;
  sp.sign = g.sign;
  sp.r2r_kind = g.r2r_kind;
  sp.n = (int)g.n;
  sp.L = L;
  sp.nin = (int)g.nin; sp.zin = (int)g.zin; sp.nout = (int)g.nout; sp.zout = (int)g.zout;
  sp.istride = g.istride; sp.iseg_stride = g.iseg_stride; sp.iblk = (int)g.iblk;
  sp.ostride = g.ostride; sp.oblk = (int)g.oblk; sp.noseg = g.noseg;
  sp.nbatch = g.nbatch;
  for (int k = 0; k < g.nbatch; k++) {
    sp.bext[k] = g.batch[k].extent;
    sp.bis[k] = g.batch[k].istride;
    sp.bos[k] = g.batch[k].ostride;
  }
  sp.tile_dim = g.tile_dim;
  sp.iblk2 = (int)g.iblk2; sp.iblk2_stride = g.iblk2_stride;
  sp.oblk2 = (int)g.oblk2; sp.oblk2_stride = g.oblk2_stride;
  sp.ntile = g.ntile;
  sp.iswz_mask = g.iswz_mask; sp.oswz_mask = g.oswz_mask; sp.oswz_shift = g.oswz_shift; sp.oswz_batch = g.oswz_batch;
  for (int l = 0; l < g.ntile; l++) { sp.tile_ioff[l] = g.tile_ioff[l]; sp.tile_ooff[l] = g.tile_ooff[l]; }
  sp.in_real = g.in_real; sp.out_real = g.out_real; sp.conj_in = g.conj_in; sp.conj_out = g.conj_out;
  sp.mod_in = {g.mod_in.on, (int)g.mod_in.start, (int)g.mod_in.half, g.mod_in.extra};
  sp.mod_out = {g.mod_out.on, (int)g.mod_out.start, (int)g.mod_out.half, g.mod_out.extra};
  return true;
}

void stage_params_tiles(const Stage &g, StageParams &sp, int tl) {
  long long lines = 1;
  for (int k = 0; k < g.nbatch; k++) lines *= g.batch[k].extent;
  if (g.in_elems == 0 && g.nin > 0) lines = 0;   // some batch extent of size 1 was dropped but another is 0
  sp.tl = tl;
  if (g.tile_dim >= 0) {
    long long others = 1;
    for (int k = 0; k < g.nbatch; k++)
      if (k != g.tile_dim) others *= g.batch[k].extent;
    sp.tiles_along = (g.batch[g.tile_dim].extent + tl - 1) / tl;
    sp.ntiles = others * sp.tiles_along;
  } else {
    sp.tiles_along = 1;
    sp.ntiles = lines;
  }
  if (g.ntile > 0) {
    // explicit tiles: batch[] enumerates them (`lines` is their number here)
    sp.tiles_along = 1;
    sp.ntiles = lines;
  }
  if (g.nout == 0 || lines == 0) sp.ntiles = 0;
}

unsigned long long &launch_counter() {
  static unsigned long long n = 0;
  return n;
}

// DCT/DST kinds as FFTW defines them (fftw3.h r2r kinds 3..10; reference api/pfft.h:45-55):
//   Y_k = 2 sum_j w_j X_j cos|sin(pi (j + a)(k + b) / D),  w = 1 except the halved end points
//   of the types I and III.  Halfcomplex kinds (R2HC, HC2R, DHT) are out of scope.
bool r2r_params(int kind, int n, int *a2, int *b2, int *D, int *sine, int *half0, int *halfn) {
  *half0 = *halfn = 0;
  switch (kind) {
    case 3: *a2 = 0; *b2 = 0; *D = n - 1; *sine = 0; *half0 = 1; *halfn = 1; break;   // REDFT00 (DCT-I)
    case 4: *a2 = 0; *b2 = 1; *D = n; *sine = 0; *half0 = 1; break;                    // REDFT01 (DCT-III)
    case 5: *a2 = 1; *b2 = 0; *D = n; *sine = 0; break;                                // REDFT10 (DCT-II)
    case 6: *a2 = 1; *b2 = 1; *D = n; *sine = 0; break;                                // REDFT11 (DCT-IV)
    case 7: *a2 = 2; *b2 = 2; *D = n + 1; *sine = 1; break;                            // RODFT00 (DST-I)
    case 8: *a2 = 2; *b2 = 1; *D = n; *sine = 1; *halfn = 1; break;                    // RODFT01 (DST-III)
    case 9: *a2 = 1; *b2 = 2; *D = n; *sine = 1; break;                                // RODFT10 (DST-II)
    case 10: *a2 = 1; *b2 = 1; *D = n; *sine = 1; break;                               // RODFT11 (DST-IV)
    default: return false;
  }
  return *D >= 1;
}

}  // namespace pfb
