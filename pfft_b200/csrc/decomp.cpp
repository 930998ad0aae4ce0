// decomp.cpp -- block decomposition of a transform over a process mesh.
// Must be bit-exact with the reference's integer layer:
//   kernel/block.c:54-95 (blocks), util/util.c:43-59,207-227 (physical sizes),
//   kernel/partrafo.c:99-199,652-701,735-834 (local blocks of a parallel transform),
//   kernel/partrafo-transposed.c:67-85,341-411 (standard / transposed layouts),
//   kernel/procmesh.c:191-198,367-391 and kernel/remap_3dto2d.c:437-457 (3-D mesh).
// Checked against golden vectors captured from the reference's own code
// (tests/golden/local_block.json).
#include <math.h>

#include <algorithm>

#include "core.h"

namespace pfb {

INT block_count(INT n, INT blk) { return blk > 0 ? (n + blk - 1) / blk : 0; }

INT block_default(INT n, INT user_blk, int nprocs) {
  return user_blk == 0 ? (n + nprocs - 1) / nprocs : user_blk;
}

INT block_extent(INT n, INT blk, int which) {
  INT nb = block_count(n, blk);
  if (which >= nb) return 0;
  return which == nb - 1 ? n - (INT)which * blk : blk;
}

INT block_offset(INT n, INT blk, int which) {
  return which >= block_count(n, blk) ? 0 : (INT)which * blk;
}

void mesh_coords(int rnk_pm, const int *np, int pid, int *coords) {
  for (int t = rnk_pm - 1; t >= 0; t--) {
    coords[t] = pid % np[t];
    pid /= np[t];
  }
}

// Search for q0*q1 == q that makes p0*q0 and p1*q1 as equal as possible.  The
// reference starts from (1, q) but measures that start with the error of (q, 1)
// (kernel/procmesh.c:367-391); the quirk is part of the contract.
static void split_third_mesh_dim(int p0, int p1, int q, int *q0_out, int *q1_out) {
  int best0 = 1, best1 = q;
  double best = fabs((double)p0 * q - (double)p1);
  for (int q1 = 1; q1 <= sqrt((double)q); q1++) {
    int q0 = q / q1;
    if (q0 * q1 != q) continue;
    double err = fabs((double)(p0 * q0 - p1 * q1));
    if (err < best) {
      best = err;
      best0 = q0;
      best1 = q1;
    }
  }
  *q0_out = best0;
  *q1_out = best1;
}

Mesh3dto2d mesh_3dto2d(const Problem &p) {
  Mesh3dto2d m;
  if (p.rnk_n == 3 && p.rnk_pm == 3) {
    m.active = true;
    split_third_mesh_dim(p.np[0], p.np[1], p.np[2], &m.q0, &m.q1);
  }
  return m;
}

bool problem_is_legal(const Problem &p, std::string *why) {
  auto fail = [&](const char *msg) {
    if (why) *why = msg;
    return false;
  };
  const unsigned tr = p.flags & (F_TRANSPOSED_IN | F_TRANSPOSED_OUT);
  // limits of this implementation (fixed-size arrays in Problem / LocalSizes); the reference takes any rnk_pm < rnk_n
  if (p.rnk_n < 1 || p.rnk_n > kMaxDims - 1) return fail("rnk_n outside 1..7 is not supported");
  if (p.rnk_pm < 1 || p.rnk_pm > kMaxMesh) return fail("process meshes of more than 3 dimensions are not supported");
  if (p.flags & (F_SHIFTED_IN | F_SHIFTED_OUT))
    for (int t = 0; t < p.rnk_n; t++)
      if (p.n[t] % 2) return fail("index shift needs even n");
  if (tr == (F_TRANSPOSED_IN | F_TRANSPOSED_OUT)) return fail("TRANSPOSED_IN and TRANSPOSED_OUT in one plan");
  if (p.kind == Kind::R2C && (tr & F_TRANSPOSED_IN)) return fail("r2c with TRANSPOSED_IN");
  if (p.kind == Kind::C2R && (tr & F_TRANSPOSED_OUT)) return fail("c2r with TRANSPOSED_OUT");
  if (p.rnk_n < p.rnk_pm) return fail("rnk_n < rnk_pm");
  if (p.rnk_n == p.rnk_pm && p.rnk_n != 3) return fail("rnk_n == rnk_pm only for 3d");
  return true;
}

namespace {

void physical(const Problem &p, Kind kind, const INT *n, INT *pn) {
  for (int t = 0; t < p.rnk_n; t++) pn[t] = n[t];
  if (kind == Kind::R2C || kind == Kind::C2R) pn[p.rnk_n - 1] = n[p.rnk_n - 1] / 2 + 1;
}

struct Blocks {
  INT iblk[kMaxMesh] = {0}, mblk[kMaxMesh] = {0}, oblk[kMaxMesh] = {0};
};

// kernel/partrafo.c:652-701 -- blocks come from the PHYSICAL sizes; the transposed
// ("middle") layout distributes dims 1..r, hence the shift by one.
Blocks evaluate_blocks(const Problem &p, int r, const int *np_pm) {
  Blocks b;
  INT pni[kMaxDims], pno[kMaxDims];
  physical(p, p.kind, p.ni, pni);
  physical(p, p.kind, p.no, pno);
  const unsigned tr = p.flags & (F_TRANSPOSED_IN | F_TRANSPOSED_OUT);
  const INT *pnm = (p.kind == Kind::C2R || (tr & F_TRANSPOSED_IN)) ? pni : pno;
  const INT *mblock_user = nullptr;
  if (tr & F_TRANSPOSED_IN) mblock_user = p.has_iblock ? p.iblock : nullptr;
  if (tr & F_TRANSPOSED_OUT) mblock_user = p.has_oblock ? p.oblock : nullptr;
  for (int t = 0; t < r; t++) {
    if (!(tr & F_TRANSPOSED_IN)) b.iblk[t] = block_default(pni[t], p.has_iblock ? p.iblock[t] : 0, np_pm[t]);
    b.mblk[t] = block_default(pnm[t + 1], mblock_user ? mblock_user[t] : 0, np_pm[t]);
    if (!(tr & F_TRANSPOSED_OUT)) b.oblk[t] = block_default(pno[t], p.has_oblock ? p.oblock[t] : 0, np_pm[t]);
  }
  return b;
}

// kernel/partrafo-transposed.c:363-411
void decompose(const Problem &p, Kind kind, const INT *n, const INT *blk, int r, const int *coords,
               bool transposed, INT *ln, INT *ls) {
  INT pn[kMaxDims];
  physical(p, kind, n, pn);
  for (int t = 0; t < p.rnk_n; t++) {
    ln[t] = pn[t];
    ls[t] = 0;
  }
  const int off = transposed ? 1 : 0;
  for (int t = 0; t < r; t++) {
    ln[t + off] = block_extent(pn[t + off], blk[t], coords[t]);
    ls[t + off] = block_offset(pn[t + off], blk[t], coords[t]);
  }
}

// The user interface counts REALS in the last dim of the real side of r2c / c2r
// (kernel/partrafo-transposed.c:67-85).
void real_side_count(const Problem &p, Kind kind, bool is_input, const INT *n, INT *ln) {
  const bool padded = p.flags & F_PADDED_R2C;
  if ((kind == Kind::R2C && is_input) || (kind == Kind::C2R && !is_input)) {
    const int l = p.rnk_n - 1;
    ln[l] = padded ? ln[l] * 2 : n[l];
  }
}

// kernel/remap_3dto2d.c:437-457
void blocks_3dto2d(const INT *n, int p0, int p1, int q0, int q1, INT *iblk, INT *mblk, INT *oblk) {
  oblk[0] = block_default(n[0], 0, p0 * q0);
  oblk[1] = block_default(n[1], 0, p1 * q1);
  oblk[2] = n[2];
  iblk[0] = oblk[0] * q0;
  iblk[1] = oblk[1] * q1;
  iblk[2] = block_default(n[2], 0, q0 * q1);
  mblk[0] = oblk[0] * q0;
  mblk[1] = oblk[1];
  mblk[2] = iblk[2] * q1;
}

}  // namespace

void local_block(const Problem &p, int pid, LocalSizes *out) {
  const int d = p.rnk_n;
  const unsigned tr = p.flags & (F_TRANSPOSED_IN | F_TRANSPOSED_OUT);
  const Mesh3dto2d m3 = mesh_3dto2d(p);
  int r = p.rnk_pm;
  int np_pm[kMaxMesh] = {1, 1, 1}, coords[kMaxMesh] = {0, 0, 0}, c3[kMaxMesh] = {0, 0, 0};
  if (m3.active) {
    mesh_coords(3, p.np, pid, c3);
    coords[0] = c3[0] * m3.q0 + c3[2] / m3.q1;
    coords[1] = c3[1] * m3.q1 + c3[2] % m3.q1;
    np_pm[0] = p.np[0] * m3.q0;
    np_pm[1] = p.np[1] * m3.q1;
    r = 2;
  } else {
    mesh_coords(r, p.np, pid, coords);
    for (int t = 0; t < r; t++) np_pm[t] = p.np[t];
  }
  const Blocks b = evaluate_blocks(p, r, np_pm);

  // sizes / kinds of the "transposed-out" (to) and "transposed-in" (ti) halves,
  // kernel/partrafo.c:735-834
  INT ni_to[kMaxDims], no_to[kMaxDims], ni_ti[kMaxDims], no_ti[kMaxDims];
  Kind kind_to = p.kind, kind_ti = p.kind;
  for (int t = 0; t < d; t++) {
    ni_to[t] = p.ni[t];
    no_to[t] = p.no[t];
    ni_ti[t] = p.no[t];
    no_ti[t] = p.no[t];
  }
  if (tr & F_TRANSPOSED_IN)
    for (int t = 0; t < d; t++) ni_ti[t] = p.ni[t];
  if (p.kind == Kind::R2C) {
    for (int t = 0; t < d; t++) ni_ti[t] = no_ti[t] = p.no[t];
    ni_ti[d - 1] = no_ti[d - 1] = p.no[d - 1] / 2 + 1;
    kind_ti = Kind::C2C;
  }
  if (p.kind == Kind::C2R) {
    for (int t = 0; t < d; t++) ni_to[t] = no_to[t] = p.ni[t];
    ni_to[d - 1] = no_to[d - 1] = p.ni[d - 1] / 2 + 1;
    kind_to = Kind::C2C;
    for (int t = 0; t < d; t++) {
      ni_ti[t] = p.ni[t];
      no_ti[t] = p.no[t];
    }
  }

  INT an[kMaxDims], as[kMaxDims], bn[kMaxDims], bs[kMaxDims];
  INT *lni = out->lni, *lis = out->lis, *lno = out->lno, *los = out->los;
  for (int t = 0; t < d; t++) lni[t] = lis[t] = lno[t] = los[t] = 0;
  auto copy = [&](const INT *src, INT *dst) { std::copy(src, src + d, dst); };

  if (!(tr & F_TRANSPOSED_IN)) {
    decompose(p, kind_to, ni_to, b.iblk, r, coords, false, an, as);
    decompose(p, kind_to, no_to, b.mblk, r, coords, true, bn, bs);
    real_side_count(p, kind_to, true, ni_to, an);
    real_side_count(p, kind_to, false, no_to, bn);
    copy(an, lni);
    copy(as, lis);
    if (tr & F_TRANSPOSED_OUT) {
      copy(bn, lno);
      copy(bs, los);
    }
    if (m3.active) {
      // the remap works on logical sizes; r2c input is handled like r2r (kernel/remap_3dto2d.c:92-94)
      INT ib3[3], mb3[3], ob3[3];
      blocks_3dto2d(ni_to, p.np[0], p.np[1], m3.q0, m3.q1, ib3, mb3, ob3);
      for (int t = 0; t < 3; t++) {
        lni[t] = block_extent(ni_to[t], ib3[t], c3[t]);
        lis[t] = block_offset(ni_to[t], ib3[t], c3[t]);
      }
    }
  }
  if (!(tr & F_TRANSPOSED_OUT)) {
    decompose(p, kind_ti, ni_ti, b.mblk, r, coords, true, an, as);
    decompose(p, kind_ti, no_ti, b.oblk, r, coords, false, bn, bs);
    real_side_count(p, kind_ti, true, ni_ti, an);
    real_side_count(p, kind_ti, false, no_ti, bn);
    copy(bn, lno);
    copy(bs, los);
    if (tr & F_TRANSPOSED_IN) {
      copy(an, lni);
      copy(as, lis);
    }
    if (m3.active) {
      INT ib3[3], mb3[3], ob3[3];
      blocks_3dto2d(no_ti, p.np[0], p.np[1], m3.q0, m3.q1, ib3, mb3, ob3);
      for (int t = 0; t < 3; t++) {
        lno[t] = block_extent(no_ti[t], ib3[t], c3[t]);
        los[t] = block_offset(no_ti[t], ib3[t], c3[t]);
      }
    }
  }
  // index shift: starts move by -n/2 (kernel/partrafo.c:178-190)
  if (p.flags & F_SHIFTED_IN)
    for (int t = 0; t < d; t++) lis[t] -= p.ni[t] / 2;
  if (p.flags & F_SHIFTED_OUT)
    for (int t = 0; t < d; t++) los[t] -= p.no[t] / 2;
}

INT alloc_local(const Problem &p, int pid) {
  // The reference returns the maximum footprint over its own (FFTW-dictated) stage
  // arrays, kernel/partrafo-transposed.c:118-155 -- a number that depends on
  // fftw_mpi_local_size_many_transposed and is not contractual.  Here the
  // user arrays only ever hold the input block and the output block (intermediates
  // live in plan-owned device scratch), so: max of both, in complex elements for
  // c2c/r2c/c2r (real rows are padded to 2*(n/2+1) like the reference's in-place
  // embed needs) and real elements for r2r; never 0 (kernel/transpose.c:122,136).
  LocalSizes ls;
  local_block(p, pid, &ls);
  const int d = p.rnk_n;
  auto prod_complex = [&](const INT *ln, bool real_side, const INT *nlog) {
    INT m = p.howmany;
    for (int t = 0; t < d; t++) {
      INT e = ln[t];
      if (t == d - 1 && real_side) e = nlog[t] / 2 + 1;   // reals -> complex pairs, padded row
      m *= e;
    }
    return m;
  };
  INT a = prod_complex(ls.lni, p.kind == Kind::R2C, p.ni);
  INT b = prod_complex(ls.lno, p.kind == Kind::C2R, p.no);
  // pruned r2c/c2r rows are embedded to n before the transform in the reference
  if (p.kind == Kind::R2C) a = std::max(a, prod_complex(ls.lni, true, p.n));
  if (p.kind == Kind::C2R) b = std::max(b, prod_complex(ls.lno, true, p.n));
  return std::max<INT>(std::max(a, b), 1);
}

}  // namespace pfb
