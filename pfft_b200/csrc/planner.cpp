// planner.cpp -- turns a Problem into the stage list one rank executes.
//
// What is contractual (and therefore follows the reference): the user-visible
// input and output layouts and the order in which dimensions are transformed,
// embedded (ni -> n, just before that dimension's transform) and truncated
// (n -> no, just after), reference kernel/partrafo-transposed.c:185-291,
// kernel/outrafo.c:68-91,118-168, kernel/ousample.c:208-339, and which mesh
// dimension exchanges which pair of array dimensions, kernel/transpose.c:87-112.
// What is NOT contractual: every intermediate layout.  The reference's are dictated
// by FFTW's transpose interface; here each stage writes the layout its successor
// wants: a stage that precedes an exchange stores its output directly as
// per-destination chunks (split dimension outermost, the dimension about to be
// gathered innermost), and the stage after the exchange reads its lines straight
// out of the received chunks.  No stand-alone pack / unpack / transpose pass exists.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <sstream>

#include "core.h"

namespace pfb {
namespace {

constexpr int kTuple = kMaxDims;   // pseudo dimension id of the howmany-tuple

struct Layout {
  int order[kMaxDims + 1];   // dim ids outer -> inner, tuple last
  int nord = 0;
  INT ext[kMaxDims + 1] = {0};
  INT pitch[kMaxDims + 1] = {0};
  INT stride[kMaxDims + 1] = {0};
  INT total = 0;             // elements of the array (or of one received chunk)
  int seg_dim = -1;          // dimension delivered in chunks by an exchange
  INT seg_blk = 0, seg_stride = 0;
  bool real = false;
  // micro-blocking: index x of dim t sits at (x / mb[t]) * stride[t] + (x % mb[t]) * lstride[t]; the low
  // parts of all blocked dims form one dense block (dims `border`, outer -> inner) that is the
  // innermost unit of the layout.  mb == 1: plain (lstride == stride).
  INT mb[kMaxDims + 1] = {1, 1, 1, 1, 1, 1, 1, 1, 1};
  INT lstride[kMaxDims + 1] = {0};
  int border[kMaxDims + 1] = {0};
  int nbord = 0;
  int swz_mask = 0;          // lines of a block are XOR-permuted by (block index along seg/gather dim) & swz_mask
  int swz_dim = -1;          // the gathered dimension (its high part is the block index)

  void finish() {
    INT s = 1;
    for (int k = nbord - 1; k >= 0; k--) {
      lstride[border[k]] = s;
      s *= mb[border[k]];
    }
    for (int k = nord - 1; k >= 0; k--) {
      const int t = order[k];
      stride[t] = s;
      if (mb[t] == 1) lstride[t] = s;
      s *= pitch[t] / mb[t];
    }
    total = s;
  }
  bool blocked() const { return nbord > 0; }
};

struct HalfSizes {        // sizes one half of the schedule works with (physical units on complex sides)
  bool active = false;
  bool trafo = false;     // false: pure re-distribution (all stages copy)
  Kind kind = Kind::C2C;
  INT ni[kMaxDims], n[kMaxDims], no[kMaxDims];
};

struct Step {
  int dim;
  bool trafo;             // belongs to the transforming half
  int half;               // 0 = forward ("transposed out"), 1 = backward ("transposed in")
  int xch_mesh = -1;      // exchange after this step over this mesh dim
  int xch_split = -1, xch_gather = -1;
  INT split_blk = 0, split_n = 0, gather_blk = 0, gather_n = 0;
  int remap3d = 0;        // 1 / 2: stage of the 3-D mesh remap in front of / behind the pencil schedule
};

struct Builder {
  const Problem &p;
  Schedule &s;
  int d, r;
  int np[kMaxGroups], c[kMaxGroups];   // [0, r): mesh the pencil schedule runs on; then the remap groups
  Mesh3dto2d m3;
  int c3[3] = {0, 0, 0};
  static constexpr int GQ1 = 2, GQ0 = 3;   // group ids of the remap sub-groups (size q1 / q0)
  unsigned tr;
  INT iblk[kMaxMesh], mblk[kMaxMesh], oblk[kMaxMesh];
  HalfSizes hs[2];
  INT cur_ext[kMaxDims + 1];   // current local extents (complex-physical / real as stored)
  // micro-blocked intermediate layouts (plan_microblocks): tile factor of step s along dim t
  bool mbk = false;
  int mbk_order = 2;           // 1: [others][gather][split] (producer writes contiguously), 2: [others][split][gather]
  INT tilef[2 * kMaxDims + 4][kMaxDims + 1];

  Builder(const Problem &pp, Schedule &ss) : p(pp), s(ss) {}

  INT phys(Kind k, const INT *n, int t) const {
    return (t == d - 1 && (k == Kind::R2C || k == Kind::C2R)) ? n[t] / 2 + 1 : n[t];
  }

  bool setup() {
    d = p.rnk_n;
    r = p.rnk_pm;
    tr = p.flags & (F_TRANSPOSED_IN | F_TRANSPOSED_OUT);
    m3 = mesh_3dto2d(p);
    if (d > kMaxDims - 1 || r > kMaxMesh) {
      s.error = "too many dimensions";
      return false;
    }
    for (int t = 0; t < kMaxGroups; t++) { np[t] = 1; c[t] = 0; }
    if (m3.active) {
      // 3-D data on a 3-D mesh P0 x P1 x (q0 q1): the transform runs as a pencil schedule on the
      // 2-D mesh (P0 q0) x (P1 q1), reference kernel/procmesh.c:191-205, with a remap in front / behind
      if (p.flags & F_PADDED_R2C) {
        s.error = "PFFT_PADDED_R2C/C2R is not supported for 3-D data on a 3-D process mesh";
        return false;
      }
      if (p.has_iblock || p.has_oblock) {
        // the reference feeds them to the blocks of the remapped mesh, where they have no meaning
        // (kernel/remap_3dto2d.c:36 "TODO: implement user blocksize"); refuse instead of misplacing data
        s.error = "user block sizes are not supported for 3-D data on a 3-D process mesh";
        return false;
      }
      mesh_coords(3, p.np, s.pid, c3);
      r = 2;
      np[0] = p.np[0] * m3.q0;
      np[1] = p.np[1] * m3.q1;
      c[0] = c3[0] * m3.q0 + c3[2] / m3.q1;
      c[1] = c3[1] * m3.q1 + c3[2] % m3.q1;
      np[GQ1] = m3.q1;
      c[GQ1] = c3[2] % m3.q1;
      np[GQ0] = m3.q0;
      c[GQ0] = c3[2] / m3.q1;
    } else {
      mesh_coords(r, p.np, s.pid, c);
      for (int t = 0; t < r; t++) np[t] = p.np[t];
    }
    for (int t = 0; t < kMaxGroups; t++)
      if (np[t] > kMaxSeg) {
        s.error = "mesh dimension larger than 32";
        return false;
      }
    s.rnk_pm_eff = r;
    for (int t = 0; t < r; t++) {
      s.np_eff[t] = np[t];
      s.coords_eff[t] = c[t];
    }
    // exchange groups as ranks of the user's Cartesian communicator (row-major, last fastest)
    if (m3.active) {
      const int P1 = p.np[1], P2 = p.np[2], q0 = m3.q0, q1 = m3.q1;
      auto rank3 = [&](int a, int b, int cc) { return (a * P1 + b) * P2 + cc; };
      s.ngroups = 4;
      for (int g = 0; g < 4; g++) { s.groups[g].size = np[g]; s.groups[g].me = c[g]; }
      for (int e = 0; e < np[0]; e++) s.groups[0].members[e] = rank3(e / q0, c3[1], (e % q0) * q1 + c3[2] % q1);
      for (int e = 0; e < np[1]; e++) s.groups[1].members[e] = rank3(c3[0], e / q1, (c3[2] / q1) * q1 + e % q1);
      for (int j = 0; j < q1; j++) s.groups[GQ1].members[j] = rank3(c3[0], c3[1], (c3[2] / q1) * q1 + j);
      for (int i = 0; i < q0; i++) s.groups[GQ0].members[i] = rank3(c3[0], c3[1], i * q1 + c3[2] % q1);
    } else {
      s.ngroups = r;
      for (int m = 0; m < r; m++) {
        s.groups[m].size = np[m];
        s.groups[m].me = c[m];
        for (int q = 0; q < np[m]; q++) {
          int rk = 0;
          for (int t = 0; t < r; t++) rk = rk * np[t] + (t == m ? q : c[t]);
          s.groups[m].members[q] = rk;
        }
      }
    }
    // blocks, reference kernel/partrafo.c:652-701
    INT pni[kMaxDims], pno[kMaxDims];
    for (int t = 0; t < d; t++) {
      pni[t] = phys(p.kind, p.ni, t);
      pno[t] = phys(p.kind, p.no, t);
    }
    const INT *pnm = (p.kind == Kind::C2R || (tr & F_TRANSPOSED_IN)) ? pni : pno;
    const INT *mu = nullptr;
    if (tr & F_TRANSPOSED_IN) mu = p.has_iblock ? p.iblock : nullptr;
    if (tr & F_TRANSPOSED_OUT) mu = p.has_oblock ? p.oblock : nullptr;
    const bool ub = !m3.active;   // user block sizes do not apply to the remapped mesh (kernel/remap_3dto2d.c:36)
    if (!ub) mu = nullptr;
    for (int t = 0; t < r; t++) {
      iblk[t] = block_default(pni[t], ub && p.has_iblock && !(tr & F_TRANSPOSED_IN) ? p.iblock[t] : 0, np[t]);
      mblk[t] = block_default(pnm[t + 1], mu ? mu[t] : 0, np[t]);
      oblk[t] = block_default(pno[t], ub && p.has_oblock && !(tr & F_TRANSPOSED_OUT) ? p.oblock[t] : 0, np[t]);
    }
    // the two halves, reference kernel/partrafo.c:735-834
    HalfSizes &to = hs[0], &ti = hs[1];
    to.active = !(tr & F_TRANSPOSED_IN);
    ti.active = !(tr & F_TRANSPOSED_OUT);
    to.kind = ti.kind = p.kind;
    for (int t = 0; t < d; t++) {
      to.ni[t] = p.ni[t]; to.n[t] = p.n[t]; to.no[t] = p.no[t];
      ti.ni[t] = ti.n[t] = ti.no[t] = p.no[t];
    }
    to.trafo = true;
    ti.trafo = false;
    if (tr & F_TRANSPOSED_IN) {
      ti.trafo = true;
      for (int t = 0; t < d; t++) { ti.ni[t] = p.ni[t]; ti.n[t] = p.n[t]; ti.no[t] = p.no[t]; }
    }
    if (p.kind == Kind::R2C) {
      ti.kind = Kind::C2C;
      ti.ni[d - 1] = ti.n[d - 1] = ti.no[d - 1] = p.no[d - 1] / 2 + 1;
    }
    if (p.kind == Kind::C2R) {
      to.trafo = false;
      to.kind = Kind::C2C;
      for (int t = 0; t < d; t++) to.ni[t] = to.n[t] = to.no[t] = p.ni[t];
      to.ni[d - 1] = to.n[d - 1] = to.no[d - 1] = p.ni[d - 1] / 2 + 1;
      ti.trafo = true;
      for (int t = 0; t < d; t++) { ti.ni[t] = p.ni[t]; ti.n[t] = p.n[t]; ti.no[t] = p.no[t]; }
    }
    return true;
  }

  bool user_skip(int dim) const { return p.has_skip && p.skip[std::min(dim, p.rnk_pm)] != 0; }

  bool mesh_trivial() const {
    for (int t = 0; t < r; t++)
      if (np[t] > 1) return false;
    return true;
  }

  // ---- user layouts -------------------------------------------------------------
  Layout user_layout(bool is_input) const {
    const bool transposed = is_input ? (tr & F_TRANSPOSED_IN) : (tr & F_TRANSPOSED_OUT);
    const INT *ln = is_input ? s.ls.lni : s.ls.lno;
    Layout L;
    int k = 0;
    if (transposed) {
      for (int t = 1; t <= r; t++) L.order[k++] = t;
      L.order[k++] = 0;
      for (int t = r + 1; t < d; t++) L.order[k++] = t;
    } else {
      for (int t = 0; t < d; t++) L.order[k++] = t;
    }
    L.order[k++] = kTuple;
    L.nord = k;
    for (int t = 0; t < d; t++) L.ext[t] = L.pitch[t] = ln[t];
    L.ext[kTuple] = L.pitch[kTuple] = p.howmany;
    L.real = p.kind == Kind::R2R || (is_input && p.kind == Kind::R2C) || (!is_input && p.kind == Kind::C2R);
    L.finish();
    return L;
  }

  // ---- one stage ------------------------------------------------------------------
  // Lin/Lout carry extents of every dimension; along `a` Lin.ext = elements present, Lout.ext = kept.
  void emit(const Step &st, const Layout &Lin, const Layout &Lout, const HalfSizes &h, bool first, bool last,
            int noseg, const INT *seg_rows, const INT *Fs = nullptr) {
    Stage g;
    const int a = st.dim;
    const bool do_trafo = st.trafo && !user_skip(a);
    g.dim = a;
    g.in_real = Lin.real;
    g.out_real = Lout.real;
    g.sign = p.sign;
    if (p.kind == Kind::C2R) g.sign = +1;
    if (p.kind == Kind::R2C) g.sign = -1;
    const bool si = p.flags & F_SHIFTED_IN, so = p.flags & F_SHIFTED_OUT;
    if (st.trafo) {
      g.n = h.n[a];
      const bool rlast = a == d - 1;
      if (h.kind == Kind::R2C && rlast) {
        g.op = OP_R2C;
        g.nin = h.ni[a];
        g.zin = si ? (h.n[a] - h.ni[a]) / 2 : 0;
        const INT pn = h.n[a] / 2 + 1, pno = h.no[a] / 2 + 1;
        g.nout = pno;
        g.zout = pn - pno;   // reference kernel/ousample.c:262-272 keeps the upper end
      } else if (h.kind == Kind::C2R && rlast) {
        g.op = OP_C2R;
        const INT pn = h.n[a] / 2 + 1, pni = h.ni[a] / 2 + 1;
        g.nin = pni;
        g.zin = pn - pni;    // reference kernel/ousample.c:292-301
        g.nout = h.no[a];
        g.zout = so ? (h.n[a] - h.no[a]) / 2 : 0;
      } else {
        g.op = h.kind == Kind::R2R ? OP_R2R : OP_C2C;
        g.r2r_kind = p.r2r_kinds[a];
        g.nin = h.ni[a];
        g.zin = si ? (h.n[a] - h.ni[a]) / 2 : 0;
        g.nout = h.no[a];
        g.zout = so ? (h.n[a] - h.no[a]) / 2 : 0;
      }
      if (!do_trafo) {
        if (g.op == OP_R2C || g.op == OP_C2R) s.error = "skipping the real-data dimension of r2c/c2r is not supported";
        g.op = OP_COPY;
      }
      // index-shift modulations, reference api/api-basic.c:1186-1285 (skipped dims untouched)
      if (do_trafo && so) {
        g.mod_in.on = 1;
        g.mod_in.start = si ? -(p.ni[a] / 2) : 0;
        g.mod_in.half = p.ni[a] / 2;
        g.mod_in.extra = (si && (p.n[a] / 2) % 2) ? -1 : 1;
      }
      if (do_trafo && si) {
        g.mod_out.on = 1;
        g.mod_out.start = so ? -(p.no[a] / 2) : 0;
        g.mod_out.half = p.no[a] / 2;
        g.mod_out.extra = 1;
      }
    } else {
      g.op = OP_COPY;
      g.n = g.nin = g.nout = Lin.ext[a];
      g.zin = g.zout = 0;
    }
    if (first && p.kind == Kind::C2R && p.sign == -1) g.conj_in = true;    // reference kernel/partrafo.c:428-443
    if (last && p.kind == Kind::R2C && p.sign == +1) g.conj_out = true;

    if (Lin.seg_dim >= 0 && Lin.seg_dim != a) {
      s.error = "internal: gathered dimension is not the one transformed next";
      return;
    }
    // fft-dim addressing
    g.istride = Lin.lstride[a];
    if (Lin.mb[a] > 1) {
      g.iblk2 = Lin.mb[a];
      g.iblk2_stride = Lin.stride[a];
      if (Lin.swz_dim == a) g.iswz_mask = Lin.swz_mask;
    }
    if (Lin.seg_dim == a) {
      g.iblk = Lin.seg_blk;
      g.iseg_stride = Lin.seg_stride;
    } else {
      g.iblk = std::max<INT>(g.nin, 1);
      g.iseg_stride = 0;
    }
    g.ostride = Lout.lstride[a];
    if (Lout.mb[a] > 1) {
      g.oblk2 = Lout.mb[a];
      g.oblk2_stride = Lout.stride[a];
    }
    if (noseg > 0) {
      g.noseg = noseg;
      g.oblk = st.split_blk;
      INT off = 0;
      for (int q = 0; q < noseg; q++) {
        // (a micro-blocked chunk layout describes ONE chunk; all chunks are equal then)
        g.oseg_cnt[q] = (noseg == 1 || Lout.blocked()) ? Lout.total : seg_rows[q] * Lout.stride[a];
        g.oseg_off[q] = off;
        off += g.oseg_cnt[q];
      }
    } else {
      g.noseg = 1;
      g.oblk = std::max<INT>(g.nout, 1);
      g.oseg_cnt[0] = Lout.total;
    }
    // batch dims: everything except `a`, merged where both layouts allow
    std::vector<BatchDim> b;
    struct TileDim { INT f, ils, ols; };
    std::vector<TileDim> td;
    INT lines = 1;
    for (int k = 0; k < Lin.nord; k++) {
      int t = Lin.order[k];
      if (t == a) continue;
      lines *= Lin.ext[t];
      if (Lin.ext[t] == 1) continue;
      const INT f = Fs ? Fs[t] : 1, mi = Lin.mb[t], mo = Lout.mb[t];
      if (f == 1) {
        if (mi != 1 || mo != 1) { s.error = "internal: micro-blocked dimension outside the tile"; return; }
        b.push_back({Lin.ext[t], Lin.stride[t], Lout.stride[t], t == kTuple ? -1 : t});
        continue;
      }
      // the low part (x % f) selects the line inside the tile, the high part enumerates tiles
      if ((mi != 1 && mi != f) || (mo != 1 && mo != f) || Lin.ext[t] % f) { s.error = "internal: tile / micro-block mismatch"; return; }
      td.push_back({f, Lin.lstride[t], Lout.lstride[t]});
      if (Lin.ext[t] / f > 1 || t == Lout.swz_dim)
        b.push_back({Lin.ext[t] / f, mi == f ? Lin.stride[t] : f * Lin.stride[t], mo == f ? Lout.stride[t] : f * Lout.stride[t], t});
    }
    if (!td.empty()) {
      // line order inside the tile: fastest along the smallest low stride of the blocked side, so that a
      // warp that walks lines first touches consecutive addresses there
      const bool by_out = Lout.blocked();
      std::sort(td.begin(), td.end(), [&](const TileDim &x, const TileDim &y) { return by_out ? x.ols > y.ols : x.ils > y.ils; });
      INT nt = 1;
      for (auto &x : td) nt *= x.f;
      if (nt > kMaxTile) { s.error = "internal: tile too large"; return; }
      g.ntile = (int)nt;
      for (INT l = 0; l < nt; l++) {
        INT rest = l, io = 0, oo = 0;
        for (int q = (int)td.size() - 1; q >= 0; q--) {
          const INT c = rest % td[q].f;
          rest /= td[q].f;
          io += c * td[q].ils;
          oo += c * td[q].ols;
        }
        g.tile_ioff[l] = io;
        g.tile_ooff[l] = oo;
      }
    }
    bool merged = true;
    while (merged) {
      merged = false;
      for (size_t x = 0; x < b.size() && !merged; x++)
        for (size_t y = 0; y < b.size() && !merged; y++) {
          if (x == y) continue;
          if (Lout.swz_mask && (b[x].dim == Lout.swz_dim || b[y].dim == Lout.swz_dim)) continue;   // its coordinate drives the swizzle
          if (b[x].istride == b[y].istride * b[y].extent && b[x].ostride == b[y].ostride * b[y].extent) {
            b[y].extent *= b[x].extent;
            b.erase(b.begin() + x);
            merged = true;
          }
        }
    }
    std::sort(b.begin(), b.end(), [](const BatchDim &x, const BatchDim &y) { return x.istride > y.istride; });
    if ((int)b.size() > kMaxBatch) {
      s.error = "more than 4 independent batch dimensions in one stage";
      return;
    }
    g.nbatch = (int)b.size();
    for (int k = 0; k < g.nbatch; k++) g.batch[k] = b[k];
    // tile dimension: lines of one tile are neighbours along it (coalescing on the strided side)
    g.tile_dim = -1;
    if (g.nbatch > 0) {
      auto pick = [&](bool by_out) {
        int best = 0;
        for (int k = 1; k < g.nbatch; k++) {
          INT sk = by_out ? g.batch[k].ostride : g.batch[k].istride;
          INT sb = by_out ? g.batch[best].ostride : g.batch[best].istride;
          if (sk < sb) best = k;
        }
        return best;
      };
      if (g.ostride != 1) g.tile_dim = pick(true);
      else g.tile_dim = pick(false);
    }
    if (g.ntile > 0) g.tile_dim = -1;
    if (g.ntile > 0 && Lout.swz_mask) {
      // producer side of the swizzle: the block index is the tile's coordinate along the gathered dimension;
      // the permuted bits are the lowest "other dimension" bits of the line id = just above the gathered low part
      for (int k = 0; k < g.nbatch; k++)
        if (g.batch[k].dim == Lout.swz_dim) g.oswz_batch = k;
      if (g.oswz_batch < 0) { s.error = "internal: swizzle dimension is not a batch dimension"; return; }
      g.oswz_mask = Lout.swz_mask;
      INT sh = 0;
      while (((INT)1 << sh) < Lout.mb[Lout.swz_dim]) sh++;
      g.oswz_shift = (int)sh;
    }
    g.in_elems = lines * g.nin;
    g.out_elems = lines * g.nout;
    s.stages.push_back(g);
  }

  // intermediate layout in front of an exchange: split dim outermost, gathered dim
  // innermost (before the always-local trailing dims and the tuple)
  // `mode` 0: as above.  Boundaries without a real exchange (one rank along the mesh
  // dimension) are free to use any order: mode 1 = [others][split][gather] keeps the dimensions
  // that neither neighbouring stage transforms outermost ("planes"), so consecutive lines of the
  // producing stage land within one plane (short strides) and the two stages around the
  // boundary can be run plane by plane through L2 (plan.cu: fused pair); mode 2 = explicit
  // order given by the caller.
  Layout chunk_layout(int split, int gather, INT gather_pitch, bool real, int mode = 0, const int *explicit_order = nullptr) const {
    Layout L;
    int k = 0;
    if (mode == 2) {
      for (int t = 0; t <= r; t++) L.order[k++] = explicit_order[t];
    } else {
      if (mode == 0) L.order[k++] = split;
      for (int t = 0; t <= r; t++)
        if (t != split && t != gather) L.order[k++] = t;
      if (mode == 1) L.order[k++] = split;
      L.order[k++] = gather;
    }
    for (int t = r + 1; t < d; t++) L.order[k++] = t;
    L.order[k++] = kTuple;
    L.nord = k;
    for (int t = 0; t < d; t++) L.ext[t] = L.pitch[t] = cur_ext[t];
    L.ext[kTuple] = L.pitch[kTuple] = p.howmany;
    L.pitch[gather] = gather_pitch;
    L.real = real;
    return L;
  }

  Layout dense_like(const Layout &src, bool real) const {
    Layout L = src;
    for (int t = 0; t < d; t++) L.ext[t] = L.pitch[t] = cur_ext[t];
    L.seg_dim = -1;
    L.real = real;
    L.finish();
    return L;
  }

  // ---- micro-blocked intermediate layouts ------------------------------------------------
  // Measured on B200 (profiles/microbench/runlen.cu): a pass whose strided side moves 128-byte
  // runs reaches 83 % of the contiguous rate (70 % with runs on both sides), 256-byte runs and
  // longer reach 99 %.  A tile of tl lines x L points can only offer tl-element runs to a plain
  // [k][...] layout, so for chains of register-resident power-of-two stages the boundary between
  // stage s and s+1 is stored as dense blocks  [A lo][O lo][G lo]  (A: dimension stage s
  // transformed, G: the one stage s+1 transforms, O: the others) and the tiles of both stages
  // are 2-D: stage s takes (G lo x O lo) lines, stage s+1 takes (A lo x O lo).  Producer tile and
  // consumer tile then share a whole block (tA * tO * tG elements, 512 / 256 bytes at 1024^3 fp64).
  // All ranks must agree, hence only global quantities enter the decision.
  void plan_microblocks(const std::vector<Step> &steps) {
    mbk = false;
    const char *env = getenv("PFFT_B200_BLOCKED");
    const int mode = env ? atoi(env) : 1;
    if (mode <= 0) return;
    mbk_order = 2;   // [others][split][gather]: the consuming tile reads one contiguous piece per source rank
    if (p.kind != Kind::C2C || p.howmany != 1 || m3.active || d != r + 1 || d < 2) return;
    if (tr != F_TRANSPOSED_IN && tr != F_TRANSPOSED_OUT) return;
    if ((int)steps.size() != d || p.has_skip || p.has_iblock || p.has_oblock) return;
    if (p.flags & (F_SHIFTED_IN | F_SHIFTED_OUT)) return;
    int tl = 0;
    for (int t = 0; t < d; t++) {
      const INT L = p.n[t];
      if (p.ni[t] != L || p.no[t] != L || L < 64 || L > 1024 || (L & (L - 1))) return;   // (longer lines: staging + exchange buffers exceed shared memory)
      const int lines = 512 / pow2_threads_per_line((int)L);
      if (tl == 0) tl = lines;
      if (lines != tl) return;
    }
    if (tl < 2 || tl > kMaxTile) return;
    // the tile's lines are spread over ALL dimensions the first stage does not transform, the larger
    // factors on the dimensions transformed first: every boundary then gets blocks of tl * tG elements
    int bits = 0;
    while ((1 << bits) < tl) bits++;
    INT fac[kMaxDims];
    INT fmax = 1;
    for (int q = 0; q < d - 1; q++) {
      fac[q] = (INT)1 << (bits / (d - 1) + (q < bits % (d - 1) ? 1 : 0));
      if (fac[q] < 2) return;
      fmax = std::max(fmax, fac[q]);
    }
    for (int t = 0; t < d; t++)
      for (int m = 0; m < r; m++) {
        if (p.n[t] % ((INT)np[m] * fmax)) return;                                      // whole blocks on every rank
        if ((p.n[t] / np[m]) % pow2_threads_per_line((int)p.n[t])) return;             // chunk boundaries on thread boundaries
      }
    for (size_t i = 0; i < steps.size(); i++) {
      if (!steps[i].trafo) return;
      if ((i + 1 < steps.size()) != (steps[i].xch_mesh >= 0)) return;
      for (int t = 0; t <= kMaxDims; t++) tilef[i][t] = 1;
    }
    // first stage: factors in the order in which the dimensions will be transformed
    for (size_t i = 0; i + 1 < steps.size(); i++) {
      if (steps[i].xch_gather != steps[i + 1].dim) return;
      tilef[0][steps[i].xch_gather] = fac[i];
    }
    // a stage inherits its predecessor's tile with the factor of its own dimension moved to the
    // dimension the predecessor transformed
    for (size_t i = 0; i + 1 < steps.size(); i++) {
      const Step &st = steps[i];
      if (st.dim != st.xch_split || steps[i + 1].dim != st.xch_gather) return;
      for (int t = 0; t <= kMaxDims; t++) tilef[i + 1][t] = tilef[i][t];
      tilef[i + 1][st.xch_split] = tilef[i][st.xch_gather];
      tilef[i + 1][st.xch_gather] = 1;
      if (tilef[i][st.xch_gather] < 2) return;     // (blocks of at least tl * 2 elements on both sides)
    }
    mbk = true;
  }

  bool build() {
    if (!setup()) return false;
    local_block(p, s.pid, &s.ls);
    const HalfSizes &to = hs[0], &ti = hs[1];

    // ---- step list
    std::vector<Step> steps;
    const bool local_only = mesh_trivial() && tr == 0;
    if (local_only) {
      // every "exchange" is the identity: transform in place of the layout, no return trip
      const HalfSizes &h = to.trafo ? to : ti;
      const int half = to.trafo ? 0 : 1;
      if (p.kind == Kind::C2R) for (int a = 0; a < d; a++) steps.push_back({a, true, half});
      else for (int a = d - 1; a >= 0; a--) steps.push_back({a, true, half});
      (void)h;
    } else {
      if (to.active) {
        if (to.trafo) for (int a = d - 1; a > r; a--) steps.push_back({a, true, 0});
        for (int a = r; a >= 0; a--) {
          Step st{a, to.trafo, 0};
          if (a >= 1) {
            const int m = a - 1;
            st.xch_mesh = m;
            st.xch_split = a;
            st.xch_gather = m;
            st.split_blk = mblk[m];
            st.split_n = phys(to.kind, to.no, a);
            st.gather_blk = iblk[m];
            st.gather_n = phys(to.kind, to.ni, m);
          }
          steps.push_back(st);
        }
      }
      if (ti.active) {
        for (int a = 0; a <= r; a++) {
          Step st{a, ti.trafo, 1};
          if (a < r) {
            const int m = a;
            st.xch_mesh = m;
            st.xch_split = a;
            st.xch_gather = a + 1;
            st.split_blk = oblk[m];
            st.split_n = phys(ti.kind, ti.no, a);
            st.gather_blk = mblk[m];
            st.gather_n = phys(ti.kind, ti.ni, a + 1);
          }
          if (a == 0 && to.active) {
            // merge with the forward half's dim-0 step (the reference's PHANTOM stage, kernel/partrafo.c:828-833)
            Step &prev = steps.back();
            const bool was_trafo = prev.trafo;
            prev.xch_mesh = st.xch_mesh; prev.xch_split = st.xch_split; prev.xch_gather = st.xch_gather;
            prev.split_blk = st.split_blk; prev.split_n = st.split_n;
            prev.gather_blk = st.gather_blk; prev.gather_n = st.gather_n;
            if (!was_trafo && st.trafo) { prev.trafo = true; prev.half = 1; }
            continue;
          }
          steps.push_back(st);
        }
        if (ti.trafo) for (int a = r + 1; a < d; a++) steps.push_back({a, true, 1});
      }
    }
    if (m3.active && !local_only) {
      // 3-D blocks <-> pencils of the 2-D mesh, reference kernel/remap_3dto2d.c:219-338: two
      // exchanges inside the sub-groups of size q1 and q0 (block sizes kernel/remap_3dto2d.c:437-457).
      // A stage splits and gathers along the dimension it walks, hence a dense re-layout stage
      // between the two exchanges.
      const int q0 = m3.q0, q1 = m3.q1;
      if (to.active) {
        // sizes of the arriving array: logical reals for r2c input, else as stored
        INT nn[3];
        for (int t = 0; t < 3; t++) nn[t] = (to.kind == Kind::R2C) ? to.ni[t] : phys(to.kind, to.ni, t);
        const INT ob0 = block_default(nn[0], 0, np[0]), ob1 = block_default(nn[1], 0, np[1]);
        const INT ib2 = block_default(nn[2], 0, q0 * q1), mb2 = ib2 * q1;
        const INT l0 = block_extent(nn[0], ob0 * q0, c3[0]), l1 = block_extent(nn[1], ob1 * q1, c3[1]);
        Step a{1, false, 0}, b{2, false, 0}, cst{0, false, 0};
        a.xch_mesh = GQ1; a.xch_split = 1; a.xch_gather = 2;
        a.split_blk = ob1; a.split_n = l1; a.gather_blk = ib2; a.gather_n = block_extent(nn[2], mb2, c[GQ0]);
        cst.xch_mesh = GQ0; cst.xch_split = 0; cst.xch_gather = 2;
        cst.split_blk = ob0; cst.split_n = l0; cst.gather_blk = mb2; cst.gather_n = nn[2];
        a.remap3d = b.remap3d = cst.remap3d = 1;
        steps.insert(steps.begin(), {a, b, cst});
      }
      if (ti.active) {
        INT nn[3];
        for (int t = 0; t < 3; t++) nn[t] = (ti.kind == Kind::C2R) ? ti.no[t] : phys(ti.kind, ti.no, t);
        const INT ob0 = block_default(nn[0], 0, np[0]), ob1 = block_default(nn[1], 0, np[1]);
        const INT ib2 = block_default(nn[2], 0, q0 * q1), mb2 = ib2 * q1;
        const INT l0 = block_extent(nn[0], ob0 * q0, c3[0]), l1 = block_extent(nn[1], ob1 * q1, c3[1]);
        Step &lastti = steps.back();   // walks dimension 2 and has no exchange yet
        if (lastti.dim != 2 || lastti.xch_mesh >= 0) {
          s.error = "internal: unexpected last stage in front of the 2d->3d remap";
          return false;
        }
        lastti.xch_mesh = GQ0; lastti.xch_split = 2; lastti.xch_gather = 0;
        lastti.split_blk = mb2; lastti.split_n = nn[2]; lastti.gather_blk = ob0; lastti.gather_n = l0;
        Step b{0, false, 1}, cst{2, false, 1}, e{1, false, 1};
        cst.xch_mesh = GQ1; cst.xch_split = 2; cst.xch_gather = 1;
        cst.split_blk = ib2; cst.split_n = block_extent(nn[2], mb2, c[GQ0]); cst.gather_blk = ob1; cst.gather_n = l1;
        b.remap3d = cst.remap3d = e.remap3d = 2;
        steps.push_back(b);
        steps.push_back(cst);
        steps.push_back(e);
      }
    }
    if (steps.empty()) {
      s.error = "empty schedule";
      return false;
    }

    plan_microblocks(steps);

    // ---- walk the steps, tracking the layout
    Layout cur = user_layout(true);
    for (int t = 0; t < d; t++) cur_ext[t] = cur.ext[t];
    // extents in complex-physical units where the user counts reals
    bool cur_real = cur.real;
    const Layout Lfinal = user_layout(false);

    int trafo_slot = 0;
    for (size_t i = 0; i < steps.size(); i++) {
      const Step &st = steps[i];
      const HalfSizes &h = hs[st.half];
      const int a = st.dim;
      const bool first = i == 0, last = i + 1 == steps.size();
      // extent of `a` on input / output of this step
      Layout Lin = cur;
      bool out_real = cur_real;
      INT out_ext;
      if (st.trafo) {
        const bool rlast = a == d - 1;
        if (h.kind == Kind::R2C && rlast) out_real = false;
        if (h.kind == Kind::C2R && rlast) out_real = true;
        if (h.kind == Kind::R2C && rlast) {
          Lin.ext[a] = h.ni[a];                 // logical reals present in a row (pitch may be padded)
          out_ext = h.no[a] / 2 + 1;
        } else if (h.kind == Kind::C2R && rlast) {
          out_ext = h.no[a];
        } else {
          out_ext = phys(h.kind, h.no, a);
        }
      } else {
        out_ext = Lin.ext[a];
      }
      cur_ext[a] = out_ext;

      Layout Lout;
      int noseg = 0;
      INT seg_rows[kMaxSeg] = {0};
      if (last) {
        Lout = Lfinal;
        if (Lout.real != out_real) { s.error = "internal: element type mismatch at the last stage"; return false; }
      } else if (st.xch_mesh >= 0) {
        const int m = st.xch_mesh;
        // local boundaries in front of / inside the last pair of stages: plane-friendly orders
        int mode = 0;
        int ord[kMaxDims + 1];
        const size_t ns = steps.size();
        const bool simple = d == r + 1 && p.howmany == 1 && ns >= 2;
        const bool pair_local = simple && steps[ns - 2].xch_mesh >= 0 && np[steps[ns - 2].xch_mesh] == 1;
        if (np[m] == 1 && simple) {
          if (i + 2 == ns) {
            mode = 1;
          } else if (i + 3 == ns && pair_local) {
            // input of the pair = final layout with the pair's two dimensions swapped: plane p of
            // the pair's input then occupies exactly the addresses of plane p of the final output
            const int x = steps[ns - 2].dim, y = steps[ns - 1].dim;
            int k = 0;
            for (int q = 0; q < Lfinal.nord; q++) {
              const int t = Lfinal.order[q];
              if (t == kTuple || t > r) continue;
              ord[k++] = t == x ? y : (t == y ? x : t);
            }
            if (k == r + 1 && ord[r] == st.xch_gather) mode = 2;
            else mode = 1;
          } else {
            mode = 1;
          }
        }
        if (mbk) {
          // outer order: everything but the pair first, then [gather][split] (the producing tile writes one
          // contiguous piece per destination) or [split][gather] (the consuming tile reads one)
          int k = 0;
          for (int t = 0; t <= r; t++)
            if (t != st.xch_split && t != st.xch_gather) ord[k++] = t;
          ord[k++] = mbk_order == 1 ? st.xch_gather : st.xch_split;
          ord[k++] = mbk_order == 1 ? st.xch_split : st.xch_gather;
          mode = 2;
        }
        Lout = chunk_layout(st.xch_split, st.xch_gather, st.gather_blk, out_real, mode, ord);
        if (mbk) {
          // one chunk = the rows of one destination (uniform by construction)
          Lout.pitch[st.xch_split] = st.split_blk;
          const INT *Fp = tilef[i], *Fc = tilef[i + 1];
          Lout.nbord = 0;
          auto add = [&](int t, INT f) {
            Lout.mb[t] = f;
            if (f > 1) Lout.border[Lout.nbord++] = t;
          };
          add(st.xch_split, Fc[st.xch_split]);
          for (int t = 0; t <= r; t++)
            if (t != st.xch_split && t != st.xch_gather) {
              if (Fp[t] != Fc[t]) { s.error = "internal: tile factors of neighbouring stages differ"; return false; }
              add(t, Fp[t]);
            }
          add(st.xch_gather, Fp[st.xch_gather]);
          // bank swizzle: a quarter-warp of the consumer reads 8 / tG consecutive blocks (16-byte words)
          INT others = 1;
          for (int t = 0; t <= r; t++)
            if (t != st.xch_split && t != st.xch_gather) others *= Fp[t];
          const INT tG = Fp[st.xch_gather];
          const INT want = tG < 8 ? 8 / tG : 1;
          Lout.swz_mask = (int)(std::min(want, others) - 1);
          Lout.swz_dim = st.xch_gather;
        }
        Lout.finish();
        noseg = np[m];
        for (int q = 0; q < noseg; q++) seg_rows[q] = block_extent(st.split_n, st.split_blk, q);
        if (out_ext != st.split_n) { s.error = "internal: split extent mismatch"; return false; }
      } else {
        Lout = dense_like(Lin, out_real);
        // keep the input's memory order; a C2R/R2C row changes element type, dense rows
      }
      if (last && out_ext != Lout.ext[a] && !(Lout.real && a == d - 1)) {
        // the user's output block must hold exactly what the last stage keeps
        s.error = "internal: output extent mismatch";
        return false;
      }
      if (last && Lout.real && a == d - 1) {
        // real rows may be padded (PFFT_PADDED_C2R): extent kept is `no`, pitch is the user's row length
        Lout.ext[a] = out_ext;
      }
      emit(st, Lin, Lout, h, first, last, noseg, seg_rows, mbk ? tilef[i] : nullptr);
      if (!s.error.empty()) return false;
      Stage &g = s.stages.back();
      g.timer_slot = st.remap3d ? -st.remap3d : trafo_slot++;

      if (st.xch_mesh >= 0 && !last) {
        const int m = st.xch_mesh;
        Exchange x;
        x.mesh_dim = m;
        x.nparts = np[m];
        x.me = c[m];
        x.elem_real = out_real;
        for (int q = 0; q < np[m]; q++) {
          x.send_cnt[q] = g.oseg_cnt[q];
          x.peer_recv_cnt[q] = g.oseg_cnt[q];
        }
        x.recv_cnt = g.oseg_cnt[c[m]];
        g.exchange = (int)s.exchanges.size();
        s.exchanges.push_back(x);
        // layout the next stage reads: my rows of the split dim, the gathered dim whole (in chunks)
        cur = Lout;
        cur_ext[st.xch_split] = seg_rows[c[m]];
        cur.ext[st.xch_split] = cur.pitch[st.xch_split] = seg_rows[c[m]];
        cur.finish();
        cur.seg_dim = st.xch_gather;
        cur.seg_blk = st.gather_blk;
        cur.seg_stride = cur.total;
        cur.ext[st.xch_gather] = st.gather_n;
        cur_ext[st.xch_gather] = st.gather_n;
      } else {
        cur = Lout;
        cur.seg_dim = -1;
      }
      cur_real = out_real;
    }

    // buffers: in -> A -> B -> A ... -> out
    INT scratch = 0;
    for (size_t i = 0; i < s.stages.size(); i++) {
      Stage &g = s.stages[i];
      g.in_buf = i == 0 ? BUF_USER_IN : s.stages[i - 1].out_buf;
      g.out_buf = i + 1 == s.stages.size() ? BUF_USER_OUT : (i % 2 == 0 ? BUF_A : BUF_B);
      if (i + 1 < s.stages.size()) {
        INT sendside = 0, recvside = 0;
        for (int q = 0; q < g.noseg; q++) sendside += g.oseg_cnt[q];
        if (g.exchange >= 0) {
          const Exchange &x = s.exchanges[g.exchange];
          recvside = x.recv_cnt * x.nparts;
        }
        const INT scal = g.out_real ? 1 : 2;
        scratch = std::max(scratch, std::max(sendside, recvside) * scal);
      }
    }
    s.scratch_elems = scratch;
    return true;
  }
};

}  // namespace

bool boundary_is_remote(const Schedule &s, int i) {
  if (i < 0 || i + 1 >= (int)s.stages.size()) return false;
  const int x = s.stages[i].exchange;
  return x >= 0 && x < (int)s.exchanges.size() && s.exchanges[x].nparts > 1;
}

std::vector<ExchangeWait> exchange_waits(const Schedule &s, const std::vector<int> &assign, int i) {
  std::vector<ExchangeWait> w;
  auto add_group = [&](int xi, int back, int code) {
    const Exchange &x = s.exchanges[xi];
    for (int q = 0; q < x.nparts; q++)
      if (q != x.me) w.push_back(ExchangeWait{s.groups[x.mesh_dim].members[q], back, code});
  };
  const int nb = (int)s.stages.size() - 1;      // boundaries
  if (i > 0 && boundary_is_remote(s, i - 1)) add_group(s.stages[i - 1].exchange, 0, i);
  if (i < nb && boundary_is_remote(s, i) && i < (int)assign.size()) {
    const int b = assign[i];
    int j = -1;
    for (int k = i - 1; k >= 0; k--)
      if (assign[k] == b) { j = k; break; }
    if (j >= 0) {
      add_group(s.stages[i].exchange, 0, j + 2);
    } else {
      for (int k = std::min(nb, (int)assign.size()) - 1; k >= 0; k--)
        if (assign[k] == b) { j = k; break; }
      add_group(s.stages[i].exchange, 1, j + 2);
    }
  }
  return w;
}

int pow2_points_per_thread(int L) {
  // three radix passes at most: 64 -> 8*8, 128 -> 8*8*2, 256 -> 16*16, 512 -> 8*8*8, 1024 -> 16*16*4, ...
  return (L == 64 || L == 128 || L == 512) ? 8 : 16;
}

bool build_schedule(const Problem &p, int pid, Schedule *sched) {
  sched->prob = p;
  sched->pid = pid;
  sched->stages.clear();
  sched->exchanges.clear();
  sched->error.clear();
  std::string why;
  if (!problem_is_legal(p, &why)) {
    sched->error = "illegal: " + why;
    return false;
  }
  Builder b(p, *sched);
  if (!b.build()) {
    if (sched->error.empty()) sched->error = "unsupported configuration";
    return false;
  }
  return sched->error.empty();
}

std::string schedule_to_json(const Schedule &s) {
  std::ostringstream o;
  auto arr = [&](const INT *v, int n) {
    o << "[";
    for (int i = 0; i < n; i++) o << (i ? "," : "") << v[i];
    o << "]";
  };
  const int d = s.prob.rnk_n;
  o << "{\"pid\":" << s.pid << ",\"rnk_n\":" << d << ",\"error\":\"" << s.error << "\"";
  o << ",\"local_ni\":"; arr(s.ls.lni, d);
  o << ",\"local_i_start\":"; arr(s.ls.lis, d);
  o << ",\"local_no\":"; arr(s.ls.lno, d);
  o << ",\"local_o_start\":"; arr(s.ls.los, d);
  o << ",\"scratch_elems\":" << s.scratch_elems;
  o << ",\"coords\":[";
  for (int t = 0; t < s.rnk_pm_eff; t++) o << (t ? "," : "") << s.coords_eff[t];
  o << "],\"np\":[";
  for (int t = 0; t < s.rnk_pm_eff; t++) o << (t ? "," : "") << s.np_eff[t];
  o << "],\"groups\":[";
  for (int g = 0; g < s.ngroups; g++) {
    o << (g ? "," : "") << "{\"size\":" << s.groups[g].size << ",\"me\":" << s.groups[g].me << ",\"members\":[";
    for (int q = 0; q < s.groups[g].size; q++) o << (q ? "," : "") << s.groups[g].members[q];
    o << "]}";
  }
  o << "],\"stages\":[";
  for (size_t i = 0; i < s.stages.size(); i++) {
    const Stage &g = s.stages[i];
    o << (i ? "," : "") << "{\"op\":" << g.op << ",\"sign\":" << g.sign << ",\"r2r_kind\":" << g.r2r_kind
      << ",\"dim\":" << g.dim << ",\"n\":" << g.n << ",\"nin\":" << g.nin << ",\"zin\":" << g.zin
      << ",\"nout\":" << g.nout << ",\"zout\":" << g.zout << ",\"istride\":" << g.istride
      << ",\"iblk\":" << g.iblk << ",\"iseg_stride\":" << g.iseg_stride << ",\"ostride\":" << g.ostride
      << ",\"oblk\":" << g.oblk << ",\"noseg\":" << g.noseg << ",\"iblk2\":" << g.iblk2 << ",\"iblk2_stride\":" << g.iblk2_stride
      << ",\"oblk2\":" << g.oblk2 << ",\"oblk2_stride\":" << g.oblk2_stride << ",\"ntile\":" << g.ntile << ",\"iswz_mask\":" << g.iswz_mask << ",\"oswz_mask\":" << g.oswz_mask
      << ",\"oswz_shift\":" << g.oswz_shift << ",\"oswz_batch\":" << g.oswz_batch << ",\"tile_ioff\":";
    arr(g.tile_ioff, g.ntile);
    o << ",\"tile_ooff\":";
    arr(g.tile_ooff, g.ntile);
    o << ",\"oseg_off\":";
    arr(g.oseg_off, g.noseg);
    o << ",\"oseg_cnt\":";
    arr(g.oseg_cnt, g.noseg);
    o << ",\"batch\":[";
    for (int k = 0; k < g.nbatch; k++)
      o << (k ? "," : "") << "[" << g.batch[k].extent << "," << g.batch[k].istride << "," << g.batch[k].ostride
        << "," << g.batch[k].dim << "]";
    o << "],\"tile_dim\":" << g.tile_dim << ",\"in_real\":" << (int)g.in_real << ",\"out_real\":" << (int)g.out_real
      << ",\"conj_in\":" << (int)g.conj_in << ",\"conj_out\":" << (int)g.conj_out << ",\"mod_in\":[" << g.mod_in.on
      << "," << g.mod_in.start << "," << g.mod_in.half << "," << g.mod_in.extra << "],\"mod_out\":[" << g.mod_out.on
      << "," << g.mod_out.start << "," << g.mod_out.half << "," << g.mod_out.extra << "],\"in_buf\":" << g.in_buf
      << ",\"out_buf\":" << g.out_buf << ",\"exchange\":" << g.exchange << ",\"in_elems\":" << g.in_elems
      << ",\"out_elems\":" << g.out_elems << ",\"timer_slot\":" << g.timer_slot << "}";
  }
  o << "],\"exchanges\":[";
  for (size_t i = 0; i < s.exchanges.size(); i++) {
    const Exchange &x = s.exchanges[i];
    o << (i ? "," : "") << "{\"mesh_dim\":" << x.mesh_dim << ",\"nparts\":" << x.nparts << ",\"me\":" << x.me
      << ",\"recv_cnt\":" << x.recv_cnt << ",\"elem_real\":" << (int)x.elem_real << ",\"send_cnt\":";
    arr(x.send_cnt, x.nparts);
    o << "}";
  }
  o << "]}";
  return o.str();
}

}  // namespace pfb
