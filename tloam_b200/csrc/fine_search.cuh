// fine_search.cuh -- exact radius-truncated kNN on the TWO-LEVEL grid, for maps that are dense relative to the search
// radius (BASELINE config 3: ~400 points in a cell of edge r, so the plain 27-cell search streams ~4 000 candidates
// per query although the K nearest ones lie within a few centimetres).
//
// Level 1 is the brick table of map_grid.cuh (cell edge == search radius).  Level 2 exists for the cells with at least
// kFineMin points only (k_map_fine, map_build.cuh): their points are ordered by a 4 x 4 x 4 grid of fine bins and a
// 128-byte table per dense cell holds the bin boundaries.  Cells below the threshold stay unordered.
//
// The search is an expanding-box scheme.  The 3 x 3 x 3 cell neighbourhood is a box of 12 x 12 x 12 fine bins; a
// pass visits every segment (row of fine bins of a dense cell, or a whole sparse cell) that intersects the box
// [q - w, q + w]^3 and was not reached by an earlier pass, skipping segments whose distance lower bound exceeds the
// current K-th best.  w starts at one bin edge (r / 4); after a pass the search is complete iff the K-th best distance
// is <= w (everything closer lies inside the box); otherwise w becomes that distance (list full) or doubles (list not
// full), up to r.  Same result as knn_search (exact, ordered by (d2, original index)): candidates are pre-filtered in
// FP32 with a rigorous margin and accepted by the same three FP64 operations as the oracle.
//
// Execution (what the measurements forced):
//   * a pass never scans inside its enumeration loops: it LISTS its segments (shared memory, one column per thread) and
//     the candidates are consumed afterwards by ONE flat loop, so the lanes of a warp -- which walk different cells --
//     stay converged where the time goes (nested: 7 of 32 lanes active on average, 2.4x slower than the plain search);
//   * the flat loop does not insert either: it keeps the K smallest FP32 distances in a branch-free min/max network and
//     QUEUES the candidates that may still belong to the K best (shared memory, ~20 of 150); the queue is drained with
//     all lanes busy (exact FP64 distance + sorted insertion).  Inserting inside the scan ran the 70-instruction
//     insertion for ~6 lanes at a time on almost every candidate;
//   * the first pass (w = one bin edge: enough whenever the K-th neighbour is closer than that) is done by every thread
//     for its own query; the queries it does not finish (outliers hanging in free space, sparse spots: 10 % in config 3,
//     each worth 300 ... 4 000 more candidates) are continued ONE AT A TIME BY THE WHOLE WARP: the rows of the box are
//     dealt to the lanes (lane l lists rows l, l + 32, ...), every lane scans its rows into a private list, and the 32
//     lists are merged by K rounds of a warp-wide argmin over the list heads.  Left to their own threads these queries
//     made every warp wait for its slowest lane.
// Replaces, like k_correspond, KDTreeFlann::SearchHybrid of ref: src/models/registration/registration.cpp:272, 444,
// 535, 588, 731.
#pragma once
#include "frame_kernels.cuh"

namespace tloam {

// per-query constants of the geometric tests and of the FP32 pre-filter (a pure function of the grid and the query,
// so the lanes of a warp that continue somebody else's query recompute identical values)
struct FineConst {
  int cx, cy, cz;
  float cellf, fe, inv_fe, geo_slack, wmax;
  float ql[3];       // query in the local frame: origin = low corner of cell (c - 1); the query sits in [cell, 2 cell)^3
  float qa[3];       // query rounded to FP32 (map-origin relative, like the stored points)
  double pf_abs;     // |FP32 squared distance - exact squared distance| <= pf_abs (+ 1e-6 relative) inside the neighbourhood
};

__device__ __forceinline__ FineConst fine_const(const GridDesc& g, double rx, double ry, double rz) {
  FineConst q;
  q.cx = (int)floor(rx * g.inv_cell); q.cy = (int)floor(ry * g.inv_cell); q.cz = (int)floor(rz * g.inv_cell);
  q.cellf = (float)g.cell;
  q.fe = q.cellf * 0.25f; q.inv_fe = 4.0f / q.cellf;
  q.wmax = q.cellf * 1.000001f;
  q.ql[0] = (float)(rx - (double)(q.cx - 1) * g.cell); q.ql[1] = (float)(ry - (double)(q.cy - 1) * g.cell);
  q.ql[2] = (float)(rz - (double)(q.cz - 1) * g.cell);
  // every rounding of the box / bin arithmetic (local query 2^-24 * 2 cell, bin edges k * fe, the squares) stays below
  // 1e-6 cell per axis, i.e. below 2e-5 cell^2 on a squared distance of at most 27 cell^2
  q.geo_slack = 2e-5f * q.cellf * q.cellf;
  // FP32 distances on the stored (map-origin relative) coordinates: per axis |fl(p - fl(q)) - (p - q)| <= e with
  // e = 2^-24 (|q| + |p - q|) and |p - q| <= 2.5 cell inside the neighbourhood, so for a squared distance d2 <= r^2
  // |df - d2| <= 2 sqrt(3) r e + 3 e^2 (+ 3 * 2^-24 relative); r = cell
  q.qa[0] = (float)rx; q.qa[1] = (float)ry; q.qa[2] = (float)rz;
  const double e_ax = (fmax(fabs(rx), fmax(fabs(ry), fabs(rz))) + 2.5 * g.cell) * 6.1e-8;
  q.pf_abs = 3.5 * g.cell * e_ax + 3.0 * e_ax * e_ax;
  return q;
}
// FP32 image of an exact bound: d2 <= bound implies df <= fine_boundf (strictly above `bound`: ties on the K-th best survive)
__device__ __forceinline__ float fine_boundf(double pf_abs, double bound) {
  return (float)((bound + pf_abs) * 1.000001) * 1.0000005f;
}

// per-thread work lists (shared memory, one column per thread)
constexpr int kFineCap = 28;          // segments: first pass <= 3 x 3 rows x 2 cells; continuation <= 14 rows x 2 per lane
constexpr int kFineQueue = 32;        // queued candidates
struct FineSmem {
  unsigned jb[kFineCap][kBlk];        // first point of the segment
  unsigned cm[kFineCap][kBlk];        // count << 16 | upper half of the FP32 lower bound of its squared distance (rounded down)
  unsigned qu[kFineQueue][kBlk];      // positions of the candidates that may belong to the K best
  unsigned done[kBlk];                // column c: the thread that adopted query c has finished scanning its first-pass list
};

__device__ __forceinline__ void fine_index_box(const FineConst& q, float w, int lo[3], int hi[3]) {
  const float we = w * 1.0001f + 1e-5f * q.cellf;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int a = (int)floorf((q.ql[d] - we) * q.inv_fe), b = (int)floorf((q.ql[d] + we) * q.inv_fe);
    lo[d] = a < 0 ? 0 : a; hi[d] = b > 11 ? 11 : b;
  }
}

// lower bound (rounded down, clamped at 0) of the squared distance from the query to the box of local bins
// [x0, x1] x [y0, y1] x [z0, z1]
__device__ __forceinline__ float fine_box_lb(const FineConst& q, int x0, int x1, int y0, int y1, int z0, int z1) {
  const float blo[3] = {(float)x0 * q.fe, (float)y0 * q.fe, (float)z0 * q.fe};
  const float bhi[3] = {(float)(x1 + 1) * q.fe, (float)(y1 + 1) * q.fe, (float)(z1 + 1) * q.fe};
  float md = 0.0f;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float v = fmaxf(0.0f, fmaxf(blo[d] - q.ql[d], q.ql[d] - bhi[d]));
    md = fmaf(v, v, md);
  }
  return fmaxf(md * 0.9999f - q.geo_slack, 0.0f);
}

__device__ __forceinline__ void fine_push(FineSmem* sm, int& m, unsigned jb, unsigned je, float mds) {
  if (je <= jb) return;
  if (m >= 0 && m < kFineCap) {
    sm->jb[m][threadIdx.x] = jb;
    sm->cm[m][threadIdx.x] = ((je - jb) << 16) | (__float_as_uint(mds) >> 16);
    ++m;
  } else {
    m = -1;                           // overflow: reported to the caller
  }
}

// the segments of one row (y, z) of a dense cell whose local bins start at (ox, oy, oz): bins [x0, x1] minus what the
// previous index box [plo, phi] already reached
__device__ __forceinline__ void fine_push_row(const FineConst& q, const unsigned short* ft, unsigned cbeg, int ox, int oy, int oz,
                                              int x0, int x1, int y, int z, const int plo[3], const int phi[3], float boundf,
                                              FineSmem* sm, int& m) {
  const int row = (z * kFineDiv + y) * kFineDiv;
  const bool cut = (oz + z >= plo[2] && oz + z <= phi[2] && oy + y >= plo[1] && oy + y <= phi[1]);
#pragma unroll
  for (int seg = 0; seg < 2; ++seg) {
    int xa = x0, xb = x1;
    if (cut) {
      if (seg == 0) xb = (plo[0] - 1 - ox < x1) ? plo[0] - 1 - ox : x1;
      else xa = (phi[0] + 1 - ox > x0) ? phi[0] + 1 - ox : x0;
    } else if (seg == 1) {
      continue;
    }
    if (xa > xb) continue;
    const unsigned jb = cbeg + ((row + xa) ? (unsigned)ft[row + xa - 1] : 0u), je = cbeg + (unsigned)ft[row + xb];
    if (je <= jb) continue;                                      // empty bins: most rows of a cell crossed by one surface
    const float mds = fine_box_lb(q, ox + xa, ox + xb, oy + y, oy + y, oz + z, oz + z);
    if (mds >= boundf) continue;
    fine_push(sm, m, jb, je, mds);
  }
}

// First pass of the calling thread's own query: lists every segment that intersects the index box [lo, hi] (previous
// box: none; at w = one bin edge the box is at most 3 bins = 2 cells wide per axis).  Two loops of FIXED trip count with
// predicated bodies, so that the 32 lanes -- 32 different queries -- stay converged (the nested brick / cell / row loops
// this replaces ran with 4 to 8 active lanes): (A) the <= 2 x 2 x 2 cells the box touches: brick probe, first point,
// count, second-level table; sparse cells are listed whole; (B) the <= 4 x 4 rows of the box x <= 2 cells along x.
__device__ __forceinline__ void fine_list_own(const GridDesc& g, const FineConst& q, const int lo[3], const int hi[3], float boundf,
                                              FineSmem* sm, int& m) {
  const int tid = threadIdx.x;
  const int plo[3] = {12, 12, 12}, phi[3] = {-1, -1, -1};
  const int c0x = lo[0] >> 2, c0y = lo[1] >> 2, c0z = lo[2] >> 2;           // first cell of the box per axis (0..2)
  // (A) cell records in the (still unused) candidate queue of this thread: [3 i] first point, [3 i + 1] count (0 = nothing
  //     to do in loop B), [3 i + 2] second-level table
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int ox3 = c0x + (i & 1), oy3 = c0y + ((i >> 1) & 1), oz3 = c0z + (i >> 2);
    unsigned cbeg = 0u, cnt = 0u, fidx = 0u;
    if (ox3 <= (hi[0] >> 2) && oy3 <= (hi[1] >> 2) && oz3 <= (hi[2] >> 2)) {
      const int gx = q.cx - 1 + ox3, gy = q.cy - 1 + oy3, gz = q.cz - 1 + oz3;
      BrickEntry e;
      if (probe_brick(g, cell_key(brick_of(gx), brick_of(gy), brick_of(gz)), e)) {
        const int sub = subcell_of(gx, gy, gz);
        unsigned beg = e.a.z, nd = 0u;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const unsigned cs = e.count(s);
          if (s < sub) { beg += cs; nd += (cs >= kFineMin) ? 1u : 0u; }
          if (s == sub) cnt = cs;
        }
        cbeg = beg;
        if (e.b.w != 0u && cnt >= kFineMin) {
          fidx = e.b.w - 1u + nd;
        } else if (cnt != 0u) {                                     // sparse cell: one segment, listed here
          const float mds = fine_box_lb(q, 4 * ox3, 4 * ox3 + 3, 4 * oy3, 4 * oy3 + 3, 4 * oz3, 4 * oz3 + 3);
          if (mds < boundf) fine_push(sm, m, cbeg, cbeg + cnt, mds);
          cnt = 0u;
        }
      }
    }
    sm->qu[3 * i][tid] = cbeg; sm->qu[3 * i + 1][tid] = cnt; sm->qu[3 * i + 2][tid] = fidx;
  }
  // (B) rows
#pragma unroll 1
  for (int it = 0; it < 32; ++it) {                                          // 3 bins per axis, 4 when the query sits on a bin face
    const int xc = it & 1, yy = (it >> 1) & 3, zz = it >> 3;
    const int gy = lo[1] + yy, gz = lo[2] + zz;                              // local bin coordinates of the row
    const int ox3 = c0x + xc;
    if (gy > hi[1] || gz > hi[2] || ox3 > (hi[0] >> 2)) continue;
    const int ci = ((gz >> 2) - c0z) * 4 + ((gy >> 2) - c0y) * 2 + xc;
    const unsigned cnt = sm->qu[3 * ci + 1][tid];
    if (cnt == 0u) continue;
    const unsigned cbeg = sm->qu[3 * ci][tid], fidx = sm->qu[3 * ci + 2][tid];
    const int ox = 4 * ox3, oy = gy & ~3, oz = gz & ~3;
    const int x0 = (lo[0] > ox ? lo[0] : ox) - ox, x1 = (hi[0] < ox + 3 ? hi[0] : ox + 3) - ox;
    if (x0 > x1) continue;
    fine_push_row(q, g.fine + (size_t)fidx * kFineBins, cbeg, ox, oy, oz, x0, x1, gy & 3, gz & 3, plo, phi, boundf, sm, m);
  }
}

// Continuation pass, the 32 lanes of a warp on ONE query (all arguments warp-uniform).  Lane l < 27 fetches cell l of the
// 3 x 3 x 3 neighbourhood; the rows of the index box [lo, hi] (bins in y x bins in z x cells in x) are dealt to the lanes,
// each lane works out the segments of its rows (minus the previous box [plo, phi], minus what the K-th best prunes) and
// the segments are cut into chunks of `cs` points that go round-robin into the warp's 32 list columns (a warp-wide
// prefix sum per step gives every chunk its slot): every lane ends up with the same number of chunks, whatever the
// shape of the surfaces inside the box.  m = chunks in my column, -1 on overflow of the 28 x 32 slots.
__device__ __forceinline__ void fine_list_coop(const GridDesc& g, const FineConst& q, const int lo[3], const int hi[3],
                                               const int plo[3], const int phi[3], float boundf, FineSmem* sm, int& m) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int wbase = (int)threadIdx.x & ~31;
  // ---- my cell ----
  unsigned c_beg = 0u, c_cnt = 0u;
  int c_fidx = -1;
  {
    const int l27 = lane < 27 ? lane : 0;
    const int o3x = l27 % 3, o3y = (l27 / 3) % 3, o3z = l27 / 9;
    const bool inbox = lane < 27 && !(4 * o3x + 3 < lo[0] || 4 * o3x > hi[0] || 4 * o3y + 3 < lo[1] || 4 * o3y > hi[1] ||
                                      4 * o3z + 3 < lo[2] || 4 * o3z > hi[2]);
    if (inbox) {
      const int gx = q.cx - 1 + o3x, gy = q.cy - 1 + o3y, gz = q.cz - 1 + o3z;
      BrickEntry e;
      if (probe_brick(g, cell_key(brick_of(gx), brick_of(gy), brick_of(gz)), e)) {
        const int sub = subcell_of(gx, gy, gz);
        unsigned beg = e.a.z, nd = 0u;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const unsigned cs_ = e.count(s);
          if (s < sub) { beg += cs_; nd += (cs_ >= kFineMin) ? 1u : 0u; }
          if (s == sub) c_cnt = cs_;
        }
        c_beg = beg;
        if (e.b.w != 0u && c_cnt >= kFineMin) c_fidx = (int)(e.b.w - 1u + nd);
      }
    }
  }
  // chunk size: 8 points, doubled until the points of the box's cells (an upper bound of the candidates) fit ~600 slots
  unsigned upper = c_cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) upper += __shfl_xor_sync(full, upper, o);
  unsigned cs = 8u;
  while (cs * 600u < upper) cs <<= 1;
  int total = 0;                                                  // chunks listed so far (warp-uniform)
  bool over = false;
  // one step: every lane offers up to two segments; chunks get consecutive slots
  auto publish = [&](unsigned jb0, unsigned n0, float md0, unsigned jb1, unsigned n1, float md1) {
    const unsigned k0 = (n0 + cs - 1u) / cs, k1 = (n1 + cs - 1u) / cs;
    unsigned incl = k0 + k1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(full, incl, o); if (lane >= o) incl += v; }
    const unsigned step_total = __shfl_sync(full, incl, 31);
    unsigned p = (unsigned)total + incl - (k0 + k1);
    if ((unsigned)total + step_total > (unsigned)(kFineCap * 32)) { over = true; }
    else {
      for (unsigned c = 0; c < k0; ++c, ++p) {
        const unsigned len = (n0 - c * cs < cs) ? n0 - c * cs : cs;
        sm->jb[p >> 5][wbase + (p & 31u)] = jb0 + c * cs;
        sm->cm[p >> 5][wbase + (p & 31u)] = (len << 16) | (__float_as_uint(md0) >> 16);
      }
      for (unsigned c = 0; c < k1; ++c, ++p) {
        const unsigned len = (n1 - c * cs < cs) ? n1 - c * cs : cs;
        sm->jb[p >> 5][wbase + (p & 31u)] = jb1 + c * cs;
        sm->cm[p >> 5][wbase + (p & 31u)] = (len << 16) | (__float_as_uint(md1) >> 16);
      }
      total += (int)step_total;
    }
  };
  // ---- sparse cells: one segment each (a cell belongs to the first pass whose box reaches it) ----
  {
    unsigned n0 = 0u;
    float md0 = 0.0f;
    if (lane < 27 && c_cnt != 0u && c_fidx < 0) {
      const int ox = 4 * (lane % 3), oy = 4 * ((lane / 3) % 3), oz = 4 * (lane / 9);
      if (ox + 3 < plo[0] || ox > phi[0] || oy + 3 < plo[1] || oy > phi[1] || oz + 3 < plo[2] || oz > phi[2]) {
        md0 = fine_box_lb(q, ox, ox + 3, oy, oy + 3, oz, oz + 3);
        if (md0 < boundf) n0 = c_cnt;
      }
    }
    if (__any_sync(full, n0 != 0u)) publish(c_beg, n0, md0, 0u, 0u, 0.0f);
  }
  // ---- rows of the dense cells ----
  const int cx0 = lo[0] >> 2, ncx = (hi[0] >> 2) - cx0 + 1;
  const int ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1;
  const int nrows = ncx * ny * nz;
#pragma unroll 1
  for (int base = 0; base < nrows; base += 32) {
    const int r = base + lane;
    const bool have = r < nrows;
    const int xc = have ? r % ncx : 0, yy = have ? (r / ncx) % ny : 0, zz = have ? r / (ncx * ny) : 0;
    const int gy = lo[1] + yy, gz = lo[2] + zz;
    const int o3x = cx0 + xc, cell = (gz >> 2) * 9 + (gy >> 2) * 3 + o3x;
    const unsigned cbeg = __shfl_sync(full, c_beg, cell), cnt = __shfl_sync(full, c_cnt, cell);
    const int fidx = __shfl_sync(full, c_fidx, cell);
    unsigned jb[2] = {0u, 0u}, n[2] = {0u, 0u};
    float md[2] = {0.0f, 0.0f};
    if (have && cnt != 0u && fidx >= 0) {
      const int ox = 4 * o3x, oy = gy & ~3, oz = gz & ~3, y = gy & 3, z = gz & 3;
      const int x0 = (lo[0] > ox ? lo[0] : ox) - ox, x1 = (hi[0] < ox + 3 ? hi[0] : ox + 3) - ox;
      const unsigned short* ft = g.fine + (size_t)fidx * kFineBins;
      const int row = (z * kFineDiv + y) * kFineDiv;
      const bool cut = (gz >= plo[2] && gz <= phi[2] && gy >= plo[1] && gy <= phi[1]);
#pragma unroll
      for (int seg = 0; seg < 2; ++seg) {
        int xa = x0, xb = x1;
        if (cut) {
          if (seg == 0) xb = (plo[0] - 1 - ox < x1) ? plo[0] - 1 - ox : x1;
          else xa = (phi[0] + 1 - ox > x0) ? phi[0] + 1 - ox : x0;
        } else if (seg == 1) {
          continue;
        }
        if (xa > xb) continue;
        const unsigned b = cbeg + ((row + xa) ? (unsigned)ft[row + xa - 1] : 0u), e = cbeg + (unsigned)ft[row + xb];
        if (e <= b) continue;
        const float mds = fine_box_lb(q, ox + xa, ox + xb, gy, gy, gz, gz);
        if (mds >= boundf) continue;
        jb[seg] = b; n[seg] = e - b; md[seg] = mds;
      }
    }
    if (__any_sync(full, (n[0] | n[1]) != 0u)) publish(jb[0], n[0], md[0], jb[1], n[1], md[1]);
  }
  m = over ? -1 : (total + 31 - lane) / 32;
}

// Consumes the calling thread's segment list: exact top-K of its candidates merged into `t` (which may already hold
// candidates of earlier passes; `bound` = min(r2, K-th best so far) on entry and on return).
template <int K>
__device__ __forceinline__ void fine_scan(const GridDesc& g, FineSmem* sm, int col, int m, double rx, double ry, double rz,
                                          double r2, const float qa[3], double pf_abs, double& bound, TopK<K>& t) {
  const int tid = threadIdx.x;               // my queue column; the segment list is column `col` (mine, or an adopted query's)
  // nearest segment first: it fills the threshold network with good candidates
  if (m > 1) {
    int best = 0;
    unsigned bm = sm->cm[0][col] & 0xFFFFu;
    for (int i = 1; i < m; ++i) { const unsigned v = sm->cm[i][col] & 0xFFFFu; if (v < bm) { bm = v; best = i; } }
    if (best != 0) {
      const unsigned a = sm->jb[0][col], b = sm->cm[0][col];
      sm->jb[0][col] = sm->jb[best][col]; sm->cm[0][col] = sm->cm[best][col];
      sm->jb[best][col] = a; sm->cm[best][col] = b;
    }
  }
  // FP32 threshold: the K smallest FP32 distances seen so far, ascending, in a branch-free network.  A candidate of the
  // exact K best has d2 <= D_K, and D_K <= F_K + E (F_K = K-th smallest FP32 distance, E = FP32 error bound), so its own
  // FP32 distance is <= F_K + 2E: everything above `thr` can be dropped without looking at it again.
  float a[K];
#pragma unroll
  for (int j = 0; j < K; ++j) a[j] = 3.0e38f;
  const float e2 = (float)(2.0 * pf_abs + 4.0e-6 * r2) * 1.000001f + 1e-37f;
  const float cap = fine_boundf(pf_abs, bound);                   // nothing beyond the incoming bound
  float thr = cap;
  int ci = 0, qn = 0;
  unsigned off = 0u, cb = 0u, cc = 0u;
  auto load_entry = [&]() {
    while (ci < m) {
      const unsigned cm = sm->cm[ci][col];
      if (!(__uint_as_float(cm << 16) >= thr)) { cb = sm->jb[ci][col]; cc = cm >> 16; off = 0u; return; }
      ++ci;                                                       // nothing in this segment can belong to the K best
    }
  };
  load_entry();
  do {
    // ---- scan: 4 loads in flight; queue what may still belong to the K best ----
    while (ci < m && qn <= kFineQueue - 4) {
      float4 p[4];
      int pos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pos[u] = -1;
        p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ci < m) {
          pos[u] = (int)(cb + off);
          p[u] = __ldg(&g.pts[pos[u]]);
          if (++off >= cc) { ++ci; load_entry(); }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float fx = p[u].x - qa[0], fy = p[u].y - qa[1], fz = p[u].z - qa[2];
        float df = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
        if (pos[u] < 0) df = 3.0e38f;
        if (df <= thr) { sm->qu[qn][tid] = (unsigned)pos[u]; ++qn; }
        float x = df;
#pragma unroll
        for (int j = 0; j < K; ++j) { const float lo_ = fminf(a[j], x); x = fmaxf(a[j], x); a[j] = lo_; }
        thr = fminf(cap, a[K - 1] * 1.000002f + e2);
      }
    }
    // ---- drain: exact FP64 distance + sorted insertion, every lane busy ----
    for (int i = 0; i < qn; ++i) {
      const int pos = (int)sm->qu[i][tid];
      const float4 p = __ldg(&g.pts[pos]);
      const float fx = p.x - qa[0], fy = p.y - qa[1], fz = p.z - qa[2];
      const float df = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
      if (df <= thr) {                                            // the threshold has tightened since it was queued
        const double ddx = (double)p.x - rx, ddy = (double)p.y - ry, ddz = (double)p.z - rz;
        const double dd = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));   // same three operations as the oracle
        if (dd < r2) t.insert(dd, __float_as_int(p.w), pos);
      }
    }
    qn = 0;
  } while (ci < m);
  if (t.d2[K - 1] < bound) bound = t.d2[K - 1];
}

// Phase 1a, every thread for its OWN query: lists the segments of the first pass (w = one bin edge) in its column and
// publishes the query for adoption (rows 24..31 of its queue column: coordinates, list length, sort key = candidates
// << 7 | thread).  The thread block then sorts the keys (fine_sort_queries) and thread t ADOPTS the query of rank t:
// the 32 queries a warp scans together have similar candidate counts (150 ... 450 on config 3: unsorted, a warp waited
// for its largest query while the average lane was idle half of the time).
__device__ __forceinline__ void fine_list_and_publish(const GridDesc& g, bool live, double rx, double ry, double rz, double r2, FineSmem* sm) {
  const int tid = threadIdx.x;
  int m = 0;
  unsigned ncand = 0u;
  if (live && g.n != 0u) {
    const FineConst q = fine_const(g, rx, ry, rz);
    int lo[3], hi[3];
    fine_index_box(q, q.fe, lo, hi);
    fine_list_own(g, q, lo, hi, fine_boundf(q.pf_abs, r2), sm, m);
    for (int i = 0; i < m; ++i) ncand += sm->cm[i][tid] >> 16;
    if (m < 0) ncand = 0xFFFFFu;                                 // list overflow: the warp does this query from scratch
  } else {
    m = -2;                                                       // no query
  }
  sm->qu[24][tid] = (unsigned)__double2loint(rx); sm->qu[25][tid] = (unsigned)__double2hiint(rx);
  sm->qu[26][tid] = (unsigned)__double2loint(ry); sm->qu[27][tid] = (unsigned)__double2hiint(ry);
  sm->qu[28][tid] = (unsigned)__double2loint(rz); sm->qu[29][tid] = (unsigned)__double2hiint(rz);
  sm->done[tid] = 0u;
  sm->qu[30][tid] = (unsigned)m;
  sm->qu[31][tid] = ((ncand > 0xFFFFFu ? 0xFFFFFu : ncand) << 7) | (unsigned)tid;
}

// bitonic sort of the kBlk keys in row 31 of the queue array, descending (all threads of the block)
__device__ __forceinline__ void fine_sort_queries(FineSmem* sm) {
  const unsigned tid = threadIdx.x;
  __syncthreads();
#pragma unroll 1
  for (unsigned k = 2; k <= (unsigned)kBlk; k <<= 1)
#pragma unroll 1
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      const unsigned partner = tid ^ j;
      if (partner > tid) {
        const unsigned a = sm->qu[31][tid], b = sm->qu[31][partner];
        const bool desc = (tid & k) == 0;                         // this run is sorted descending
        if (desc ? (a < b) : (a > b)) { sm->qu[31][tid] = b; sm->qu[31][partner] = a; }
      }
      __syncthreads();
    }
}

// Phase 1b + 2 for the ADOPTED query (thread `col` listed it): flat scan of its first-pass segments, then the
// warp-cooperative continuation.  ALL 32 lanes of a warp must call it together (lanes without a query: live = false).
template <int K>
__device__ __forceinline__ void knn_search_fine(const GridDesc& g, bool live, int col, int m, double rx, double ry, double rz, double r2,
                                                TopK<K>& t, FineSmem* sm) {
  const unsigned lane = threadIdx.x & 31u;
  t.init();
  bool need = false;
  double bound = r2;
  float wdone = 0.0f;                                            // half-width of the box searched so far
  int lo[3] = {12, 12, 12}, hi[3] = {-1, -1, -1};
  // ---- phase 1: first pass of my (adopted) query ----
  {
    float qa[3] = {0.f, 0.f, 0.f};
    double pf_abs = 0.0;
    if (live && g.n != 0u) {
      const FineConst q = fine_const(g, rx, ry, rz);
      fine_index_box(q, q.fe, lo, hi);
      need = true;
      wdone = q.fe;
      qa[0] = q.qa[0]; qa[1] = q.qa[1]; qa[2] = q.qa[2]; pf_abs = q.pf_abs;
      if (m < 0) {                                               // list overflow (cannot happen at w = fe): the warp does it all
#pragma unroll
        for (int d = 0; d < 3; ++d) { lo[d] = 12; hi[d] = -1; }
        wdone = 0.5f * q.fe;
        m = 0;
      }
    } else {
      m = 0;
    }
    fine_scan<K>(g, sm, col, m, rx, ry, rz, r2, qa, pf_abs, bound, t);
    // column `col` (another thread's, possibly another warp's) may be overwritten from here on
    __threadfence_block();
    reinterpret_cast<volatile unsigned*>(sm->done)[col] = 1u;
  }
  // complete iff everything within sqrt(bound) of the query lies inside the box just searched
  need = need && !((double)wdone * (double)wdone >= bound);
  // ---- phase 2: unfinished queries, one at a time, by the whole warp ----
  unsigned todo = __ballot_sync(0xffffffffu, need);
  if (todo) {
    // the continuation passes deal their chunks into THIS warp's 32 columns: wait until their adopters are done with them
    // (a spin on shared-memory flags inside one resident block; adopters never wait for anything before they set theirs)
    const volatile unsigned* dv = reinterpret_cast<const volatile unsigned*>(sm->done);
    while (!__all_sync(0xffffffffu, dv[(threadIdx.x & ~31u) + lane] != 0u)) {}
    __threadfence_block();
  }
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1u;
    const double bx = __shfl_sync(0xffffffffu, rx, src), by = __shfl_sync(0xffffffffu, ry, src), bz = __shfl_sync(0xffffffffu, rz, src);
    const FineConst q = fine_const(g, bx, by, bz);
    double bnd = __shfl_sync(0xffffffffu, bound, src);
    int blo[3], bhi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { blo[d] = __shfl_sync(0xffffffffu, lo[d], src); bhi[d] = __shfl_sync(0xffffffffu, hi[d], src); }
    float w = __shfl_sync(0xffffffffu, wdone, src);
#pragma unroll 1
    for (int pass = 1; pass < 6; ++pass) {
      const float wn = (bnd < r2) ? (float)sqrt(bnd) * 1.000001f + 1e-7f * q.cellf : 2.0f * w;
      w = wn < q.wmax ? wn : q.wmax;
      int nlo[3], nhi[3], m = 0;
      fine_index_box(q, w, nlo, nhi);
      __syncwarp();                                              // the previous pass's lists have been consumed
      fine_list_coop(g, q, nlo, nhi, blo, bhi, fine_boundf(q.pf_abs, bnd), sm, m);
      __syncwarp();                                              // chunks were written into each other's columns
      const bool over = __any_sync(0xffffffffu, m < 0);          // more chunks than slots: pathological (tens of thousands of candidates)
      if (over) m = 0;
      TopK<K> tl;
      tl.init();
      if ((int)lane == src) tl = t;
      double pb = bnd;
      fine_scan<K>(g, sm, (int)threadIdx.x, m, bx, by, bz, r2, q.qa, q.pf_abs, pb, tl);
      // merge: K rounds of a warp-wide argmin over the heads of the 32 sorted private lists; lane src collects
      double kth = __longlong_as_double(0x7FF0000000000000ll);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        double bd = tl.d2[0];
        int bi = tl.idx[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double od = __shfl_xor_sync(0xffffffffu, bd, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        const bool win = bi != 0x7FFFFFFF && tl.idx[0] == bi;          // original indices are unique
        const unsigned wm = __ballot_sync(0xffffffffu, win);
        const int wp = __shfl_sync(0xffffffffu, tl.pos[0], wm ? __ffs(wm) - 1 : 0);
        if ((int)lane == src) { t.d2[j] = bd; t.idx[j] = bi; t.pos[j] = wm ? wp : -1; }
        if (j == K - 1) kth = bd;
        if (win) {                                                      // pop my head
#pragma unroll
          for (int k = 0; k + 1 < K; ++k) { tl.d2[k] = tl.d2[k + 1]; tl.idx[k] = tl.idx[k + 1]; tl.pos[k] = tl.pos[k + 1]; }
          tl.d2[K - 1] = __longlong_as_double(0x7FF0000000000000ll); tl.idx[K - 1] = 0x7FFFFFFF; tl.pos[K - 1] = -1;
        }
      }
      if (kth < bnd) bnd = kth;
      if (over) {
        // segment list overflow: redo this query from scratch with the plain search (keeps the result exact)
        TopK<K> ref;
        knn_search<K>(g, bx, by, bz, r2, ref);
        if ((int)lane == src) t = ref;
        break;
      }
      if ((double)w * (double)w >= bnd || w >= q.wmax) break;
#pragma unroll
      for (int d = 0; d < 3; ++d) { blo[d] = nlo[d]; bhi[d] = nhi[d]; }
    }
  }
}

// Correspondence search + fit + lazy GNC weight update for the clouds of ctx.dense_mask, one THREAD per feature
// (k_correspond returns early for these clouds and only clears the flags of their padding lanes).  Same outputs as
// k_correspond (w, slot, prim, flags, blk_count).  dbg (TLOAM_B200_DENSE_CHECK=1): [0] queries, [1] kNN lists that
// differ from the plain thread-per-query search.
template <bool kBatched>
__global__ void __launch_bounds__(kBlk, 4) k_correspond_fine(const __grid_constant__ DeviceCtx one, const __grid_constant__ BatchTab tab, unsigned* dbg) {
  TL_RESOLVE_CTX(one, tab);
  (void)nblk;
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  const int fb = lb;
  if (fb >= ctx.blk_off[4]) return;
  const int c = cloud_of_block(ctx, fb);
  if (!((ctx.dense_mask >> c) & 1)) return;
  extern __shared__ __align__(16) unsigned char fine_raw[];
  FineSmem* s_fine = reinterpret_cast<FineSmem*>(fine_raw);
  unsigned char flag = 0;
  double rx = 0.0, ry = 0.0, rz = 0.0;
  {
    // my OWN feature: transform, list the first pass, publish
    const int il0 = (fb - ctx.blk_off[c]) * kBlk + (int)threadIdx.x;
    const int gi0 = ctx.pad_off[c] + il0;
    const bool live0 = il0 < ctx.n[c] && cloud_enabled(ctx, c);
    if (live0) {
      const Rt T = pose_to_rt(st->xq);
      double qx, qy, qz;
      rt_apply(T, ctx.px[gi0], ctx.py[gi0], ctx.pz[gi0], qx, qy, qz);
      rx = qx - ctx.origin[0]; ry = qy - ctx.origin[1]; rz = qz - ctx.origin[2];
    }
    fine_list_and_publish(ctx.grid[c], live0, rx, ry, rz, ctx.r2[c], s_fine);
  }
  fine_sort_queries(s_fine);
  // the feature I ADOPT: rank threadIdx.x by candidate count
  const int col = (int)(s_fine->qu[31][threadIdx.x] & 127u);
  const int m_ad = (int)s_fine->qu[30][col];
  rx = __hiloint2double((int)s_fine->qu[25][col], (int)s_fine->qu[24][col]);
  ry = __hiloint2double((int)s_fine->qu[27][col], (int)s_fine->qu[26][col]);
  rz = __hiloint2double((int)s_fine->qu[29][col], (int)s_fine->qu[28][col]);
  __syncthreads();                                    // everybody has read its adoption record: the queue rows are free
  const int il = (fb - ctx.blk_off[c]) * kBlk + col;
  const int gi = ctx.pad_off[c] + il;
  const bool live = m_ad != -2;
  double prim[6];
  TopK<1> t1;
  TopK<5> t5;
  if (c == kSphere) knn_search_fine<1>(ctx.grid[c], live, col, m_ad, rx, ry, rz, ctx.r2[c], t1, s_fine);    // warp-cooperative: every lane calls
  else knn_search_fine<5>(ctx.grid[c], live, col, m_ad, rx, ry, rz, ctx.r2[c], t5, s_fine);
  if (live) {
    if (c == kSphere) {
      TopK<1>& t = t1;
      if (dbg) {
        TopK<1> ref;
        knn_search<1>(ctx.grid[c], rx, ry, rz, ctx.r2[c], ref);
        atomicAdd(&dbg[0], 1u);
        if (ref.idx[0] != t.idx[0] || (ref.d2[0] != t.d2[0] && ref.pos[0] >= 0)) atomicAdd(&dbg[1], 1u);
      }
      flag = fit_one<1>(ctx, c, t, prim);
    } else {
      TopK<5>& t = t5;
      if (dbg) {
        TopK<5> ref;
        knn_search<5>(ctx.grid[c], rx, ry, rz, ctx.r2[c], ref);
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 5; ++j) bad = bad || (ref.idx[j] != t.idx[j]) || (ref.d2[j] != t.d2[j] && ref.pos[j] >= 0);
        atomicAdd(&dbg[0], 1u);
        if (bad) { if (atomicAdd(&dbg[1], 1u) == 0u) { dbg[4] = (unsigned)gi; dbg[5] = (unsigned)c; dbg[6] = (unsigned)t.count(); dbg[7] = (unsigned)ref.count(); } }
      }
      flag = fit_one<5>(ctx, c, t, prim);
    }
    double wv = 1.0;                                                              // ref: registration.cpp:931-949
    if (st->outer != 0) {
      wv = ctx.w[gi];
      const double res = ctx.slot[gi];
      if (res != 0.0) {                                                           // Q13: res == 0 leaves the weight alone
        if (res >= st->th1) wv = 0.0;
        else if (res <= st->th2) wv = 1.0;
        else wv = sqrt(st->c2 * st->mu_used * (st->mu_used + 1.0) / res) - st->mu_used;
      }
    }
    ctx.w[gi] = wv;
    ctx.slot[gi] = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) ctx.prim[j][gi] = prim[j];
    ctx.flags[gi] = flag;
  }
  const int buf = st->outer & 1;
  const int cnt = __syncthreads_count(live && (flag & kFlagCounted) != 0);
  if (threadIdx.x == 0 && cnt > 0) atomicAdd(&ctx.blk_count[buf * ctx.blk_cap + fb], cnt);
}
constexpr size_t kFineSmemBytes = sizeof(FineSmem);

}  // namespace tloam
