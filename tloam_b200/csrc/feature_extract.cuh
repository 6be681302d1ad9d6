// feature_extract.cuh -- "next" row (f)-2: PCA feature extraction on the device.
//
// Replaces featureExtract::calculatePCAInfo / extractPlanarSphere
// (ref: src/models/feature_extraction/feature_extract.cpp:47-122, 133-197; config/mapping/feature.yaml;
//  caller FrontEnd::processCloud, src/front_end/front_end.cpp:181-199): for every point of the "general" cloud the
// K = 20 nearest neighbours within r = 0.2 m, the 3x3 covariance from raw cumulants, its eigen-decomposition, the
// curvature / flatness / sphericity measures, then the planar / sphere selection lists.
//
// Exactness: the output is INDEX LISTS, so this row is held to bit-exact parity with the CPU restatement
// (oracle/feature_oracle.cpp).  Points are therefore stored as FP64 (32 B / point: x, y, z, original index) in the
// brick-sorted order of map_grid.cuh, distances are the same three FP64 operations as the oracle's, neighbours are
// ordered by (d2, index), and the PCA arithmetic below is written with explicit round-to-nearest intrinsics in the
// oracle's operation order (no FMA contraction on either side).
#pragma once
#include "map_grid.cuh"

namespace tloam {

constexpr int kFeK = 20;                 // neighbour list capacity (feature.yaml: K = 20); smaller K = a prefix

struct alignas(32) FePoint { double x, y, z; long long idx; };

struct FeGrid {
  const FePoint* pts;    // [n] brick-sorted
  const uint4* table;    // brick table (map_grid.cuh layout)
  unsigned mask, n;
  double inv_cell, cell;
  const double* origin;  // MapHeader::origin (device)
};

struct FeParams {
  unsigned long long* dbg;   // profiling mode: SM cycles of thread 0 of every block, [0] search [1] cumulants [2] eigen [3] blocks
  double r2;
  int K, min_neigh;
  double cvr_submap, planar_submap_thres, planar_vertic_thres;
};

struct FeOut {           // indexed by ORIGINAL point index
  double *cvr, *flatness, *sphericity, *normal;   // normal: [n][3]
  int *num_sum, *neigh;                           // neigh: [n][kFeK], ascending (d2, index), -1 padded
};

__device__ __forceinline__ FePoint fe_load(const FePoint* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  FePoint r;
  r.x = a.x; r.y = a.y; r.z = b.x; r.idx = __double_as_longlong(b.y);
  return r;
}

__device__ __forceinline__ void fe_cell(const double* origin, double inv, double x, double y, double z, int& cx, int& cy, int& cz) {
  cx = (int)floor((x - origin[0]) * inv); cy = (int)floor((y - origin[1]) * inv); cz = (int)floor((z - origin[2]) * inv);
}

// ---- grid build (bbox / origin / offsets are the map kernels; cells come from the FP64 coordinates here) ----
struct FeBuildArgs {
  const double* stage;   // AoS xyz
  unsigned n;
  unsigned char* blob;   // MapHeader + FePoint[n] + brick table (cloud slot 0 of the header)
  unsigned* slot_of;
  unsigned* rank_of;
};

__global__ void k_fe_insert(FeBuildArgs a) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  int cx, cy, cz;
  fe_cell(h->origin, 1.0 / h->cell[0], a.stage[3ull * i], a.stage[3ull * i + 1], a.stage[3ull * i + 2], cx, cy, cz);
  const unsigned long long key = cell_key(brick_of(cx), brick_of(cy), brick_of(cz));
  const int sub = subcell_of(cx, cy, cz);
  uint4* table = reinterpret_cast<uint4*>(a.blob + h->table_off[0]);
  const unsigned mask = h->tsize[0] - 1u;
  unsigned s = hash_key(key) & mask;
  while (true) {
    unsigned long long* kp = reinterpret_cast<unsigned long long*>(&table[2u * s]);
    const unsigned long long prev = atomicCAS(kp, 0ull, key);
    if (prev == 0ull || prev == key) break;
    s = (s + 1u) & mask;
  }
  unsigned* words = reinterpret_cast<unsigned*>(&table[2u * s]) + 3;
  const unsigned old = atomicAdd(&words[sub >> 1], (sub & 1) ? 0x10000u : 1u);
  const unsigned rank = (sub & 1) ? (old >> 16) : (old & 0xFFFFu);
  if (rank >= kMaxCellPoints) atomicOr(&h->build_flags, 1ull);
  a.slot_of[i] = s;
  a.rank_of[i] = rank;
}

__global__ void k_fe_scatter(FeBuildArgs a) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  if (h->build_flags & 1ull) return;
  const double x = a.stage[3ull * i], y = a.stage[3ull * i + 1], z = a.stage[3ull * i + 2];
  int cx, cy, cz;
  fe_cell(h->origin, 1.0 / h->cell[0], x, y, z, cx, cy, cz);
  const int sub = subcell_of(cx, cy, cz);
  const uint4* table = reinterpret_cast<const uint4*>(a.blob + h->table_off[0]);
  FePoint* pts = reinterpret_cast<FePoint*>(a.blob + h->pts_off[0]);
  const unsigned slot = a.slot_of[i];
  const uint4 ea = table[2u * slot], eb = table[2u * slot + 1u];
  const unsigned w[4] = {ea.w, eb.x, eb.y, eb.z};
  unsigned dst = ea.z + a.rank_of[i];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < sub) dst += (w[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;
  FePoint p;
  p.x = x; p.y = y; p.z = z; p.idx = (long long)i;
  pts[dst] = p;
}

// ---- explicit round-to-nearest arithmetic (mirrors oracle/feature_oracle.cpp built with -ffp-contract=off) ----
__device__ __forceinline__ double fM(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fA(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fS(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fD(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double fQ(double a) { return __dsqrt_rn(a); }

// cyclic Jacobi, same sweep / rotation / update order as jacobi3() of the oracle
__device__ __forceinline__ void fe_jacobi3(const double c[6], double eig[3], double nvec[3]) {
  double a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = fA(fA(fM(a[0][1], a[0][1]), fM(a[0][2], a[0][2])), fM(a[1][2], a[1][2]));
    const double diag = fA(fA(fM(a[0][0], a[0][0]), fM(a[1][1], a[1][1])), fM(a[2][2], a[2][2]));
    if (off <= fM(1e-32, diag) || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = fD(fS(a[q][q], a[p][p]), fM(2.0, a[p][q]));
        const double t = fD(theta >= 0 ? 1.0 : -1.0, fA(fabs(theta), fQ(fA(fM(theta, theta), 1.0))));
        const double cs = fD(1.0, fQ(fA(fM(t, t), 1.0))), sn = fM(t, cs);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = fS(fM(cs, akp), fM(sn, akq));
          a[k][q] = fA(fM(sn, akp), fM(cs, akq));
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = fS(fM(cs, apk), fM(sn, aqk));
          a[q][k] = fA(fM(sn, apk), fM(cs, aqk));
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = fS(fM(cs, vkp), fM(sn, vkq));
          v[k][q] = fA(fM(sn, vkp), fM(cs, vkq));
        }
      }
  }
  // ascending order by three compare-exchanges (ties keep the lower axis first), without dynamic indexing
  double d0 = a[0][0], d1 = a[1][1], d2 = a[2][2];
  double n0[3] = {v[0][0], v[1][0], v[2][0]}, n1[3] = {v[0][1], v[1][1], v[2][1]}, n2[3] = {v[0][2], v[1][2], v[2][2]};
  auto cswap = [](double& da, double& db, double* na, double* nb) {
    if (db < da) {
      const double t = da; da = db; db = t;
      for (int r = 0; r < 3; ++r) { const double u = na[r]; na[r] = nb[r]; nb[r] = u; }
    }
  };
  cswap(d0, d1, n0, n1);
  cswap(d1, d2, n1, n2);
  cswap(d0, d1, n0, n1);
  eig[0] = d0; eig[1] = d1; eig[2] = d2;
  nvec[0] = n0[0]; nvec[1] = n0[1]; nvec[2] = n0[2];
}

// One thread per point, taken in brick-sorted order (neighbouring threads query neighbouring bricks).
// calculatePCAInfo, ref: feature_extract.cpp:47-122.
//
// Search: stage 1 lists the non-empty cells of the 27-cell neighbourhood (run start, length, lower-bound distance;
// the query's own cell first) in shared memory; stage 2 walks them as one flattened sequence in a WARP-UNIFORM loop,
// 4 loads in flight per thread.  A candidate that could enter the list is only queued (shared memory, 8 per
// thread); the queues are drained into the sorted 20-entry register lists when one of them runs full, all lanes
// together: the 20-wide sorted insertion (~250 instructions) then runs at full SIMT efficiency, ~3x fewer times
// than when it is executed whenever any lane has a candidate.
constexpr int kFeBlk = 128, kFeCells = 27, kFeQueue = 8;
constexpr size_t kFeSmemBytes = (size_t)kFeBlk * (kFeCells * 12 + kFeQueue * 16);

__global__ void __launch_bounds__(kFeBlk) k_fe_pca(FeGrid g, FeParams prm, FeOut out) {
  extern __shared__ __align__(16) unsigned char fe_smem[];
  double (*q_d)[kFeBlk] = reinterpret_cast<double (*)[kFeBlk]>(fe_smem);
  int (*q_i)[kFeBlk] = reinterpret_cast<int (*)[kFeBlk]>(fe_smem + (size_t)kFeQueue * kFeBlk * 8);
  int (*q_p)[kFeBlk] = reinterpret_cast<int (*)[kFeBlk]>(fe_smem + (size_t)kFeQueue * kFeBlk * 12);
  unsigned (*c_beg)[kFeBlk] = reinterpret_cast<unsigned (*)[kFeBlk]>(fe_smem + (size_t)kFeQueue * kFeBlk * 16);
  unsigned (*c_cnt)[kFeBlk] = c_beg + kFeCells;
  float (*c_md)[kFeBlk] = reinterpret_cast<float (*)[kFeBlk]>(c_cnt + kFeCells);
  const int tid = threadIdx.x;
  const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < g.n;
  FePoint me;
  me.x = me.y = me.z = 0.0; me.idx = 0;
  if (live) me = fe_load(&g.pts[p]);
  const long long i = me.idx;
  long long tc0 = 0, tc1 = 0, tc2 = 0;
  if (prm.dbg) tc0 = clock64();
  TopK<kFeK> t;
  t.init();
  int nc = 0;                                       // cells listed by this thread
  if (live) {
    int cx, cy, cz;
    fe_cell(g.origin, g.inv_cell, me.x, me.y, me.z, cx, cy, cz);
    const int bx0 = brick_of(cx - 1), by0 = brick_of(cy - 1), bz0 = brick_of(cz - 1);
    // bricks and sub-cells are listed starting from the query's own (XOR order): the list fills with near points first
    const int ibc = (brick_of(cx) - bx0) | ((brick_of(cy) - by0) << 1) | ((brick_of(cz) - bz0) << 2);
    const int scc = subcell_of(cx, cy, cz);
    const double qx = me.x - g.origin[0], qy = me.y - g.origin[1], qz = me.z - g.origin[2];
    const float r2f = (float)prm.r2 * 1.0001f;
#pragma unroll 1
    for (int iv = 0; iv < 8; ++iv) {
      const int ib = iv ^ ibc;
      const int bx = bx0 + (ib & 1), by = by0 + ((ib >> 1) & 1), bz = bz0 + (ib >> 2);
      const unsigned long long key = cell_key(bx, by, bz);
      unsigned s = hash_key(key) & g.mask;
      BrickEntry e;
      bool found = false;
      while (true) {
        e.a = __ldg(&g.table[2u * s]); e.b = __ldg(&g.table[2u * s + 1u]);
        const unsigned long long k = e.key();
        if (k == key) { found = true; break; }
        if (k == 0ull) break;
        s = (s + 1u) & g.mask;
      }
      if (!found) continue;
      unsigned begs[8];
      {
        unsigned run = e.a.z;
#pragma unroll
        for (int q = 0; q < 8; ++q) { begs[q] = run; run += e.count(q); }
      }
#pragma unroll 1
      for (int sv = 0; sv < 8; ++sv) {
        const int sc = sv ^ scc;
        const unsigned cnt = e.count(sc);
        const int gx = 2 * bx + (sc & 1), gy = 2 * by + ((sc >> 1) & 1), gz = 2 * bz + (sc >> 2);
        if (cnt == 0u || abs(gx - cx) > 1 || abs(gy - cy) > 1 || abs(gz - cz) > 1) continue;
        // lower bound of the distance to any point of that cell (slab distances, shrunk by a safety margin that
        // covers the rounding of the cell assignment and of the float it is stored in)
        const double lx = gx < cx ? qx - (double)(gx + 1) * g.cell : (gx > cx ? (double)gx * g.cell - qx : 0.0);
        const double ly = gy < cy ? qy - (double)(gy + 1) * g.cell : (gy > cy ? (double)gy * g.cell - qy : 0.0);
        const double lz = gz < cz ? qz - (double)(gz + 1) * g.cell : (gz > cz ? (double)gz * g.cell - qz : 0.0);
        const double ex = fmax(lx - 1e-9, 0.0), ey = fmax(ly - 1e-9, 0.0), ez = fmax(lz - 1e-9, 0.0);
        const float md = (float)(ex * ex + ey * ey + ez * ez) * 0.9999f;
        if (!(md < r2f)) continue;
        unsigned beg = begs[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) beg = (q == sc) ? begs[q] : beg;
        c_beg[nc][tid] = beg; c_cnt[nc][tid] = cnt; c_md[nc][tid] = md;
        ++nc;
      }
    }
  }
  // stage 2 (warp-uniform): all 32 lanes stay in the loop until the slowest has no candidates left
  int ci = 0, qn = 0;
  unsigned off = 0u, cb = 0u, cc = 0u;
  bool more = nc > 0;
  if (more) { cb = c_beg[0][tid]; cc = c_cnt[0][tid]; }
  while (__any_sync(0xffffffffu, more)) {
    FePoint pt[4];
    int pos[4];
    const double worst = t.d2[kFeK - 1];            // +inf until the list is full
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pos[q] = -1;
      if (more) {
        pos[q] = (int)(cb + off);
        pt[q] = fe_load(&g.pts[pos[q]]);
        if (++off == cc) {
          ++ci; off = 0u;
          while (ci < nc && (double)c_md[ci][tid] > worst) ++ci;       // no point of that cell can enter the list
          if (ci < nc) { cb = c_beg[ci][tid]; cc = c_cnt[ci][tid]; } else more = false;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (pos[q] >= 0) {
        const double ddx = pt[q].x - me.x, ddy = pt[q].y - me.y, ddz = pt[q].z - me.z;
        const double d = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));     // same three operations as the oracle
        if (d < prm.r2 && !(d > worst)) { q_d[qn][tid] = d; q_i[qn][tid] = (int)pt[q].idx; q_p[qn][tid] = pos[q]; ++qn; }
      }
    }
    if (__any_sync(0xffffffffu, qn > kFeQueue - 4)) {                   // the next batch might not fit: drain together
#pragma unroll 1
      for (int u = 0; u < kFeQueue; ++u) {
        if (!__any_sync(0xffffffffu, u < qn)) break;
        if (u < qn) t.insert(q_d[u][tid], q_i[u][tid], q_p[u][tid]);
      }
      qn = 0;
    }
  }
#pragma unroll 1
  for (int u = 0; u < kFeQueue; ++u) {
    if (!__any_sync(0xffffffffu, u < qn)) break;
    if (u < qn) t.insert(q_d[u][tid], q_i[u][tid], q_p[u][tid]);
  }
  if (!live) return;
  if (prm.dbg) tc1 = clock64();
  // SearchHybrid(cur_pt, r, K): the K nearest of the (up to kFeK) found
  int m = 0;
#pragma unroll
  for (int j = 0; j < kFeK; ++j) m += (j < prm.K && t.pos[j] >= 0) ? 1 : 0;
  double cvr = 0.0, flat = 0.0, sph = 0.0, nv[3] = {0.0, 0.0, 0.0};
  const bool keep = m > 0 && m > prm.min_neigh;                     // :67-71
  if (keep) {
    double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int j = 0; j < m; ++j) {                                    // ascending distance, :77-88
      int pos = t.pos[0];
#pragma unroll
      for (int q = 1; q < kFeK; ++q) pos = (q == j) ? t.pos[q] : pos;
      const FePoint nb = fe_load(&g.pts[pos]);
      cum[0] = fA(cum[0], nb.x); cum[1] = fA(cum[1], nb.y); cum[2] = fA(cum[2], nb.z);
      cum[3] = fA(cum[3], fM(nb.x, nb.x)); cum[4] = fA(cum[4], fM(nb.x, nb.y)); cum[5] = fA(cum[5], fM(nb.x, nb.z));
      cum[6] = fA(cum[6], fM(nb.y, nb.y)); cum[7] = fA(cum[7], fM(nb.y, nb.z)); cum[8] = fA(cum[8], fM(nb.z, nb.z));
    }
    const double dm = (double)m;
#pragma unroll
    for (int k = 0; k < 9; ++k) cum[k] = fD(cum[k], dm);             // :89
    const double cov[6] = {fS(cum[3], fM(cum[0], cum[0])), fS(cum[4], fM(cum[0], cum[1])), fS(cum[5], fM(cum[0], cum[2])),
                           fS(cum[6], fM(cum[1], cum[1])), fS(cum[7], fM(cum[1], cum[2])), fS(cum[8], fM(cum[2], cum[2]))};
    double ev[3];
    if (prm.dbg) tc2 = clock64();
    fe_jacobi3(cov, ev, nv);                                         // :100-104
    const double sum = fA(fA(ev[0], ev[1]), ev[2]);
    cvr = (sum == 0.0) ? 0.0 : fD(ev[0], sum);                       // :106-111
    flat = fD(fS(ev[1], ev[0]), ev[2]);                              // :113
    sph = fD(ev[0], ev[2]);                                          // :114
  }
  if (prm.dbg && threadIdx.x == 0 && keep) {
    atomicAdd(&prm.dbg[0], (unsigned long long)(tc1 - tc0));
    atomicAdd(&prm.dbg[1], (unsigned long long)(tc2 - tc1));
    atomicAdd(&prm.dbg[2], (unsigned long long)(clock64() - tc2));
    atomicAdd(&prm.dbg[3], 1ull);
  }
  out.cvr[i] = cvr; out.flatness[i] = flat; out.sphericity[i] = sph;
  out.normal[3 * i] = nv[0]; out.normal[3 * i + 1] = nv[1]; out.normal[3 * i + 2] = nv[2];
  out.num_sum[i] = keep ? m : 0;
#pragma unroll
  for (int j = 0; j < kFeK; ++j) out.neigh[i * kFeK + j] = (keep && j < m) ? t.idx[j] : -1;
}

// Classification of extractPlanarSphere, ref: feature_extract.cpp:148-164.  The candidates of each list are COMPACTED
// (key = order-preserving encoding of the flatness, value = point index; the order they land in does not matter, the
// sort that follows orders them by (flatness descending, index ascending) -- a total order).
__global__ void k_fe_classify(unsigned n, FeParams prm, FeOut out, unsigned long long* key_planar,
                              unsigned long long* key_sphere, unsigned* val_planar, unsigned* val_sphere, unsigned* counts) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  bool planar = false, sphere = false;
  double f = 0.0;
  if (i < n) {
    f = out.flatness[i];
    const double c = out.cvr[i];
    if (f > prm.planar_submap_thres && fabs(out.normal[3ull * i + 2]) < prm.planar_vertic_thres) {
      planar = true;
    } else if (c > prm.cvr_submap) {
      bool max_uniform = true;
      for (int j = 0; j < kFeK; ++j) {
        const int item = out.neigh[(size_t)i * kFeK + j];
        if (item < 0) break;
        if (c < out.cvr[item]) { max_uniform = false; break; }
      }
      sphere = max_uniform;
    }
  }
  // block-level compaction: one atomic per block and list
  __shared__ unsigned s_wp[8], s_ws[8], s_bp, s_bs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bp = __ballot_sync(0xffffffffu, planar), bs = __ballot_sync(0xffffffffu, sphere);
  if (lane == 0) { s_wp[warp] = __popc(bp); s_ws[warp] = __popc(bs); }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tp = 0, ts = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { const unsigned a = s_wp[w], b = s_ws[w]; s_wp[w] = tp; s_ws[w] = ts; tp += a; ts += b; }
    s_bp = tp ? atomicAdd(&counts[0], tp) : 0u;
    s_bs = ts ? atomicAdd(&counts[1], ts) : 0u;
  }
  __syncthreads();
  if (planar) { const unsigned d = s_bp + s_wp[warp] + __popc(bp & ((1u << lane) - 1u)); key_planar[d] = enc_ordered(f); val_planar[d] = i; }
  if (sphere) { const unsigned d = s_bs + s_ws[warp] + __popc(bs & ((1u << lane) - 1u)); key_sphere[d] = enc_ordered(f); val_sphere[d] = i; }
}

// Sort of the compacted candidates by (key descending, index ascending), first form: RANK sort.  Lists of up to
// kFeRankMax candidates (a 50k-point cloud has ~10k) are ordered by counting, for every candidate, how many come
// before it -- n^2 comparisons, but spread over the whole GPU (one thread per candidate, the list streamed through shared
// memory in tiles that every thread of the block reads at the same address; 16 blocks share a candidate block's list and
// add their partial ranks atomically), where the bitonic network below keeps ONE SM per list busy for ~200 us.  (key, index) pairs are distinct, so the ranks are a permutation.  Longer lists are copied
// through unchanged and sorted by k_fe_sort.
constexpr unsigned kFeRankMax = 32768u;
constexpr int kFeRankTile = 1024;

constexpr int kFeRankSplit = 16;       // slices of the list per candidate block: 16 x more blocks than candidates / 256

// grid (blocks over the candidate bound, 2 lists, kFeRankSplit slices): partial rank of my 256 candidates against one slice
__global__ void __launch_bounds__(256) k_fe_rank(const unsigned long long* key_planar_raw, const unsigned* val_planar_raw,
                                                 const unsigned long long* key_sphere_raw, const unsigned* val_sphere_raw,
                                                 unsigned* rank_planar, unsigned* rank_sphere, const unsigned* counts) {
  __shared__ unsigned long long s_key[kFeRankTile];
  __shared__ unsigned s_val[kFeRankTile];
  const int list = blockIdx.y;
  const unsigned long long* src_k = list == 0 ? key_planar_raw : key_sphere_raw;
  const unsigned* src_v = list == 0 ? val_planar_raw : val_sphere_raw;
  unsigned* rank_out = list == 0 ? rank_planar : rank_sphere;
  const unsigned total = counts[list];
  if (blockIdx.x * 256u >= total || total > kFeRankMax) return;
  const unsigned e = blockIdx.x * 256u + threadIdx.x;
  const bool live = e < total;
  const unsigned long long mk = live ? src_k[e] : 0ull;
  const unsigned mv = live ? src_v[e] : 0xFFFFFFFFu;
  const unsigned per = (total + kFeRankSplit - 1u) / kFeRankSplit;
  const unsigned lo = min(total, blockIdx.z * per), hi = min(total, lo + per);
  unsigned r0 = 0u, r1 = 0u;
  for (unsigned base = lo; base < hi; base += kFeRankTile) {
    const unsigned cnt = min((unsigned)kFeRankTile, hi - base);
    __syncthreads();
    for (unsigned t = threadIdx.x; t < cnt; t += 256u) { s_key[t] = src_k[base + t]; s_val[t] = src_v[base + t]; }
    for (unsigned t = cnt + threadIdx.x; t < ((cnt + 1u) & ~1u); t += 256u) { s_key[t] = 0ull; s_val[t] = 0xFFFFFFFFu; }   // pad to even: sorts last
    __syncthreads();
#pragma unroll 4
    for (unsigned t = 0u; t < cnt; t += 2u) {                              // two independent counters; the index is only
      const unsigned long long k0 = s_key[t], k1 = s_key[t + 1];           // looked at on a key tie (rare)
      r0 += k0 > mk ? 1u : 0u;
      r1 += k1 > mk ? 1u : 0u;
      if (k0 == mk) r0 += s_val[t] < mv ? 1u : 0u;
      if (k1 == mk) r1 += s_val[t + 1] < mv ? 1u : 0u;
    }
  }
  if (live && (r0 + r1)) atomicAdd(rank_out + e, r0 + r1);
}

// grid (blocks over the candidate bound, 2 lists): candidates to their ranks (or straight through when the list is too long)
__global__ void __launch_bounds__(256) k_fe_rank_scatter(const unsigned long long* key_planar_raw, const unsigned* val_planar_raw,
                                                         const unsigned long long* key_sphere_raw, const unsigned* val_sphere_raw,
                                                         const unsigned* rank_planar, const unsigned* rank_sphere,
                                                         unsigned long long* key_planar, unsigned* val_planar, unsigned long long* key_sphere,
                                                         unsigned* val_sphere, const unsigned* counts) {
  const int list = blockIdx.y;
  const unsigned total = counts[list];
  const unsigned e = blockIdx.x * 256u + threadIdx.x;
  if (e >= total) return;
  const unsigned long long k = (list == 0 ? key_planar_raw : key_sphere_raw)[e];
  const unsigned v = (list == 0 ? val_planar_raw : val_sphere_raw)[e];
  const unsigned d = total > kFeRankMax ? e : (list == 0 ? rank_planar : rank_sphere)[e];
  (list == 0 ? key_planar : key_sphere)[d] = k;
  (list == 0 ? val_planar : val_sphere)[d] = v;
}

// Second form, for lists longer than kFeRankMax: bitonic network over (key descending, index ascending), one block per
// list (blockIdx.x: 0 planar, 1 sphere), padded to a power of two with sentinels that sort last.  The network runs in
// SHARED memory: a tile of 16 384 (key, index) pairs (192 KB) is loaded once per merge level, all compare-exchange
// distances below the tile size are done there, and only the distances >= the tile size (lists longer than 16 384
// candidates) touch global memory.  ~10k candidates of a 50k-point cloud: one load, 105 steps in shared memory, one
// store.  (First version of this round: every step in global memory, 295 us; round 1: two cub::DeviceRadixSort passes
// over ALL n points with 64-bit keys, 160 us.)
constexpr unsigned kFeSortTile = 16384u;
constexpr size_t kFeSortSmemBytes = (size_t)kFeSortTile * (sizeof(unsigned long long) + sizeof(unsigned));

__global__ void __launch_bounds__(1024) k_fe_sort(unsigned long long* key_planar, unsigned* val_planar, unsigned long long* key_sphere,
                                                  unsigned* val_sphere, const unsigned* counts) {
  extern __shared__ __align__(16) unsigned char fe_sort_raw[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(fe_sort_raw);
  unsigned* sval = reinterpret_cast<unsigned*>(fe_sort_raw + (size_t)kFeSortTile * sizeof(unsigned long long));
  unsigned long long* key = blockIdx.x == 0 ? key_planar : key_sphere;
  unsigned* val = blockIdx.x == 0 ? val_planar : val_sphere;
  const unsigned total = counts[blockIdx.x];
  if (total <= kFeRankMax) return;                                         // ordered by k_fe_ranksort already
  unsigned m = 1u;
  while (m < total) m <<= 1;
  for (unsigned i = total + threadIdx.x; i < m; i += blockDim.x) { key[i] = 0ull; val[i] = 0xFFFFFFFFu; }   // sentinels: last
  __syncthreads();
  auto before = [](unsigned long long ka, unsigned va, unsigned long long kb, unsigned vb) { return ka > kb || (ka == kb && va < vb); };
  const unsigned tile = m < kFeSortTile ? m : kFeSortTile;
  // compare-exchange distances j0, j0 / 2, ..., 1 of merge level k on the tile that starts at `base`, in shared memory
  auto tile_steps = [&](unsigned base, unsigned k, unsigned j0) {
    for (unsigned j = j0; j > 0u; j >>= 1) {
      for (unsigned t = threadIdx.x; t < (tile >> 1); t += blockDim.x) {
        const unsigned i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));      // element whose bit j is clear
        const unsigned l = i | j;
        const bool up = ((base + i) & k) == 0u;                           // this sub-sequence ends up in final order
        const unsigned long long ki = skey[i], kl = skey[l];
        const unsigned vi = sval[i], vl = sval[l];
        const bool swap = up ? before(kl, vl, ki, vi) : before(ki, vi, kl, vl);
        if (swap) { skey[i] = kl; skey[l] = ki; sval[i] = vl; sval[l] = vi; }
      }
      __syncthreads();
    }
  };
  if (m <= kFeSortTile) {                                                  // the whole list fits: one load, one store
    for (unsigned i = threadIdx.x; i < m; i += blockDim.x) { skey[i] = key[i]; sval[i] = val[i]; }
    __syncthreads();
    for (unsigned k = 2u; k <= m; k <<= 1) tile_steps(0u, k, k >> 1);
    for (unsigned i = threadIdx.x; i < m; i += blockDim.x) { key[i] = skey[i]; val[i] = sval[i]; }
    return;
  }
  for (unsigned k = 2u; k <= m; k <<= 1) {
    unsigned j = k >> 1;
    for (; j >= kFeSortTile; j >>= 1) {                                    // distances that span tiles: global memory
      for (unsigned t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
        const unsigned i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
        const unsigned l = i | j;
        const bool up = (i & k) == 0u;
        const unsigned long long ki = key[i], kl = key[l];
        const unsigned vi = val[i], vl = val[l];
        const bool swap = up ? before(kl, vl, ki, vi) : before(ki, vi, kl, vl);
        if (swap) { key[i] = kl; key[l] = ki; val[i] = vl; val[l] = vi; }
      }
      __syncthreads();
    }
    for (unsigned base = 0u; base < m; base += kFeSortTile) {              // the rest of the level, tile by tile
      for (unsigned i = threadIdx.x; i < kFeSortTile; i += blockDim.x) { skey[i] = key[base + i]; sval[i] = val[base + i]; }
      __syncthreads();
      tile_steps(base, k, j);
      for (unsigned i = threadIdx.x; i < kFeSortTile; i += blockDim.x) { key[base + i] = skey[i]; val[base + i] = sval[i]; }
      __syncthreads();
    }
  }
}

// scan-list lengths (:177-188): the lists are sorted by descending flatness, so "rank < num || flatness > thres" is a
// prefix of length max(min(num, total), #{flatness > thres}).  counts: [0] planar, [1] sphere -> [2] planar_scan,
// [3] sphere_scan.
__global__ void k_fe_counts(const unsigned long long* key_planar, const unsigned long long* key_sphere, unsigned* counts,
                            int planar_num, int sphere_num, double planar_scan_thres, double cvr_scan) {
  if (threadIdx.x >= 2) return;
  const unsigned long long* keys = threadIdx.x == 0 ? key_planar : key_sphere;
  const unsigned total = counts[threadIdx.x];
  const unsigned long long thr = enc_ordered(threadIdx.x == 0 ? planar_scan_thres : cvr_scan);
  unsigned lo = 0, hi = total;              // first position whose key is NOT > thr
  while (lo < hi) {
    const unsigned mid = (lo + hi) / 2;
    if (keys[mid] > thr) lo = mid + 1; else hi = mid;
  }
  const unsigned num = (unsigned)max(threadIdx.x == 0 ? planar_num : sphere_num, 0);
  const unsigned base = num < total ? num : total;
  counts[2 + threadIdx.x] = lo > base ? lo : base;
}

}  // namespace tloam
