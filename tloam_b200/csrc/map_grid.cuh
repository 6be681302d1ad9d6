// map_grid.cuh -- GPU-resident voxel-hash local map + exact radius-truncated kNN.
//
// Replaces the four Open3D KDTreeFlann objects the reference rebuilds on every scanMatching call
// (ref: src/models/registration/registration.cpp:888-915) and their SearchHybrid(query, r, k) queries
// (ref: registration.cpp:272, 444, 535, 588, 731).
//
// One grid per cloud, cell edge == that cloud's search radius, so the 27 cells around a query contain
// every point with |m - q| < r: the search is EXACT for radius-truncated kNN (no approximation).
// Cells are grouped in 2x2x2 BRICKS and the hash table is keyed by brick: the 3x3x3 cell neighbourhood of
// any query always lies in exactly 2x2x2 bricks, so a query costs 8 table probes instead of 27, and the
// points of a brick are contiguous (sorted by brick, then by sub-cell), so the cells a query needs from one
// brick share cache lines.  Points are float4 (x,y,z relative to a per-map origin, w = original index
// bits): 16 B / point, one LDG.128 per candidate.  Table: open addressing, 32 B / entry (one sector)
// {key 8 B, base 4 B, 8 x u16 sub-cell counts, pad}.  All distance arithmetic is FP64 on the FP32 stored
// coordinates, ordering is (d2, original index) => results do not depend on the in-cell order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tloam {

struct GridDesc {
  const float4* pts;    // [n] cell-sorted
  const uint4* table;   // [mask+1] bricks, 2 x uint4 each (see brick_* below); key == 0 => empty
  unsigned mask;
  unsigned n;
  double inv_cell;
  double cell;
  const unsigned short* fine;   // [n / kFineMin + 1][64] second-level index of the dense cells (see knn_search_fine)
};

// 256-byte header at the start of the map blob (host + device visible layout)
struct MapHeader {
  unsigned long long magic;
  unsigned n[4];
  unsigned tsize[4];        // table entries = bricks slots (power of two)
  unsigned long long pts_off[4], table_off[4];  // byte offsets inside the blob
  double origin[3];
  double cell[4];
  unsigned long long bbox_enc[6];  // ordered-uint encodings of min xyz / max xyz (build scratch)
  unsigned cursor[4];              // bump allocators (build scratch)
  unsigned long long build_flags;  // bit 0: a cell holds more than kMaxCellPoints points (map unusable)
  unsigned bbox_ticket;            // blocks of k_map_bbox that have contributed (build scratch)
  unsigned pad32;
  unsigned nbricks[4];             // occupied bricks per cloud (k_map_offsets): points / bricks picks the search path
  unsigned long long fine_off[4];  // byte offset of each cloud's second-level tables (64 x u16 per dense cell)
  unsigned fine_cursor[4];         // dense cells per cloud (bump allocator of k_map_offsets, work list of k_map_fine)
  unsigned fine_build;             // bit c: cloud c was built with the second level (k_map_fine ran)
  unsigned pad2[3];
};
static_assert(sizeof(MapHeader) == 320, "MapHeader must be 320 bytes");
constexpr unsigned long long kMapMagic = 0x544C4F414D423230ull;  // "TLOAMB20"

__host__ __device__ __forceinline__ unsigned long long cell_key(int cx, int cy, int cz) {
  const unsigned long long ux = (unsigned long long)((unsigned)(cx + (1 << 20)) & 0x1FFFFFu);
  const unsigned long long uy = (unsigned long long)((unsigned)(cy + (1 << 20)) & 0x1FFFFFu);
  const unsigned long long uz = (unsigned long long)((unsigned)(cz + (1 << 20)) & 0x1FFFFFu);
  return (1ull << 63) | (ux << 42) | (uy << 21) | uz;
}

__host__ __device__ __forceinline__ unsigned hash_key(unsigned long long k) {
  k *= 0x9E3779B97F4A7C15ull;
  k ^= k >> 32;
  k *= 0xD6E8FEB86659FD93ull;
  k ^= k >> 29;
  return (unsigned)k;
}

__host__ __device__ __forceinline__ unsigned long long enc_ordered(double v) {
  unsigned long long u;
#ifdef __CUDA_ARCH__
  u = (unsigned long long)__double_as_longlong(v);
#else
  memcpy(&u, &v, 8);
#endif
  return (u >> 63) ? ~u : (u | (1ull << 63));
}
__host__ __device__ __forceinline__ double dec_ordered(unsigned long long u) {
  u = (u >> 63) ? (u & ~(1ull << 63)) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double v;
  memcpy(&v, &u, 8);
  return v;
#endif
}

// ------------------------------------------------------------------------------------------------
// Sorted top-K list in registers, ordered by (d2, idx). Unfilled entries: d2 = +inf, idx = INT_MAX.
// ------------------------------------------------------------------------------------------------
template <int K>
struct TopK {
  double d2[K];
  int idx[K];    // original point index
  int pos[K];    // position in the cell-sorted array (to re-load the coordinates)
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < K; ++j) { d2[j] = __longlong_as_double(0x7FF0000000000000ll); idx[j] = 0x7FFFFFFF; pos[j] = -1; }
  }
  __device__ __forceinline__ int count() const {
    int c = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) c += (pos[j] >= 0) ? 1 : 0;
    return c;
  }
  // branch-free sorted insertion
  __device__ __forceinline__ void insert(double d, int i, int p) {
    bool lt[K];
#pragma unroll
    for (int j = 0; j < K; ++j) lt[j] = (d < d2[j]) || (d == d2[j] && i < idx[j]);
    if (!lt[K - 1]) return;
#pragma unroll
    for (int j = K - 1; j >= 0; --j) {
      const bool shift = (j > 0) ? lt[j - 1] : false;   // new element lands before slot j-1 -> take j-1
      if (j > 0 && shift) { d2[j] = d2[j - 1]; idx[j] = idx[j - 1]; pos[j] = pos[j - 1]; }
      else if (lt[j]) { d2[j] = d; idx[j] = i; pos[j] = p; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Brick entry: uint4 A = {key.lo, key.hi, base, cnt01}, uint4 B = {cnt23, cnt45, cnt67, fine}; cntXY packs the
// u16 point counts of sub-cells X (low half) and Y (high half).  Sub-cell s = (cx&1) | (cy&1)<<1 | (cz&1)<<2,
// the brick's points are stored sub-cell after sub-cell from `base`.  fine = 0: no second level in this brick; else
// (index of the second-level table of the brick's FIRST dense sub-cell) + 1 -- its dense sub-cells (count >= kFineMin)
// own consecutive tables in sub-cell order.
// ------------------------------------------------------------------------------------------------
constexpr unsigned kMaxCellPoints = 65535u;
constexpr unsigned kBrickBytes = 32u;
// Second level: a cell with >= kFineMin points has them ordered by a 4 x 4 x 4 grid of fine bins (edge = cell / 4);
// its table holds the INCLUSIVE prefix sums of the 64 bin counts (bin = bx + 4 by + 16 bz), so bin i is
// [i ? table[i-1] : 0, table[i]) relative to the first point of the cell.
constexpr unsigned kFineMin = 64u;
constexpr int kFineDiv = 4;
constexpr int kFineBins = kFineDiv * kFineDiv * kFineDiv;
constexpr unsigned kFineEntryBytes = kFineBins * 2u;

__host__ __device__ __forceinline__ int brick_of(int c) { return c >> 1; }    // floor(c / 2), also for c < 0
__host__ __device__ __forceinline__ int subcell_of(int cx, int cy, int cz) { return (cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2); }

struct BrickEntry {
  uint4 a, b;
  __device__ __forceinline__ unsigned long long key() const { return ((unsigned long long)a.y << 32) | a.x; }
  __device__ __forceinline__ unsigned word(int i) const { return i == 0 ? a.w : (i == 1 ? b.x : (i == 2 ? b.y : b.z)); }
  __device__ __forceinline__ unsigned count(int s) const { return (word(s >> 1) >> (16 * (s & 1))) & 0xFFFFu; }
};

__device__ __forceinline__ BrickEntry load_brick(const GridDesc& g, unsigned slot) {
  BrickEntry e;
  e.a = __ldg(&g.table[2u * slot]);
  e.b = __ldg(&g.table[2u * slot + 1u]);
  return e;
}

// returns false if the brick is absent
__device__ __forceinline__ bool probe_brick(const GridDesc& g, unsigned long long key, BrickEntry& e) {
  unsigned s = hash_key(key) & g.mask;
  while (true) {
    e = load_brick(g, s);
    const unsigned long long k = e.key();
    if (k == key) return true;
    if (k == 0ull) return false;
    s = (s + 1u) & g.mask;
  }
}

// Exact kNN of the query (rx,ry,rz) [coordinates RELATIVE to the map origin, FP64] restricted to
// d2 < r2 (strict, like the std::lower_bound truncation in KDTreeFlann::SearchHybrid).  Plain
// thread-per-query form (k_knn, k_fitness).
template <int K>
__device__ __forceinline__ void knn_search(const GridDesc& g, double rx, double ry, double rz, double r2, TopK<K>& t) {
  t.init();
  if (g.n == 0u) return;
  const int cx = (int)floor(rx * g.inv_cell), cy = (int)floor(ry * g.inv_cell), cz = (int)floor(rz * g.inv_cell);
  const int bx0 = brick_of(cx - 1), by0 = brick_of(cy - 1), bz0 = brick_of(cz - 1);
#pragma unroll 1
  for (int ib = 0; ib < 8; ++ib) {
    const int bx = bx0 + (ib & 1), by = by0 + ((ib >> 1) & 1), bz = bz0 + (ib >> 2);
    BrickEntry e;
    if (!probe_brick(g, cell_key(bx, by, bz), e)) continue;
    unsigned beg = e.a.z;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const unsigned cnt = e.count(s);
      const int gx = 2 * bx + (s & 1), gy = 2 * by + ((s >> 1) & 1), gz = 2 * bz + (s >> 2);
      const bool need = (abs(gx - cx) <= 1) && (abs(gy - cy) <= 1) && (abs(gz - cz) <= 1);
      if (need) {
        for (unsigned j = 0; j < cnt; ++j) {
          const float4 m = __ldg(&g.pts[beg + j]);
          const double ddx = (double)m.x - rx, ddy = (double)m.y - ry, ddz = (double)m.z - rz;
          const double d = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));   // same three operations as the oracle
          if (d < r2) t.insert(d, __float_as_int(m.w), (int)(beg + j));
        }
      }
      beg += cnt;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Lane-pair search used by k_correspond: lanes 2q and 2q+1 share query q; the lane takes the four bricks of one
// z-layer of the 2x2x2 brick neighbourhood:
//   stage 1: the lane's 4 brick probes (8 LDG.128) are issued together before any is consumed; the needed,
//            non-empty sub-cells are compacted into a per-thread list in shared memory (s_beg/s_cnt/s_md,
//            [kPairCells][blockDim]) together with a lower bound of the distance from the query to the cell; the
//            nearest cell is moved to the front;
//   stage 2: both lanes stream the pair's unified cell list as one flattened sequence, alternating points
//            (adjacent 16 B loads), 8 loads in flight per thread; cells whose lower bound exceeds the K-th best
//            distance known to either lane are skipped;
//   merge  : lane 2q pulls lane 2q+1's sorted list (K shuffle rounds).
// Same result as knn_search (exact, ordered by (d2, original index)).  Twice the warps of the thread-per-query
// form for the same work: the kernel is latency-bound at F = 40k, so the extra warps buy issue slots.
// All 32 lanes must call this together.  On return the even lane holds the exact top-K.
// ------------------------------------------------------------------------------------------------
constexpr bool kMergeRuns = true;     // measured: 28.2 us per launch with merging, 29.2 without (config 2)
// (an FP32 pre-filter in front of the FP64 distance was measured in round 2: 186 vs 175 us per batched launch of 8
//  sequences, 2343 vs 2387 frames/s single stream -- the extra instructions cost more than the skipped FP64 ones)
constexpr unsigned kMergeMax = 16u;   // merged entry: at most this many points
constexpr int kPairCells = 18;    // 3 x 3 x 2 cells at most in one z-layer of bricks
// candidate queue in front of the sorted insertion: built and measured in round 2 (second half), SLOWER for K = 5 --
// 222 vs 171 us per batched launch of 8 sequences, 2 443 vs 2 562 frames/s single stream (the two-pass mask, the
// shared-memory round trip and the warp-uniform loop cost more than the ~40-instruction insertion they avoid; at K = 20,
// k_fe_pca, the same idea wins).  Kept behind the macro; poses are bit-identical either way.
#ifndef TLOAM_SEARCH_QUEUE
#define TLOAM_SEARCH_QUEUE 0
#endif
constexpr int kSearchQueue = 8;   // queued candidates per thread = one batch of loads

template <int K, int kThreads>
__device__ __forceinline__ void knn_search_pair(const GridDesc& g, bool live, double rx, double ry, double rz, double r2,
                                                unsigned (*s_beg)[kThreads], unsigned (*s_cnt)[kThreads],
                                                float (*s_md)[kThreads], TopK<K>& t, long long* stamps = nullptr,
                                                double (*q_d)[kThreads] = nullptr, int (*q_i)[kThreads] = nullptr,
                                                int (*q_p)[kThreads] = nullptr) {
  t.init();
  const int tid = threadIdx.x;
  const int half = tid & 1;
  if (stamps) stamps[0] = clock64();
  int m = 0;
  if (live && g.n != 0u) {
    const int cx = (int)floor(rx * g.inv_cell), cy = (int)floor(ry * g.inv_cell), cz = (int)floor(rz * g.inv_cell);
    const float cellf = (float)g.cell, r2f = (float)r2;
    const float fx = (float)(rx - (double)cx * g.cell), fy = (float)(ry - (double)cy * g.cell), fz = (float)(rz - (double)cz * g.cell);
    const int bx0 = brick_of(cx - 1), by0 = brick_of(cy - 1), bz = brick_of(cz - 1) + half;
    // per-axis lower bounds (squared, rounded DOWN) from the query to each of the 4 cell slabs the two bricks
    // span; slabs outside the 3-cell neighbourhood get +big, so "md < r2" also rejects the cells not needed
    float ax[4], ay[4], az[2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ox = 2 * bx0 + u - cx, oy = 2 * by0 + u - cy;
      const float vx = ox < 0 ? fx : cellf - fx, vy = oy < 0 ? fy : cellf - fy;
      ax[u] = (ox < -1 || ox > 1) ? 1.0e30f : (ox == 0 ? 0.0f : vx * vx * 0.9999f);
      ay[u] = (oy < -1 || oy > 1) ? 1.0e30f : (oy == 0 ? 0.0f : vy * vy * 0.9999f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int oz = 2 * bz + u - cz;
      const float vz = oz < 0 ? fz : cellf - fz;
      az[u] = ((oz < -1 || oz > 1) ? 1.0e30f : (oz == 0 ? 0.0f : vz * vz * 0.9999f)) - 1e-12f;
    }
    float md_min = 3.0e38f;
    int m_min = 0;
    unsigned long long key[4];
    unsigned slot[4];
    BrickEntry e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      key[j] = cell_key(bx0 + (j & 1), by0 + (j >> 1), bz);
      slot[j] = hash_key(key[j]) & g.mask;
      e[j] = load_brick(g, slot[j]);
    }
    // short runs that are contiguous in memory (neighbouring sub-cells of one brick) are merged into one list entry;
    // long runs stay separate so that dense cells can still be pruned one by one (config 3: 400 points per cell)
    bool open = false;
    unsigned run_beg = 0u, run_cnt = 0u;
    float run_md = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned long long k = e[j].key();
      while (k != key[j] && k != 0ull) {          // collision (rare): keep probing
        slot[j] = (slot[j] + 1u) & g.mask;
        e[j] = load_brick(g, slot[j]);
        k = e[j].key();
      }
      if (stamps && j == 3) stamps[1] = stamps[2] = clock64();
      if (k != key[j]) continue;
      unsigned run = e[j].a.z;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const unsigned cnt = e[j].count(s), beg = run;
        run += cnt;
        const float md = ax[2 * (j & 1) + (s & 1)] + ay[2 * (j >> 1) + ((s >> 1) & 1)] + az[s >> 2];
        if (cnt == 0u || !(md < r2f)) continue;
        if (kMergeRuns && open && run_beg + run_cnt == beg && run_cnt + cnt <= kMergeMax) {
          run_cnt += cnt;
          run_md = fminf(run_md, md);
        } else {
          if (open) {
            if (run_md < md_min) { md_min = run_md; m_min = m; }
            s_md[m][tid] = run_md; s_beg[m][tid] = run_beg; s_cnt[m][tid] = run_cnt;
            ++m;
          }
          open = true; run_beg = beg; run_cnt = cnt; run_md = md;
        }
      }
    }
    if (open) {
      if (run_md < md_min) { md_min = run_md; m_min = m; }
      s_md[m][tid] = run_md; s_beg[m][tid] = run_beg; s_cnt[m][tid] = run_cnt;
      ++m;
    }
    if (m_min > 0) {                               // nearest cell first: it fills the list with good candidates
      const float a0 = s_md[0][tid]; const unsigned b0 = s_beg[0][tid], c0 = s_cnt[0][tid];
      s_md[0][tid] = s_md[m_min][tid]; s_beg[0][tid] = s_beg[m_min][tid]; s_cnt[0][tid] = s_cnt[m_min][tid];
      s_md[m_min][tid] = a0; s_beg[m_min][tid] = b0; s_cnt[m_min][tid] = c0;
    }
  }
  // stage 2 walks the UNIFIED cell list of the pair (even lane's entries, then the odd lane's); in every run
  // the even lane takes points 0,2,4.. and the odd lane 1,3,5..: the pair's two loads are adjacent (one 32 B
  // sector), so a warp-wide load touches 16 lines instead of 32 -- this stage is bound by L1TEX wavefronts.
  __syncwarp();
  const int col0 = tid & ~1, col1 = tid | 1;
  const unsigned pairmask = 3u << ((tid & 31) & ~1);
  const int m_other = __shfl_xor_sync(0xffffffffu, m, 1);
  const int m0 = half ? m_other : m;
  const int mt = m + m_other;
  int ci = 0;
  unsigned off = (unsigned)half, cb = 0u, cc = 0u;
  if (mt > 0) { const int col = (0 < m0) ? col0 : col1; cb = s_beg[0][col]; cc = s_cnt[0][col]; }
#if TLOAM_SEARCH_QUEUE
  // Candidate queue (K > 1): once a lane's list is full, a candidate that beats the lane's K-th best (as of the last drain)
  // is only QUEUED (shared memory, kSearchQueue per thread); the queues are drained into the sorted lists by all lanes
  // together, when one of them could overflow on the next batch and once at the end.  The sorted insertion (~40
  // instructions) then runs a few times per query with most lanes busy, instead of once per candidate batch slot
  // whenever ANY lane of the warp has a candidate (k_fe_pca does the same for K = 20).  The set of candidates that
  // reaches insert() is a superset of what the direct form inserts, and insert() orders by (d2, index): same result.
  int qn = 0;
  double kth = t.d2[K - 1];
  auto drain = [&]() {
    const int most = __reduce_max_sync(0xffffffffu, qn);
    for (int e = 0; e < most; ++e)
      if (e < qn) t.insert(q_d[e][tid], q_i[e][tid], q_p[e][tid]);
    qn = 0;
    kth = t.d2[K - 1];
  };
#endif
#if TLOAM_SEARCH_QUEUE
  // warp-uniform loop (the drain uses warp-wide collectives): lanes whose pair is done load nothing
  while (__any_sync(0xffffffffu, ci < mt)) {
#else
  while (ci < mt) {
#endif
    float4 pt[8];
    int pos[8];
    double worst = t.d2[K - 1];                   // +inf until the list is full
    worst = fmin(worst, __shfl_xor_sync(pairmask, worst, 1));   // K candidates <= worst exist in one of the lists
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      pos[q] = -1;
      if (ci < mt) {
        if (off < cc) {
          pos[q] = (int)(cb + off);
          pt[q] = __ldg(&g.pts[pos[q]]);
        }
        off += 2u;
        if (off - (unsigned)half >= cc) {          // pair-uniform: the run is exhausted
          ++ci; off = (unsigned)half;
          while (ci < mt) {                        // skip cells no point of which can enter the list
            const float mdv = (ci < m0) ? s_md[ci][col0] : s_md[ci - m0][col1];
            if (!((double)mdv > worst)) break;
            ++ci;
          }
          if (ci < mt) {
            const int col = (ci < m0) ? col0 : col1, row = (ci < m0) ? ci : ci - m0;
            cb = s_beg[row][col]; cc = s_cnt[row][col];
          }
        }
      }
    }
#if TLOAM_SEARCH_QUEUE
    if (K > 1 && q_d != nullptr) {
      const bool full = t.pos[K - 1] >= 0;
      const double lim = full ? kth : r2;                      // list not full yet: everything inside the radius goes in directly
      // pass 1: which of the 8 would be queued (a bit mask: the distances are recomputed for the few that are)
      unsigned pass = 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (pos[q] >= 0) {
          const double ddx = (double)pt[q].x - rx, ddy = (double)pt[q].y - ry, ddz = (double)pt[q].z - rz;
          const double d = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));   // same three operations as the oracle
          if (d < r2) {
            if (!full) t.insert(d, __float_as_int(pt[q].w), pos[q]);
            else if (d <= lim) pass |= 1u << q;
          }
        }
      }
      if (__any_sync(0xffffffffu, qn + __popc(pass) > kSearchQueue)) drain();   // room first (the mask stays a superset)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if ((pass >> q) & 1u) {
          const double ddx = (double)pt[q].x - rx, ddy = (double)pt[q].y - ry, ddz = (double)pt[q].z - rz;
          q_d[qn][tid] = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));
          q_i[qn][tid] = __float_as_int(pt[q].w); q_p[qn][tid] = pos[q];
          ++qn;
        }
      }
    } else
#endif
    {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (pos[q] >= 0) {
          const double ddx = (double)pt[q].x - rx, ddy = (double)pt[q].y - ry, ddz = (double)pt[q].z - rz;
          const double d = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));   // same three operations as the oracle
          if (d < r2) t.insert(d, __float_as_int(pt[q].w), pos[q]);
        }
      }
    }
  }
#if TLOAM_SEARCH_QUEUE
  if (K > 1 && q_d != nullptr && __any_sync(0xffffffffu, qn > 0)) drain();
#endif
  if (stamps) stamps[3] = clock64();
  // merge: the even lane pulls the odd lane's list (sorted) and inserts its entries until one fails
  __syncwarp();
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double od = __shfl_xor_sync(0xffffffffu, t.d2[j], 1);
    const int oi = __shfl_xor_sync(0xffffffffu, t.idx[j], 1);
    const int op = __shfl_xor_sync(0xffffffffu, t.pos[j], 1);
    if (half == 0 && op >= 0) t.insert(od, oi, op);
  }
  if (stamps) stamps[4] = clock64();
}

}  // namespace tloam
