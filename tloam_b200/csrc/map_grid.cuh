// map_grid.cuh -- GPU-resident voxel-hash local map + exact radius-truncated kNN.
//
// Replaces the four Open3D KDTreeFlann objects the reference rebuilds on every scanMatching call
// (ref: src/models/registration/registration.cpp:888-915) and their SearchHybrid(query, r, k) queries
// (ref: registration.cpp:272, 444, 535, 588, 731).
//
// One grid per cloud, cell edge == that cloud's search radius, so the 27 cells around a query contain
// every point with |m - q| < r: the search is EXACT for radius-truncated kNN (no approximation).
// Points are stored cell-sorted as float4 (x,y,z relative to a per-map origin, w = original index bits):
// 16 B / point, one LDG.128 per candidate.  The hash table is open-addressing, 16 B / entry
// {key 8 B, start 4 B, count 4 B}: one LDG.128 per probe.  All distance arithmetic is FP64 on the FP32
// stored coordinates, ordering is (d2, original index) => results do not depend on the in-cell order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tloam {

struct GridDesc {
  const float4* pts;    // [n] cell-sorted
  const uint4* table;   // [mask+1] {key.lo, key.hi, start, count}; key == 0 => empty
  unsigned mask;
  unsigned n;
  double inv_cell;
  double cell;
};

// 256-byte header at the start of the map blob (host + device visible layout)
struct MapHeader {
  unsigned long long magic;
  unsigned n[4];
  unsigned tsize[4];        // table entries (power of two)
  unsigned long long pts_off[4], table_off[4];  // byte offsets inside the blob
  double origin[3];
  double cell[4];
  unsigned long long bbox_enc[6];  // ordered-uint encodings of min xyz / max xyz (build scratch)
  unsigned cursor[4];              // bump allocators (build scratch)
  unsigned long long pad[4];
};
static_assert(sizeof(MapHeader) == 256, "MapHeader must be 256 bytes");
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

__device__ __forceinline__ uint4 probe_cell(const GridDesc& g, unsigned long long key) {
  unsigned s = hash_key(key) & g.mask;
  uint4 e = __ldg(&g.table[s]);
  while (true) {
    const unsigned long long k = ((unsigned long long)e.y << 32) | e.x;
    if (k == key) return e;
    if (k == 0ull) { e.w = 0u; return e; }
    s = (s + 1u) & g.mask;
    e = __ldg(&g.table[s]);
  }
}

// Exact kNN of the query (rx,ry,rz) [coordinates RELATIVE to the map origin, FP64] restricted to
// d2 < r2 (strict, like the std::lower_bound truncation in KDTreeFlann::SearchHybrid).
template <int K>
__device__ __forceinline__ void knn_search(const GridDesc& g, double rx, double ry, double rz, double r2, TopK<K>& t) {
  t.init();
  if (g.n == 0u) return;
  const int cx = (int)floor(rx * g.inv_cell), cy = (int)floor(ry * g.inv_cell), cz = (int)floor(rz * g.inv_cell);
#pragma unroll 1
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll 1
    for (int dy = -1; dy <= 1; ++dy) {
      uint4 e[3];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) e[dx] = probe_cell(g, cell_key(cx + dx - 1, cy + dy, cz + dz));
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const unsigned beg = e[dx].z, cnt = e[dx].w;
        for (unsigned j = 0; j < cnt; ++j) {
          const float4 m = __ldg(&g.pts[beg + j]);
          const double ddx = (double)m.x - rx, ddy = (double)m.y - ry, ddz = (double)m.z - rz;
          const double d = ddx * ddx + ddy * ddy + ddz * ddz;
          if (d < r2) t.insert(d, __float_as_int(m.w), (int)(beg + j));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Lane-pair search used by k_correspond: lanes 2q and 2q+1 share query q, each visits every other cell of the
// 27-cell neighbourhood (14 / 13 cells, centre first):
//   stage 1: the lane's hash probes are issued as two batches of 7 independent LDG.128 before any is consumed;
//            the non-empty cells are compacted into a per-thread list in shared memory (s_beg/s_cnt/s_md,
//            [14][blockDim]) together with a lower bound of the distance from the query to the cell;
//   stage 2: the cells' point runs are streamed as one flattened sequence, 8 loads in flight per thread, cells
//            whose lower bound exceeds the current K-th best distance are skipped;
//   merge  : lane 2q pulls lane 2q+1's sorted list (K shuffle rounds).
// Same result as knn_search (exact, ordered by (d2, original index)).  Twice the warps of the thread-per-query form for the same work: the kernel is latency-bound
// at F = 40k (8.5 warps per SM), so the extra warps buy issue slots.  All 32 lanes must call this together.
// On return the even lane holds the exact top-K.
// ------------------------------------------------------------------------------------------------
template <int K, int kThreads>
__device__ __forceinline__ void knn_search_pair(const GridDesc& g, bool live, double rx, double ry, double rz, double r2,
                                                unsigned (*s_beg)[kThreads], unsigned (*s_cnt)[kThreads],
                                                float (*s_md)[kThreads], TopK<K>& t, long long* stamps = nullptr) {
  t.init();
  const int tid = threadIdx.x;
  const int half = tid & 1;
  if (stamps) stamps[0] = clock64();
  int m = 0;
  if (live && g.n != 0u) {
    const int cx = (int)floor(rx * g.inv_cell), cy = (int)floor(ry * g.inv_cell), cz = (int)floor(rz * g.inv_cell);
    const float cellf = (float)g.cell, r2f = (float)r2;
    const float fx = (float)(rx - (double)cx * g.cell), fy = (float)(ry - (double)cy * g.cell), fz = (float)(rz - (double)cz * g.cell);
#pragma unroll
    for (int batch = 0; batch < 2; ++batch) {
      unsigned long long key[7];
      uint4 e[7];
      unsigned slot[7];
      int off3[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int n = 2 * (batch * 7 + j) + half;               // 0..27 ; 27 = none
        const int mcell = (n + 13) % 27;                        // visiting order starts at the centre cell (13)
        off3[j] = (n < 27) ? mcell : -1;
        key[j] = 0ull; slot[j] = 0u; e[j] = make_uint4(0u, 0u, 0u, 0u);
        if (n < 27) {
          key[j] = cell_key(cx + (mcell % 3) - 1, cy + ((mcell / 3) % 3) - 1, cz + (mcell / 9) - 1);
          slot[j] = hash_key(key[j]) & g.mask;
          e[j] = __ldg(&g.table[slot[j]]);
        }
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        if (off3[j] < 0) continue;
        unsigned long long k = ((unsigned long long)e[j].y << 32) | e[j].x;
        while (k != key[j] && k != 0ull) {          // collision (rare): keep probing
          slot[j] = (slot[j] + 1u) & g.mask;
          e[j] = __ldg(&g.table[slot[j]]);
          k = ((unsigned long long)e[j].y << 32) | e[j].x;
        }
        if (stamps && j == 6) stamps[1 + batch] = clock64();
        if (k == key[j] && e[j].w != 0u) {
          const int ox = (off3[j] % 3) - 1, oy = ((off3[j] / 3) % 3) - 1, oz = (off3[j] / 9) - 1;
          const float mx = ox < 0 ? fx : (ox > 0 ? cellf - fx : 0.0f);
          const float my = oy < 0 ? fy : (oy > 0 ? cellf - fy : 0.0f);
          const float mz = oz < 0 ? fz : (oz > 0 ? cellf - fz : 0.0f);
          const float md = fmaxf(mx * mx + my * my + mz * mz, 0.0f) * 0.9999f - 1e-12f;   // rounded DOWN
          if (md < r2f) {
            s_md[m][tid] = md;
            s_beg[m][tid] = e[j].z;
            s_cnt[m][tid] = e[j].w;
            ++m;
          }
        }
      }
    }
  }
  int ci = 0;
  unsigned off = 0u, cb = 0u, cc = 0u;
  if (m > 0) { cb = s_beg[0][tid]; cc = s_cnt[0][tid]; }
  while (ci < m) {
    float4 pt[8];
    int pos[8];
    const double worst = t.d2[K - 1];             // +inf until the list is full
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      pos[q] = -1;
      if (ci < m) {
        pos[q] = (int)(cb + off);
        pt[q] = __ldg(&g.pts[pos[q]]);
        if (++off == cc) {
          ++ci; off = 0u;
          while (ci < m && (double)s_md[ci][tid] > worst) ++ci;   // no point of that cell can enter the list
          if (ci < m) { cb = s_beg[ci][tid]; cc = s_cnt[ci][tid]; }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (pos[q] >= 0) {
        const double ddx = (double)pt[q].x - rx, ddy = (double)pt[q].y - ry, ddz = (double)pt[q].z - rz;
        const double d = ddx * ddx + ddy * ddy + ddz * ddz;
        if (d < r2) t.insert(d, __float_as_int(pt[q].w), pos[q]);
      }
    }
  }
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
