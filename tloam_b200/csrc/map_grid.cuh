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

}  // namespace tloam
