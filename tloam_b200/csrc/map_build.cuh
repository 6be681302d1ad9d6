// map_build.cuh -- voxel-hash map build kernels (brick-keyed), once per set_target.
// Replaces the four KDTreeFlann::SetGeometry calls of ref: src/models/registration/registration.cpp:888-915.
#pragma once
#include "registration.cuh"

namespace tloam {

struct MapBuildArgs {
  const double* src[4];         // AoS xyz of each cloud (the staging buffer for host input, the caller's arrays for device input)
  unsigned stage_off[5];        // global point index of each cloud's first point
  unsigned char* blob;          // MapHeader + pts + tables
  unsigned* slot_of;            // scratch [total]
  unsigned* rank_of;            // scratch [total]
  unsigned i_beg, i_end;        // global point range this launch works on (one cloud in the pipelined host path)
  int only_cloud;               // k_map_offsets: table of this cloud only (-1 = all four)
  // device-resident point counts (sync-free submap chain): stage_off / i_end are then CAPACITY bounds known to the
  // host, and cloud c really holds *n_dev[c] points (nullptr = the bound is exact)
  const unsigned* n_dev[4];
  // second level (map_grid.cuh: kFineMin): clouds of fine_mask get their dense cells ordered by fine bin (k_map_fine)
  int fine_mask;
  float4* fine_tmp;             // scratch [total]: k_map_fine reorders a cell out of place
};

// false for the slack between a cloud's device-side count and its host-side bound
__device__ __forceinline__ bool map_point_live(const MapBuildArgs& a, unsigned i, int c) {
  const unsigned* nd = c == 0 ? a.n_dev[0] : c == 1 ? a.n_dev[1] : c == 2 ? a.n_dev[2] : a.n_dev[3];
  return nd == nullptr || (i - a.stage_off[c]) < *nd;
}

__device__ __forceinline__ int cloud_of_point(const MapBuildArgs& a, unsigned i) {
  return (i >= a.stage_off[3]) ? 3 : (i >= a.stage_off[2]) ? 2 : (i >= a.stage_off[1]) ? 1 : 0;
}
__device__ __forceinline__ const double* point_of(const MapBuildArgs& a, unsigned i) {
  const int c = cloud_of_point(a, i);
  const double* base = c == 0 ? a.src[0] : c == 1 ? a.src[1] : c == 2 ? a.src[2] : a.src[3];
  return base + 3ull * (i - a.stage_off[c]);
}

// bounding box of the points [i_beg, i_end) -- the first non-empty cloud: the origin must be known before any
// point can be inserted, and the host path inserts cloud c while cloud c+1 is still crossing PCIe
__global__ void k_map_bbox(MapBuildArgs a) {
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (unsigned i = a.i_beg + blockIdx.x * blockDim.x + threadIdx.x; i < a.i_end; i += gridDim.x * blockDim.x) {
    if (!map_point_live(a, i, cloud_of_point(a, i))) break;         // one cloud per launch: the live points come first
    const double* pt = point_of(a, i);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double v = pt[d];
      mn[d] = fmin(mn[d], v); mx[d] = fmax(mx[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  __shared__ double s_mn[8][3], s_mx[8][3];
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_mn[threadIdx.x >> 5][d] = mn[d]; s_mx[threadIdx.x >> 5][d] = mx[d]; }
  }
  __syncthreads();
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  if (threadIdx.x < 3) {                       // 6 atomics per block (same-address atomics serialise in L2)
    const int d = threadIdx.x;
    double lo = s_mn[0][d], hi = s_mx[0][d];
    for (int wi = 1; wi < (int)(blockDim.x >> 5); ++wi) { lo = fmin(lo, s_mn[wi][d]); hi = fmax(hi, s_mx[wi][d]); }
    atomicMin(&h->bbox_enc[d], enc_ordered(lo));
    atomicMax(&h->bbox_enc[3 + d], enc_ordered(hi));
  }
  // the last block to arrive turns the bounding box into the map origin = integer-rounded centre of the box
  // (exactly representable; |rel| stays small so the FP32 storage keeps ~8e-6 m resolution at 100 m)
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&h->bbox_ticket, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (s_last && threadIdx.x < 3) {
    __threadfence();
    const double lo = dec_ordered(atomicMin(&h->bbox_enc[threadIdx.x], ~0ull)), hi = dec_ordered(atomicMax(&h->bbox_enc[3 + threadIdx.x], 0ull));
    h->origin[threadIdx.x] = (lo <= hi) ? rint(0.5 * (lo + hi)) : 0.0;      // empty (device-counted) cloud: origin 0
  }
}

// empty map: origin 0
__global__ void k_map_origin(MapBuildArgs a) {
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  if (threadIdx.x < 3) h->origin[threadIdx.x] = 0.0;
}

__device__ __forceinline__ float3 rel_of(const MapBuildArgs& a, const MapHeader* h, unsigned i) {
  const double* pt = point_of(a, i);
  return make_float3((float)(pt[0] - h->origin[0]), (float)(pt[1] - h->origin[1]), (float)(pt[2] - h->origin[2]));
}

// cell of the STORED (FP32-rounded) coordinates, so that the 27-cell search is exact for what is stored
__device__ __forceinline__ void cell_of_rel(const float3 r, double inv, int& cx, int& cy, int& cz) {
  cx = (int)floor((double)r.x * inv); cy = (int)floor((double)r.y * inv); cz = (int)floor((double)r.z * inv);
}

// claims the point's brick (CAS on the key) and takes a rank inside its sub-cell (packed u16 counters)
__global__ void k_map_insert(MapBuildArgs a) {
  const unsigned i = a.i_beg + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.i_end) return;
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  const int c = cloud_of_point(a, i);
  if (!map_point_live(a, i, c)) return;
  const float3 r = rel_of(a, h, i);
  int cx, cy, cz;
  cell_of_rel(r, 1.0 / h->cell[c], cx, cy, cz);
  const unsigned long long key = cell_key(brick_of(cx), brick_of(cy), brick_of(cz));
  const int sub = subcell_of(cx, cy, cz);
  uint4* table = reinterpret_cast<uint4*>(a.blob + h->table_off[c]);
  const unsigned mask = h->tsize[c] - 1u;
  unsigned s = hash_key(key) & mask;
  while (true) {
    unsigned long long* kp = reinterpret_cast<unsigned long long*>(&table[2u * s]);
    const unsigned long long prev = atomicCAS(kp, 0ull, key);
    if (prev == 0ull || prev == key) break;
    s = (s + 1u) & mask;
  }
  // words 3..6 of the entry hold the 8 u16 counts
  unsigned* words = reinterpret_cast<unsigned*>(&table[2u * s]) + 3;
  const unsigned old = atomicAdd(&words[sub >> 1], (sub & 1) ? 0x10000u : 1u);
  const unsigned rank = (sub & 1) ? (old >> 16) : (old & 0xFFFFu);
  if (rank >= kMaxCellPoints) atomicOr(&h->build_flags, 1ull);    // the packed counter would wrap: map unusable
  a.slot_of[i] = s;
  a.rank_of[i] = rank;
}

// base offsets of the occupied bricks: block-level exclusive scan of the brick totals + ONE atomic per block on
// the cloud's bump allocator (table sizes are multiples of the block size, so a block never straddles two clouds)
__global__ void __launch_bounds__(256) k_map_offsets(MapBuildArgs a) {
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  if (a.only_cloud >= 0) c = a.only_cloud;
  else while (c < 3 && s >= h->tsize[c]) { s -= h->tsize[c]; ++c; }
  uint4* table = reinterpret_cast<uint4*>(a.blob + h->table_off[c]);
  unsigned cnt = 0u;
  if (s < h->tsize[c]) {
    const uint4 ea = table[2u * s];
    if ((ea.x | ea.y) != 0u) {
      const uint4 eb = table[2u * s + 1u];
      cnt = (ea.w & 0xFFFFu) + (ea.w >> 16) + (eb.x & 0xFFFFu) + (eb.x >> 16) + (eb.y & 0xFFFFu) + (eb.y >> 16) +
            (eb.z & 0xFFFFu) + (eb.z >> 16);
    }
  }
  // warp inclusive scan
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = cnt;
  for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  __shared__ unsigned s_w[8];
  __shared__ unsigned s_base;
  if (lane == 31) s_w[warp] = incl;
  const int nocc = __syncthreads_count(cnt > 0u);
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int wi = 0; wi < 8; ++wi) { const unsigned v = s_w[wi]; s_w[wi] = tot; tot += v; }
    s_base = (tot > 0u) ? atomicAdd(&h->cursor[c], tot) : 0u;
    if (nocc > 0) atomicAdd(&h->nbricks[c], (unsigned)nocc);
  }
  __syncthreads();
  if (cnt > 0u) table[2u * s].z = s_base + s_w[warp] + (incl - cnt);
  // second level: consecutive tables for the brick's dense sub-cells; the first 8 bytes of a table carry its cell
  // (slot, sub-cell) to k_map_fine, which overwrites them with the bin boundaries
  if (!((a.fine_mask >> c) & 1)) return;           // block-uniform
  unsigned nd = 0u, dmask = 0u;
  if (cnt >= kFineMin) {
    const uint4 ea = table[2u * s], eb = table[2u * s + 1u];
    const unsigned w[4] = {ea.w, eb.x, eb.y, eb.z};
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (((w[q >> 1] >> (16 * (q & 1))) & 0xFFFFu) >= kFineMin) { ++nd; dmask |= 1u << q; }
  }
  if (__syncthreads_count(nd > 0u) == 0) return;
  unsigned incl2 = nd;
  for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl2, o); if (lane >= o) incl2 += v; }
  if (lane == 31) s_w[warp] = incl2;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int wi = 0; wi < 8; ++wi) { const unsigned v = s_w[wi]; s_w[wi] = tot; tot += v; }
    s_base = atomicAdd(&h->fine_cursor[c], tot);
  }
  __syncthreads();
  if (nd > 0u) {
    unsigned fi = s_base + s_w[warp] + (incl2 - nd);
    table[2u * s + 1u].w = fi + 1u;
    unsigned char* fine = a.blob + h->fine_off[c];
    for (int q = 0; q < 8; ++q)
      if ((dmask >> q) & 1u) {
        uint2* tag = reinterpret_cast<uint2*>(fine + (size_t)fi * kFineEntryBytes);
        *tag = make_uint2(s, (unsigned)q);
        ++fi;
      }
  }
}

// Second level of the dense cells: one block per dense cell (grid-stride over the cloud's work list = its tables).
// Pass A copies the cell to scratch and counts the fine bins, pass B writes the points back bin after bin and the
// table of inclusive bin boundaries.  Bin of a point: floor(4 * frac(stored coordinate / cell)) per axis, FP64.
__global__ void __launch_bounds__(128) k_map_fine(MapBuildArgs a) {
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  if (h->build_flags & 1ull) return;
  __shared__ unsigned s_cnt[kFineBins], s_cur[kFineBins];
  for (int c = 0; c < 4; ++c) {
    if (!((a.fine_mask >> c) & 1)) continue;
    const unsigned ncell = h->fine_cursor[c];
    if (ncell == 0u) continue;
    uint4* table = reinterpret_cast<uint4*>(a.blob + h->table_off[c]);
    float4* pts = reinterpret_cast<float4*>(a.blob + h->pts_off[c]);
    float4* tmp = a.fine_tmp + a.stage_off[c];
    unsigned char* fine = a.blob + h->fine_off[c];
    const double inv = 1.0 / h->cell[c];
    for (unsigned e = blockIdx.x; e < ncell; e += gridDim.x) {
      unsigned short* ft = reinterpret_cast<unsigned short*>(fine + (size_t)e * kFineEntryBytes);
      const uint2 tag = *reinterpret_cast<const uint2*>(ft);
      const unsigned slot = tag.x;
      const int sub = (int)tag.y;
      const uint4 ea = table[2u * slot], eb = table[2u * slot + 1u];
      const unsigned w[4] = {ea.w, eb.x, eb.y, eb.z};
      unsigned beg = ea.z, cnt = 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned cq = (w[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;
        if (q < sub) beg += cq;
        if (q == sub) cnt = cq;
      }
      // integer cell coordinates from the brick key (21-bit fields biased by 2^20) and the sub-cell bits
      const unsigned long long key = ((unsigned long long)ea.y << 32) | ea.x;
      const int gx = 2 * ((int)((key >> 42) & 0x1FFFFFu) - (1 << 20)) + (sub & 1);
      const int gy = 2 * ((int)((key >> 21) & 0x1FFFFFu) - (1 << 20)) + ((sub >> 1) & 1);
      const int gz = 2 * ((int)(key & 0x1FFFFFu) - (1 << 20)) + (sub >> 2);
      __syncthreads();                               // previous cell done with s_cnt / s_cur / its tag
      if (threadIdx.x < kFineBins) s_cnt[threadIdx.x] = 0u;
      __syncthreads();
      auto bin_of = [&](const float4 p) {
        int bx = (int)(((double)p.x * inv - (double)gx) * (double)kFineDiv), by = (int)(((double)p.y * inv - (double)gy) * (double)kFineDiv),
            bz = (int)(((double)p.z * inv - (double)gz) * (double)kFineDiv);
        bx = bx < 0 ? 0 : (bx > kFineDiv - 1 ? kFineDiv - 1 : bx);
        by = by < 0 ? 0 : (by > kFineDiv - 1 ? kFineDiv - 1 : by);
        bz = bz < 0 ? 0 : (bz > kFineDiv - 1 ? kFineDiv - 1 : bz);
        return (bz * kFineDiv + by) * kFineDiv + bx;
      };
      for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) {
        const float4 p = pts[beg + i];
        tmp[beg + i] = p;
        atomicAdd(&s_cnt[bin_of(p)], 1u);
      }
      __syncthreads();
      if (threadIdx.x < 32) {                        // inclusive scan of the 64 bins: 2 per lane
        const unsigned c0 = s_cnt[2 * threadIdx.x], c1 = s_cnt[2 * threadIdx.x + 1];
        unsigned incl = c0 + c1;
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += v; }
        const unsigned excl = incl - (c0 + c1);
        s_cur[2 * threadIdx.x] = excl; s_cur[2 * threadIdx.x + 1] = excl + c0;
        ft[2 * threadIdx.x] = (unsigned short)(excl + c0); ft[2 * threadIdx.x + 1] = (unsigned short)incl;
      }
      __syncthreads();
      for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) {
        const float4 p = tmp[beg + i];
        pts[beg + atomicAdd(&s_cur[bin_of(p)], 1u)] = p;
      }
    }
  }
}

__global__ void k_map_scatter(MapBuildArgs a) {
  const unsigned i = a.i_beg + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.i_end) return;
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  if (h->build_flags & 1ull) return;               // counters wrapped: destinations are meaningless
  const int c = cloud_of_point(a, i);
  if (!map_point_live(a, i, c)) return;
  const float3 r = rel_of(a, h, i);
  int cx, cy, cz;
  cell_of_rel(r, 1.0 / h->cell[c], cx, cy, cz);
  const int sub = subcell_of(cx, cy, cz);
  const uint4* table = reinterpret_cast<const uint4*>(a.blob + h->table_off[c]);
  float4* pts = reinterpret_cast<float4*>(a.blob + h->pts_off[c]);
  const unsigned slot = a.slot_of[i];
  const uint4 ea = table[2u * slot], eb = table[2u * slot + 1u];
  const unsigned w[4] = {ea.w, eb.x, eb.y, eb.z};
  unsigned dst = ea.z + a.rank_of[i];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < sub) dst += (w[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;
  pts[dst] = make_float4(r.x, r.y, r.z, __int_as_float((int)(i - a.stage_off[c])));
}

// AoS FP64 staging -> padded SoA feature arrays
__global__ void k_stage_source(const double* s0, const double* s1, const double* s2, const double* s3, DeviceCtx ctx,
                               double* px, double* py, double* pz) {
  const int b = blockIdx.x;
  const int c = cloud_of_block(ctx, b);
  const int il = (b - ctx.blk_off[c]) * kBlk + threadIdx.x;
  const int gi = ctx.pad_off[c] + il;
  const double* src = (c == 0) ? s0 : (c == 1) ? s1 : (c == 2) ? s2 : s3;
  double x = 0, y = 0, z = 0;
  if (il < ctx.n[c]) {
    const double* p = src + 3ull * il;
    x = p[0]; y = p[1]; z = p[2];
  }
  px[gi] = x; py[gi] = y; pz[gi] = z;
}

}  // namespace tloam
