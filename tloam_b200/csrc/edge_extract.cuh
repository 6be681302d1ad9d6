// edge_extract.cuh -- "next" row (f)-4, second part: LOAM-style edge extraction of the segmentation nodelet,
// Segmentation::extractEdgePoint + extractFromSection (ref: src/models/segmentation/segmentation.cpp:1144-1304).
//
// Input: a cloud (AoS xyz) whose intensity channel holds the beam id of every point (what groundRemove leaves there).
// Output: index lists into the input -- edge points in (beam, sector, descending curvature) order, non-edge points in
// (beam, sector, ascending curvature) order, exactly the order in which the reference appends the points.
//
//   k_ee_key      beam of every point + per-chunk histogram (64 keys)          \  stable partition by beam =
//   k_ee_scan     per-beam exclusive scan over the chunks, beam bases           >  the reference's ringScans[beam]
//   k_ee_scatter  order[base + rank] = point                                   /   (input order kept inside a beam)
//   k_ee_section  one block per (beam, sector): curvature of the sector's ring positions (11-point stencil, summed
//                 left to right with round-to-nearest adds / multiplies: bit-exact against the -ffp-contract=off
//                 oracle), bitonic sort by (curvature, ring position) in shared memory, the serial pick loop (<= 20
//                 picks above 0.1, +-5 neighbour suppression while the gap^2 <= 0.05) on one thread with a bitmap,
//                 stable compaction of the unpicked points in ascending-curvature order
//   k_ee_offsets  exclusive scan of the 64 x 6 section counts;  k_ee_copy  section lists -> final positions
// Reference quirks kept: the last curvature of every sector belongs to neither output (exclusive iterator end, :1292);
// the 21st candidate of a sector is marked picked before the loop ends (:1165-1175).  Fixed where the reference leaves
// it to the implementation: ties of std::sort by ring position; beam id == sensorModel (one past the end of ringScans
// in the reference) is dropped.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tloam {

constexpr int kEeChunk = 256;
constexpr int kEeKeys = 64;           // beams (sensorModel <= 64)
constexpr int kEeSectors = 6;
constexpr int kEeMaxSection = 4096;   // curvature values per sector (ring <= 24 586 points)
constexpr int kEeThreads = 256;

struct EeArgs {
  const double* pts;                   // AoS xyz
  const double* intensity;             // beam id per point
  unsigned n, nchunk;
  int sensor_model, ring_min;
  unsigned char* key;                  // [n] beam, 255 = dropped
  unsigned* chunk_cnt;                 // [nchunk][kEeKeys] counts, then exclusive offsets per beam
  unsigned* ring_base;                 // [kEeKeys + 1]
  unsigned* order;                     // [n] ring-sorted position -> input index
  unsigned* sec_edge;                  // [n] per-section edge lists (input indices) at the section's ring-sorted offset
  unsigned* sec_non;                   // [n] per-section non-edge lists
  unsigned* sec_cnt;                   // [kEeKeys * kEeSectors][2] edge / non-edge counts
  unsigned* sec_off;                   // [kEeKeys * kEeSectors + 1][2] exclusive offsets; last row = totals
  unsigned long long* out_edge;        // [n]
  unsigned long long* out_non;         // [n]
  int* status;                         // 1 = a sector exceeds kEeMaxSection
};

__device__ __forceinline__ double ee_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ee_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ee_mul(double a, double b) { return __dmul_rn(a, b); }

// rank of this thread's key among the earlier threads of the block with the same key; optionally the block's histogram
__device__ __forceinline__ unsigned ee_rank_in_block(int key, unsigned (*s_cnt)[kEeKeys], unsigned* hist_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = lane; k < kEeKeys; k += 32) s_cnt[warp][k] = 0u;
  __syncwarp();
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const unsigned rank_w = __popc(peers & ((1u << lane) - 1u));
  if (key >= 0 && rank_w == 0u) s_cnt[warp][key] = __popc(peers);
  __syncthreads();
  unsigned before = 0u;
  if (key >= 0)
    for (int w = 0; w < warp; ++w) before += s_cnt[w][key];
  if (hist_out)
    for (int k = threadIdx.x; k < kEeKeys; k += blockDim.x) {
      unsigned t = 0u;
      for (int w = 0; w < kEeChunk / 32; ++w) t += s_cnt[w][k];
      hist_out[k] = t;
    }
  return before + rank_w;
}

__global__ void __launch_bounds__(kEeChunk) k_ee_key(const __grid_constant__ EeArgs a) {
  const unsigned i = blockIdx.x * kEeChunk + threadIdx.x;
  __shared__ unsigned s_cnt[kEeChunk / 32][kEeKeys];
  int key = -1;
  if (i < a.n) {
    const int beam = (int)a.intensity[i];                                  // static_cast<int>, :1232
    if (beam >= 0 && beam < a.sensor_model) key = beam;
    a.key[i] = (unsigned char)(key < 0 ? 255 : key);
  }
  ee_rank_in_block(key, s_cnt, a.chunk_cnt + (size_t)blockIdx.x * kEeKeys);
}

// one block; thread k < 64 scans beam k over the chunks (nchunk ~ 500: a serial loop per beam, 64 beams in parallel)
__global__ void __launch_bounds__(kEeKeys) k_ee_scan(const __grid_constant__ EeArgs a) {
  __shared__ unsigned s_tot[kEeKeys];
  const int k = threadIdx.x;
  unsigned run = 0u;
  for (unsigned c = 0; c < a.nchunk; ++c) {
    const unsigned v = a.chunk_cnt[(size_t)c * kEeKeys + k];
    a.chunk_cnt[(size_t)c * kEeKeys + k] = run;
    run += v;
  }
  s_tot[k] = run;
  __syncthreads();
  if (k == 0) {
    unsigned off = 0u;
    for (int b = 0; b < kEeKeys; ++b) { a.ring_base[b] = off; off += s_tot[b]; }
    a.ring_base[kEeKeys] = off;
    *a.status = 0;
  }
}

__global__ void __launch_bounds__(kEeChunk) k_ee_scatter(const __grid_constant__ EeArgs a) {
  const unsigned i = blockIdx.x * kEeChunk + threadIdx.x;
  __shared__ unsigned s_cnt[kEeChunk / 32][kEeKeys];
  const int key = (i < a.n && a.key[i] != 255) ? (int)a.key[i] : -1;
  const unsigned rank = ee_rank_in_block(key, s_cnt, nullptr);
  if (key >= 0) a.order[a.ring_base[key] + a.chunk_cnt[(size_t)blockIdx.x * kEeKeys + key] + rank] = i;
}

struct EeSmem {
  double key[kEeMaxSection];
  int id[kEeMaxSection];
  unsigned bitmap[(kEeSectors * kEeMaxSection + 10 + 31) / 32 + 1];
  unsigned scan[kEeThreads / 32];
  int picks[20];
  int npick;
};
constexpr size_t kEeSmemBytes = sizeof(EeSmem);

// grid (kEeSectors, sensor_model)
__global__ void __launch_bounds__(kEeThreads) k_ee_section(const __grid_constant__ EeArgs a) {
  extern __shared__ __align__(16) unsigned char ee_raw[];
  EeSmem& sm = *reinterpret_cast<EeSmem*>(ee_raw);
  const int sct = blockIdx.x, ring = blockIdx.y;
  const int tid = threadIdx.x;
  const unsigned base = a.ring_base[ring];
  const int total = (int)(a.ring_base[ring + 1] - base);
  unsigned* cnt_out = a.sec_cnt + 2 * (ring * kEeSectors + sct);
  if (tid == 0) { cnt_out[0] = 0u; cnt_out[1] = 0u; }
  if (total < a.ring_min) return;                                           // :1240
  const int total_points = total - 10;                                       // :1246
  if (total_points <= 0) return;
  const int sector_length = total_points / kEeSectors;                        // :1288
  const int start = sector_length * sct;
  const int end = sct != kEeSectors - 1 ? sector_length * (sct + 1) - 1 : total_points - 1;
  const int cnt = end - start;                                                // exclusive end (:1292)
  if (cnt <= 0) return;
  if (cnt > kEeMaxSection) { if (tid == 0) atomicExch(a.status, 1); return; }
  int npad = 1;
  while (npad < cnt) npad <<= 1;
  auto P = [&](int j, int d) { return a.pts[3ull * a.order[base + (unsigned)j] + (unsigned)d]; };
  // ---- curvature of ring positions start + 5 ... start + cnt + 4 ----
  for (int c = tid; c < npad; c += kEeThreads) {
    if (c < cnt) {
      const int j = start + c + 5;
      double diff[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double s = ee_add(P(j - 5, d), P(j - 4, d));
        s = ee_add(s, P(j - 3, d)); s = ee_add(s, P(j - 2, d)); s = ee_add(s, P(j - 1, d));
        s = ee_sub(s, ee_mul(10.0, P(j, d)));
        s = ee_add(s, P(j + 1, d)); s = ee_add(s, P(j + 2, d)); s = ee_add(s, P(j + 3, d)); s = ee_add(s, P(j + 4, d));
        s = ee_add(s, P(j + 5, d));
        diff[d] = s;
      }
      sm.key[c] = ee_add(ee_add(ee_mul(diff[0], diff[0]), ee_mul(diff[1], diff[1])), ee_mul(diff[2], diff[2]));
      sm.id[c] = j;
    } else {
      sm.key[c] = __longlong_as_double(0x7FF0000000000000ll);                 // +inf padding sorts to the end
      sm.id[c] = 0x7FFFFFFF;
    }
  }
  for (int w = tid; w < (total + 31) / 32 + 1; w += kEeThreads) sm.bitmap[w] = 0u;
  __syncthreads();
  // ---- bitonic sort, ascending by (curvature, ring position) ----
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < npad; t += kEeThreads) {
        const int partner = t ^ j;
        if (partner > t) {
          const double ka = sm.key[t], kb = sm.key[partner];
          const int ia = sm.id[t], ib = sm.id[partner];
          const bool gt = ka > kb || (ka == kb && ia > ib);
          const bool asc = (t & k) == 0;
          if (asc ? gt : !gt) { sm.key[t] = kb; sm.key[partner] = ka; sm.id[t] = ib; sm.id[partner] = ia; }
        }
      }
      __syncthreads();
    }
  // ---- pick loop (serial, :1153-1199) ----
  if (tid == 0) {
    int largest = 0, np = 0;
    auto picked = [&](int id) { return (sm.bitmap[id >> 5] >> (id & 31)) & 1u; };
    auto mark = [&](int id) { sm.bitmap[id >> 5] |= 1u << (id & 31); };
    for (int i = cnt - 1; i >= 0; --i) {
      const int id = sm.id[i];
      if (picked(id)) continue;
      if (sm.key[i] <= 0.1) break;
      ++largest;
      mark(id);
      if (largest <= 20) sm.picks[np++] = id;
      else break;
      for (int k = 1; k <= 5; ++k) {
        const double dx = ee_sub(P(id + k, 0), P(id + k - 1, 0)), dy = ee_sub(P(id + k, 1), P(id + k - 1, 1)), dz = ee_sub(P(id + k, 2), P(id + k - 1, 2));
        if (ee_add(ee_add(ee_mul(dx, dx), ee_mul(dy, dy)), ee_mul(dz, dz)) > 0.05) break;
        mark(id + k);
      }
      for (int k = -1; k >= -5; --k) {
        const double dx = ee_sub(P(id + k, 0), P(id + k + 1, 0)), dy = ee_sub(P(id + k, 1), P(id + k + 1, 1)), dz = ee_sub(P(id + k, 2), P(id + k + 1, 2));
        if (ee_add(ee_add(ee_mul(dx, dx), ee_mul(dy, dy)), ee_mul(dz, dz)) > 0.05) break;
        mark(id + k);
      }
    }
    sm.npick = np;
  }
  __syncthreads();
  const unsigned sec_base = base + (unsigned)start;                           // this section's slice of the scratch lists
  for (int k = tid; k < sm.npick; k += kEeThreads) a.sec_edge[sec_base + (unsigned)k] = a.order[base + (unsigned)sm.picks[k]];
  // ---- unpicked points in ascending-curvature order: stable block compaction ----
  unsigned run = 0u;
  for (int c0 = 0; c0 < cnt; c0 += kEeThreads) {
    const int c = c0 + tid;
    const bool keep = c < cnt && !((sm.bitmap[sm.id[c < cnt ? c : 0] >> 5] >> (sm.id[c < cnt ? c : 0] & 31)) & 1u);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) sm.scan[warp] = __popc(b);
    __syncthreads();
    unsigned pre = run, tot = 0u;
    for (int w = 0; w < kEeThreads / 32; ++w) { if (w < warp) pre += sm.scan[w]; tot += sm.scan[w]; }
    if (keep) a.sec_non[sec_base + pre + __popc(b & ((1u << lane) - 1u))] = a.order[base + (unsigned)sm.id[c]];
    run += tot;
    __syncthreads();
  }
  if (tid == 0) { cnt_out[0] = (unsigned)sm.npick; cnt_out[1] = run; }
}

// one warp: exclusive scan of the section counts in (beam, sector) order (each lane owns a run of consecutive sections)
__global__ void __launch_bounds__(32) k_ee_offsets(const __grid_constant__ EeArgs a) {
  const int lane = threadIdx.x;
  const int nsec = a.sensor_model * kEeSectors;
  const int per = (nsec + 31) / 32;
  const int s0 = min(lane * per, nsec), s1 = min(s0 + per, nsec);
  unsigned e = 0u, nn = 0u;
  for (int s = s0; s < s1; ++s) { e += a.sec_cnt[2 * s]; nn += a.sec_cnt[2 * s + 1]; }
  unsigned ie = e, in = nn;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned ve = __shfl_up_sync(0xffffffffu, ie, o), vn = __shfl_up_sync(0xffffffffu, in, o);
    if (lane >= o) { ie += ve; in += vn; }
  }
  unsigned oe = ie - e, on = in - nn;
  for (int s = s0; s < s1; ++s) {
    a.sec_off[2 * s] = oe; a.sec_off[2 * s + 1] = on;
    oe += a.sec_cnt[2 * s]; on += a.sec_cnt[2 * s + 1];
  }
  if (lane == 31) { a.sec_off[2 * nsec] = ie; a.sec_off[2 * nsec + 1] = in; }
}

// grid (kEeSectors, sensor_model)
__global__ void __launch_bounds__(kEeThreads) k_ee_copy(const __grid_constant__ EeArgs a) {
  const int sct = blockIdx.x, ring = blockIdx.y, s = ring * kEeSectors + sct;
  const unsigned ne = a.sec_cnt[2 * s], nn = a.sec_cnt[2 * s + 1];
  if (ne == 0u && nn == 0u) return;
  const unsigned base = a.ring_base[ring];
  const int total_points = (int)(a.ring_base[ring + 1] - base) - 10;
  const unsigned sec_base = base + (unsigned)((total_points / kEeSectors) * sct);
  for (unsigned k = threadIdx.x; k < ne; k += kEeThreads) a.out_edge[a.sec_off[2 * s] + k] = a.sec_edge[sec_base + k];
  for (unsigned k = threadIdx.x; k < nn; k += kEeThreads) a.out_non[a.sec_off[2 * s + 1] + k] = a.sec_non[sec_base + k];
}

}  // namespace tloam
