// object_segment.cuh -- "next" row (f)-4, third part: object segmentation of the segmentation nodelet = Dynamic
// Curved-Voxel Clustering (ref: src/models/segmentation/segmentation.cpp:772-1112: getPolarIndex, convertToPolar,
// createHashTable, searchKNN, DCVC, labelAnalysis, colorSegmentation, objectSegmentation).
//
// The reference's DCVC is a sequential, order-dependent labelling: point i is visited only if it is still unlabelled;
// it gathers the points of the <= 27 listed voxels (searchKNN), and walks them in list order -- neighbours seen before
// the first labelled one stay unlabelled, everything after it receives / merges with the visitor's label.  What that
// loop computes can be stated per VOXEL (proved in DESIGN.md 4e, checked against the literal oracle by
// tests/test_object_segmentation.py::structural_dcvc):
//   * the labelled points of a voxel always form a prefix of its point list: 0 points, the first point only, or all;
//   * the only visits that do anything (EVENTS) are: the first point of a voxel whose state is 0, the second point of a
//     voxel whose state is "first only", and every point of a voxel that is not in its own list (pitch layer height + 1,
//     azimuth index > 300), which is in nobody's list;
//   * an event reads the states of its listed voxels, finds the first labelled entry p, sets every listed voxel at or
//     after p to "all" (everything listed when there is no p) and joins those voxels' classes with the visitor's.
// So the device runs: parallel polar conversion + extrema; the polar-bound table (one thread: it is a running FP64 sum);
// voxel hashing; per-voxel first / second point; the ordered EVENT list (stable compaction); one neighbour row per event
// (parallel hash lookups); ONE WARP that replays the events in order against 2-bit voxel states in shared memory (rows
// streamed with cp.async, one event = a ballot + a few shared-memory atomics); a parallel lock-free union-find over the
// recorded (event, entry >= p) pairs; class sizes, the size-ordered cluster table, a stable LSD partition of the points
// by cluster rank (= the segmented scan, index order inside a cluster) and the bounding boxes.
// Integer outputs are bit-exact against oracle/segmentation_oracle.cpp given the same polar triples; the triples use
// libdevice's asin / atan2 (triples within 4 ulp of libm's after the conversion to degrees, tests state the tolerance), every other FP64 operation is spelled
// with round-to-nearest intrinsics in the oracle's order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace tloam {

constexpr int kOsChunk = 256;
constexpr int kOsMaxBounds = 4096;        // polar rings (default configuration: ~500 at 120 m)
constexpr int kOsKeys = 256;              // digit of the stable partition
constexpr int kOsRow = 27;                // searchKNN entries
constexpr int kOsRowStride = 32;          // ints per event row: 27 neighbours, [27] last own position, [28] vid * 4 + kind, padding
constexpr int kOsRowOwn = 27, kOsRowInfo = 28;
constexpr int kOsSeqStages = 8;           // chunks of 32 event rows in flight towards shared memory (cp.async ring)
constexpr int kOsSeqSmemBytes = 200 * 1024;

struct OsArgs {
  const double* pts;                      // AoS xyz
  unsigned n, nchunk;
  double start_r, delta_r, delta_p, delta_a, min_range, max_range;
  double init[4];                         // minPitch, maxPitch, minPolar, maxPolar before the scan
  int min_seg;
  double* polar;                          // [n][3]
  unsigned long long* ext_enc;            // [4] order-preserving encodings (atomicMin / atomicMax)
  double* ext;                            // [4] decoded
  int* params;                            // [0] polar_num [1] width [2] height [3] status [4] nvox [5] nevents [6] nclusters [7] nseg
  double* bounds;                         // [kOsMaxBounds]
  int* key;                               // [n] the reference's voxelIndex
  int* coord;                             // [n][3] polar / pitch / azimuth index
  unsigned long long* table;              // [cap] (1 << 63) | (uint32)key, 0 = empty
  unsigned cap_mask;
  int* slot_vid;                          // [cap]
  int* pslot;                             // [n]
  int* vid;                               // [n]
  int* f1;                                // [n] by vid: smallest point index (the voxel's node in the union-find)
  int* f2;                                // [n] by vid: second smallest
  int* evkey;                             // [n] partition key of the event compaction: 0 = event, -1 = not
  int* ev_pt;                             // [n] events in point order (output of the compaction)
  int* ev_info;                           // [n] vid * 4 + kind (0 first point, 1 second point, 2 voxel not in its own list)
  int* rows;                              // [n32 + 32 * kOsSeqStages][32] per event: 27 neighbour vids (-1 = skipped / absent), own position, info
  signed char* ev_p;                      // [n] first labelled entry (27 = none: everything listed), -1 = the event did nothing
  unsigned* state_g;                      // 2-bit voxel states when they do not fit in shared memory
  int* parent;                            // [n]
  int* root;                              // [n]
  int* cnt;                               // [n] class size at its root
  int* cl_root;                           // [n] roots of the classes with > min_seg points (arbitrary order)
  int* cl_rank_of_root;                   // [n] by root: rank + 1, 0 = filtered
  int* cluster;                           // [n] per point: rank + 1 or 0
  int* sizes;                             // [n] by rank
  int* pkey;                              // [n] partition key: rank or -1
  double* boxes;                          // [n][6] by rank
  unsigned long long* box_enc;            // [n][6] by rank: encoded min xyz, max xyz
  // stable partition scratch
  unsigned* chunk_cnt;                    // [nchunk][kOsKeys]
  unsigned* key_base;                     // [kOsKeys + 1]
  int* perm_a;                            // [n]
  int* perm_b;                            // [n]
  unsigned long long* out_seg;            // [n]
};

__device__ __forceinline__ unsigned long long os_encode(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double os_decode(unsigned long long u) {
  return __longlong_as_double((long long)((u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u));
}

// ---- 0. reset ---------------------------------------------------------------------------------------------------
__global__ void k_os_init(const __grid_constant__ OsArgs a) {
  if (threadIdx.x < 4) a.ext_enc[threadIdx.x] = os_encode(a.init[threadIdx.x] + 0.0);
  if (threadIdx.x < 8) a.params[threadIdx.x] = 0;
}

// ---- 1. convertToPolar, first half (:790-822) -------------------------------------------------------------------
__global__ void __launch_bounds__(kOsChunk) k_os_polar(const __grid_constant__ OsArgs a) {
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  const double kInf = __longlong_as_double(0x7FF0000000000000ll);
  double mn_p = kInf, mx_p = -kInf, mn_r = kInf, mx_r = -kInf;
  if (i < a.n) {
    const double x = a.pts[3ull * i], y = a.pts[3ull * i + 1], z = a.pts[3ull * i + 2];
    const double r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double pitch = __ddiv_rn(__dmul_rn(asin(__ddiv_rn(z, r)), 180.0), M_PI);
    const double angle = atan2(y, x);
    const double az = angle > 0.0 ? __ddiv_rn(__dmul_rn(angle, 180.0), M_PI)
                                  : __ddiv_rn(__dmul_rn(__dadd_rn(angle, 2 * M_PI), 180.0), M_PI);
    double o0 = 0.0, o1 = 0.0, o2 = 0.0;
    if (!(r >= a.max_range || r <= a.min_range)) {                           // :808-809
      o0 = r; o1 = pitch; o2 = az;
      if (pitch == pitch) { mn_p = pitch + 0.0; mx_p = pitch + 0.0; }        // (+0.0: -0 and +0 compare equal in the reference)
      if (r == r) { mn_r = r; mx_r = r; }
    }
    a.polar[3ull * i] = o0; a.polar[3ull * i + 1] = o1; a.polar[3ull * i + 2] = o2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn_p = fmin(mn_p, __shfl_xor_sync(0xffffffffu, mn_p, o)); mx_p = fmax(mx_p, __shfl_xor_sync(0xffffffffu, mx_p, o));
    mn_r = fmin(mn_r, __shfl_xor_sync(0xffffffffu, mn_r, o)); mx_r = fmax(mx_r, __shfl_xor_sync(0xffffffffu, mx_r, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mn_p != kInf) { atomicMin(a.ext_enc + 0, os_encode(mn_p)); atomicMax(a.ext_enc + 1, os_encode(mx_p)); }
    if (mn_r != kInf) { atomicMin(a.ext_enc + 2, os_encode(mn_r)); atomicMax(a.ext_enc + 3, os_encode(mx_r)); }
  }
}

// ---- 2. convertToPolar, second half (:825-836): one thread, the table is a running sum -----------------------------
__global__ void k_os_bounds(const __grid_constant__ OsArgs a) {
  if (threadIdx.x != 0) return;
  double e[4];
  for (int k = 0; k < 4; ++k) { e[k] = os_decode(a.ext_enc[k]); a.ext[k] = e[k]; }
  const int width = (int)__dadd_rn(round(__ddiv_rn(360.0, a.delta_a)), 1.0);
  const int height = (int)__ddiv_rn(__dsub_rn(e[1], e[0]), a.delta_p);
  int polar_num = 0, step = 1, status = 0;
  double range = e[2];
  while (range <= e[3]) {
    range = __dadd_rn(range, __dsub_rn(a.start_r, __dmul_rn((double)step, a.delta_r)));
    if (polar_num >= kOsMaxBounds) { status = 1; break; }
    a.bounds[polar_num] = range;
    polar_num++, step++;
  }
  a.params[0] = polar_num; a.params[1] = width; a.params[2] = height; a.params[3] = status;
}

__device__ __forceinline__ int os_voxel_key(int ax, int y, int z, int polar_num, int width) {
  return (ax * (polar_num + 1) + y) + z * (polar_num + 1) * (width + 1);
}
__device__ __forceinline__ unsigned os_hash(int key) {
  unsigned h = (unsigned)key * 0x9E3779B1u;
  return h ^ (h >> 15);
}

// ---- 3. createHashTable (:843-874): voxel index of every point, hash insert ----------------------------------------
__global__ void __launch_bounds__(kOsChunk) k_os_key(const __grid_constant__ OsArgs a) {
  __shared__ double s_b[kOsMaxBounds];
  const int polar_num = a.params[0], width = a.params[1];
  for (int k = threadIdx.x; k < polar_num; k += kOsChunk) s_b[k] = a.bounds[k];
  __syncthreads();
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  if (i >= a.n || a.params[3] != 0) return;
  const double r = a.polar[3ull * i], pitch = a.polar[3ull * i + 1], az = a.polar[3ull * i + 2];
  // getPolarIndex (:777-784): first ring with radius < bound; the table is increasing whenever its loop terminated
  int lo = 0, hi = polar_num;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (r < s_b[mid]) hi = mid; else lo = mid + 1;
  }
  const int pi = lo < polar_num ? lo : polar_num - 1;
  const int ti = (int)round(__ddiv_rn(__dsub_rn(pitch, a.ext[0]), a.delta_p));
  const int ai = (int)round(__ddiv_rn(az, a.delta_a));
  const int key = os_voxel_key(ai, pi, ti, polar_num, width);
  a.key[i] = key;
  a.coord[3ull * i] = pi; a.coord[3ull * i + 1] = ti; a.coord[3ull * i + 2] = ai;
  a.parent[i] = (int)i;
  a.cnt[i] = 0;
  a.f1[i] = 0x7FFFFFFF; a.f2[i] = 0x7FFFFFFF;
  const unsigned long long want = 0x8000000000000000ull | (unsigned)key;
  unsigned s = os_hash(key) & a.cap_mask;
  while (true) {
    const unsigned long long cur = a.table[s];
    if (cur == want) break;
    if (cur == 0ull) {
      const unsigned long long old = atomicCAS(a.table + s, 0ull, want);
      if (old == 0ull || old == want) break;
    }
    s = (s + 1) & a.cap_mask;
  }
  a.pslot[i] = (int)s;
}

__device__ __forceinline__ int os_lookup(const OsArgs& a, int key) {
  const unsigned long long want = 0x8000000000000000ull | (unsigned)key;
  unsigned s = os_hash(key) & a.cap_mask;
  while (true) {
    const unsigned long long cur = a.table[s];
    if (cur == want) return a.slot_vid[s];
    if (cur == 0ull) return -1;
    s = (s + 1) & a.cap_mask;
  }
}

// ---- 4. dense voxel ids --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_os_vid(const __grid_constant__ OsArgs a) {
  const unsigned s = blockIdx.x * 256 + threadIdx.x;
  if (s > a.cap_mask) return;
  if (a.table[s] != 0ull) a.slot_vid[s] = atomicAdd(a.params + 4, 1);
}

// ---- 5. first point of every voxel ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOsChunk) k_os_first(const __grid_constant__ OsArgs a) {
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  if (i >= a.n || a.params[3] != 0) return;
  const int v = a.slot_vid[a.pslot[i]];
  a.vid[i] = v;
  atomicMin(a.f1 + v, (int)i);
}

// ---- 6. second point -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOsChunk) k_os_second(const __grid_constant__ OsArgs a) {
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  if (i >= a.n || a.params[3] != 0) return;
  const int v = a.vid[i];
  if ((int)i != a.f1[v]) atomicMin(a.f2 + v, (int)i);
}

// a voxel is in its own searchKNN list iff its pitch layer is listed and its azimuth index survives the clamp
__device__ __forceinline__ bool os_own_listed(int ti, int ai, int height) { return ti >= 0 && ti <= height && ai >= 0 && ai <= 300; }

// ---- 7. which points are events ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOsChunk) k_os_evkey(const __grid_constant__ OsArgs a) {
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  if (i >= a.n) return;
  int k = -1;
  if (a.params[3] == 0) {
    const int v = a.vid[i];
    if (!os_own_listed(a.coord[3ull * i + 1], a.coord[3ull * i + 2], a.params[2]) || (int)i == a.f1[v] || (int)i == a.f2[v]) k = 0;
  }
  a.evkey[i] = k;
}

// ---- stable partition by an 8-bit digit (3 kernels per pass) ----------------------------------------------------------
// items: perm_in[j] (j < *n_items, or j < n_fixed when n_items == nullptr; perm_in == nullptr: identity);
// digit = (keys[item] >> shift) & 255, items with keys[item] < 0 are dropped.
struct OsPart {
  const int* keys; const int* perm_in; int* perm_out; const int* n_items; unsigned n_fixed; int shift;
  unsigned* chunk_cnt; unsigned* key_base; unsigned nchunk; int* total_out;
};

__device__ __forceinline__ unsigned os_rank_in_block(int key, unsigned (*s_cnt)[kOsKeys], unsigned* hist_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = lane; k < kOsKeys; k += 32) s_cnt[warp][k] = 0u;
  __syncwarp();
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const unsigned rank_w = __popc(peers & ((1u << lane) - 1u));
  if (key >= 0 && rank_w == 0u) s_cnt[warp][key] = __popc(peers);
  __syncthreads();
  unsigned before = 0u;
  if (key >= 0)
    for (int w = 0; w < warp; ++w) before += s_cnt[w][key];
  if (hist_out)
    for (int k = threadIdx.x; k < kOsKeys; k += blockDim.x) {
      unsigned t = 0u;
      for (int w = 0; w < kOsChunk / 32; ++w) t += s_cnt[w][k];
      hist_out[k] = t;
    }
  return before + rank_w;
}

__device__ __forceinline__ int os_part_digit(const OsPart& p, unsigned j, int* item_out) {
  const unsigned cnt = p.n_items ? (unsigned)*p.n_items : p.n_fixed;
  if (j >= cnt) return -1;
  const int item = p.perm_in ? p.perm_in[j] : (int)j;
  *item_out = item;
  const int k = p.keys[item];
  return k < 0 ? -1 : ((k >> p.shift) & (kOsKeys - 1));
}

__global__ void __launch_bounds__(kOsChunk) k_os_part_hist(const __grid_constant__ OsPart p) {
  __shared__ unsigned s_cnt[kOsChunk / 32][kOsKeys];
  int item = 0;
  const int d = os_part_digit(p, blockIdx.x * kOsChunk + threadIdx.x, &item);
  os_rank_in_block(d, s_cnt, p.chunk_cnt + (size_t)blockIdx.x * kOsKeys);
}

__global__ void __launch_bounds__(kOsKeys) k_os_part_scan(const __grid_constant__ OsPart p) {
  __shared__ unsigned s_tot[kOsKeys];
  const int k = threadIdx.x;
  unsigned run = 0u;
  for (unsigned c = 0; c < p.nchunk; ++c) {
    const unsigned v = p.chunk_cnt[(size_t)c * kOsKeys + k];
    p.chunk_cnt[(size_t)c * kOsKeys + k] = run;
    run += v;
  }
  s_tot[k] = run;
  __syncthreads();
  if (k == 0) {
    unsigned off = 0u;
    for (int b = 0; b < kOsKeys; ++b) { p.key_base[b] = off; off += s_tot[b]; }
    p.key_base[kOsKeys] = off;
    if (p.total_out) *p.total_out = (int)off;
  }
}

__global__ void __launch_bounds__(kOsChunk) k_os_part_scatter(const __grid_constant__ OsPart p) {
  __shared__ unsigned s_cnt[kOsChunk / 32][kOsKeys];
  int item = 0;
  const int d = os_part_digit(p, blockIdx.x * kOsChunk + threadIdx.x, &item);
  const unsigned rank = os_rank_in_block(d, s_cnt, nullptr);
  if (d >= 0) p.perm_out[p.key_base[d] + p.chunk_cnt[(size_t)blockIdx.x * kOsKeys + d] + rank] = item;
}

// ---- 8. searchKNN (:886-908): one neighbour row per event ----------------------------------------------------------
__global__ void __launch_bounds__(256) k_os_rows(const __grid_constant__ OsArgs a) {
  const unsigned t = blockIdx.x * 256 + threadIdx.x;
  const unsigned j = t / 32u, l = t % 32u;
  if (j >= (unsigned)a.params[5]) return;
  const int pt = a.ev_pt[j];
  const int v = a.vid[pt];
  const int pi = a.coord[3ull * pt], ti = a.coord[3ull * pt + 1], ai = a.coord[3ull * pt + 2];
  const int polar_num = a.params[0], width = a.params[1], height = a.params[2];
  int nb = -1;
  if (l < (unsigned)kOsRow) {
    const int z = ti - 1 + (int)(l / 9), y = pi - 1 + (int)((l / 3) % 3), x = ai - 1 + (int)(l % 3);
    if (!(z < 0 || z > height) && !(y < 0 || y > polar_num)) {
      int ax = x;
      if (ax < 0) ax = width - 1;
      if (ax > 300) ax = 300;
      nb = os_lookup(a, os_voxel_key(ax, y, z, polar_num, width));
    }
  }
  const unsigned own = __ballot_sync(0xffffffffu, nb == v);                 // the 32 threads of an event are one warp
  const int kind = !os_own_listed(ti, ai, height) ? 2 : (pt == a.f1[v] ? 0 : 1);
  int word = nb;                                                            // entries 0..26: neighbour voxel ids
  if (l == kOsRowOwn) word = own ? 31 - __clz(own) : -1;                    // LAST position of the own voxel in its row
  if (l == kOsRowInfo) word = v * 4 + kind;
  a.rows[(size_t)j * kOsRowStride + l] = word;
  if (l == 31) { a.ev_info[j] = v * 4 + kind; a.ev_p[j] = -1; }
}

// ---- 9. the sequential part: one warp replays the events in point order ----------------------------------------------
// state of a voxel: 0 = no point labelled, 1 = the first point is labelled, 2 (or 3) = all points are labelled.
// kBytes: one byte per voxel in shared memory (plain stores: lanes that hit the same voxel write the same value);
// otherwise 2 bits per voxel (shared memory up to ~800k voxels, else global memory), updated with atomicOr.
template <bool kBytes>
__device__ __forceinline__ unsigned os_state(const void* st, int v) {
  if (kBytes) return reinterpret_cast<const volatile unsigned char*>(st)[v];
  return (reinterpret_cast<const volatile unsigned*>(st)[v >> 4] >> ((v & 15) * 2)) & 3u;
}

template <bool kBytes>
__global__ void __launch_bounds__(32) k_os_seq(const __grid_constant__ OsArgs a, const int use_smem) {
  extern __shared__ __align__(16) unsigned char os_raw[];
  constexpr int kChunkInts = 32 * kOsRowStride;                              // one chunk of 32 events: 4 KB
  int* s_rows = reinterpret_cast<int*>(os_raw);                              // [kOsSeqStages][kChunkInts]
  void* st = (kBytes || use_smem) ? static_cast<void*>(os_raw + kOsSeqStages * kChunkInts * sizeof(int)) : static_cast<void*>(a.state_g);
  const int lane = threadIdx.x;
  const int nev = a.params[5], nvox = a.params[4];
  if (a.params[3] != 0) return;
  const int nchunks = (nev + 31) / 32;
  // rows of a chunk are contiguous (and the buffer is padded to whole chunks): 256 x 16 B per chunk, 8 per lane.  A
  // group is committed for every stage, empty past the end, so that wait_group counts stay uniform.
  auto prefetch = [&](int c) {
    if (c < nchunks) {
      const int* src = a.rows + (size_t)c * kChunkInts;
      int* dst_base = s_rows + (c % kOsSeqStages) * kChunkInts;
#pragma unroll
      for (int q = 0; q < kChunkInts / 4 / 32; ++q) {
        const int e = (q * 32 + lane) * 4;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(dst_base + e);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + e) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  for (int c = 0; c < kOsSeqStages - 1; ++c) prefetch(c);
  if (kBytes) {
    for (int w = lane; w < (nvox + 3) / 4 + 1; w += 32) reinterpret_cast<unsigned*>(st)[w] = 0u;
  } else if (use_smem) {
    for (int w = lane; w < (nvox + 15) / 16 + 1; w += 32) reinterpret_cast<unsigned*>(st)[w] = 0u;
  }
  __syncwarp();
  for (int c = 0; c < nchunks; ++c) {
    prefetch(c + kOsSeqStages - 1);                                           // its slot was consumed in iteration c - 1
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kOsSeqStages - 1) : "memory");
    __syncwarp();
    const int* rows = s_rows + (c % kOsSeqStages) * kChunkInts;
    const int j = c * 32 + lane;
    const int info = j < nev ? rows[lane * kOsRowStride + kOsRowInfo] : -1;
    const int my_own = rows[lane * kOsRowStride + kOsRowOwn];
    const int my_v = info >> 2, my_kind = info & 3;
    bool cand = info >= 0;
    int my_p = -1;
    while (true) {
      // which events of the chunk would act now?  (one shared-memory read of the own voxel's state per lane)
      bool fire = false;
      if (cand) {
        if (my_kind == 2) fire = true;
        else {
          const unsigned s = os_state<kBytes>(st, my_v);
          fire = my_kind == 0 ? (s == 0u) : (s == 1u);
          if (s & 2u) cand = false;                                           // a full voxel never fires again
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, fire);
      if (m == 0u) break;
      const int src = __ffs(m) - 1;                                           // the earliest one is the next in point order
      const int nb = lane < kOsRow ? rows[src * kOsRowStride + lane] : -1;
      const unsigned s_nb = nb >= 0 ? os_state<kBytes>(st, nb) : 0u;
      const unsigned lab = __ballot_sync(0xffffffffu, nb >= 0 && s_nb != 0u);   // (all state reads are done once this returns)
      const int p = lab ? __ffs(lab) - 1 : 0;
      const bool take = nb >= 0 && lane >= p && !(s_nb & 2u);
      const bool own_first = lane == src && my_kind != 2 && my_own < p;        // own voxel listed before p only: first point labelled
      __syncwarp();                                                           // every state read above precedes the writes below
      if (kBytes) {
        if (take) reinterpret_cast<volatile unsigned char*>(st)[nb] = 2;
        if (own_first) reinterpret_cast<volatile unsigned char*>(st)[my_v] = 1;
      } else {
        if (take) atomicOr(reinterpret_cast<unsigned*>(st) + (nb >> 4), 2u << ((nb & 15) * 2));
        if (own_first) atomicOr(reinterpret_cast<unsigned*>(st) + (my_v >> 4), 1u << ((my_v & 15) * 2));
      }
      if (lane == src) { my_p = p; cand = false; }
      __syncwarp();
    }
    if (j < nev && my_p >= 0) a.ev_p[j] = (signed char)my_p;
    __syncwarp();                                                             // the slot may be refilled from the next iteration on
  }
}

// ---- 10. lock-free union-find over the recorded (event, entry >= p) pairs ------------------------------------------
__device__ __forceinline__ int os_find(int* parent, int x) {
  while (true) {
    const int p = parent[x];
    if (p == x) return x;
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;                                              // path halving (parents only ever decrease)
    x = p;
  }
}
__device__ __forceinline__ void os_unite(int* parent, int x, int y) {
  while (true) {
    x = os_find(parent, x); y = os_find(parent, y);
    if (x == y) return;
    const int hi = x > y ? x : y, lo = x > y ? y : x;
    if (atomicCAS(parent + hi, hi, lo) == hi) return;
  }
}

__global__ void __launch_bounds__(256) k_os_union(const __grid_constant__ OsArgs a) {
  const unsigned t = blockIdx.x * 256 + threadIdx.x;
  const unsigned j = t / 32u, l = t % 32u;
  if (j >= (unsigned)a.params[5] || l >= (unsigned)kOsRow) return;
  const int p = a.ev_p[j];
  if (p < 0 || (int)l < p) return;
  const int nb = a.rows[(size_t)j * kOsRowStride + l];
  if (nb < 0) return;
  const int info = a.ev_info[j];
  const int seed = (info & 3) == 2 ? a.ev_pt[j] : a.f1[info >> 2];
  os_unite(a.parent, seed, a.f1[nb]);
}

// ---- 11. classes: root (= smallest point index of the class) and size ----------------------------------------------
__global__ void __launch_bounds__(kOsChunk) k_os_label(const __grid_constant__ OsArgs a) {
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  if (i >= a.n || a.params[3] != 0) return;
  const int v = a.vid[i];
  const int node = os_own_listed(a.coord[3ull * i + 1], a.coord[3ull * i + 2], a.params[2]) ? a.f1[v] : (int)i;
  const int r = os_find(a.parent, node);
  a.root[i] = r;
  atomicAdd(a.cnt + r, 1);
  a.cl_rank_of_root[i] = 0;
}

// ---- 12. labelAnalysis (:998-1025): classes with > minSeg points, by (size descending, smallest index ascending) -------
__global__ void __launch_bounds__(kOsChunk) k_os_clusters(const __grid_constant__ OsArgs a) {
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  if (i >= a.n || a.params[3] != 0) return;
  if (a.root[i] == (int)i && a.cnt[i] > a.min_seg) a.cl_root[atomicAdd(a.params + 6, 1)] = (int)i;
}

__global__ void __launch_bounds__(256) k_os_rank(const __grid_constant__ OsArgs a) {
  __shared__ int s_cnt[256], s_root[256];
  const int nc = a.params[6];
  const int c = blockIdx.x * 256 + threadIdx.x;
  if ((int)(blockIdx.x * 256) >= nc) return;
  const int my_root = c < nc ? a.cl_root[c] : 0;
  const int my_cnt = c < nc ? a.cnt[my_root] : 0;
  int rank = 0;
  for (int base = 0; base < nc; base += 256) {
    const int o = base + threadIdx.x;
    if (o < nc) { s_root[threadIdx.x] = a.cl_root[o]; s_cnt[threadIdx.x] = a.cnt[s_root[threadIdx.x]]; }
    __syncthreads();
    const int lim = min(256, nc - base);
    for (int k = 0; k < lim; ++k) rank += (s_cnt[k] > my_cnt || (s_cnt[k] == my_cnt && s_root[k] < my_root)) ? 1 : 0;
    __syncthreads();
  }
  if (c < nc) {
    a.cl_rank_of_root[my_root] = rank + 1; a.sizes[rank] = my_cnt;
    for (int d = 0; d < 3; ++d) { a.box_enc[6ull * rank + d] = ~0ull; a.box_enc[6ull * rank + 3 + d] = 0ull; }
  }
}

__global__ void __launch_bounds__(kOsChunk) k_os_pkey(const __grid_constant__ OsArgs a) {
  const unsigned i = blockIdx.x * kOsChunk + threadIdx.x;
  if (i >= a.n) return;
  const int cl = a.params[3] == 0 ? a.cl_rank_of_root[a.root[i]] : 0;
  a.cluster[i] = cl;
  a.pkey[i] = cl - 1;
}

// ---- 13. colorSegmentation (:1032-1078): the segmented scan as 64-bit indices + one box per cluster --------------------
// One thread per position of the segmented scan; min / max through order-preserving 64-bit encodings (warp-aggregated when
// the whole warp sits in one cluster, which is the common case: the scan is grouped by cluster).
__global__ void __launch_bounds__(256) k_os_boxes_acc(const __grid_constant__ OsArgs a, const int* __restrict__ seg) {
  const unsigned j = blockIdx.x * 256 + threadIdx.x;
  const bool on = j < (unsigned)a.params[7];
  int r = -1;
  double v[3] = {0.0, 0.0, 0.0};
  if (on) {
    const int id = seg[j];
    a.out_seg[j] = (unsigned long long)id;
    r = a.cluster[id] - 1;
    v[0] = a.pts[3ull * id]; v[1] = a.pts[3ull * id + 1]; v[2] = a.pts[3ull * id + 2];
  }
  const unsigned active = __ballot_sync(0xffffffffu, on);
  if (active == 0u) return;
  const int r0 = __shfl_sync(0xffffffffu, r, __ffs(active) - 1);
  const bool uniform = __all_sync(0xffffffffu, !on || r == r0);
  if (uniform) {
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double lo = on ? v[d] : kInf, hi = on ? v[d] : -kInf;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
      if ((threadIdx.x & 31) == 0) { atomicMin(a.box_enc + 6ull * r0 + d, os_encode(lo)); atomicMax(a.box_enc + 6ull * r0 + 3 + d, os_encode(hi)); }
    }
  } else if (on) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { atomicMin(a.box_enc + 6ull * r + d, os_encode(v[d])); atomicMax(a.box_enc + 6ull * r + 3 + d, os_encode(v[d])); }
  }
}

__global__ void __launch_bounds__(256) k_os_boxes_fin(const __grid_constant__ OsArgs a) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= a.params[6]) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double lo = os_decode(a.box_enc[6ull * r + d]), hi = os_decode(a.box_enc[6ull * r + 3 + d]);
    const double len = __dsub_rn(hi, lo);
    a.boxes[6ull * r + d] = __dadd_rn(lo, __ddiv_rn(len, 2.0));
    a.boxes[6ull * r + 3 + d] = len < 0 ? __dmul_rn(-1.0, len) : len;
  }
}

}  // namespace tloam
