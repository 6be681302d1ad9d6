// tloam_b200.cu -- kernels + C ABI of libtloam_b200.so (sm_100a). See include/tloam_b200.h for the boundary
// and registration.cuh for the execution model.  No CPU fallback exists anywhere in this file.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "registration.cuh"
#include "solver.cuh"
#include "submap.cuh"
#include "feature_extract.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace tloam {

// =================================================================================================
// Map build kernels (brick-keyed voxel hash, once per set_target)
// =================================================================================================
struct MapBuildArgs {
  const double* src[4];         // AoS xyz of each cloud (the staging buffer for host input, the caller's arrays for device input)
  unsigned stage_off[5];        // global point index of each cloud's first point
  unsigned char* blob;          // MapHeader + pts + tables
  unsigned* slot_of;            // scratch [total]
  unsigned* rank_of;            // scratch [total]
  unsigned i_beg, i_end;        // global point range this launch works on (one cloud in the pipelined host path)
  int only_cloud;               // k_map_offsets: table of this cloud only (-1 = all four)
};

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
    h->origin[threadIdx.x] = rint(0.5 * (lo + hi));
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
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int wi = 0; wi < 8; ++wi) { const unsigned v = s_w[wi]; s_w[wi] = tot; tot += v; }
    s_base = (tot > 0u) ? atomicAdd(&h->cursor[c], tot) : 0u;
  }
  __syncthreads();
  if (cnt > 0u) table[2u * s].z = s_base + s_w[warp] + (incl - cnt);
}

__global__ void k_map_scatter(MapBuildArgs a) {
  const unsigned i = a.i_beg + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.i_end) return;
  MapHeader* h = reinterpret_cast<MapHeader*>(a.blob);
  if (h->build_flags & 1ull) return;               // counters wrapped: destinations are meaningless
  const int c = cloud_of_point(a, i);
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

// =================================================================================================
// Frame kernels
// =================================================================================================
struct Predict { double m[16]; double from_state; };   // from_state != 0: constant-velocity prediction on the device

// Programmatic dependent launch (opt-in, TLOAM_B200_PDL=1): a frame kernel waits for its predecessor before it
// touches anything the predecessor may have written; the block that runs the serial tail of a k_eval (partial sum
// + solver, ~10 us on one SM) releases the successor first, so that the successor's blocks are already resident
// on the idle SMs when the tail ends.  (Releasing at kernel entry was measured slower: the successor's waiting
// blocks take residency away from the running grid.)
__device__ __forceinline__ void pdl_prologue() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// thread-block cluster barrier, split in its two halves (all threads of the block execute both, convergently)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
// store a double into the same shared-memory variable of block `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ void dsmem_store(double* local_smem, unsigned rank, double v) {
  const unsigned laddr = (unsigned)__cvta_generic_to_shared(local_smem);
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}

// scanMatching prologue, ref: registration.cpp:879-886, 961-964, 1027-1033.
// 4x4 column-major helpers for the constant-velocity prediction (ref: src/front_end/front_end.cpp:329-330)
__device__ __forceinline__ void mat4_mul(const double* A, const double* B, double* C) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k];
      C[c * 4 + r] = s;
    }
}
__device__ __forceinline__ void isometry_inverse(const double* T, double* out) {   // Eigen::Isometry3d::inverse(): R^T, -R^T t
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out[c * 4 + r] = T[r * 4 + c];
  for (int r = 0; r < 3; ++r) out[12 + r] = -(out[r] * T[12] + out[4 + r] * T[13] + out[8 + r] * T[14]);
  out[3] = out[7] = out[11] = 0.0; out[15] = 1.0;
}

__global__ void k_begin_frame(DeviceCtx ctx, const Predict* prp) {
  Predict pr = *prp;
  if (pr.from_state != 0.0) {
    // step = last^-1 * curr ; predict = curr * step, from the two last results kept in the frame state
    // (every thread computes the same 16 values: the state is rewritten further down by thread 0 only)
    double inv[16], step[16];
    isometry_inverse(ctx.st->last_pose, inv);
    mat4_mul(inv, ctx.st->curr_pose, step);
    mat4_mul(ctx.st->curr_pose, step, pr.m);
  }
  // zero the trace
  {
    unsigned* w = reinterpret_cast<unsigned*>(ctx.stats);
    if (w) for (unsigned i = threadIdx.x; i < sizeof(tloam_b200_stats) / 4; i += blockDim.x) w[i] = 0u;
    for (int i = threadIdx.x; i < ctx.blk_off[4]; i += blockDim.x) ctx.blk_count[i] = 0;   // buffer 0 (outer 0)
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  FrameState* st = ctx.st;
  st->status = TLOAM_B200_OK;
  st->frame_done = 0;
  *ctx.counter = 0u;
  for (int i = 0; i < 16; ++i) st->last_pose[i] = st->curr_pose[i];            // :882
  Pose7 p;
  if (*ctx.map_flags & 1ull) {                   // a map cell overflowed its u16 counter at build time
    st->status = TLOAM_B200_ERR_MAP_DENSITY;
    for (int i = 0; i < 16; ++i) st->result[i] = pr.m[i];
    st->frame_done = 1;
    return;
  }
  if (!pose_from_matrix(pr.m, p)) {
    st->status = TLOAM_B200_ERR_BAD_POSE;
    for (int i = 0; i < 16; ++i) st->result[i] = pr.m[i];
    st->frame_done = 1;
    return;
  }
  se3_log(p, st->x);                                                            // :881
  const double wn = sqrt(st->x[3] * st->x[3] + st->x[4] * st->x[4] + st->x[5] * st->x[5]);
  if (wn < 1e-2) {                                                              // :884-886 (explicit direction)
    const double* d = ctx.reinit_dir;
    const double nn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int j = 0; j < 3; ++j) st->x[3 + j] = d[j] / nn * 1e-4;
  }
  if (ctx.stats) for (int i = 0; i < 6; ++i) ctx.stats->x_init[i] = st->x[i];
  st->xq = se3_exp(st->x);
  st->evalq = st->xq;
  st->phase = kPhaseIter0;
  st->outer = 0;
  st->planar_prev = __longlong_as_double(0x7FF0000000000000ll);                 // +inf, :956
  double c2 = ctx.noise_bound * ctx.noise_bound;                                // :962-964
  if (c2 < 1e-16) c2 = 1e-2;
  st->c2 = c2;
  // :1027-1033 -- mu is derived from the residual slots BEFORE the first solve, when they are all zero
  const double max_residual = 0.0;
  double mu = 1.0 / (2.0 * max_residual / c2 - 1.0);
  if (mu <= 0.0) mu = 1e-10;
  st->mu = mu; st->mu_used = mu; st->th1 = 0.0; st->th2 = 0.0;
  for (int k = 0; k < 4; ++k) st->slot_sum[k] = 0.0;
}

// Primitive fit + validity tests of one feature given its (already merged) neighbour list.
// ref: registration.cpp:445-493 (edge), 536-551 (sphere), 589-625 (planar), 732-768 (ground).
template <int K>
__device__ __forceinline__ unsigned char fit_one(const DeviceCtx& ctx, int c, const TopK<K>& t, double prim[6]) {
  const GridDesc& g = ctx.grid[c];
  const double o0 = ctx.origin[0], o1 = ctx.origin[1], o2 = ctx.origin[2];
#pragma unroll
  for (int j = 0; j < 6; ++j) prim[j] = 0.0;
  if (K == 1) {                                            // sphere
    if (t.pos[0] < 0) return kFlagCounted;                 // not found: sphere_sum++ (:551)
    if (t.d2[0] > 0.2) return 0;                           // squared distance vs 0.2, `continue` (:536)
    const float4 m = __ldg(&g.pts[t.pos[0]]);
    prim[0] = o0 + (double)m.x; prim[1] = o1 + (double)m.y; prim[2] = o2 + (double)m.z;
    return kFlagCand | kFlagCounted;
  }
  const int k = t.count();
  if (k <= 0) return 0;
  double nb[K][3];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < k) {
      const float4 m = __ldg(&g.pts[t.pos[j]]);
      nb[j][0] = o0 + (double)m.x; nb[j][1] = o1 + (double)m.y; nb[j][2] = o2 + (double)m.z;
    } else {
      nb[j][0] = nb[j][1] = nb[j][2] = 0.0;
    }
  }
  if (c == kEdge) {
    if (k <= 3) return 0;                                  // :445
    // mean + covariance from raw cumulants (:451-474)
    double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < k; ++j) {
      const double x = nb[j][0], y = nb[j][1], z = nb[j][2];
      cu[0] += x; cu[1] += y; cu[2] += z;
      cu[3] += x * x; cu[4] += x * y; cu[5] += x * z; cu[6] += y * y; cu[7] += y * z; cu[8] += z * z;
    }
    const double kn = (double)k;
#pragma unroll
    for (int j = 0; j < 9; ++j) cu[j] /= kn;
    double ev[3], dir[3];
    sym_eig3_max(cu[3] - cu[0] * cu[0], cu[4] - cu[0] * cu[1], cu[5] - cu[0] * cu[2], cu[6] - cu[1] * cu[1],
                 cu[7] - cu[1] * cu[2], cu[8] - cu[2] * cu[2], ev, dir);
    if (!(ev[2] > 3.0 * ev[1] && fabs(dir[2]) > ctx.edge_dir_thres)) return 0;   // :481
    prim[0] = 0.1 * dir[0] + cu[0]; prim[1] = 0.1 * dir[1] + cu[1]; prim[2] = 0.1 * dir[2] + cu[2];
    prim[3] = -0.1 * dir[0] + cu[0]; prim[4] = -0.1 * dir[1] + cu[1]; prim[5] = -0.1 * dir[2] + cu[2];
    return kFlagCand | kFlagCounted;                        // edge_num++ (:492)
  }
  if (k <= 4) return 0;                                     // :589 / :732
  double nd[4];
  fit_best_plane(nb, K, nd);                                // :600 / :743
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (nd[0] * nb[j][0] + nd[1] * nb[j][1] + nd[2] * nb[j][2] + nd[3] > 0.2) return 0;   // one-sided, :605-613
  prim[0] = nd[0]; prim[1] = nd[1]; prim[2] = nd[2]; prim[3] = nd[3];
  return kFlagCand | kFlagCounted;                          // surf_num++ / ground_num++
}

// Correspondence search + primitive fit.  A LANE PAIR serves one feature (knn_search_pair: each lane probes one
// z-layer of bricks, both lanes stream the candidate runs interleaved, the even lane merges and fits), so a
// 128-thread block serves 64 features and a 128-feature block of one cloud is served by two thread blocks.  Also
// applies the lazy GNC weight update of the previous outer iteration (ref: registration.cpp:858-876) and resets the
// residual slot (:1118-1121).  Measured alternatives: DESIGN.md section 4.
// 5 blocks per SM (<= 102 registers): measured best; 6 (80 regs) and 8 (64 regs) spill and are 8% / 55% slower,
// 4 (114 regs, what ptxas picks when unconstrained) is 28% slower
__global__ void __launch_bounds__(kBlk, 5) k_correspond(const __grid_constant__ DeviceCtx ctx) {
  pdl_prologue();
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  __shared__ unsigned s_beg[kPairCells][kBlk];
  __shared__ unsigned s_cnt[kPairCells][kBlk];
  __shared__ float s_md[kPairCells][kBlk];
  constexpr int kQ = kBlk / 2;                 // features per thread block
  const int fb = blockIdx.x / 2, sb = blockIdx.x % 2;
  const int c = cloud_of_block(ctx, fb);
  const int il = (fb - ctx.blk_off[c]) * kBlk + sb * kQ + (int)(threadIdx.x >> 1);
  const int gi = ctx.pad_off[c] + il;
  const bool live = (il < ctx.n[c]) && cloud_enabled(ctx, c);
  const bool even = (threadIdx.x & 1) == 0;
  const int buf = st->outer & 1;
  if (threadIdx.x == 0 && sb == 0) ctx.blk_count[(buf ^ 1) * ctx.blk_cap + fb] = 0;   // for the next outer iteration
  long long tk0 = 0, tk1 = 0;
  if (ctx.dbg) tk0 = clock64();
  double rx = 0.0, ry = 0.0, rz = 0.0;
  if (live) {
    const Rt T = pose_to_rt(st->xq);
    double qx, qy, qz;
    rt_apply(T, ctx.px[gi], ctx.py[gi], ctx.pz[gi], qx, qy, qz);
    rx = qx - ctx.origin[0]; ry = qy - ctx.origin[1]; rz = qz - ctx.origin[2];
  }
  unsigned char flag = 0;
  double prim[6];
  if (c == kSphere) {
    TopK<1> t;
    knn_search_pair<1, kBlk>(ctx.grid[c], live, rx, ry, rz, ctx.r2[c], s_beg, s_cnt, s_md, t);
    if (ctx.dbg) tk1 = clock64();
    if (live && even) flag = fit_one<1>(ctx, c, t, prim);
  } else {
    TopK<5> t;
    long long stamps[5] = {0, 0, 0, 0, 0};
    knn_search_pair<5, kBlk>(ctx.grid[c], live, rx, ry, rz, ctx.r2[c], s_beg, s_cnt, s_md, t,
                             (ctx.dbg && threadIdx.x == 0) ? stamps : nullptr);
    if (ctx.dbg) tk1 = clock64();
    if (ctx.dbg && threadIdx.x == 0 && live) {
      atomicAdd(&ctx.dbg[11], (unsigned long long)(stamps[1] - stamps[0]));   // brick probes + cell list
      atomicAdd(&ctx.dbg[12], (unsigned long long)(stamps[2] - stamps[1]));   // (unused)
      atomicAdd(&ctx.dbg[13], (unsigned long long)(stamps[3] - stamps[2]));   // candidate streaming
      atomicAdd(&ctx.dbg[14], (unsigned long long)(stamps[4] - stamps[3]));   // merge
    }
    if (live && even) flag = fit_one<5>(ctx, c, t, prim);
  }
  if (even) {
    if (live) {
      if (st->outer == 0) {
        ctx.w[gi] = 1.0;                                                          // :931-949
      } else {
        const double res = ctx.slot[gi];
        if (res != 0.0) {
          double wv;
          if (res >= st->th1) wv = 0.0;
          else if (res <= st->th2) wv = 1.0;
          else wv = sqrt(st->c2 * st->mu_used * (st->mu_used + 1.0) / res) - st->mu_used;
          ctx.w[gi] = wv;
        }
      }
      ctx.slot[gi] = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) ctx.prim[j][gi] = prim[j];
    }
    ctx.flags[gi] = flag;
  }
  if (ctx.dbg && threadIdx.x == 0 && live) {
    atomicAdd(&ctx.dbg[8], (unsigned long long)(tk1 - tk0));
    atomicAdd(&ctx.dbg[10], (unsigned long long)(clock64() - tk1));
    atomicAdd(&ctx.dbg[9], 1ull);
  }
  const int cnt = __syncthreads_count(even && (flag & kFlagCounted) != 0);
  if (threadIdx.x == 0 && cnt > 0) atomicAdd(&ctx.blk_count[buf * ctx.blk_cap + fb], cnt);
}

// `*_maxnum` caps in feature-index order (Q8): factor i is active iff it is a candidate and the number of
// counted features before it is < maxnum (the reference `return`s at the first cap-checked feature
// after the counter reached the cap, ref: :448-449, 538-539, 592-593, 735-736).
__device__ __forceinline__ bool compute_active(const DeviceCtx& ctx, int b, int c, unsigned char flag, int* s_warp) {
  if (ctx.maxnum[c] >= ctx.n[c]) return (flag & kFlagCand) != 0;   // the cap cannot bind (block-uniform branch)
  // counted features in previous blocks of this cloud
  int before = 0;
  if (threadIdx.x < 32) {
    const int* cntbuf = ctx.blk_count + (ctx.st->outer & 1) * ctx.blk_cap;
    for (int bb = ctx.blk_off[c] + (int)threadIdx.x; bb < b; bb += 32) before += cntbuf[bb];
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, (flag & kFlagCounted) != 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[1 + warp] = __popc(ballot);
  if (threadIdx.x == 0) s_warp[0] = before;
  __syncthreads();
  int prefix = s_warp[0];
  for (int wi = 0; wi < warp; ++wi) prefix += s_warp[1 + wi];
  prefix += __popc(ballot & ((1u << lane) - 1u));
  __syncthreads();                                   // s_warp is reused by the next feature block of the caller's loop
  return (flag & kFlagCand) && (prefix < ctx.maxnum[c]);
}

// One evaluation pass of the whole problem at st->evalq (all blocks) + reduction + solver (last block).
// Returns true in the block that ran the solver (the state has been written back to global memory).
template <bool kFirst>
__device__ __forceinline__ bool eval_body(const DeviceCtx& ctx) {
  FrameState* st = ctx.st;
  __shared__ double s_red[kBlk / 32][32];
  __shared__ int s_cnt[kBlk / 32][4];
  __shared__ double s_tot[kNRed];
  __shared__ double s_gather[kEvalCluster][kNRed];   // only the cluster's block 0 receives (DSMEM stores of its peers)
  __shared__ int s_warp[1 + kBlk / 32];
  __shared__ bool s_last;
  cluster_arrive();                                  // "I have started": peers may store into my shared memory
  unsigned long long tg0 = 0;
  if (ctx.dbg && threadIdx.x == 0) { tg0 = gtime_ns(); atomicMin(&ctx.dbg[0], tg0); }
  const int b = blockIdx.x;
  // per-thread accumulators: H (21), g (6), cost, slot sum per cloud (4); factor count per cloud (ints)
  double v[32];
  int nact[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.0;
  Pose7 ev;                      // read around L1 (written by the previous launch's solver block)
  ev.qw = __ldcg(&st->evalq.qw); ev.qx = __ldcg(&st->evalq.qx); ev.qy = __ldcg(&st->evalq.qy); ev.qz = __ldcg(&st->evalq.qz);
  ev.tx = __ldcg(&st->evalq.tx); ev.ty = __ldcg(&st->evalq.ty); ev.tz = __ldcg(&st->evalq.tz);
  const Rt T = pose_to_rt(ev);
  // grid-stride over the 128-feature blocks (the grid is capped so that the final partial sum stays short)
  for (int fb = b; fb < ctx.blk_off[4]; fb += gridDim.x) {
    const int c = cloud_of_block(ctx, fb);
    const int il = (fb - ctx.blk_off[c]) * kBlk + threadIdx.x;
    const int gi = ctx.pad_off[c] + il;
    bool act;
    if (kFirst) {
      act = compute_active(ctx, fb, c, ctx.flags[gi], s_warp);
      ctx.active[gi] = act ? 1 : 0;
    } else {
      act = __ldcg(&ctx.active[gi]) != 0;
    }
    if (!act) continue;
    double cpt[3];
    rt_apply(T, ctx.px[gi], ctx.py[gi], ctx.pz[gi], cpt[0], cpt[1], cpt[2]);
    const double w = ctx.w[gi];
    double r[3], J[18];
    int nr;
    double slot;
    if (c == kPlanar || c == kGround) {
      const double n[3] = {ctx.prim[0][gi], ctx.prim[1][gi], ctx.prim[2][gi]};
      functor_plane(cpt, n, ctx.prim[3][gi], w, r[0], J);
      nr = 1;
      slot = r[0] * r[0];                                                    // :101
    } else if (c == kEdge) {
      const double a[3] = {ctx.prim[0][gi], ctx.prim[1][gi], ctx.prim[2][gi]};
      const double bb[3] = {ctx.prim[3][gi], ctx.prim[4][gi], ctx.prim[5][gi]};
      functor_line(cpt, a, bb, w, r, J);
      nr = 3;
      const double s3 = r[0] + r[1] + r[2];
      slot = s3 * s3;                                                        // :69 (Q3)
    } else {
      const double q[3] = {ctx.prim[0][gi], ctx.prim[1][gi], ctx.prim[2][gi]};
      functor_point(cpt, q, w, r, J);
      nr = 3;
      const double s3 = r[0] + r[1] + r[2];
      slot = s3 * s3;                                                        // :32 (Q3)
    }
    ctx.slot[gi] = slot;                                                     // *cost side effect (Q5)
    double sq = 0.0;
    for (int k = 0; k < nr; ++k) sq += r[k] * r[k];
    // CauchyLoss(1.0): rho = log(1+s), rho' = 1/(1+s), rho'' < 0 => residual and Jacobian scaled by sqrt(rho')
    const double sum = 1.0 + sq;
    const double rho1 = fmax(DBL_MIN, 1.0 / sum);
    const double sc = sqrt(rho1);
    v[27] += 0.5 * log(sum);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[28 + k] += (k == c) ? slot : 0.0; nact[k] += (k == c) ? 1 : 0; }
    for (int k = 0; k < nr; ++k) {
      double row[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) row[j] = J[k * 6 + j] * sc;
      const double rk = r[k] * sc;
      int t = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) v[t++] += row[i] * row[j];
        v[21 + i] += row[i] * rk;
      }
    }
  }
  // ---- block reduction (fixed shape => run-to-run bit-reproducible) ----
  // warp level: butterfly "transpose" reduction -- 32 values across 32 lanes in 16+8+4+2+1 = 31 shuffles
  // (instead of 5 per value); afterwards lane L holds the warp total of value L.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const double send = up ? v[i] : v[i + o];
        const double keep = up ? v[i + o] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    s_red[warp][lane] = v[0];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int tot_k = __reduce_add_sync(0xffffffffu, nact[k]);
      if (lane == 0) s_cnt[warp][k] = tot_k;
    }
  }
  __syncthreads();
  // ---- cluster level: the 8 blocks of a cluster push their 36 block totals into block 0's shared memory ----
  cluster_wait();                                     // every block of the cluster is running
  const unsigned crank = cluster_rank();
  if (threadIdx.x < kNRed) {
    const int t = threadIdx.x;
    double s = 0.0;
    if (t < 32) {
      for (int wi = 0; wi < kBlk / 32; ++wi) s += s_red[wi][t];
    } else {
      int n = 0;
      for (int wi = 0; wi < kBlk / 32; ++wi) n += s_cnt[wi][t - 32];
      s = (double)n;
    }
    dsmem_store(&s_gather[crank][t], 0u, s);
  }
  cluster_arrive();                                   // release: my stores are visible to whoever waits
  cluster_wait();
  if (crank != 0u) return false;
  const unsigned nclusters = gridDim.x / kEvalCluster;
  if (threadIdx.x < kNRed) {
    const int t = threadIdx.x;
    double s = s_gather[0][t];
#pragma unroll
    for (int r = 1; r < kEvalCluster; ++r) s += s_gather[r][t];     // fixed order => deterministic
    ctx.partial[(size_t)cluster_id() * kNRed + t] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(ctx.counter, 1u);
    s_last = (ticket == nclusters - 1u);
  }
  __syncthreads();
  if (!s_last) return false;
  pdl_release();
  // ---- last cluster leader: deterministic sum of the per-cluster partials, then the solver state machine ----
  __threadfence();
  unsigned long long tg1 = 0, tg2 = 0, tg3 = 0;
  if (ctx.dbg && threadIdx.x == 0) tg1 = gtime_ns();
  __shared__ double s_part[3][kNRed];
  __shared__ FrameState s_state;
  __shared__ SolverShared s_solver;
  {
    // the state machine is a long dependent chain: run it on a shared-memory copy of the state
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(st);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&s_state);
    for (unsigned i = threadIdx.x; i < sizeof(FrameState) / 8; i += kBlk) dst[i] = __ldcg(src + i);
  }
  if (threadIdx.x < 3 * kNRed) {
    // 3 row groups x 36 columns, 8 independent accumulators each; the summation tree is fixed => deterministic
    const int col = threadIdx.x % kNRed, grp = threadIdx.x / kNRed;
    const int nb = (int)nclusters;
    const double* P = ctx.partial + col;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = grp; r < nb; r += 24) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = r + 3 * u;
        if (row < nb) a[u] += __ldcg(P + (size_t)row * kNRed);
      }
    }
    s_part[grp][col] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  __syncthreads();
  if (threadIdx.x < kNRed) s_tot[threadIdx.x] = (s_part[0][threadIdx.x] + s_part[1][threadIdx.x]) + s_part[2][threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    *ctx.counter = 0u;
    if (ctx.dbg) tg2 = gtime_ns();
  }
  solver_on_eval(ctx, &s_state, s_tot, &s_solver);      // all threads enter (3 helper warps + thread 0)
  if (threadIdx.x == 0) {
    if (ctx.dbg) {
      tg3 = gtime_ns();
      ctx.dbg[1] += tg1 - ctx.dbg[0];   // parallel phase: first block start -> last block arrives
      ctx.dbg[2] += tg2 - tg1;          // final partial sum
      ctx.dbg[3] += tg3 - tg2;          // solver state machine
      ctx.dbg[4] += 1ull;
      ctx.dbg[0] = ~0ull;
    }
  }
  __syncthreads();
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&s_state);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(st);
    for (unsigned i = threadIdx.x; i < sizeof(FrameState) / 8; i += kBlk) dst[i] = src[i];
  }
  __syncthreads();
  return true;
}

// Clusters of 8 blocks: the block totals are combined through distributed shared memory before they reach
// global memory, so the serial tail sums 1/8 of the rows (40 instead of 315 at F = 40k).
template <bool kFirst>
__global__ void __cluster_dims__(kEvalCluster, 1, 1) __launch_bounds__(kBlk, 3) k_eval(const __grid_constant__ DeviceCtx ctx) {
  pdl_prologue();
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != (kFirst ? kPhaseIter0 : kPhaseCand)) return;     // grid-uniform
  eval_body<kFirst>(ctx);
}

// ---- standalone caps kernel (used by the build_factors test entry point) ----
__global__ void __launch_bounds__(kBlk) k_caps(const __grid_constant__ DeviceCtx ctx) {
  __shared__ int s_warp[1 + kBlk / 32];
  const int b = blockIdx.x;
  const int c = cloud_of_block(ctx, b);
  const int gi = ctx.pad_off[c] + (b - ctx.blk_off[c]) * kBlk + threadIdx.x;
  ctx.active[gi] = compute_active(ctx, b, c, ctx.flags[gi], s_warp) ? 1 : 0;
}

__global__ void k_set_pose(DeviceCtx ctx, Predict x6) {   // x6.m[0..5] = tangent
  for (int i = threadIdx.x; i < ctx.blk_off[4]; i += blockDim.x) ctx.blk_count[i] = 0;
  if (threadIdx.x != 0) return;
  FrameState* st = ctx.st;
  for (int i = 0; i < 6; ++i) st->x[i] = x6.m[i];
  st->xq = se3_exp(st->x);
  st->evalq = st->xq;
  st->frame_done = 0; st->phase = kPhaseIter0; st->outer = 0; st->status = 0;
}

// ---- piecewise test kernels ----
template <int K>
__global__ void k_knn(GridDesc g, const double* origin, const double* q, unsigned nq, double r2, int* idx,
                      double* d2, int* count) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  TopK<K> t;
  knn_search<K>(g, q[3ull * i] - origin[0], q[3ull * i + 1] - origin[1], q[3ull * i + 2] - origin[2], r2, t);
#pragma unroll
  for (int j = 0; j < K; ++j) {
    idx[(size_t)i * K + j] = (t.pos[j] >= 0) ? t.idx[j] : -1;
    d2[(size_t)i * K + j] = t.d2[j];
  }
  count[i] = t.count();
}

// getFitnessScore, ref: registration.cpp:257-296: 1-NN of the UNTRANSFORMED scan points within fitness_thres.
__global__ void __launch_bounds__(kBlk) k_fitness(const __grid_constant__ DeviceCtx ctx, double r2, double* out /*[blocks][2]*/) {
  const int b = blockIdx.x;
  const int c = cloud_of_block(ctx, b);
  const int il = (b - ctx.blk_off[c]) * kBlk + threadIdx.x;
  const int gi = ctx.pad_off[c] + il;
  double err = 0.0, cnt = 0.0;
  if (il < ctx.n[c]) {
    TopK<1> t;
    knn_search<1>(ctx.grid[c], ctx.px[gi] - ctx.origin[0], ctx.py[gi] - ctx.origin[1], ctx.pz[gi] - ctx.origin[2], r2, t);
    if (t.pos[0] >= 0) { err = t.d2[0]; cnt = 1.0; }
  }
  __shared__ double s_e[kBlk / 32], s_c[kBlk / 32];
  for (int o = 16; o > 0; o >>= 1) { err += __shfl_xor_sync(0xffffffffu, err, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  if ((threadIdx.x & 31) == 0) { s_e[threadIdx.x >> 5] = err; s_c[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0, n = 0;
    for (int wi = 0; wi < kBlk / 32; ++wi) { e += s_e[wi]; n += s_c[wi]; }
    out[2 * b] = e; out[2 * b + 1] = n;
  }
}

__global__ void k_functor(int type, Predict x6, unsigned m, const double* p, const double* a, const double* bq,
                          const double* w, double* r, double* J, double* cost) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const Rt T = pose_to_rt(se3_exp(x6.m));
  double c[3];
  rt_apply(T, p[3ull * i], p[3ull * i + 1], p[3ull * i + 2], c[0], c[1], c[2]);
  if (type == 0) {          // point-to-point: a = target q
    double rr[3], JJ[18];
    const double q[3] = {a[3ull * i], a[3ull * i + 1], a[3ull * i + 2]};
    functor_point(c, q, w[i], rr, JJ);
    for (int k = 0; k < 3; ++k) r[3ull * i + k] = rr[k];
    for (int k = 0; k < 18; ++k) J[18ull * i + k] = JJ[k];
    cost[i] = (rr[0] + rr[1] + rr[2]) * (rr[0] + rr[1] + rr[2]);
  } else if (type == 1) {   // point-to-line: a, bq = line points
    double rr[3], JJ[18];
    const double la[3] = {a[3ull * i], a[3ull * i + 1], a[3ull * i + 2]};
    const double lb[3] = {bq[3ull * i], bq[3ull * i + 1], bq[3ull * i + 2]};
    functor_line(c, la, lb, w[i], rr, JJ);
    for (int k = 0; k < 3; ++k) r[3ull * i + k] = rr[k];
    for (int k = 0; k < 18; ++k) J[18ull * i + k] = JJ[k];
    cost[i] = (rr[0] + rr[1] + rr[2]) * (rr[0] + rr[1] + rr[2]);
  } else {                  // point-to-plane: a = normal, bq = d
    double rr, JJ[6];
    const double n[3] = {a[3ull * i], a[3ull * i + 1], a[3ull * i + 2]};
    functor_plane(c, n, bq[i], w[i], rr, JJ);
    r[i] = rr;
    for (int k = 0; k < 6; ++k) J[6ull * i + k] = JJ[k];
    cost[i] = rr * rr;
  }
}

__global__ void k_se3(int op, Predict in, double* out) {
  if (threadIdx.x != 0) return;
  if (op == 0) { pose_to_matrix(se3_exp(in.m), out); }
  else if (op == 1) { Pose7 p; const bool ok = pose_from_matrix(in.m, p); se3_log(p, out); out[6] = ok ? 1.0 : 0.0; }
  else { se3_log(se3_mul(se3_exp(in.m + 6), se3_exp(in.m)), out); }   // plus: in.m[0..5] = x, in.m[6..11] = delta
}

}  // namespace tloam

// =================================================================================================
// Host side: handle + C ABI
// =================================================================================================
using namespace tloam;

#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      snprintf(h->last_error, sizeof(h->last_error), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                                    \
      return TLOAM_B200_ERR_CUDA;                                                                      \
    }                                                                                                  \
  } while (0)

struct tloam_b200_handle {
  tloam_tls_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t copy_stream = nullptr;              // H2D of the map clouds, overlapped with the build (set_target)
  cudaEvent_t ev_copy[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int last_uploaded = 0;
  cudaEvent_t ev_src = nullptr;                    // set_source: the H2D copies have landed
  char last_error[512] = {0};
  long long launches = 0;
  int launches_frame = 0;
  // source
  size_t n_src[4] = {0, 0, 0, 0};
  bool have_src = false, have_tgt = false, frame_pending = false;
  double* d_stage_src = nullptr; size_t cap_stage_src = 0;      // points
  double* d_feat = nullptr; size_t cap_pad = 0;                 // 3 + 2 + 6 arrays of cap_pad doubles
  unsigned char* d_flags = nullptr;                             // flags + active
  int* d_blk_count = nullptr; double* d_partial = nullptr; size_t cap_blocks = 0;
  unsigned* d_counter = nullptr;
  FrameState* d_state = nullptr;
  tloam_b200_stats* d_stats = nullptr;
  // target
  size_t n_tgt[4] = {0, 0, 0, 0};
  double* d_stage_tgt = nullptr; size_t cap_stage_tgt = 0;
  unsigned* d_scratch = nullptr; size_t cap_scratch = 0;        // slot_of + rank_of
  unsigned char* d_blob = nullptr; size_t cap_blob = 0, blob_bytes = 0;
  MapHeader hdr;                                                // host copy of the layout (origin filled lazily)
  bool origin_known = false;
  // pinned host staging for results
  double* h_result = nullptr;                                   // 16 result + 2 (status, done)
  tloam_b200_stats* h_stats = nullptr;
  DeviceCtx ctx;
  int total_blocks = 0;
  // whole-frame CUDA graph (re-captured only when the device context changes)
  Predict* h_predict = nullptr; Predict* d_predict = nullptr;
  cudaGraphExec_t gexec = nullptr; DeviceCtx gctx; bool gvalid = false; int glaunches = 0; bool use_graph = true;
  bool use_pdl = false;     // programmatic dependent launch between the frame kernels: measured no faster inside the graph (opt-in)
  // optional per-kernel-class timing (CUDA events around every launch; off by default)
  bool profiling = false;
  bool traced_last = false;
  bool trace = false;                                           // record tloam_b200_stats traces
  std::vector<cudaEvent_t> ev_pool;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t ev_next = 0;
  tloam_b200_profile prof;
  unsigned long long* d_dbg = nullptr;
  // ---- device-resident submap ((f)-1) ----
  tloam_submap_config scfg;
  bool submap_ready = false;
  double* d_acc[2] = {nullptr, nullptr};   size_t cap_acc[2] = {0, 0}, n_acc[2] = {0, 0};     // edge, ground accumulators
  double* d_acc_tmp = nullptr;             size_t cap_acc_tmp = 0;
  std::vector<double*> ring;               std::vector<size_t> ring_n, ring_cap;               // planar sliding window
  double* d_cat = nullptr;                 size_t cap_cat = 0, n_cat = 0;                       // concatenated planar window
  double* d_sphere0 = nullptr;             size_t n_sphere0 = 0; bool sphere_is_init = false;  // frame-0 sphere submap
  double* d_up = nullptr;                  size_t cap_up = 0;                                   // upload staging
  unsigned char* d_vox = nullptr;          size_t cap_vox = 0;                                  // voxel hash scratch
  // ---- PCA feature extraction ((f)-2): one arena, carved up per call ----
  unsigned char* d_fe = nullptr;           size_t cap_fe = 0;  bool fe_attr_set = false;
  double* d_pose = nullptr;
};

// launch bookkeeping: counts the kernel and, in profiling mode, brackets it with events
struct LaunchScope {
  tloam_b200_handle* h; int cls; cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(tloam_b200_handle* hh, int c) : h(hh), cls(c) {
    h->launches++;
    if (h->profiling) {
      while (h->ev_pool.size() < h->ev_next + 2) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
      a = h->ev_pool[h->ev_next++]; b = h->ev_pool[h->ev_next++];
      cudaEventRecord(a, h->stream);
    }
  }
  ~LaunchScope() {
    if (h->profiling) { cudaEventRecord(b, h->stream); h->spans.push_back({cls, a, b}); }
  }
};
#define TL_LAUNCH(cls, ...) do { LaunchScope ls__(h, cls); __VA_ARGS__; } while (0)

static size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

template <typename K>
static cudaError_t launch_pdl(K kernel, int grid, int block, cudaStream_t stream, const DeviceCtx& c, bool pdl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, c);
}



extern "C" {

void tloam_b200_default_config(tloam_tls_config* c) {   // ref: config/mapping/lidar_odometry.yaml:23-39
  c->k_corr = 10; c->factor_num = 4;
  c->edge_dist_thres = 1.0; c->sphere_dist_thres = 0.5; c->planar_dist_thres = 0.5; c->ground_dist_thres = 0.5;
  c->edge_dir_thres = 0.85;
  c->edge_maxnum = 1200; c->sphere_maxnum = 200; c->planar_maxnum = 2500; c->ground_maxnum = 2000;
  c->max_iterations = 4; c->cost_threshold = 0.000000005; c->gnc_factor = 11.8; c->noise_bound = 0.01;
  c->fitness_thres = 0.02;
  c->ceres_max_num_iterations = 4;
  c->reinit_dir[0] = 1.0; c->reinit_dir[1] = 1.0; c->reinit_dir[2] = 1.0;
}

const char* tloam_b200_status_string(int s) {
  switch (s) {
    case TLOAM_B200_OK: return "ok";
    case TLOAM_B200_ERR_INVALID_ARG: return "invalid argument";
    case TLOAM_B200_ERR_TOO_FEW_POINTS: return "a cloud has fewer than 10 points";
    case TLOAM_B200_ERR_BAD_POSE: return "predicted pose is not a rigid transform";
    case TLOAM_B200_ERR_CUDA: return "CUDA error";
    case TLOAM_B200_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case TLOAM_B200_ERR_NOT_READY: return "source or target not set";
    case TLOAM_B200_ERR_NUMERIC: return "non-finite value in the solve";
    case TLOAM_B200_ERR_MAP_DENSITY: return "a map cell holds more than 65535 points";
    default: return "unknown status";
  }
}

static double radius_of(const tloam_tls_config& c, int cloud) {
  return cloud == 0 ? c.edge_dist_thres : cloud == 1 ? c.sphere_dist_thres : cloud == 2 ? c.planar_dist_thres : c.ground_dist_thres;
}

int tloam_b200_create(const tloam_tls_config* cfg, int device, void* stream, tloam_b200_handle** out) {
  if (!cfg || !out) return TLOAM_B200_ERR_INVALID_ARG;
  *out = nullptr;
  if (cfg->factor_num < 2 || cfg->factor_num > 4 || cfg->max_iterations < 1 ||
      cfg->max_iterations > TLOAM_B200_MAX_OUTER || cfg->ceres_max_num_iterations < 0 ||
      cfg->ceres_max_num_iterations > TLOAM_B200_MAX_INNER)
    return TLOAM_B200_ERR_INVALID_ARG;
  for (int c = 0; c < 4; ++c) if (!(radius_of(*cfg, c) > 0.0)) return TLOAM_B200_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return TLOAM_B200_ERR_NO_DEVICE;
  }
  tloam_b200_handle* h = new (std::nothrow) tloam_b200_handle();
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  h->cfg = *cfg;
  h->device = device;
  auto fail = [&](int code) { tloam_b200_destroy(h); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
    h->own_stream = true;
  }
  if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  for (int i = 0; i < 5; ++i)
    if (cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_src, cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_state, sizeof(FrameState)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_stats, sizeof(tloam_b200_stats)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_counter, 256) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMallocHost(&h->h_result, 32 * sizeof(double)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMallocHost(&h->h_stats, sizeof(tloam_b200_stats)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMallocHost(&h->h_predict, sizeof(Predict)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_predict, sizeof(Predict)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  { const char* e = getenv("TLOAM_B200_NO_GRAPH"); h->use_graph = !(e && e[0] == '1'); }
  { const char* e = getenv("TLOAM_B200_PDL"); h->use_pdl = (e && e[0] == '1'); }
  if (kEvalCluster > 8) {                          // cluster sizes above 8 are "non-portable": opt in per kernel
    if (cudaFuncSetAttribute(k_eval<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_eval<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return fail(TLOAM_B200_ERR_CUDA);
  }
  // identity curr/last pose (the reference leaves them uninitialised until the first scanMatching)
  FrameState init;
  memset(&init, 0, sizeof(init));
  for (int i = 0; i < 16; ++i) init.curr_pose[i] = init.last_pose[i] = init.result[i] = (i % 5 == 0) ? 1.0 : 0.0;
  init.frame_done = 1;
  if (cudaMemcpy(h->d_state, &init, sizeof(init), cudaMemcpyHostToDevice) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMemset(h->d_counter, 0, 256) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  memset(&h->ctx, 0, sizeof(h->ctx));
  memset(&h->hdr, 0, sizeof(h->hdr));
  *out = h;
  return TLOAM_B200_OK;
}

int tloam_b200_destroy(tloam_b200_handle* h) {
  if (!h) return TLOAM_B200_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_stage_src); cudaFree(h->d_feat); cudaFree(h->d_flags); cudaFree(h->d_blk_count);
  cudaFree(h->d_partial); cudaFree(h->d_counter); cudaFree(h->d_state); cudaFree(h->d_stats);
  cudaFree(h->d_stage_tgt); cudaFree(h->d_scratch); cudaFree(h->d_blob); cudaFree(h->d_dbg);
  cudaFree(h->d_acc[0]); cudaFree(h->d_acc[1]); cudaFree(h->d_acc_tmp); cudaFree(h->d_cat); cudaFree(h->d_sphere0);
  cudaFree(h->d_up); cudaFree(h->d_vox); cudaFree(h->d_pose); cudaFree(h->d_fe);
  for (double* p : h->ring) cudaFree(p);
  if (h->h_result) cudaFreeHost(h->h_result);
  if (h->h_stats) cudaFreeHost(h->h_stats);
  if (h->h_predict) cudaFreeHost(h->h_predict);
  cudaFree(h->d_predict);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < 5; ++i) if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
  if (h->ev_src) cudaEventDestroy(h->ev_src);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return TLOAM_B200_OK;
}

// fills the configuration part of the device context
static void fill_ctx_config(tloam_b200_handle* h) {
  DeviceCtx& c = h->ctx;
  const tloam_tls_config& f = h->cfg;
  for (int k = 0; k < 4; ++k) { const double r = radius_of(f, k); c.r2[k] = r * r; }
  c.maxnum[0] = f.edge_maxnum; c.maxnum[1] = f.sphere_maxnum; c.maxnum[2] = f.planar_maxnum; c.maxnum[3] = f.ground_maxnum;
  c.factor_num = f.factor_num; c.max_iterations = f.max_iterations; c.ceres_max_it = f.ceres_max_num_iterations;
  c.edge_dir_thres = f.edge_dir_thres; c.cost_threshold = f.cost_threshold; c.gnc_factor = f.gnc_factor;
  c.noise_bound = f.noise_bound; c.fitness_thres = f.fitness_thres;
  for (int k = 0; k < 3; ++k) c.reinit_dir[k] = f.reinit_dir[k];
  c.st = h->d_state; c.stats = h->d_stats; c.counter = h->d_counter;
}

static int set_source_impl(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4], bool on_device) {
  if (!h || !xyz || !n) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  size_t total = 0, pad = 0;
  int blocks = 0;
  for (int c = 0; c < 4; ++c) {
    if (n[c] > 0 && !xyz[c]) return TLOAM_B200_ERR_INVALID_ARG;
    if (n[c] > (size_t)1 << 30) return TLOAM_B200_ERR_INVALID_ARG;
    total += n[c];
    pad += round_up(n[c], kBlk);
  }
  if (pad == 0) pad = kBlk;
  blocks = (int)(pad / kBlk);
  // host input is staged; device input is read in place -- unless the device-side submap is in use, whose update
  // appends the staged edge / ground features after the frame (tloam_b200_submap_update)
  const bool stage = !on_device || h->submap_ready;
  if (stage && total > h->cap_stage_src) {
    cudaFree(h->d_stage_src);
    h->cap_stage_src = total + total / 4 + 1024;
    CU_TRY(cudaMalloc(&h->d_stage_src, h->cap_stage_src * 3 * sizeof(double)));
  }
  if (pad > h->cap_pad) {
    cudaFree(h->d_feat); cudaFree(h->d_flags);
    h->cap_pad = round_up(pad + pad / 4, kBlk);
    CU_TRY(cudaMalloc(&h->d_feat, h->cap_pad * 11 * sizeof(double)));
    CU_TRY(cudaMalloc(&h->d_flags, h->cap_pad * 2));
  }
  if ((size_t)blocks > h->cap_blocks) {
    cudaFree(h->d_blk_count); cudaFree(h->d_partial);
    h->cap_blocks = h->cap_pad / kBlk + 8;
    CU_TRY(cudaMalloc(&h->d_blk_count, 2 * h->cap_blocks * sizeof(int)));
    CU_TRY(cudaMemsetAsync(h->d_blk_count, 0, 2 * h->cap_blocks * sizeof(int), h->stream));
    CU_TRY(cudaMalloc(&h->d_partial, h->cap_blocks * kNRed * sizeof(double)));
  }
  DeviceCtx& c = h->ctx;
  fill_ctx_config(h);
  size_t off = 0, poff = 0;
  const double* src[4];
  c.blk_off[0] = 0;
  for (int k = 0; k < 4; ++k) {
    src[k] = stage ? h->d_stage_src + 3 * off : xyz[k];
    if (n[k] > 0 && stage)
      CU_TRY(cudaMemcpyAsync(h->d_stage_src + 3 * off, xyz[k], n[k] * 3 * sizeof(double),
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    c.n[k] = (int)n[k];
    c.pad_off[k] = (int)poff;
    off += n[k];
    poff += round_up(n[k], kBlk);
    c.blk_off[k + 1] = (int)(poff / kBlk);
    h->n_src[k] = n[k];
  }
  double* f = h->d_feat;
  const size_t cp = h->cap_pad;
  c.px = f; c.py = f + cp; c.pz = f + 2 * cp; c.w = f + 3 * cp; c.slot = f + 4 * cp;
  for (int j = 0; j < 6; ++j) c.prim[j] = f + (5 + j) * cp;
  c.flags = h->d_flags; c.active = h->d_flags + cp;
  c.blk_count = h->d_blk_count; c.blk_cap = (int)h->cap_blocks; c.partial = h->d_partial;
  h->total_blocks = c.blk_off[4];
  if (!on_device) CU_TRY(cudaEventRecord(h->ev_src, h->stream));      // the uploads are in; the caller's buffers are free
  if (h->total_blocks > 0) {
    TL_LAUNCH(TLOAM_B200_K_STAGE_SOURCE, (k_stage_source<<<h->total_blocks, kBlk, 0, h->stream>>>(src[0], src[1], src[2], src[3], c, f, f + cp, f + 2 * cp)));
    CU_TRY(cudaGetLastError());
  }
  h->have_src = true;
  if (!on_device) CU_TRY(cudaEventSynchronize(h->ev_src));    // caller buffers may be freed on return
  return TLOAM_B200_OK;
}

int tloam_b200_set_source(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_source_impl(h, xyz, n, false);
}
int tloam_b200_set_source_device(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_source_impl(h, xyz, n, true);
}

static unsigned next_pow2(size_t v) { unsigned p = 256; while ((size_t)p < v) p <<= 1; return p; }

static int layout_map(tloam_b200_handle* h, const size_t n[4]) {
  MapHeader& hd = h->hdr;
  memset(&hd, 0, sizeof(hd));
  hd.magic = kMapMagic;
  size_t off = sizeof(MapHeader);
  for (int c = 0; c < 4; ++c) {
    hd.n[c] = (unsigned)n[c];
    hd.tsize[c] = next_pow2(n[c] + 1);      // >= bricks + 1 even if every point sits in its own brick
    hd.cell[c] = radius_of(h->cfg, c);
    hd.pts_off[c] = off; off += round_up(n[c] * sizeof(float4), 256);
  }
  for (int c = 0; c < 4; ++c) { hd.table_off[c] = off; off += (size_t)hd.tsize[c] * kBrickBytes; }
  for (int d = 0; d < 3; ++d) { hd.bbox_enc[d] = ~0ull; hd.bbox_enc[3 + d] = 0ull; }
  h->blob_bytes = off;
  if (off > h->cap_blob) {
    cudaFree(h->d_blob);
    h->cap_blob = off + off / 4;
    CU_TRY(cudaMalloc(&h->d_blob, h->cap_blob));
  }
  return TLOAM_B200_OK;
}

static void bind_map(tloam_b200_handle* h) {
  DeviceCtx& c = h->ctx;
  for (int k = 0; k < 4; ++k) {
    c.grid[k].pts = reinterpret_cast<const float4*>(h->d_blob + h->hdr.pts_off[k]);
    c.grid[k].table = reinterpret_cast<const uint4*>(h->d_blob + h->hdr.table_off[k]);
    c.grid[k].mask = h->hdr.tsize[k] - 1u;
    c.grid[k].n = h->hdr.n[k];
    c.grid[k].cell = h->hdr.cell[k];
    c.grid[k].inv_cell = 1.0 / h->hdr.cell[k];
  }
  c.origin = reinterpret_cast<const double*>(h->d_blob + offsetof(MapHeader, origin));
  c.map_flags = reinterpret_cast<const unsigned long long*>(h->d_blob + offsetof(MapHeader, build_flags));
}

// the origin lives in the device header; fetch it once per map (tiny D2H) so that it can be passed by value
// (and the build flags with it: a map whose cell counters overflowed must not be searched)
static int fetch_origin(tloam_b200_handle* h) {
  if (!h->origin_known) {
    CU_TRY(cudaMemcpyAsync(h->h_result + 24, h->d_blob + offsetof(MapHeader, origin), 3 * sizeof(double),
                           cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaMemcpyAsync(h->h_result + 27, h->d_blob + offsetof(MapHeader, build_flags), sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (int d = 0; d < 3; ++d) h->hdr.origin[d] = h->h_result[24 + d];
    memcpy(&h->hdr.build_flags, h->h_result + 27, sizeof(unsigned long long));
    h->origin_known = true;
  }
  return (h->hdr.build_flags & 1ull) ? TLOAM_B200_ERR_MAP_DENSITY : TLOAM_B200_OK;
}

static int set_target_impl(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4], bool on_device) {
  if (!h || !xyz || !n) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  size_t total = 0;
  for (int c = 0; c < 4; ++c) {
    if (n[c] > 0 && !xyz[c]) return TLOAM_B200_ERR_INVALID_ARG;
    if (n[c] > (size_t)1 << 30) return TLOAM_B200_ERR_INVALID_ARG;
    total += n[c];
  }
  if (total > h->cap_scratch) {
    cudaFree(h->d_scratch);
    h->cap_scratch = total + total / 4 + 1024;
    CU_TRY(cudaMalloc(&h->d_scratch, h->cap_scratch * 2 * sizeof(unsigned)));
  }
  if (!on_device && total > h->cap_stage_tgt) {
    cudaFree(h->d_stage_tgt);
    h->cap_stage_tgt = total + total / 4 + 1024;
    CU_TRY(cudaMalloc(&h->d_stage_tgt, h->cap_stage_tgt * 3 * sizeof(double)));
  }
  int rc = layout_map(h, n);
  if (rc != TLOAM_B200_OK) return rc;
  MapBuildArgs a;
  a.blob = h->d_blob; a.slot_of = h->d_scratch; a.rank_of = h->d_scratch + h->cap_scratch;
  a.only_cloud = -1;
  size_t off = 0;
  int first = -1;                                  // first non-empty cloud: its bounding box defines the origin
  for (int c = 0; c < 4; ++c) {
    a.stage_off[c] = (unsigned)off;
    // host input is staged; device input is read in place by the build kernels (stream-ordered, no extra copy)
    a.src[c] = on_device ? xyz[c] : h->d_stage_tgt + 3 * off;
    if (first < 0 && n[c] > 0) first = c;
    off += n[c];
    h->n_tgt[c] = n[c];
  }
  a.stage_off[4] = (unsigned)off;
  // header + zeroed tables (key 0 == empty)
  CU_TRY(cudaMemcpyAsync(h->d_blob, &h->hdr, sizeof(MapHeader), cudaMemcpyHostToDevice, h->stream));
  const size_t tables_bytes = h->blob_bytes - h->hdr.table_off[0];
  CU_TRY(cudaMemsetAsync(h->d_blob + h->hdr.table_off[0], 0, tables_bytes, h->stream));
  const unsigned tb = 256;
  auto blocks_for = [&](size_t cnt) { return (unsigned)((cnt + tb - 1) / tb); };
  if (total == 0) {
    TL_LAUNCH(TLOAM_B200_K_MAP_ORIGIN, (k_map_origin<<<1, 32, 0, h->stream>>>(a)));
  } else if (on_device) {
    // inputs already in HBM: one pass over all clouds per kernel
    MapBuildArgs ab = a;
    ab.i_beg = a.stage_off[first]; ab.i_end = a.stage_off[first + 1];
    const unsigned gbb = blocks_for(n[first]);
    TL_LAUNCH(TLOAM_B200_K_MAP_BBOX, (k_map_bbox<<<(gbb < 592u ? gbb : 592u), tb, 0, h->stream>>>(ab)));   // + origin
    a.i_beg = 0; a.i_end = (unsigned)total;
    const unsigned gb = blocks_for(total);
    unsigned tslots = 0;
    for (int c = 0; c < 4; ++c) tslots += h->hdr.tsize[c];
    TL_LAUNCH(TLOAM_B200_K_MAP_INSERT, (k_map_insert<<<gb, tb, 0, h->stream>>>(a)));
    TL_LAUNCH(TLOAM_B200_K_MAP_OFFSETS, (k_map_offsets<<<(tslots + tb - 1) / tb, tb, 0, h->stream>>>(a)));
    TL_LAUNCH(TLOAM_B200_K_MAP_SCATTER, (k_map_scatter<<<gb, tb, 0, h->stream>>>(a)));
    CU_TRY(cudaGetLastError());
  } else {
    // host inputs: the clouds cross PCIe one after the other on a copy stream; cloud c is inserted, offset and
    // scattered on the compute stream while cloud c+1 is still in flight.  The origin cloud goes first, then the
    // others by descending size, so that the build left exposed after the last copy is the smallest one.
    int order[4], m = 0;
    order[m++] = first;
    for (int c = 0; c < 4; ++c) if (c != first && n[c] > 0) order[m++] = c;
    for (int i = 2; i < m; ++i)
      for (int j = i; j > 1 && n[order[j]] > n[order[j - 1]]; --j) { const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    CU_TRY(cudaEventRecord(h->ev_copy[0], h->stream));                  // the staging buffer is free once earlier work is done
    CU_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_copy[0], 0));
    for (int k = 0; k < m; ++k) {
      const int c = order[k];
      CU_TRY(cudaMemcpyAsync(h->d_stage_tgt + 3 * (size_t)a.stage_off[c], xyz[c], n[c] * 3 * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
      CU_TRY(cudaEventRecord(h->ev_copy[1 + c], h->copy_stream));
      h->last_uploaded = c;
    }
    for (int k = 0; k < m; ++k) {
      const int c = order[k];
      CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_copy[1 + c], 0));
      MapBuildArgs ac = a;
      ac.i_beg = a.stage_off[c]; ac.i_end = a.stage_off[c + 1]; ac.only_cloud = c;
      const unsigned gb = blocks_for(n[c]);
      if (k == 0) TL_LAUNCH(TLOAM_B200_K_MAP_BBOX, (k_map_bbox<<<(gb < 592u ? gb : 592u), tb, 0, h->stream>>>(ac)));   // + origin
      TL_LAUNCH(TLOAM_B200_K_MAP_INSERT, (k_map_insert<<<gb, tb, 0, h->stream>>>(ac)));
      TL_LAUNCH(TLOAM_B200_K_MAP_OFFSETS, (k_map_offsets<<<(h->hdr.tsize[c] + tb - 1) / tb, tb, 0, h->stream>>>(ac)));
      TL_LAUNCH(TLOAM_B200_K_MAP_SCATTER, (k_map_scatter<<<gb, tb, 0, h->stream>>>(ac)));
    }
    CU_TRY(cudaGetLastError());
  }
  bind_map(h);
  h->origin_known = false;
  h->have_tgt = true;
  // host path: the caller's buffers are free once the LAST upload has landed; the build of the last cloud may
  // still be running on the compute stream (everything that follows is ordered behind it on that stream)
  if (!on_device && total > 0) CU_TRY(cudaEventSynchronize(h->ev_copy[1 + h->last_uploaded]));
  return TLOAM_B200_OK;
}

int tloam_b200_set_target(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_target_impl(h, xyz, n, false);
}
int tloam_b200_set_target_device(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_target_impl(h, xyz, n, true);
}

int tloam_b200_get_map_origin(tloam_b200_handle* h, double origin[3]) {
  if (!h || !origin) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  CU_TRY(cudaSetDevice(h->device));
  int rc = fetch_origin(h);
  if (rc != TLOAM_B200_OK) return rc;
  for (int d = 0; d < 3; ++d) origin[d] = h->hdr.origin[d];
  return TLOAM_B200_OK;
}

int tloam_b200_map_blob_size(tloam_b200_handle* h, size_t* bytes) {
  if (!h || !bytes) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  *bytes = h->blob_bytes;
  return TLOAM_B200_OK;
}

int tloam_b200_map_export(tloam_b200_handle* h, void* d_dst, size_t bytes) {
  if (!h || !d_dst) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (bytes < h->blob_bytes) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaMemcpyAsync(d_dst, h->d_blob, h->blob_bytes, cudaMemcpyDeviceToDevice, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

int tloam_b200_map_import(tloam_b200_handle* h, const void* d_src, size_t bytes) {
  if (!h || !d_src || bytes < sizeof(MapHeader)) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  MapHeader hd;
  CU_TRY(cudaMemcpy(&hd, d_src, sizeof(hd), cudaMemcpyDeviceToHost));
  if (hd.magic != kMapMagic) return TLOAM_B200_ERR_INVALID_ARG;
  size_t need = hd.table_off[3] + (size_t)hd.tsize[3] * kBrickBytes;
  if (bytes < need) return TLOAM_B200_ERR_INVALID_ARG;
  for (int c = 0; c < 4; ++c)
    if (hd.cell[c] != radius_of(h->cfg, c)) return TLOAM_B200_ERR_INVALID_ARG;   // grid cell must equal this handle's radius
  if (need > h->cap_blob) {
    cudaFree(h->d_blob);
    h->cap_blob = need + need / 4;
    CU_TRY(cudaMalloc(&h->d_blob, h->cap_blob));
  }
  CU_TRY(cudaMemcpyAsync(h->d_blob, d_src, need, cudaMemcpyDeviceToDevice, h->stream));
  h->hdr = hd;
  h->blob_bytes = need;
  for (int c = 0; c < 4; ++c) h->n_tgt[c] = hd.n[c];
  fill_ctx_config(h);
  bind_map(h);
  h->origin_known = true;
  h->have_tgt = true;
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

static int check_ready(tloam_b200_handle* h) {
  if (!h->have_src || !h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  for (int c = 0; c < 4; ++c)
    if (h->n_src[c] < 10 || h->n_tgt[c] < 10) return TLOAM_B200_ERR_TOO_FEW_POINTS;   // ref: :928-929
  return TLOAM_B200_OK;
}

// enqueues the frame's fixed launch sequence on h->stream (also used under stream capture)
static int enqueue_frame(tloam_b200_handle* h, const DeviceCtx& c) {
  const int nb = h->total_blocks;
  // k_eval is grid-stride over the feature blocks; its grid is a whole number of clusters
  const int ne = ((nb < kEvalGridCap ? nb : kEvalGridCap) + kEvalCluster - 1) / kEvalCluster * kEvalCluster;
  CU_TRY(cudaMemcpyAsync(h->d_predict, h->h_predict, sizeof(Predict), cudaMemcpyHostToDevice, h->stream));
  TL_LAUNCH(TLOAM_B200_K_BEGIN_FRAME, (k_begin_frame<<<1, 256, 0, h->stream>>>(c, h->d_predict)));
  for (int outer = 0; outer < h->cfg.max_iterations; ++outer) {
    const bool pdl = h->use_pdl && !h->profiling;
    TL_LAUNCH(TLOAM_B200_K_CORRESPOND, (launch_pdl(k_correspond, nb * 2, kBlk, h->stream, c, pdl)));
    TL_LAUNCH(TLOAM_B200_K_EVAL_FIRST, (launch_pdl(k_eval<true>, ne, kBlk, h->stream, c, pdl)));
    for (int it = 0; it < h->cfg.ceres_max_num_iterations; ++it)
      TL_LAUNCH(TLOAM_B200_K_EVAL, (launch_pdl(k_eval<false>, ne, kBlk, h->stream, c, pdl)));
  }
  // result[16] + {frame_done, status}: contiguous in FrameState
  CU_TRY(cudaMemcpyAsync(h->h_result, (const char*)h->d_state + offsetof(FrameState, result), 16 * sizeof(double) + 2 * sizeof(int),
                         cudaMemcpyDeviceToHost, h->stream));
  return TLOAM_B200_OK;
}

static int scan_match_enqueue(tloam_b200_handle* h, const double* predict);

int tloam_b200_scan_match_async(tloam_b200_handle* h, const double predict[16]) {
  if (!h || !predict) return TLOAM_B200_ERR_INVALID_ARG;
  return scan_match_enqueue(h, predict);
}

// (f)-3: the frame is predicted on the device from the two last results (no host input at all)
int tloam_b200_scan_match_predicted_async(tloam_b200_handle* h) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  return scan_match_enqueue(h, nullptr);
}

int tloam_b200_scan_match_predicted(tloam_b200_handle* h, double result[16], tloam_b200_stats* stats) {
  const int rc = tloam_b200_scan_match_predicted_async(h);
  if (rc != TLOAM_B200_OK) return rc;
  return tloam_b200_get_result(h, result, stats);
}

int tloam_b200_set_pose_history(tloam_b200_handle* h, const double last_pose[16], const double curr_pose[16]) {
  if (!h || !last_pose || !curr_pose) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  CU_TRY(cudaMemcpy((char*)h->d_state + offsetof(FrameState, last_pose), last_pose, 16 * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy((char*)h->d_state + offsetof(FrameState, curr_pose), curr_pose, 16 * sizeof(double), cudaMemcpyHostToDevice));
  return TLOAM_B200_OK;
}

static int scan_match_enqueue(tloam_b200_handle* h, const double* predict) {
  const int rc = check_ready(h);
  if (rc != TLOAM_B200_OK) return rc;
  CU_TRY(cudaSetDevice(h->device));
  if (predict) { memcpy(h->h_predict->m, predict, 16 * sizeof(double)); h->h_predict->from_state = 0.0; }
  else h->h_predict->from_state = 1.0;
  DeviceCtx c = h->ctx;
  if (!h->trace) c.stats = nullptr;             // skip the per-iteration trace (fewer instructions in the serial solver)
  const int per_frame = 1 + h->cfg.max_iterations * (2 + h->cfg.ceres_max_num_iterations);
  CU_TRY(cudaEventRecord(h->ev0, h->stream));
  if (h->use_graph && !h->profiling) {
    // one graph launch per frame; the graph is re-captured only when the device context changed
    if (!h->gvalid || memcmp(&h->gctx, &c, sizeof(DeviceCtx)) != 0) {
      h->gvalid = false;
      cudaGraph_t graph = nullptr;
      CU_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      const long long l0 = h->launches;
      const int erc = enqueue_frame(h, c);
      h->launches = l0;                           // capture does not execute anything
      cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      if (erc != TLOAM_B200_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        snprintf(h->last_error, sizeof(h->last_error), "graph capture failed: %s", cudaGetErrorString(ce));
        return TLOAM_B200_ERR_CUDA;
      }
      // same topology, new kernel parameters / grid sizes (cloud sizes change from frame to frame in a real
      // stream): update the instantiated graph in place, which is much cheaper than instantiating a new one
      bool updated = false;
      if (h->gexec) {
        cudaGraphExecUpdateResultInfo info;
        updated = cudaGraphExecUpdate(h->gexec, graph, &info) == cudaSuccess;
        if (!updated) { cudaGetLastError(); cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
      }
      if (!updated) ce = cudaGraphInstantiate(&h->gexec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "graph instantiate: %s", cudaGetErrorString(ce)); return TLOAM_B200_ERR_CUDA; }
      h->gctx = c;
      h->gvalid = true;
    }
    CU_TRY(cudaGraphLaunch(h->gexec, h->stream));
    h->launches += per_frame;
  } else {
    const int erc = enqueue_frame(h, c);
    if (erc != TLOAM_B200_OK) return erc;
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaEventRecord(h->ev1, h->stream));
  h->launches_frame = per_frame;
  h->traced_last = h->trace;
  h->frame_pending = true;
  return TLOAM_B200_OK;
}

int tloam_b200_get_result(tloam_b200_handle* h, double result[16], tloam_b200_stats* stats) {
  if (!h || !result) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->frame_pending) return TLOAM_B200_ERR_NOT_READY;
  CU_TRY(cudaSetDevice(h->device));
  if (stats && !h->traced_last) memset(h->h_stats, 0, sizeof(tloam_b200_stats));
  if (stats && h->traced_last) CU_TRY(cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(tloam_b200_stats), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->frame_pending = false;
  memcpy(result, h->h_result, 16 * sizeof(double));
  int flags[2];
  memcpy(flags, h->h_result + 16, sizeof(flags));   // frame_done, status
  if (stats) {
    *stats = *h->h_stats;
    stats->gpu_launches = h->launches_frame;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) stats->gpu_ms = ms;
  }
  if (flags[1] != TLOAM_B200_OK) return flags[1];
  if (!flags[0]) { snprintf(h->last_error, sizeof(h->last_error), "frame did not complete"); return TLOAM_B200_ERR_CUDA; }
  return TLOAM_B200_OK;
}

int tloam_b200_scan_match(tloam_b200_handle* h, const double predict[16], double result[16], tloam_b200_stats* stats) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  const bool saved = h->trace;
  if (stats) h->trace = true;                   // the blocking form knows whether the caller wants the trace
  int rc = tloam_b200_scan_match_async(h, predict);
  h->trace = saved;
  if (rc != TLOAM_B200_OK) return rc;
  return tloam_b200_get_result(h, result, stats);
}

int tloam_b200_set_trace(tloam_b200_handle* h, int on) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  h->trace = on != 0;
  return TLOAM_B200_OK;
}

int tloam_b200_synchronize(tloam_b200_handle* h) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

long long tloam_b200_launch_count(tloam_b200_handle* h) { return h ? h->launches : 0; }

static int read_state_pose(tloam_b200_handle* h, size_t offset, double out[16]) {
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaMemcpyAsync(h->h_result, (const char*)h->d_state + offset, 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  memcpy(out, h->h_result, 16 * sizeof(double));
  return TLOAM_B200_OK;
}

int tloam_b200_get_transform(tloam_b200_handle* h, double pose[16]) {   // ref: registration.cpp:370-372
  if (!h || !pose) return TLOAM_B200_ERR_INVALID_ARG;
  return read_state_pose(h, offsetof(FrameState, curr_pose), pose);
}

int tloam_b200_get_pose_increment(tloam_b200_handle* h, double pose[16]) {   // ref: registration.cpp:374-376
  if (!h || !pose) return TLOAM_B200_ERR_INVALID_ARG;
  double L[16], C[16];
  int rc = read_state_pose(h, offsetof(FrameState, last_pose), L);
  if (rc != TLOAM_B200_OK) return rc;
  rc = read_state_pose(h, offsetof(FrameState, curr_pose), C);
  if (rc != TLOAM_B200_OK) return rc;
  double Li[16];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Li[c * 4 + r] = L[r * 4 + c];
  for (int r = 0; r < 3; ++r) Li[12 + r] = -(Li[0 * 4 + r] * L[12] + Li[1 * 4 + r] * L[13] + Li[2 * 4 + r] * L[14]);
  Li[3] = Li[7] = Li[11] = 0.0; Li[15] = 1.0;
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s += Li[k * 4 + r] * C[c * 4 + k];
    pose[c * 4 + r] = s;
  }
  return TLOAM_B200_OK;
}

int tloam_b200_fitness(tloam_b200_handle* h, double* fitness, double* rmse) {
  if (!h || !fitness || !rmse) return TLOAM_B200_ERR_INVALID_ARG;
  *fitness = 0.0; *rmse = 0.0;
  if (!h->have_src || !h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (h->cfg.fitness_thres <= 0.0) return TLOAM_B200_OK;                 // ref: :258-261
  for (int c = 0; c < 4; ++c) if (h->cfg.fitness_thres > radius_of(h->cfg, c)) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  { const int rc = fetch_origin(h); if (rc != TLOAM_B200_OK) return rc; }
  const int nb = h->total_blocks;
  double* d_out = nullptr;
  CU_TRY(cudaMalloc(&d_out, (size_t)nb * 2 * sizeof(double)));
  k_fitness<<<nb, kBlk, 0, h->stream>>>(h->ctx, h->cfg.fitness_thres * h->cfg.fitness_thres, d_out);
  h->launches++;
  std::vector<double> out((size_t)nb * 2);
  cudaError_t e = cudaMemcpyAsync(out.data(), d_out, out.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_out);
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "fitness: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  for (int c = 0; c < 4; ++c) {
    double err = 0.0, cnt = 0.0;
    for (int b = h->ctx.blk_off[c]; b < h->ctx.blk_off[c + 1]; ++b) { err += out[2 * b]; cnt += out[2 * b + 1]; }
    if (cnt > 0.0) { *fitness += cnt / (double)h->n_src[c]; *rmse += sqrt(err / cnt); }   // ref: :278-284, 292-293
  }
  return TLOAM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// piecewise entry points (tests)
// ---------------------------------------------------------------------------------------------
int tloam_b200_knn(tloam_b200_handle* h, int cloud, const double* queries, size_t nq, double radius, int k, int* idx,
                   double* d2, int* count) {
  if (!h || cloud < 0 || cloud > 3 || !queries || !idx || !d2 || !count) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (!(radius > 0.0) || radius > h->hdr.cell[cloud] || (k != 1 && k != 3 && k != 5)) return TLOAM_B200_ERR_INVALID_ARG;
  if (nq == 0) return TLOAM_B200_OK;
  CU_TRY(cudaSetDevice(h->device));
  { const int rc = fetch_origin(h); if (rc != TLOAM_B200_OK) return rc; }
  double *dq = nullptr, *dd = nullptr; int *di = nullptr, *dc = nullptr;
  cudaError_t e = cudaMalloc(&dq, nq * 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dd, nq * k * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&di, nq * k * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&dc, nq * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(dq, queries, nq * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    const unsigned tb = 128, gb = (unsigned)((nq + tb - 1) / tb);
    const GridDesc g = h->ctx.grid[cloud];
    const double* o = h->ctx.origin;
    if (k == 1) k_knn<1><<<gb, tb, 0, h->stream>>>(g, o, dq, (unsigned)nq, radius * radius, di, dd, dc);
    else if (k == 3) k_knn<3><<<gb, tb, 0, h->stream>>>(g, o, dq, (unsigned)nq, radius * radius, di, dd, dc);
    else k_knn<5><<<gb, tb, 0, h->stream>>>(g, o, dq, (unsigned)nq, radius * radius, di, dd, dc);
    h->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(idx, di, nq * k * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d2, dd, nq * k * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(count, dc, nq * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(dq); cudaFree(dd); cudaFree(di); cudaFree(dc);
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "knn: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  return TLOAM_B200_OK;
}

int tloam_b200_build_factors(tloam_b200_handle* h, int cloud, const double x[6], int* valid, double* prim, size_t n) {
  if (!h || cloud < 0 || cloud > 3 || !x || !valid || !prim) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_src || !h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (n != h->n_src[cloud]) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  { const int rc = fetch_origin(h); if (rc != TLOAM_B200_OK) return rc; }
  Predict pr;
  memset(&pr, 0, sizeof(pr));
  memcpy(pr.m, x, 6 * sizeof(double));
  DeviceCtx c = h->ctx;
  c.factor_num = 4;   // build every cloud regardless of the configured subset
  k_set_pose<<<1, 256, 0, h->stream>>>(c, pr);
  k_correspond<<<h->total_blocks * 2, kBlk, 0, h->stream>>>(c);
  k_caps<<<h->total_blocks, kBlk, 0, h->stream>>>(c);
  h->launches += 3;
  CU_TRY(cudaGetLastError());
  std::vector<unsigned char> act(n);
  std::vector<double> col(n);
  const size_t off = (size_t)h->ctx.pad_off[cloud];
  CU_TRY(cudaMemcpyAsync(act.data(), h->ctx.active + off, n, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < n; ++i) valid[i] = act[i];
  for (int j = 0; j < 6; ++j) {
    CU_TRY(cudaMemcpyAsync(col.data(), h->ctx.prim[j] + off, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) prim[6 * i + j] = act[i] ? col[i] : 0.0;
  }
  // leave the handle idle
  int one = 1;
  CU_TRY(cudaMemcpyAsync((char*)h->d_state + offsetof(FrameState, frame_done), &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

static int run_functor(tloam_b200_handle* h, int type, const double x[6], size_t m, const double* p, const double* a,
                       const double* b, size_t b_stride, const double* w, double* r, size_t r_stride, double* J,
                       size_t j_stride, double* cost) {
  if (!h || !x || !p || !a || !b || !w || !r || !J || !cost) return TLOAM_B200_ERR_INVALID_ARG;
  if (m == 0) return TLOAM_B200_OK;
  CU_TRY(cudaSetDevice(h->device));
  Predict pr;
  memset(&pr, 0, sizeof(pr));
  memcpy(pr.m, x, 6 * sizeof(double));
  const size_t in_doubles = m * (3 + 3 + b_stride + 1), out_doubles = m * (r_stride + j_stride + 1);
  double* d = nullptr;
  cudaError_t e = cudaMalloc(&d, (in_doubles + out_doubles) * sizeof(double));
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "functor: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  double *dp = d, *da = dp + 3 * m, *db = da + 3 * m, *dw = db + b_stride * m, *dr = dw + m, *dJ = dr + r_stride * m, *dc = dJ + j_stride * m;
  e = cudaMemcpyAsync(dp, p, 3 * m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(da, a, 3 * m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(db, b, b_stride * m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dw, w, m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    k_functor<<<(unsigned)((m + 127) / 128), 128, 0, h->stream>>>(type, pr, (unsigned)m, dp, da, db, dw, dr, dJ, dc);
    h->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(r, dr, r_stride * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(J, dJ, j_stride * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cost, dc, m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "functor: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  return TLOAM_B200_OK;
}

int tloam_b200_eval_point_to_point(tloam_b200_handle* h, const double x[6], size_t m, const double* p, const double* q,
                                   const double* w, double* r, double* J, double* cost) {
  return run_functor(h, 0, x, m, p, q, q, 3, w, r, 3, J, 18, cost);
}
int tloam_b200_eval_point_to_line(tloam_b200_handle* h, const double x[6], size_t m, const double* p, const double* a,
                                  const double* b, const double* w, double* r, double* J, double* cost) {
  return run_functor(h, 1, x, m, p, a, b, 3, w, r, 3, J, 18, cost);
}
int tloam_b200_eval_point_to_plane(tloam_b200_handle* h, const double x[6], size_t m, const double* p, const double* n,
                                   const double* d, const double* w, double* r, double* J, double* cost) {
  return run_functor(h, 2, x, m, p, n, d, 1, w, r, 1, J, 6, cost);
}

static int run_se3(tloam_b200_handle* h, int op, const double* in, int nin, double* out, int nout, double* extra) {
  if (!h || !in || !out) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  Predict pr;
  memset(&pr, 0, sizeof(pr));
  memcpy(pr.m, in, nin * sizeof(double));
  double* d = nullptr;
  CU_TRY(cudaMalloc(&d, 32 * sizeof(double)));
  k_se3<<<1, 32, 0, h->stream>>>(op, pr, d);
  h->launches++;
  double tmp[32];
  cudaError_t e = cudaMemcpyAsync(tmp, d, 32 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "se3: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  memcpy(out, tmp, nout * sizeof(double));
  if (extra) *extra = tmp[6];
  return TLOAM_B200_OK;
}

int tloam_b200_se3_exp(tloam_b200_handle* h, const double a[6], double T[16]) { return run_se3(h, 0, a, 6, T, 16, nullptr); }
int tloam_b200_se3_log(tloam_b200_handle* h, const double T[16], double a[6]) {
  double ok = 1.0;
  int rc = run_se3(h, 1, T, 16, a, 6, &ok);
  if (rc != TLOAM_B200_OK) return rc;
  return ok != 0.0 ? TLOAM_B200_OK : TLOAM_B200_ERR_BAD_POSE;
}
int tloam_b200_se3_plus(tloam_b200_handle* h, const double x[6], const double delta[6], double out[6]) {
  double in[12];
  memcpy(in, x, 48); memcpy(in + 6, delta, 48);
  return run_se3(h, 2, in, 12, out, 6, nullptr);
}

// ---------------------------------------------------------------------------------------------
// (f)-1 device-side submap maintenance
// ---------------------------------------------------------------------------------------------
void tloam_b200_submap_default_config(tloam_submap_config* c) {   // ref: config/mapping/lidar_odometry.yaml:6-17
  c->ground_down_sample = 0.3; c->ground_down_sample_submap = 0.45; c->edge_down_sample_submap = 0.3;
  c->planar_frame_size = 3; c->sphere_frame_size = 3;
  c->edge_crop_box_length = 100.0; c->ground_crop_box_length = 100.0;
}

static int ensure_dev(tloam_b200_handle* h, double** p, size_t* cap, size_t need_points, bool keep) {
  if (need_points <= *cap) return TLOAM_B200_OK;
  const size_t ncap = need_points + need_points / 2 + 1024;
  double* q = nullptr;
  CU_TRY(cudaMalloc(&q, ncap * 3 * sizeof(double)));
  if (keep && *p && *cap) CU_TRY(cudaMemcpyAsync(q, *p, *cap * 3 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (*p) { CU_TRY(cudaStreamSynchronize(h->stream)); cudaFree(*p); }
  *p = q; *cap = ncap;
  return TLOAM_B200_OK;
}

// crop (inclusive box, lo > hi = no crop) + VoxelDownSample of d_in[0..n) into d_out; returns the voxel count.
static int voxel_pipeline(tloam_b200_handle* h, const double* d_in, size_t n, const double* lo, const double* hi, double voxel,
                          double* d_out, size_t* n_out) {
  *n_out = 0;
  if (n == 0) return TLOAM_B200_OK;
  if (!(voxel > 0.0)) return TLOAM_B200_ERR_INVALID_ARG;
  const unsigned tsize = next_pow2(2 * n + 1);
  const size_t bytes = 256 + (size_t)tsize * (8 + 24 + 4);
  if (bytes > h->cap_vox) {
    cudaFree(h->d_vox);
    h->cap_vox = bytes + bytes / 2;
    CU_TRY(cudaMalloc(&h->d_vox, h->cap_vox));
  }
  VoxArgs a;
  a.in = d_in; a.n = (unsigned)n; a.voxel = voxel;
  for (int d = 0; d < 3; ++d) { a.lo[d] = lo ? lo[d] : -DBL_MAX; a.hi[d] = hi ? hi[d] : DBL_MAX; }
  a.minenc = reinterpret_cast<unsigned long long*>(h->d_vox);            // [0..2] min, [3] out_count
  a.out_count = reinterpret_cast<unsigned*>(h->d_vox + 32);
  a.keys = reinterpret_cast<unsigned long long*>(h->d_vox + 256);
  a.sums = reinterpret_cast<long long*>(h->d_vox + 256 + (size_t)tsize * 8);
  a.cnt = reinterpret_cast<unsigned*>(h->d_vox + 256 + (size_t)tsize * 32);
  a.mask = tsize - 1u;
  a.out = d_out;
  CU_TRY(cudaMemsetAsync(h->d_vox, 0xFF, 24, h->stream));               // min encodings = +max
  CU_TRY(cudaMemsetAsync(h->d_vox + 24, 0, 256 - 24 + (size_t)tsize * 36, h->stream));
  const unsigned tb = 256, gb = (unsigned)((n + tb - 1) / tb);
  TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_vox_min<<<(gb < 592u ? gb : 592u), tb, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_vox_accum<<<gb, tb, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_vox_emit<<<(tsize + tb - 1) / tb, tb, 0, h->stream>>>(a)));
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(h->h_result + 28, a.out_count, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  unsigned cnt;
  memcpy(&cnt, h->h_result + 28, sizeof(cnt));
  *n_out = cnt;
  return TLOAM_B200_OK;
}

static int upload_points(tloam_b200_handle* h, const double* host, size_t n) {
  int rc = ensure_dev(h, &h->d_up, &h->cap_up, n, false);
  if (rc != TLOAM_B200_OK) return rc;
  if (n) CU_TRY(cudaMemcpyAsync(h->d_up, host, n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return TLOAM_B200_OK;
}

static int submap_set_target(tloam_b200_handle* h) {
  // order at the ABI: edge, sphere, planar, ground.  After the first update the sphere map IS the planar window
  // (ref: front_end.cpp:220-230 iterates submap_planar_buffer).
  const double* xyz[4] = {h->d_acc[0], h->sphere_is_init ? h->d_sphere0 : h->d_cat, h->d_cat, h->d_acc[1]};
  const size_t n[4] = {h->n_acc[0], h->sphere_is_init ? h->n_sphere0 : h->n_cat, h->n_cat, h->n_acc[1]};
  return set_target_impl(h, xyz, n, true);
}

// ---------------------------------------------------------------------------------------------
// "next" row (f)-2: PCA feature extraction (feature_extract.cuh)
// ---------------------------------------------------------------------------------------------
void tloam_b200_feature_default_config(tloam_feature_config* c) {   // ref: config/mapping/feature.yaml
  c->radius = 0.2; c->K = 20; c->min_neigh = 10; c->planar_num = 500; c->sphere_num = 300;
  c->cvr_scan = 0.25; c->cvr_submap = 0.15; c->planar_scan_thres = 0.75; c->planar_submap_thres = 0.65;
  c->planar_vertic_thres = 0.25;
}

namespace {
struct FeArena {
  double* stage; unsigned* scratch; unsigned char* blob; MapHeader hdr;
  FeOut out;
  unsigned long long *key_p, *key_s, *key_p_sorted, *key_s_sorted;
  unsigned *val, *val_p_sorted, *val_s_sorted, *counts;
  void* cub_tmp; size_t cub_bytes;
  unsigned *host_val_p = nullptr, *host_val_s = nullptr;   // optional: the sorted index lists land here (one sync)
};
}  // namespace

// carves the arena for n points and enqueues grid build + k_fe_pca (+ classification and sorts when `select`)
static int fe_run(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n, bool select, FeArena& A) {
  if (n == 0 || n > ((size_t)1 << 30)) return TLOAM_B200_ERR_INVALID_ARG;
  if (!(cfg->radius > 0.0) || cfg->K < 3 || cfg->K > kFeK) return TLOAM_B200_ERR_INVALID_ARG;   // :56 asserts K >= 3
  CU_TRY(cudaSetDevice(h->device));
  MapHeader& hd = A.hdr;
  memset(&hd, 0, sizeof(hd));
  hd.magic = kMapMagic;
  size_t boff = sizeof(MapHeader);
  for (int c = 0; c < 4; ++c) {
    hd.n[c] = c == 0 ? (unsigned)n : 0u;
    hd.tsize[c] = next_pow2((c == 0 ? n : 0) + 1);
    hd.cell[c] = cfg->radius;                       // cell edge == search radius: the 27-cell search is exact
    hd.pts_off[c] = boff; boff += round_up(hd.n[c] * sizeof(FePoint), 256);
  }
  for (int c = 0; c < 4; ++c) { hd.table_off[c] = boff; boff += (size_t)hd.tsize[c] * kBrickBytes; }
  for (int d = 0; d < 3; ++d) { hd.bbox_enc[d] = ~0ull; hd.bbox_enc[3 + d] = 0ull; }
  A.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, A.cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                            (const unsigned*)nullptr, (unsigned*)nullptr, (int)n, 0, 64, h->stream);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_stage = take(n * 3 * sizeof(double)), o_scr = take(n * 2 * sizeof(unsigned)), o_blob = take(boff);
  const size_t o_cvr = take(n * 8), o_flat = take(n * 8), o_sph = take(n * 8), o_nrm = take(n * 24), o_num = take(n * 4),
               o_nei = take(n * kFeK * 4);
  const size_t o_kp = take(n * 8), o_ks = take(n * 8), o_kps = take(n * 8), o_kss = take(n * 8), o_val = take(n * 4),
               o_vps = take(n * 4), o_vss = take(n * 4), o_cnt = take(256), o_cub = take(A.cub_bytes);
  if (off > h->cap_fe) {
    cudaFree(h->d_fe);
    h->cap_fe = off + off / 4;
    CU_TRY(cudaMalloc(&h->d_fe, h->cap_fe));
  }
  unsigned char* b = h->d_fe;
  A.stage = (double*)(b + o_stage); A.scratch = (unsigned*)(b + o_scr); A.blob = b + o_blob;
  A.out.cvr = (double*)(b + o_cvr); A.out.flatness = (double*)(b + o_flat); A.out.sphericity = (double*)(b + o_sph);
  A.out.normal = (double*)(b + o_nrm); A.out.num_sum = (int*)(b + o_num); A.out.neigh = (int*)(b + o_nei);
  A.key_p = (unsigned long long*)(b + o_kp); A.key_s = (unsigned long long*)(b + o_ks);
  A.key_p_sorted = (unsigned long long*)(b + o_kps); A.key_s_sorted = (unsigned long long*)(b + o_kss);
  A.val = (unsigned*)(b + o_val); A.val_p_sorted = (unsigned*)(b + o_vps); A.val_s_sorted = (unsigned*)(b + o_vss);
  A.counts = (unsigned*)(b + o_cnt); A.cub_tmp = b + o_cub;

  CU_TRY(cudaMemcpyAsync(A.stage, xyz, n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(A.blob, &hd, sizeof(MapHeader), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemsetAsync(A.blob + hd.table_off[0], 0, boff - hd.table_off[0], h->stream));
  CU_TRY(cudaMemsetAsync(A.counts, 0, 256, h->stream));
  MapBuildArgs ma;
  ma.blob = A.blob; ma.slot_of = A.scratch; ma.rank_of = A.scratch + n;
  ma.i_beg = 0; ma.i_end = (unsigned)n; ma.only_cloud = -1;
  for (int c = 0; c < 4; ++c) ma.src[c] = A.stage;
  ma.stage_off[0] = 0;
  for (int c = 1; c <= 4; ++c) ma.stage_off[c] = (unsigned)n;
  FeBuildArgs fa;
  fa.stage = A.stage; fa.n = (unsigned)n; fa.blob = A.blob; fa.slot_of = A.scratch; fa.rank_of = A.scratch + n;
  const unsigned tb = 256, gb = (unsigned)((n + tb - 1) / tb);
  unsigned tslots = 0;
  for (int c = 0; c < 4; ++c) tslots += hd.tsize[c];
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_map_bbox<<<(gb < 592u ? gb : 592u), tb, 0, h->stream>>>(ma)));   // + origin
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_insert<<<gb, tb, 0, h->stream>>>(fa)));
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_map_offsets<<<(tslots + tb - 1) / tb, tb, 0, h->stream>>>(ma)));
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_scatter<<<gb, tb, 0, h->stream>>>(fa)));
  FeGrid g;
  g.pts = reinterpret_cast<const FePoint*>(A.blob + hd.pts_off[0]);
  g.table = reinterpret_cast<const uint4*>(A.blob + hd.table_off[0]);
  g.mask = hd.tsize[0] - 1u; g.n = (unsigned)n; g.cell = cfg->radius; g.inv_cell = 1.0 / cfg->radius;
  g.origin = reinterpret_cast<const double*>(A.blob + offsetof(MapHeader, origin));
  FeParams prm;
  prm.dbg = h->profiling ? h->d_dbg + 8 : nullptr;     // slots 8..11 (k_correspond uses them in frame profiling)
  prm.r2 = cfg->radius * cfg->radius; prm.K = cfg->K; prm.min_neigh = cfg->min_neigh;
  prm.cvr_submap = cfg->cvr_submap; prm.planar_submap_thres = cfg->planar_submap_thres;
  prm.planar_vertic_thres = cfg->planar_vertic_thres;
  if (!h->fe_attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_fe_pca, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFeSmemBytes));
    h->fe_attr_set = true;
  }
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_pca<<<(unsigned)((n + kFeBlk - 1) / kFeBlk), kFeBlk, kFeSmemBytes, h->stream>>>(g, prm, A.out)));
  if (select) {
    TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_classify<<<gb, tb, 0, h->stream>>>((unsigned)n, prm, A.out, A.key_p, A.key_s, A.val, A.counts)));
    // stable descending sorts: candidates first (by flatness, ties in ascending point index), the rest (key 0) last
    size_t tmp = A.cub_bytes;
    CU_TRY(cub::DeviceRadixSort::SortPairsDescending(A.cub_tmp, tmp, A.key_p, A.key_p_sorted, A.val, A.val_p_sorted, (int)n, 0, 64, h->stream));
    tmp = A.cub_bytes;
    CU_TRY(cub::DeviceRadixSort::SortPairsDescending(A.cub_tmp, tmp, A.key_s, A.key_s_sorted, A.val, A.val_s_sorted, (int)n, 0, 64, h->stream));
    h->launches += 2;
    TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_counts<<<1, 32, 0, h->stream>>>(A.key_p_sorted, A.key_s_sorted, A.counts, cfg->planar_num,
                                                                           cfg->sphere_num, cfg->planar_scan_thres, cfg->cvr_scan)));
  }
  CU_TRY(cudaGetLastError());
  // counts + build flags -> pinned scratch (h_result[28..31])
  CU_TRY(cudaMemcpyAsync(h->h_result + 28, A.counts, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaMemcpyAsync(h->h_result + 30, A.blob + offsetof(MapHeader, build_flags), sizeof(unsigned long long),
                         cudaMemcpyDeviceToHost, h->stream));
  if (select && A.host_val_p) CU_TRY(cudaMemcpyAsync(A.host_val_p, A.val_p_sorted, n * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  if (select && A.host_val_s) CU_TRY(cudaMemcpyAsync(A.host_val_s, A.val_s_sorted, n * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  unsigned long long flags;
  memcpy(&flags, h->h_result + 30, sizeof(flags));
  return (flags & 1ull) ? TLOAM_B200_ERR_MAP_DENSITY : TLOAM_B200_OK;
}

int tloam_b200_extract_planar_sphere(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n,
                                     size_t* planar_scan_index, size_t* n_planar_scan, size_t* planar_submap_index,
                                     size_t* n_planar_submap, size_t* sphere_scan_index, size_t* n_sphere_scan,
                                     size_t* sphere_submap_index, size_t* n_sphere_submap, size_t* sphere_candidates) {
  if (!h || !cfg || !planar_scan_index || !n_planar_scan || !planar_submap_index || !n_planar_submap || !sphere_scan_index ||
      !n_sphere_scan || !sphere_submap_index || !n_sphere_submap)
    return TLOAM_B200_ERR_INVALID_ARG;
  *n_planar_scan = *n_planar_submap = *n_sphere_scan = *n_sphere_submap = 0;
  if (n == 0) return TLOAM_B200_OK;              // calculatePCAInfo fails on an empty cloud: nothing is selected (:49-54, :141)
  if (!xyz) return TLOAM_B200_ERR_INVALID_ARG;
  FeArena A;
  std::vector<unsigned> vp(n), vs(n);            // sorted point indices (candidates first), fetched with the counts
  A.host_val_p = vp.data(); A.host_val_s = vs.data();
  const int rc = fe_run(h, cfg, xyz, n, true, A);
  if (rc != TLOAM_B200_OK) return rc;
  unsigned counts[4];
  memcpy(counts, h->h_result + 28, sizeof(counts));
  const unsigned np = counts[0], ns = counts[1], nps = counts[2], nss = counts[3];
  for (unsigned i = 0; i < np; ++i) planar_submap_index[i] = vp[i];                    // :180
  for (unsigned i = 0; i < nps; ++i) planar_scan_index[i] = vp[i];                     // :177-178
  *n_planar_submap = np; *n_planar_scan = nps;
  for (unsigned i = 0; i < ns; ++i) sphere_submap_index[i] = i;                          // :187 (rank, not index)
  for (unsigned i = 0; i < nss; ++i) sphere_scan_index[i] = i;                           // :184-185
  *n_sphere_submap = ns; *n_sphere_scan = nss;
  if (sphere_candidates) for (unsigned i = 0; i < ns; ++i) sphere_candidates[i] = vs[i];
  return TLOAM_B200_OK;
}

int tloam_b200_pca_info(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n, double* cvr,
                        double* flatness, double* sphericity, double* normal, int* num_sum, int* neigh) {
  if (!h || !cfg || !xyz || n == 0) return TLOAM_B200_ERR_INVALID_ARG;
  FeArena A;
  const int rc = fe_run(h, cfg, xyz, n, false, A);
  if (rc != TLOAM_B200_OK) return rc;
  if (cvr) CU_TRY(cudaMemcpyAsync(cvr, A.out.cvr, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (flatness) CU_TRY(cudaMemcpyAsync(flatness, A.out.flatness, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (sphericity) CU_TRY(cudaMemcpyAsync(sphericity, A.out.sphericity, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (normal) CU_TRY(cudaMemcpyAsync(normal, A.out.normal, n * 24, cudaMemcpyDeviceToHost, h->stream));
  if (num_sum) CU_TRY(cudaMemcpyAsync(num_sum, A.out.num_sum, n * 4, cudaMemcpyDeviceToHost, h->stream));
  std::vector<int> nb;
  if (neigh) {
    nb.resize(n * kFeK);
    CU_TRY(cudaMemcpyAsync(nb.data(), A.out.neigh, n * kFeK * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (neigh)
    for (size_t i = 0; i < n; ++i)
      for (int j = 0; j < cfg->K; ++j) neigh[i * cfg->K + j] = nb[i * kFeK + j];
  return TLOAM_B200_OK;
}

int tloam_b200_voxel_down_sample(tloam_b200_handle* h, const double* pts, size_t n, double voxel, double* out, size_t* n_out) {
  if (!h || (!pts && n) || !out || !n_out) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  int rc = upload_points(h, pts, n);
  if (rc != TLOAM_B200_OK) return rc;
  rc = ensure_dev(h, &h->d_acc_tmp, &h->cap_acc_tmp, n, false);
  if (rc != TLOAM_B200_OK) return rc;
  rc = voxel_pipeline(h, h->d_up, n, nullptr, nullptr, voxel, h->d_acc_tmp, n_out);
  if (rc != TLOAM_B200_OK) return rc;
  if (*n_out) CU_TRY(cudaMemcpyAsync(out, h->d_acc_tmp, *n_out * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

int tloam_b200_submap_init(tloam_b200_handle* h, const tloam_submap_config* cfg, const double* edge, size_t ne,
                           const double* ground_raw, size_t ng, const double* planar_sub, size_t np,
                           const double* sphere_sub, size_t ns) {
  if (!h || !cfg || (!edge && ne) || (!ground_raw && ng) || (!planar_sub && np) || (!sphere_sub && ns)) return TLOAM_B200_ERR_INVALID_ARG;
  if (cfg->planar_frame_size < 1 || cfg->planar_frame_size > 64) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  h->scfg = *cfg;
  if (!h->d_pose) CU_TRY(cudaMalloc(&h->d_pose, 16 * sizeof(double)));
  int rc;
  // edge: raw copy (front_end.cpp:286)
  if ((rc = ensure_dev(h, &h->d_acc[0], &h->cap_acc[0], ne, false)) != TLOAM_B200_OK) return rc;
  if (ne) CU_TRY(cudaMemcpyAsync(h->d_acc[0], edge, ne * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->n_acc[0] = ne;
  // ground: VoxelDownSample(ground_down_sample) (:287)
  if ((rc = upload_points(h, ground_raw, ng)) != TLOAM_B200_OK) return rc;
  if ((rc = ensure_dev(h, &h->d_acc[1], &h->cap_acc[1], ng, false)) != TLOAM_B200_OK) return rc;
  if ((rc = voxel_pipeline(h, h->d_up, ng, nullptr, nullptr, cfg->ground_down_sample, h->d_acc[1], &h->n_acc[1])) != TLOAM_B200_OK) return rc;
  // planar / sphere: the submap-index selections (:291-292); the sliding-window buffers stay empty (:285-305)
  if ((rc = ensure_dev(h, &h->d_cat, &h->cap_cat, np, false)) != TLOAM_B200_OK) return rc;
  if (np) CU_TRY(cudaMemcpyAsync(h->d_cat, planar_sub, np * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->n_cat = np;
  cudaFree(h->d_sphere0); h->d_sphere0 = nullptr;
  if (ns) {
    CU_TRY(cudaMalloc(&h->d_sphere0, ns * 3 * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(h->d_sphere0, sphere_sub, ns * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  h->n_sphere0 = ns; h->sphere_is_init = true;
  for (double* p : h->ring) cudaFree(p);
  h->ring.clear(); h->ring_n.clear(); h->ring_cap.clear();
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->submap_ready = true;
  return submap_set_target(h);
}

int tloam_b200_submap_update(tloam_b200_handle* h, const double pose[16], const double* planar_sub, size_t np,
                             const double* sphere_sub, size_t ns) {
  (void)sphere_sub; (void)ns;   // stored but never used by the reference (front_end.cpp:202-205, 220-230)
  if (!h || !pose || (!planar_sub && np)) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->submap_ready || !h->have_src) return TLOAM_B200_ERR_NOT_READY;
  CU_TRY(cudaSetDevice(h->device));
  const tloam_submap_config& cf = h->scfg;
  int rc;
  memcpy(h->h_predict->m, pose, 16 * sizeof(double));
  CU_TRY(cudaMemcpyAsync(h->d_pose, h->h_predict, 16 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const unsigned tb = 256;
  // ---- planar sliding window (:207-217, 232-242): newest frame transformed by its pose ----
  if ((rc = upload_points(h, planar_sub, np)) != TLOAM_B200_OK) return rc;
  double* slot = nullptr; size_t slot_cap = 0;
  if ((int)h->ring.size() >= cf.planar_frame_size) {            // recycle the oldest buffer
    slot = h->ring.front(); slot_cap = h->ring_cap.front();
    h->ring.erase(h->ring.begin()); h->ring_n.erase(h->ring_n.begin()); h->ring_cap.erase(h->ring_cap.begin());
  }
  if ((rc = ensure_dev(h, &slot, &slot_cap, np, false)) != TLOAM_B200_OK) return rc;
  if (np) TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_transform_append<<<(unsigned)((np + tb - 1) / tb), tb, 0, h->stream>>>(h->d_up, (unsigned)np, slot, h->d_pose)));
  h->ring.push_back(slot); h->ring_n.push_back(np); h->ring_cap.push_back(slot_cap);
  size_t tot = 0;
  for (size_t k : h->ring_n) tot += k;
  if ((rc = ensure_dev(h, &h->d_cat, &h->cap_cat, tot, false)) != TLOAM_B200_OK) return rc;
  size_t off = 0;
  for (size_t f = 0; f < h->ring.size(); ++f) {
    if (h->ring_n[f]) CU_TRY(cudaMemcpyAsync(h->d_cat + 3 * off, h->ring[f], h->ring_n[f] * 3 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    off += h->ring_n[f];
  }
  h->n_cat = tot;
  h->sphere_is_init = false;
  // ---- edge / ground: append the current source features in the world frame (:245-246), crop (:248-264),
  //      VoxelDownSample ----
  const int src_cloud[2] = {0, 3};
  const double vox[2] = {cf.edge_down_sample_submap, cf.ground_down_sample_submap};
  const double len[2] = {cf.edge_crop_box_length, cf.ground_crop_box_length};
  size_t soff[4], o = 0;
  for (int c = 0; c < 4; ++c) { soff[c] = o; o += h->n_src[c]; }
  for (int k = 0; k < 2; ++k) {
    const size_t nadd = h->n_src[src_cloud[k]];
    if ((rc = ensure_dev(h, &h->d_acc[k], &h->cap_acc[k], h->n_acc[k] + nadd, true)) != TLOAM_B200_OK) return rc;
    if (nadd) TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_transform_append<<<(unsigned)((nadd + tb - 1) / tb), tb, 0, h->stream>>>(
        h->d_stage_src + 3 * soff[src_cloud[k]], (unsigned)nadd, h->d_acc[k] + 3 * h->n_acc[k], h->d_pose)));
    const size_t nall = h->n_acc[k] + nadd;
    const double lo[3] = {pose[12] - len[k], pose[13] - len[k], pose[14] - len[k]};
    const double hi[3] = {pose[12] + len[k], pose[13] + len[k], pose[14] + len[k]};
    if ((rc = ensure_dev(h, &h->d_acc_tmp, &h->cap_acc_tmp, nall, false)) != TLOAM_B200_OK) return rc;
    size_t nout = 0;
    if ((rc = voxel_pipeline(h, h->d_acc[k], nall, lo, hi, vox[k], h->d_acc_tmp, &nout)) != TLOAM_B200_OK) return rc;
    // the down-sampled cloud becomes the accumulator (swap buffers)
    double* t = h->d_acc[k]; h->d_acc[k] = h->d_acc_tmp; h->d_acc_tmp = t;
    size_t tc = h->cap_acc[k]; h->cap_acc[k] = h->cap_acc_tmp; h->cap_acc_tmp = tc;
    h->n_acc[k] = nout;
  }
  return submap_set_target(h);                                   // :267
}

int tloam_b200_submap_sizes(tloam_b200_handle* h, size_t n[4]) {
  if (!h || !n) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->submap_ready) return TLOAM_B200_ERR_NOT_READY;
  n[0] = h->n_acc[0]; n[1] = h->sphere_is_init ? h->n_sphere0 : h->n_cat; n[2] = h->n_cat; n[3] = h->n_acc[1];
  return TLOAM_B200_OK;
}

int tloam_b200_submap_download(tloam_b200_handle* h, int cloud, double* out, size_t capacity_points) {
  if (!h || !out || cloud < 0 || cloud > 3) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->submap_ready) return TLOAM_B200_ERR_NOT_READY;
  size_t n[4];
  tloam_b200_submap_sizes(h, n);
  if (capacity_points < n[cloud]) return TLOAM_B200_ERR_INVALID_ARG;
  const double* src = cloud == 0 ? h->d_acc[0] : cloud == 3 ? h->d_acc[1] : (cloud == 1 && h->sphere_is_init) ? h->d_sphere0 : h->d_cat;
  CU_TRY(cudaSetDevice(h->device));
  if (n[cloud]) CU_TRY(cudaMemcpyAsync(out, src, n[cloud] * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

int tloam_b200_host_alloc(void** p, size_t bytes) {
  if (!p) return TLOAM_B200_ERR_INVALID_ARG;
  return cudaMallocHost(p, bytes) == cudaSuccess ? TLOAM_B200_OK : TLOAM_B200_ERR_CUDA;
}
int tloam_b200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? TLOAM_B200_OK : TLOAM_B200_ERR_CUDA; }

const char* tloam_b200_last_error(tloam_b200_handle* h) { return h ? h->last_error : ""; }

int tloam_b200_set_profiling(tloam_b200_handle* h, int on) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->profiling = on != 0;
  h->spans.clear(); h->ev_next = 0;
  memset(&h->prof, 0, sizeof(h->prof));
  if (on && !h->d_dbg) CU_TRY(cudaMalloc(&h->d_dbg, 16 * sizeof(unsigned long long)));
  if (h->d_dbg) {
    unsigned long long init[16];
    memset(init, 0, sizeof(init));
    init[0] = ~0ull;
    CU_TRY(cudaMemcpy(h->d_dbg, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  h->ctx.dbg = on ? h->d_dbg : nullptr;
  return TLOAM_B200_OK;
}

int tloam_b200_get_profile(tloam_b200_handle* h, tloam_b200_profile* out) {
  if (!h || !out) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  for (const auto& sp : h->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.cls >= 0 && sp.cls < TLOAM_B200_K_COUNT) {
      h->prof.launches[sp.cls] += 1;
      h->prof.total_ms[sp.cls] += ms;
    }
  }
  h->spans.clear(); h->ev_next = 0;
  if (h->d_dbg) CU_TRY(cudaMemcpy(h->prof.dbg, h->d_dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  *out = h->prof;
  return TLOAM_B200_OK;
}

}  // extern "C"
