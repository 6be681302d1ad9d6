// frame_kernels.cuh -- per-frame kernels of the registration path (see registration.cuh for the execution model).
//
// Every frame kernel exists in two instantiations of ONE body:
//   kBatched = false   one sequence; its DeviceCtx travels as a __grid_constant__ kernel parameter;
//   kBatched = true    S sequences in ONE launch (tloam_b200_batch_*): the grid is the concatenation of the
//                      per-sequence grids (BatchTab::off), a block copies the DeviceCtx of its sequence from a device
//                      array into shared memory and runs the same body.  The per-sequence grid (and therefore the
//                      reduction tree) is the same function of the cloud sizes in both modes, so a sequence
//                      registered in a batch gets bit-identical poses to the same sequence registered alone.
// Thread-block clusters never straddle sequences (per-sequence grids are multiples of the cluster size), tickets
// and partial rows are per sequence, and the LAST cluster leader of a sequence runs that sequence's solver --
// S solver tails run concurrently on S SMs instead of one after the other.
#pragma once
#include "registration.cuh"
#include "solver.cuh"

namespace tloam {

struct Predict { double m[16]; double from_state; };   // from_state != 0: constant-velocity prediction on the device

constexpr int kMaxBatch = 32;
struct BatchTab {
  const DeviceCtx* ctxs;        // device array [S]
  int S;
  int off[kMaxBatch + 1];       // first block of each sequence in THIS launch's grid
};

// resolves the block's sequence: ctx (parameter or shared-memory copy), block index inside the sequence, number of
// blocks of the sequence
#define TL_RESOLVE_CTX(one, tab)                                                                    \
  __shared__ DeviceCtx s_ctx__;                                                                     \
  int lb = (int)blockIdx.x, nblk = (int)gridDim.x;                                                  \
  if (kBatched) {                                                                                   \
    int s__ = 0;                                                                                    \
    while (s__ + 1 < (tab).S && (int)blockIdx.x >= (tab).off[s__ + 1]) ++s__;                      \
    lb = (int)blockIdx.x - (tab).off[s__];                                                          \
    nblk = (tab).off[s__ + 1] - (tab).off[s__];                                                     \
    const unsigned* src__ = reinterpret_cast<const unsigned*>((tab).ctxs + s__);                    \
    unsigned* dst__ = reinterpret_cast<unsigned*>(&s_ctx__);                                        \
    for (unsigned i__ = threadIdx.x; i__ < sizeof(DeviceCtx) / 4; i__ += blockDim.x) dst__[i__] = __ldg(src__ + i__); \
    __syncthreads();                                                                                \
  }                                                                                                 \
  const DeviceCtx& ctx = kBatched ? s_ctx__ : (one);
static_assert(sizeof(DeviceCtx) % 4 == 0, "DeviceCtx is copied in 4-byte words");

// thread-block cluster barrier, split in its two halves (all threads of the block execute both, convergently)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
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

__device__ __forceinline__ void begin_frame_body(const DeviceCtx& ctx, const Predict* prp) {
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
  for (int c = 0; c < 4; ++c) st->map_bricks[c] = ctx.map_bricks[c];   // map statistics ride home with the result
  for (int i = 0; i < 16; ++i) st->last_pose[i] = st->curr_pose[i];            // :882
  Pose7 p;
  if (*ctx.map_flags & 1ull) {                   // a map cell overflowed its u16 counter at build time
    st->status = TLOAM_B200_ERR_MAP_DENSITY;
    for (int i = 0; i < 16; ++i) st->result[i] = pr.m[i];
    st->frame_done = 1;
    return;
  }
  for (int c = 0; c < 4; ++c)
    if (ctx.tgt_cnt[c] && *ctx.tgt_cnt[c] < 10u) {          // ref: :928-929, for map clouds counted on the device
      st->status = TLOAM_B200_ERR_TOO_FEW_POINTS;
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

// single: <<<1, 256>>>; batched: <<<S, 256>>>, block s = sequence s
template <bool kBatched>
__global__ void k_begin_frame(const __grid_constant__ DeviceCtx one, const __grid_constant__ BatchTab tab, const Predict* prp) {
  if (kBatched) {
    __shared__ DeviceCtx s_ctx;
    const unsigned* src = reinterpret_cast<const unsigned*>(tab.ctxs + blockIdx.x);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_ctx);
    for (unsigned i = threadIdx.x; i < sizeof(DeviceCtx) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    begin_frame_body(s_ctx, prp + blockIdx.x);
  } else {
    begin_frame_body(one, prp);
  }
}

// Primitive fit + validity tests of one feature given its k nearest neighbours (world coordinates, ascending
// (d2, index); rows >= k are zero) and the squared distance of the nearest one.
// ref: registration.cpp:445-493 (edge), 536-551 (sphere), 589-625 (planar), 732-768 (ground).
template <int K>
__device__ __forceinline__ unsigned char fit_neighbours(const DeviceCtx& ctx, int c, int k, double d2_first, const double (*nb)[3],
                                                        double prim[6]) {
#pragma unroll
  for (int j = 0; j < 6; ++j) prim[j] = 0.0;
  if (K == 1) {                                            // sphere
    if (k <= 0) return kFlagCounted;                       // not found: sphere_sum++ (:551)
    if (d2_first > 0.2) return 0;                          // squared distance vs 0.2, `continue` (:536)
    prim[0] = nb[0][0]; prim[1] = nb[0][1]; prim[2] = nb[0][2];
    return kFlagCand | kFlagCounted;
  }
  if (k <= 0) return 0;
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

// the same, with the neighbours re-loaded from the map by their position in the cell-sorted array
template <int K>
__device__ __forceinline__ unsigned char fit_one(const DeviceCtx& ctx, int c, const TopK<K>& t, double prim[6]) {
  const GridDesc& g = ctx.grid[c];
  const double o0 = ctx.origin[0], o1 = ctx.origin[1], o2 = ctx.origin[2];
  const int k = t.count();
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
  return fit_neighbours<K>(ctx, c, k, t.d2[0], nb, prim);
}

// shared memory of the lane-pair search (per-thread cell lists)
struct SearchSmem {
  unsigned beg[kPairCells][kBlk];
  unsigned cnt[kPairCells][kBlk];
  float md[kPairCells][kBlk];
#if TLOAM_SEARCH_QUEUE
  double q_d[kSearchQueue][kBlk];            // candidate queue of knn_search_pair (K = 5)
  int q_i[kSearchQueue][kBlk];
  int q_p[kSearchQueue][kBlk];
#endif
};

// Correspondence search + primitive fit of the feature served by this lane pair (64 features per 128-thread block:
// pair q of block-half `sb` of feature block `fb`) and the lazy GNC weight update of the previous outer iteration
// (ref: registration.cpp:858-876).  On return the EVEN lane holds flag / prim / w of its feature (gi).
// Measured alternatives: DESIGN.md section 4.
__device__ __forceinline__ void search_and_fit(const DeviceCtx& ctx, const FrameState* st, SearchSmem* sm, int fb, int sb,
                                               int& c_out, int& gi_out, bool& live_out, unsigned char& flag, double prim[6],
                                               double& w_out) {
  constexpr int kQ = kBlk / 2;                 // features per thread block
  const bool dead = fb >= ctx.blk_off[4];      // padding block of a cluster-rounded grid
  const int c = dead ? 3 : cloud_of_block(ctx, fb);
  const int il = dead ? 0 : (fb - ctx.blk_off[c]) * kBlk + sb * kQ + (int)(threadIdx.x >> 1);
  const int gi = ctx.pad_off[c] + il;
  const bool live = !dead && (il < ctx.n[c]) && cloud_enabled(ctx, c);
  const bool even = (threadIdx.x & 1) == 0;
  long long tk0 = 0, tk1 = 0;
  if (ctx.dbg) tk0 = clock64();
  double rx = 0.0, ry = 0.0, rz = 0.0;
  if (live) {
    const Rt T = pose_to_rt(st->xq);
    double qx, qy, qz;
    rt_apply(T, ctx.px[gi], ctx.py[gi], ctx.pz[gi], qx, qy, qz);
    rx = qx - ctx.origin[0]; ry = qy - ctx.origin[1]; rz = qz - ctx.origin[2];
  }
  flag = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) prim[j] = 0.0;
  if (c == kSphere) {
    TopK<1> t;
    knn_search_pair<1, kBlk>(ctx.grid[c], live, rx, ry, rz, ctx.r2[c], sm->beg, sm->cnt, sm->md, t);
    if (ctx.dbg) tk1 = clock64();
    if (live && even) flag = fit_one<1>(ctx, c, t, prim);
  } else {
    TopK<5> t;
#if TLOAM_SEARCH_QUEUE
    knn_search_pair<5, kBlk>(ctx.grid[c], live, rx, ry, rz, ctx.r2[c], sm->beg, sm->cnt, sm->md, t, nullptr, sm->q_d, sm->q_i, sm->q_p);
#else
    knn_search_pair<5, kBlk>(ctx.grid[c], live, rx, ry, rz, ctx.r2[c], sm->beg, sm->cnt, sm->md, t);
#endif
    if (ctx.dbg) tk1 = clock64();
    if (live && even) flag = fit_one<5>(ctx, c, t, prim);
  }
  double wv = 1.0;                                                              // :931-949
  if (even && live && st->outer != 0) {
    wv = ctx.w[gi];
    const double res = ctx.slot[gi];
    if (res != 0.0) {                                                           // Q13: res == 0 leaves the weight alone
      if (res >= st->th1) wv = 0.0;
      else if (res <= st->th2) wv = 1.0;
      else wv = sqrt(st->c2 * st->mu_used * (st->mu_used + 1.0) / res) - st->mu_used;
    }
  }
  if (ctx.dbg && threadIdx.x == 0 && live) {
    atomicAdd(&ctx.dbg[8], (unsigned long long)(tk1 - tk0));
    atomicAdd(&ctx.dbg[10], (unsigned long long)(clock64() - tk1));
    atomicAdd(&ctx.dbg[9], 1ull);
  }
  c_out = c; gi_out = gi; live_out = live; w_out = wv;
}

// Correspondence search + primitive fit as a kernel of its own (the path with binding `*_maxnum` caps, and the
// build_factors test entry point).  A LANE PAIR serves one feature, so a 128-thread block serves 64 features and a
// 128-feature block of one cloud is served by two thread blocks.  Resets the residual slot (:1118-1121).
// 5 blocks per SM (<= 102 registers): measured best; 6 (80 regs) and 8 (64 regs) spill and are 8% / 55% slower,
// 4 (114 regs, what ptxas picks when unconstrained) is 28% slower
template <bool kBatched>
__global__ void __launch_bounds__(kBlk, 5) k_correspond(const __grid_constant__ DeviceCtx one, const __grid_constant__ BatchTab tab) {
  TL_RESOLVE_CTX(one, tab);
  (void)nblk;
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  __shared__ SearchSmem s_search;
  const int fb = lb / 2, sb = lb % 2;
  const int buf = st->outer & 1;
  if (threadIdx.x == 0 && sb == 0) ctx.blk_count[(buf ^ 1) * ctx.blk_cap + fb] = 0;   // for the next outer iteration
  if (ctx.dense_mask) {                       // clouds served by k_correspond_dense: only the padding lanes are ours
    const int cd = cloud_of_block(ctx, fb);
    if ((ctx.dense_mask >> cd) & 1) {
      const int ild = (fb - ctx.blk_off[cd]) * kBlk + sb * (kBlk / 2) + (int)(threadIdx.x >> 1);
      if ((threadIdx.x & 1) == 0 && (ild >= ctx.n[cd] || !cloud_enabled(ctx, cd))) ctx.flags[ctx.pad_off[cd] + ild] = 0;
      return;
    }
  }
  int c, gi; bool live; unsigned char flag; double prim[6], wv;
  search_and_fit(ctx, st, &s_search, fb, sb, c, gi, live, flag, prim, wv);
  const bool even = (threadIdx.x & 1) == 0;
  if (even) {
    if (live) {
      ctx.w[gi] = wv;
      ctx.slot[gi] = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) ctx.prim[j][gi] = prim[j];
    }
    ctx.flags[gi] = flag;
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

// One factor: functor, *cost side effect, Cauchy scaling and its 28 normal-equation terms.
// v: H (21) | g (6) | cost | slot sum per cloud (4); nact: factors per cloud.  Returns the slot value (Q3/Q5).
__device__ __forceinline__ double accumulate_factor(int c, const double cpt[3], double w, const double* prim, double v[32],
                                                    int nact[4]) {
  double r[3], J[18];
  int nr;
  double slot;
  if (c == kPlanar || c == kGround) {
    const double n[3] = {prim[0], prim[1], prim[2]};
    functor_plane(cpt, n, prim[3], w, r[0], J);
    nr = 1;
    slot = r[0] * r[0];                                                    // :101
  } else if (c == kEdge) {
    const double a[3] = {prim[0], prim[1], prim[2]};
    const double bb[3] = {prim[3], prim[4], prim[5]};
    functor_line(cpt, a, bb, w, r, J);
    nr = 3;
    const double s3 = r[0] + r[1] + r[2];
    slot = s3 * s3;                                                        // :69 (Q3)
  } else {
    const double q[3] = {prim[0], prim[1], prim[2]};
    functor_point(cpt, q, w, r, J);
    nr = 3;
    const double s3 = r[0] + r[1] + r[2];
    slot = s3 * s3;                                                        // :32 (Q3)
  }
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
  return slot;
}

// shared memory of the reduction + solver tail.  `gather` receives distributed-shared-memory stores from the peer
// blocks of the cluster at any time after this block has started, so it must never alias anything else.
struct ReduceSmem {
  double red[kBlk / 32][32];
  int cnt[kBlk / 32][4];
  double tot[kNRed];
  double part[3][kNRed];
  FrameState state;
  SolverShared solver;
  bool last;
};
struct GatherSmem { double g[kEvalCluster][kNRed]; };

// Block -> cluster -> sequence reduction of the per-thread accumulators (fixed shape => run-to-run bit-reproducible)
// and, in the LAST cluster leader of the sequence, the deterministic sum of the per-cluster rows followed by the
// trust-region state machine (solver.cuh).  cidx = cluster index inside the sequence, nclusters = clusters of the
// sequence.  The caller has executed cluster_arrive() once ("I have started").  Returns true in the block that ran
// the solver.
__device__ __forceinline__ bool reduce_and_solve(const DeviceCtx& ctx, double v[32], const int nact[4], int cidx, int nclusters,
                                                 ReduceSmem* rs, GatherSmem* gs) {
  FrameState* st = ctx.st;
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
    rs->red[warp][lane] = v[0];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int tot_k = __reduce_add_sync(0xffffffffu, nact[k]);
      if (lane == 0) rs->cnt[warp][k] = tot_k;
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
      for (int wi = 0; wi < kBlk / 32; ++wi) s += rs->red[wi][t];
    } else {
      int n = 0;
      for (int wi = 0; wi < kBlk / 32; ++wi) n += rs->cnt[wi][t - 32];
      s = (double)n;
    }
    dsmem_store(&gs->g[crank][t], 0u, s);
  }
  cluster_arrive();                                   // release: my stores are visible to whoever waits
  cluster_wait();
  if (crank != 0u) return false;
  if (threadIdx.x < kNRed) {
    const int t = threadIdx.x;
    double s = gs->g[0][t];
#pragma unroll
    for (int r = 1; r < kEvalCluster; ++r) s += gs->g[r][t];         // fixed order => deterministic
    ctx.partial[(size_t)cidx * kNRed + t] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(ctx.counter, 1u);
    rs->last = (ticket == (unsigned)nclusters - 1u);
  }
  __syncthreads();
  if (!rs->last) return false;
  // ---- last cluster leader: deterministic sum of the per-cluster partials, then the solver state machine ----
  __threadfence();
  unsigned long long tg1 = 0, tg2 = 0, tg3 = 0;
  if (ctx.dbg && threadIdx.x == 0) tg1 = gtime_ns();
  {
    // the state machine is a long dependent chain: run it on a shared-memory copy of the state
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(st);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&rs->state);
    for (unsigned i = threadIdx.x; i < sizeof(FrameState) / 8; i += kBlk) dst[i] = __ldcg(src + i);
  }
  if (threadIdx.x < 3 * kNRed) {
    // 3 row groups x 36 columns, 8 independent accumulators each; the summation tree is fixed => deterministic
    const int col = threadIdx.x % kNRed, grp = threadIdx.x / kNRed;
    const int nb = nclusters;
    const double* P = ctx.partial + col;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = grp; r < nb; r += 24) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = r + 3 * u;
        if (row < nb) a[u] += __ldcg(P + (size_t)row * kNRed);
      }
    }
    rs->part[grp][col] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  __syncthreads();
  if (threadIdx.x < kNRed) rs->tot[threadIdx.x] = (rs->part[0][threadIdx.x] + rs->part[1][threadIdx.x]) + rs->part[2][threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    *ctx.counter = 0u;
    if (ctx.dbg) tg2 = gtime_ns();
  }
  solver_on_eval(ctx, &rs->state, rs->tot, &rs->solver);      // all threads enter (3 helper warps + thread 0)
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
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&rs->state);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(st);
    for (unsigned i = threadIdx.x; i < sizeof(FrameState) / 8; i += kBlk) dst[i] = src[i];
  }
  __syncthreads();
  return true;
}

__device__ __forceinline__ Pose7 load_eval_pose(const FrameState* st) {   // read around L1 (written by the previous launch's solver block)
  Pose7 ev;
  ev.qw = __ldcg(&st->evalq.qw); ev.qx = __ldcg(&st->evalq.qx); ev.qy = __ldcg(&st->evalq.qy); ev.qz = __ldcg(&st->evalq.qz);
  ev.tx = __ldcg(&st->evalq.tx); ev.ty = __ldcg(&st->evalq.ty); ev.tz = __ldcg(&st->evalq.tz);
  return ev;
}

// One evaluation pass of the whole problem at st->evalq + reduction + solver (last cluster leader).
// Grid-stride over the 128-feature blocks of the sequence.  Clusters of 8 blocks: the block totals are combined
// through distributed shared memory before they reach global memory, so the serial tail sums 1/8 of the rows.
#ifndef TLOAM_EVAL_MINBLOCKS_BATCHED
#define TLOAM_EVAL_MINBLOCKS_BATCHED 3
#endif
template <bool kFirst, bool kBatched>
__global__ void __cluster_dims__(kEvalCluster, 1, 1) __launch_bounds__(kBlk, kBatched ? TLOAM_EVAL_MINBLOCKS_BATCHED : 3)
k_eval(const __grid_constant__ DeviceCtx one, const __grid_constant__ BatchTab tab) {
  TL_RESOLVE_CTX(one, tab);
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != (kFirst ? kPhaseIter0 : kPhaseCand)) return;     // uniform per sequence (= per cluster)
  __shared__ ReduceSmem s_rs;
  __shared__ GatherSmem s_gs;
  __shared__ int s_warp[1 + kBlk / 32];
  cluster_arrive();                                  // "I have started": peers may store into my shared memory
  if (ctx.dbg && threadIdx.x == 0) atomicMin(&ctx.dbg[0], gtime_ns());
  double v[32];
  int nact[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.0;
  const Rt T = pose_to_rt(load_eval_pose(st));
  for (int fb = lb; fb < ctx.blk_off[4]; fb += nblk) {
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
    double prim[6];
    prim[0] = ctx.prim[0][gi]; prim[1] = ctx.prim[1][gi]; prim[2] = ctx.prim[2][gi];
    prim[3] = (c != kSphere) ? ctx.prim[3][gi] : 0.0;
    prim[4] = (c == kEdge) ? ctx.prim[4][gi] : 0.0;
    prim[5] = (c == kEdge) ? ctx.prim[5][gi] : 0.0;
    ctx.slot[gi] = accumulate_factor(c, cpt, w, prim, v, nact);             // *cost side effect (Q5)
  }
  reduce_and_solve(ctx, v, nact, lb / kEvalCluster, nblk / kEvalCluster, &s_rs, &s_gs);
}

// FUSED first evaluation of an outer iteration: correspondence search + fit + weight update (k_correspond) and the
// evaluation at the same pose (k_eval<first>) in ONE kernel -- the primitive, the weight and T*p stay in registers
// (-1 launch and -64 B per feature of traffic per outer iteration).  Valid when no `*_maxnum` cap can bind (the cap
// needs a prefix over ALL earlier features of the cloud, i.e. a grid-wide dependency); the host picks the unfused
// sequence otherwise.  Per-sequence grid = 2 x feature blocks rounded up to the cluster size (padding blocks only
// take part in the reduction).  The search arrays are dead when the reduction starts and share its memory.
#ifndef TLOAM_FIRST_MINBLOCKS_BATCHED
#define TLOAM_FIRST_MINBLOCKS_BATCHED 5
#endif
template <bool kBatched>
__global__ void __cluster_dims__(kEvalCluster, 1, 1) __launch_bounds__(kBlk, kBatched ? TLOAM_FIRST_MINBLOCKS_BATCHED : 5)
k_first(const __grid_constant__ DeviceCtx one, const __grid_constant__ BatchTab tab) {
  TL_RESOLVE_CTX(one, tab);
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  constexpr size_t kUnion = sizeof(SearchSmem) > sizeof(ReduceSmem) ? sizeof(SearchSmem) : sizeof(ReduceSmem);
  __shared__ __align__(16) unsigned char s_union[kUnion];
  __shared__ GatherSmem s_gs;
  SearchSmem* ss = reinterpret_cast<SearchSmem*>(s_union);
  ReduceSmem* rs = reinterpret_cast<ReduceSmem*>(s_union);
  cluster_arrive();
  if (ctx.dbg && threadIdx.x == 0) atomicMin(&ctx.dbg[0], gtime_ns());
  int c, gi; bool live; unsigned char flag; double prim[6], wv;
  search_and_fit(ctx, st, ss, lb / 2, lb % 2, c, gi, live, flag, prim, wv);
  const bool even = (threadIdx.x & 1) == 0;
  double v[32];
  int nact[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.0;
  if (even && live) {
    const bool act = (flag & kFlagCand) != 0;
    double slot = 0.0;
    if (act) {
      const Rt T = pose_to_rt(st->xq);               // iteration zero evaluates at the accepted pose (evalq == xq)
      double cpt[3];
      rt_apply(T, ctx.px[gi], ctx.py[gi], ctx.pz[gi], cpt[0], cpt[1], cpt[2]);
      slot = accumulate_factor(c, cpt, wv, prim, v, nact);
    }
    ctx.w[gi] = wv;
    ctx.slot[gi] = slot;                             // reset (:1118-1121) or the *cost side effect of this evaluation
#pragma unroll
    for (int j = 0; j < 6; ++j) ctx.prim[j][gi] = prim[j];
    ctx.active[gi] = act ? 1 : 0;
  } else if (even && lb / 2 < ctx.blk_off[4]) {
    ctx.active[gi] = 0;                              // padding lanes / disabled clouds
  }
  __syncthreads();                                   // the search arrays are dead: the reduction reuses their memory
  reduce_and_solve(ctx, v, nact, lb / kEvalCluster, nblk / kEvalCluster, rs, &s_gs);
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

// sums the per-block partials of k_fitness per cloud and leaves (fitness, rmse) in the frame state
// (ref: registration.cpp:278-284, 292-293: per type matched / total and sqrt(sum d2 / matched), summed over the types)
// `out` = {fitness, rmse}: NOT the frame state when the kernel runs beside the registration -- the solver block carries the
// whole FrameState through shared memory and writes it back, which would clobber the two words.
__global__ void __launch_bounds__(128) k_fitness_reduce(const __grid_constant__ DeviceCtx ctx, const double* part /*[blocks][2]*/, double* out) {
  // one warp per cloud: lane-strided partial sums, then a butterfly (a fixed tree: the result does not depend on timing)
  __shared__ double s_fit[4], s_rm[4];
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double err = 0.0, cnt = 0.0;
  for (int b = ctx.blk_off[c] + lane; b < ctx.blk_off[c + 1]; b += 32) { err += part[2 * b]; cnt += part[2 * b + 1]; }
  for (int o = 16; o > 0; o >>= 1) { err += __shfl_xor_sync(0xffffffffu, err, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  if (lane == 0) {
    const bool any = cnt > 0.0;
    s_fit[c] = any ? cnt / (double)ctx.n[c] : 0.0;
    s_rm[c] = any ? sqrt(err / cnt) : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = ((s_fit[0] + s_fit[1]) + s_fit[2]) + s_fit[3];
    out[1] = ((s_rm[0] + s_rm[1]) + s_rm[2]) + s_rm[3];
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
  else if (op == 2) { se3_log(se3_mul(se3_exp(in.m + 6), se3_exp(in.m)), out); }   // plus: in.m[0..5] = x, in.m[6..11] = delta
  else { min_on_boundary_2d(in.m, in.m + 4, in.m[6], out); }              // in.m[0..3] = B, [4..5] = g, [6] = radius
}

}  // namespace tloam
