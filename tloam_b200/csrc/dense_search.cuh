// dense_search.cuh -- correspondence search for DENSE map clouds (BASELINE config 3: ~400 points per 0.5 m cell),
// with TMA-staged point tiles.
//
// The lane-pair search of map_grid.cuh streams every candidate of the 27 neighbour cells through L1/L2: at 400
// points per cell that is ~4000 candidates per query (3.0 ms per launch at F = 500k).  Here the QUERIES are binned
// by map cell first (exact FP64 cell of T*p, counting sort through a small hash table: k_qbin_*), so that a thread
// block owns up to 128 queries of ONE cell and therefore ONE 27-cell neighbourhood:
//   1. warp 0 probes the 8 bricks of the neighbourhood and lists the <= 27 contiguous point runs;
//   2. the runs are copied global -> shared memory by the TMA engine (cp.async.bulk ... mbarrier::complete_tx::bytes,
//      one bulk copy per run, SASS UBLKCP), in passes of at most kDenseCap points;
//   3. the staged points are counting-sorted IN SHARED MEMORY into a 12 x 12 x 12 grid of fine cells (edge = r / 4)
//      covering the 3 x 3 x 3 cells;
//   4. every thread searches its query nearest-first: the 27 fine cells around it, then only the fine cells that can
//      still hold a point closer than the K-th best (box of the current reach, per-row lower bounds);
//   5. fit + weight update exactly as k_correspond.
// The result is the same exact radius-truncated kNN, ordered by (d2, original index), with the same FP64 distance
// expression -- the path taken (dense or lane-pair) never changes a pose.  A point is visited in exactly one pass,
// the top-K list persists across passes, and pruning only ever uses the current K-th best distance, so any number of
// passes is exact.
#pragma once
#include "frame_kernels.cuh"

namespace tloam {

constexpr int kDenseBlk = 128;             // queries per work item
constexpr int kDenseThreads = 512;          // 16 warps, one block per SM (128 registers, ~215 KB): a GROUP of L = 4..32
                                           // lanes serves one query, L = 512 / queries of the work item (power of two)
constexpr int kFine = 4;                   // fine cells per cell edge
constexpr int kBox = 3 * kFine;            // fine cells per axis of the staged box
constexpr int kBoxCells = kBox * kBox * kBox;
constexpr int kDenseCap = 9216;            // staged points per pass (144 KB; one block per SM)

struct DenseWork { int cx, cy, cz; unsigned qbeg; unsigned short qcnt; unsigned short cloud; };

struct DenseCloud {
  unsigned long long* keys;    // [tmask+1] cell keys (0 = empty)
  unsigned* cnt;               // [tmask+1] queries per cell
  unsigned* base;              // [tmask+1] first position of the cell in q_order
  unsigned* q_slot;            // [n]
  unsigned* q_rank;            // [n]
  unsigned* q_order;           // [n] query indices grouped by cell
  unsigned tmask;
  unsigned toff;               // first slot of this cloud in the concatenated slot space of k_qbin_offsets
};
struct DenseArgs {
  DenseCloud cl[4];
  DenseWork* work;             // [sum n]
  unsigned* ctl;               // [0] nwork, [1] next work item, [2..5] q_order bump allocators
  unsigned tslots;             // total slots over the dense clouds
  unsigned* dbg;               // self-check counters (TLOAM_B200_DENSE_CHECK=1), else nullptr: [0] queries, [1] kNN
                               // mismatches vs the plain search, [2] work items, [3] staging passes, [4] first bad gi
  int mask;                    // bit c: cloud c takes the dense path
};

__device__ __forceinline__ void decode_cell_key(unsigned long long key, int& cx, int& cy, int& cz) {
  cx = (int)((key >> 42) & 0x1FFFFFull) - (1 << 20);
  cy = (int)((key >> 21) & 0x1FFFFFull) - (1 << 20);
  cz = (int)(key & 0x1FFFFFull) - (1 << 20);
}

// ---- query binning: counting sort of the queries by the map cell of T*p ----
__global__ void __launch_bounds__(kBlk) k_qbin_count(const __grid_constant__ DeviceCtx ctx, const __grid_constant__ DenseArgs a) {
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  const int fb = blockIdx.x;
  const int c = cloud_of_block(ctx, fb);
  if (!((a.mask >> c) & 1) || !cloud_enabled(ctx, c)) return;
  const int il = (fb - ctx.blk_off[c]) * kBlk + threadIdx.x;
  if (il >= ctx.n[c]) return;
  const int gi = ctx.pad_off[c] + il;
  const Rt T = pose_to_rt(st->xq);
  double qx, qy, qz;
  rt_apply(T, ctx.px[gi], ctx.py[gi], ctx.pz[gi], qx, qy, qz);
  const GridDesc& g = ctx.grid[c];
  const double rx = qx - ctx.origin[0], ry = qy - ctx.origin[1], rz = qz - ctx.origin[2];
  const int cx = (int)floor(rx * g.inv_cell), cy = (int)floor(ry * g.inv_cell), cz = (int)floor(rz * g.inv_cell);   // as knn_search
  const unsigned long long key = cell_key(cx, cy, cz);
  const DenseCloud& q = a.cl[c];
  unsigned s = hash_key(key) & q.tmask;
  while (true) {
    const unsigned long long prev = atomicCAS(&q.keys[s], 0ull, key);
    if (prev == 0ull || prev == key) break;
    s = (s + 1u) & q.tmask;
  }
  q.q_slot[il] = s;
  q.q_rank[il] = atomicAdd(&q.cnt[s], 1u);
}

// base offsets of the occupied query cells + the work list (one item per <= kDenseBlk queries of a cell)
__global__ void __launch_bounds__(256) k_qbin_offsets(const __grid_constant__ DeviceCtx ctx, const __grid_constant__ DenseArgs a) {
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.tslots) return;
  int c = 0;
  for (int k = 0; k < 4; ++k)
    if (((a.mask >> k) & 1) && s >= a.cl[k].toff) c = k;
  const DenseCloud& q = a.cl[c];
  s -= q.toff;
  const unsigned cnt = q.cnt[s];
  if (cnt == 0u) return;
  const unsigned base = atomicAdd(&a.ctl[2 + c], cnt);
  q.base[s] = base;
  const unsigned nitems = (cnt + kDenseBlk - 1) / kDenseBlk;
  const unsigned w0 = atomicAdd(&a.ctl[0], nitems);
  int cx, cy, cz;
  decode_cell_key(q.keys[s], cx, cy, cz);
  for (unsigned i = 0; i < nitems; ++i) {
    DenseWork wk;
    wk.cx = cx; wk.cy = cy; wk.cz = cz;
    wk.qbeg = base + i * kDenseBlk;
    const unsigned rem = cnt - i * kDenseBlk;
    wk.qcnt = (unsigned short)(rem < (unsigned)kDenseBlk ? rem : (unsigned)kDenseBlk);
    wk.cloud = (unsigned short)c;
    a.work[w0 + i] = wk;
  }
}

__global__ void __launch_bounds__(kBlk) k_qbin_scatter(const __grid_constant__ DeviceCtx ctx, const __grid_constant__ DenseArgs a) {
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  const int fb = blockIdx.x;
  const int c = cloud_of_block(ctx, fb);
  if (!((a.mask >> c) & 1) || !cloud_enabled(ctx, c)) return;
  const int il = (fb - ctx.blk_off[c]) * kBlk + threadIdx.x;
  if (il >= ctx.n[c]) return;
  const DenseCloud& q = a.cl[c];
  q.q_order[q.base[q.q_slot[il]] + q.q_rank[il]] = (unsigned)il;
}

// ---- top-K list that keeps the neighbours' stored coordinates (no re-load for the fit) ----
template <int K>
struct TopKP {
  double d2[K];
  int idx[K];
  float x[K], y[K], z[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < K; ++j) { d2[j] = __longlong_as_double(0x7FF0000000000000ll); idx[j] = 0x7FFFFFFF; x[j] = y[j] = z[j] = 0.f; }
  }
  __device__ __forceinline__ int count() const {
    int c = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) c += (idx[j] != 0x7FFFFFFF) ? 1 : 0;
    return c;
  }
  __device__ __forceinline__ void insert(double d, int i, float px, float py, float pz) {
    bool lt[K];
#pragma unroll
    for (int j = 0; j < K; ++j) lt[j] = (d < d2[j]) || (d == d2[j] && i < idx[j]);
    if (!lt[K - 1]) return;
#pragma unroll
    for (int j = K - 1; j >= 0; --j) {
      const bool shift = (j > 0) ? lt[j - 1] : false;
      if (j > 0 && shift) { d2[j] = d2[j - 1]; idx[j] = idx[j - 1]; x[j] = x[j - 1]; y[j] = y[j - 1]; z[j] = z[j - 1]; }
      else if (lt[j]) { d2[j] = d; idx[j] = i; x[j] = px; y[j] = py; z[j] = pz; }
    }
  }
};

// a query that the 27 fine cells around it could not finish: position + its top-K so far
struct HardRec { double q[3]; double d2[5]; int idx[5]; float x[5], y[5], z[5]; int pad; };

struct DenseSmem {
  float4 pts[kDenseCap];
  unsigned short order[kDenseCap];
  unsigned short fid[kDenseCap];          // fine cell of each staged point (computed once, used by count and scatter)
  unsigned start[kBoxCells + 1];
  unsigned cursor[kBoxCells];
  unsigned run_beg[27], run_cnt[27];
  unsigned cp_src[32], cp_dst[32], cp_n[32];
  unsigned scan_tmp[kDenseThreads / 32];
  int ncopy, fill, run_i;
  unsigned run_o, work, nhard;
  HardRec hard[kDenseBlk];
  unsigned long long mbar;
};
constexpr size_t kDenseSmemBytes = sizeof(DenseSmem) + 128;

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TL_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TL_DONE_%=;\n"
      "bra TL_WAIT_%=;\n"
      "TL_DONE_%=:\n"
      "}\n" ::"r"(addr), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// Persistent blocks; one work item (<= 128 queries of one map cell) per iteration.
__global__ void __launch_bounds__(kDenseThreads, 1) k_correspond_dense(const __grid_constant__ DeviceCtx ctx, const __grid_constant__ DenseArgs a) {
  const FrameState* st = ctx.st;
  if (st->frame_done || st->phase != kPhaseIter0) return;
  // NOTE: no integer round trip on this pointer -- an address rebuilt from a uintptr_t loses the shared state space and
  // every access of the kernel becomes a generic LD / ATOM (measured: 0 LDS in the SASS, the search phases ran ~10x slower)
  extern __shared__ __align__(128) unsigned char dense_raw[];
  DenseSmem& sm = *reinterpret_cast<DenseSmem*>(dense_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) mbar_init(&sm.mbar, 1u);
  __syncthreads();
  unsigned parity = 0u;
  const Rt T = pose_to_rt(st->xq);
  const int buf = st->outer & 1;
  const unsigned nwork = a.ctl[0];
  while (true) {
    if (tid == 0) sm.work = atomicAdd(&a.ctl[1], 1u);
    __syncthreads();
    const unsigned w = sm.work;
    if (w >= nwork) break;
    const DenseWork wk = a.work[w];
    const int c = wk.cloud;
    long long tk0 = 0, tk1 = 0, tk_wait = 0, tk_sort = 0, tk_search = 0, tk_p1 = 0, tk_p2 = 0, tk2 = 0;     // self-check mode: SM cycles per phase (thread 0)
    if (a.dbg && tid == 0) tk0 = clock64();
    const GridDesc& g = ctx.grid[c];
    // ---- the <= 27 point runs of the neighbourhood (warp 0: 8 brick probes, then one cell per lane) ----
    if (warp == 0) {
      const int bx0 = brick_of(wk.cx - 1), by0 = brick_of(wk.cy - 1), bz0 = brick_of(wk.cz - 1);
      BrickEntry e;
      e.a = make_uint4(0, 0, 0, 0); e.b = make_uint4(0, 0, 0, 0);
      int have = 0;
      if (lane < 8 && g.n != 0u) have = probe_brick(g, cell_key(bx0 + (lane & 1), by0 + ((lane >> 1) & 1), bz0 + (lane >> 2)), e) ? 1 : 0;
      const int l27 = lane < 27 ? lane : 0;
      const int gx = wk.cx + (l27 % 3) - 1, gy = wk.cy + ((l27 / 3) % 3) - 1, gz = wk.cz + (l27 / 9) - 1;
      const int ib = (brick_of(gx) - bx0) | ((brick_of(gy) - by0) << 1) | ((brick_of(gz) - bz0) << 2);
      const int sub = subcell_of(gx, gy, gz);
      const int h = __shfl_sync(0xffffffffu, have, ib);
      const unsigned base = __shfl_sync(0xffffffffu, e.a.z, ib);
      const unsigned w0 = __shfl_sync(0xffffffffu, e.a.w, ib), w1 = __shfl_sync(0xffffffffu, e.b.x, ib);
      const unsigned w2 = __shfl_sync(0xffffffffu, e.b.y, ib), w3 = __shfl_sync(0xffffffffu, e.b.z, ib);
      const unsigned wds[4] = {w0, w1, w2, w3};
      unsigned beg = base, cnt = 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned cq = (wds[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;
        if (q < sub) beg += cq;
        if (q == sub) cnt = cq;
      }
      if (lane < 27) { sm.run_beg[lane] = beg; sm.run_cnt[lane] = h ? cnt : 0u; }
      if (lane == 0) { sm.run_i = 0; sm.run_o = 0u; }
    }
    // ---- my query: a QUAD of lanes serves one query (the quad splits every candidate run four ways; more warps per
    //      SM for the same shared-memory footprint -- the search is a latency-bound dependent chain per lane) ----
    int lsh = 2;                                       // log2(L): as many lanes per query as 512 threads allow
    while (lsh < 5 && ((int)wk.qcnt << (lsh + 1)) <= kDenseThreads) ++lsh;
    const int L = 1 << lsh;
    const int qid = tid >> lsh, sub = tid & (L - 1);
    const unsigned quadmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (lane & ~(L - 1)));
    const bool hasq = qid < (int)wk.qcnt;
    int il = 0, gi = 0;
    double rx = 0.0, ry = 0.0, rz = 0.0;
    if (hasq) {
      il = (int)a.cl[c].q_order[wk.qbeg + qid];
      gi = ctx.pad_off[c] + il;
      double qx, qy, qz;
      rt_apply(T, ctx.px[gi], ctx.py[gi], ctx.pz[gi], qx, qy, qz);
      rx = qx - ctx.origin[0]; ry = qy - ctx.origin[1]; rz = qz - ctx.origin[2];
    }
    const double e = g.cell * (1.0 / kFine), inv_e = g.inv_cell * (double)kFine;
    const double bx = (double)(wk.cx - 1) * g.cell, by = (double)(wk.cy - 1) * g.cell, bz = (double)(wk.cz - 1) * g.cell;
    const double r2 = ctx.r2[c];
    auto fclamp = [](double v) { int i = (int)floor(v); return i < 0 ? 0 : (i > kBox - 1 ? kBox - 1 : i); };
    const int qi = fclamp((rx - bx) * inv_e), qj = fclamp((ry - by) * inv_e), qk = fclamp((rz - bz) * inv_e);
    const float bxf = (float)bx, byf = (float)by, bzf = (float)bz;                    // exact (multiples of the cell edge)
    const float qlx = (float)(rx - bx), qly = (float)(ry - by), qlz = (float)(rz - bz);
    const float inv_ef = (float)inv_e;
    TopKP<5> t;
    t.init();
    __syncthreads();
    // ---- passes over the neighbourhood ----
    while (true) {
      if (tid == 0) {
        int n = 0, fill = 0, ri = sm.run_i;
        unsigned ro = sm.run_o;
        while (ri < 27 && fill < kDenseCap && n < 32) {
          const unsigned rem = sm.run_cnt[ri] - ro;
          if (rem == 0u) { ++ri; ro = 0u; continue; }
          const unsigned take = rem < (unsigned)(kDenseCap - fill) ? rem : (unsigned)(kDenseCap - fill);
          sm.cp_src[n] = sm.run_beg[ri] + ro; sm.cp_dst[n] = (unsigned)fill; sm.cp_n[n] = take;
          ++n; fill += (int)take; ro += take;
          if (ro == sm.run_cnt[ri]) { ++ri; ro = 0u; }
        }
        sm.run_i = ri; sm.run_o = ro; sm.ncopy = n; sm.fill = fill; sm.nhard = 0u;
        if (fill > 0) mbar_arrive_expect_tx(&sm.mbar, (unsigned)fill * 16u);
      }
      __syncthreads();
      const int fill = sm.fill;
      if (fill == 0) break;
      if (a.dbg && tid == 0) atomicAdd(&a.dbg[3], 1u);
      if (tid < sm.ncopy) tma_bulk_g2s(&sm.pts[sm.cp_dst[tid]], g.pts + sm.cp_src[tid], sm.cp_n[tid] * 16u, &sm.mbar);
      if (a.dbg && tid == 0) tk1 = clock64();
      for (int i = tid; i < kBoxCells; i += kDenseThreads) sm.cursor[i] = 0u;      // while the copies are in flight
      mbar_wait(&sm.mbar, parity);
      parity ^= 1u;
      __syncthreads();
      if (a.dbg && tid == 0) { const long long t = clock64(); tk_wait += t - tk1; tk1 = t; }
      // counting sort of the staged points into the fine grid (indices only: the points stay where TMA put them).
      // Fine cell from coordinates LOCAL to the box in FP32 (p - box0 is exact); a point within ~1e-7 m of a fine-cell
      // face may land on either side, which the 1e-6 m slack of the lower bounds covers.
      for (int i = tid; i < fill; i += kDenseThreads) {
        const float4 p = sm.pts[i];
        int fi = (int)floorf((p.x - bxf) * inv_ef), fj = (int)floorf((p.y - byf) * inv_ef), fk = (int)floorf((p.z - bzf) * inv_ef);
        fi = fi < 0 ? 0 : (fi > kBox - 1 ? kBox - 1 : fi); fj = fj < 0 ? 0 : (fj > kBox - 1 ? kBox - 1 : fj); fk = fk < 0 ? 0 : (fk > kBox - 1 ? kBox - 1 : fk);
        const int id = (fk * kBox + fj) * kBox + fi;
        sm.fid[i] = (unsigned short)id;
        atomicAdd(&sm.cursor[id], 1u);
      }
      __syncthreads();
      {
        constexpr int kPer = (kBoxCells + kDenseThreads - 1) / kDenseThreads;      // consecutive entries per thread
        const int i0 = tid * kPer;
        unsigned local = 0u;
        for (int i = i0; i < i0 + kPer && i < kBoxCells; ++i) local += sm.cursor[i];
        unsigned incl = local;
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) sm.scan_tmp[warp] = incl;
        __syncthreads();
        unsigned run = incl - local;
        for (int wi = 0; wi < warp; ++wi) run += sm.scan_tmp[wi];
        for (int i = i0; i < i0 + kPer && i < kBoxCells; ++i) { const unsigned cn = sm.cursor[i]; sm.start[i] = run; sm.cursor[i] = run; run += cn; }
        if (tid == kDenseThreads - 1) sm.start[kBoxCells] = (unsigned)fill;
      }
      __syncthreads();
      for (int i = tid; i < fill; i += kDenseThreads) sm.order[atomicAdd(&sm.cursor[sm.fid[i]], 1u)] = (unsigned short)i;
      __syncthreads();
      if (a.dbg && tid == 0) { const long long t = clock64(); tk_sort += t - tk1; tk1 = t; }
      // ---- phase 1: the 3 x 3 x 3 fine cells around the query, by the query's group of L lanes ----
      // Lane 0 of the group carries the running top-K across passes; the other lanes start every pass empty.  Every row
      // of the 27-cell neighbourhood is split L ways (lane `sub` takes candidates sub, sub + L, ...).  Candidates are
      // pre-filtered in FP32 on coordinates LOCAL to the staged box (p - box0 is exact in FP32: both are multiples of the
      // point's ulp and the difference is < 2 cells), with a margin that covers the FP32 rounding of the query and of the
      // arithmetic (< 1e-6 m^2 at cell <= 1 m); survivors get the exact FP64 expression.
      if (sub != 0) t.init();
      double bound = t.d2[4] < r2 ? t.d2[4] : r2;                 // group-uniform after the broadcast below
      bound = __shfl_sync(quadmask, bound, (lane & ~(L - 1)));
      if (hasq) {
        float boundf = (float)bound * 1.00001f + 2e-6f;
        for (int dk = -1; dk <= 1; ++dk) {
          const int k = qk + dk;
          if (k < 0 || k >= kBox) continue;
          for (int dj = -1; dj <= 1; ++dj) {
            const int j = qj + dj;
            if (j < 0 || j >= kBox) continue;
            const int row = (k * kBox + j) * kBox;
            const int i0 = qi - 1 > 0 ? qi - 1 : 0, i1 = qi + 1 < kBox - 1 ? qi + 1 : kBox - 1;
            const unsigned en = sm.start[row + i1 + 1];
#pragma unroll 2
            for (unsigned u = sm.start[row + i0] + (unsigned)sub; u < en; u += (unsigned)L) {
              const float4 p = sm.pts[sm.order[u]];
              const float fx = (p.x - bxf) - qlx, fy = (p.y - byf) - qly, fz = (p.z - bzf) - qlz;
              const float df = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
              if (df <= boundf) {
                const double ddx = (double)p.x - rx, ddy = (double)p.y - ry, ddz = (double)p.z - rz;
                const double d = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));   // same three operations as the oracle
                if (d < r2) {
                  t.insert(d, __float_as_int(p.w), p.x, p.y, p.z);
                  bound = t.d2[4] < bound ? t.d2[4] : bound;
                  boundf = (float)bound * 1.00001f + 2e-6f;
                }
              }
            }
          }
        }
      }
      // merge the group's L sorted lists into its lane 0 (log2 L rounds of pull + insert)
      for (int step = 1; step < L; step <<= 1) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const double od = __shfl_xor_sync(0xffffffffu, t.d2[j], step);
          const int oi = __shfl_xor_sync(0xffffffffu, t.idx[j], step);
          const float ox = __shfl_xor_sync(0xffffffffu, t.x[j], step), oy = __shfl_xor_sync(0xffffffffu, t.y[j], step);
          const float oz = __shfl_xor_sync(0xffffffffu, t.z[j], step);
          if ((sub & (2 * step - 1)) == 0 && oi != 0x7FFFFFFF) t.insert(od, oi, ox, oy, oz);
        }
      }
      // Everything OUTSIDE the 27 fine cells is at least one fine-cell edge away: the query is finished with this pass iff
      // its list is full and the K-th best is closer than that.  The rest (outliers hanging in free space, sparse spots)
      // go to the block's hard list and are finished by a whole warp each, brute force over the staged points.
      int hslot = -1;
      if (hasq && sub == 0) {
        const double bq = t.d2[4] < r2 ? t.d2[4] : r2;
        const double lb = e - 1e-6;
        if (!(lb * lb > bq)) {
          hslot = (int)atomicAdd(&sm.nhard, 1u);
          HardRec& hr = sm.hard[hslot];
          hr.q[0] = rx; hr.q[1] = ry; hr.q[2] = rz;
#pragma unroll
          for (int j = 0; j < 5; ++j) { hr.d2[j] = t.d2[j]; hr.idx[j] = t.idx[j]; hr.x[j] = t.x[j]; hr.y[j] = t.y[j]; hr.z[j] = t.z[j]; }
        }
      }
      __syncthreads();
      if (a.dbg && tid == 0) { tk2 = clock64(); tk_p1 += tk2 - tk1; }
      // ---- phase 2: one warp per hard query, all staged points of this pass, 32 ways ----
      const unsigned nhard = sm.nhard;
      if (a.dbg && tid == 0) atomicAdd(&a.dbg[13], nhard);
      for (unsigned hq = (unsigned)warp; hq < nhard; hq += kDenseThreads / 32) {
        HardRec& hr = sm.hard[hq];
        const double hx = hr.q[0], hy = hr.q[1], hz = hr.q[2];
        const float hlx = (float)(hx - bx), hly = (float)(hy - by), hlz = (float)(hz - bz);
        TopKP<5> tl;
        tl.init();
        double hb = hr.d2[4] < r2 ? hr.d2[4] : r2;
        float hbf = (float)hb * 1.00001f + 2e-6f;
        // software-pipelined by hand: 4 coalesced 16-byte loads in flight per lane, the 4 FP32 distances, then the tests
        for (int i0 = lane; i0 < fill; i0 += 128) {
          float4 p[4];
          float df[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            p[u] = sm.pts[i < fill ? i : i0];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float fx = (p[u].x - bxf) - hlx, fy = (p[u].y - byf) - hly, fz = (p[u].z - bzf) - hlz;
            df[u] = (i0 + 32 * u < fill) ? fmaf(fz, fz, fmaf(fy, fy, fx * fx)) : 3.0e38f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (df[u] <= hbf) {
              const double ddx = (double)p[u].x - hx, ddy = (double)p[u].y - hy, ddz = (double)p[u].z - hz;
              const double d = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));
              if (d < r2) {
                tl.insert(d, __float_as_int(p[u].w), p[u].x, p[u].y, p[u].z);
                hb = tl.d2[4] < hb ? tl.d2[4] : hb;
                hbf = (float)hb * 1.00001f + 2e-6f;
              }
            }
          }
        }
        for (int step = 1; step < 32; step <<= 1) {
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const double od = __shfl_xor_sync(0xffffffffu, tl.d2[j], step);
            const int oi = __shfl_xor_sync(0xffffffffu, tl.idx[j], step);
            const float ox = __shfl_xor_sync(0xffffffffu, tl.x[j], step), oy = __shfl_xor_sync(0xffffffffu, tl.y[j], step);
            const float oz = __shfl_xor_sync(0xffffffffu, tl.z[j], step);
            if ((lane & (2 * step - 1)) == 0 && oi != 0x7FFFFFFF) tl.insert(od, oi, ox, oy, oz);
          }
        }
        if (lane == 0) {
          // fold the phase-1 list in; the 27 cells were scanned twice, so equal (d2, index) pairs are dropped
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const int oi = hr.idx[j];
            if (oi == 0x7FFFFFFF) continue;
            bool dup = false;
#pragma unroll
            for (int m = 0; m < 5; ++m) dup = dup || tl.idx[m] == oi;
            if (!dup) tl.insert(hr.d2[j], oi, hr.x[j], hr.y[j], hr.z[j]);
          }
#pragma unroll
          for (int j = 0; j < 5; ++j) { hr.d2[j] = tl.d2[j]; hr.idx[j] = tl.idx[j]; hr.x[j] = tl.x[j]; hr.y[j] = tl.y[j]; hr.z[j] = tl.z[j]; }
        }
      }
      __syncthreads();
      if (a.dbg && tid == 0) tk_p2 += clock64() - tk2;
      if (hslot >= 0) {
        const HardRec& hr = sm.hard[hslot];
#pragma unroll
        for (int j = 0; j < 5; ++j) { t.d2[j] = hr.d2[j]; t.idx[j] = hr.idx[j]; t.x[j] = hr.x[j]; t.y[j] = hr.y[j]; t.z[j] = hr.z[j]; }
      }
      __syncthreads();                         // everybody is done with the staged points before the next pass
      if (a.dbg && tid == 0) { const long long tt = clock64(); tk_search += tt - tk1; tk1 = tt; }
    }
    const bool owner = hasq && sub == 0;
    if (a.dbg && owner) {                      // self-check against the thread-per-query search of map_grid.cuh
      TopK<5> ref;
      knn_search<5>(g, rx, ry, rz, r2, ref);
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 5; ++j) bad = bad || (ref.idx[j] != t.idx[j]) || (ref.d2[j] != t.d2[j] && !(ref.pos[j] < 0 && t.idx[j] == 0x7FFFFFFF));
      atomicAdd(&a.dbg[0], 1u);
      if (bad) { if (atomicAdd(&a.dbg[1], 1u) == 0u) { a.dbg[4] = (unsigned)gi; a.dbg[5] = (unsigned)c; a.dbg[6] = (unsigned)t.count(); a.dbg[7] = (unsigned)ref.count(); } }
    }
    if (a.dbg && tid == 0) {
      atomicAdd(&a.dbg[2], 1u);
      // [8] total item cycles / 64 (before the self-check), [9] TMA wait, [10] fine sort, [11] search, [12] queries <= 8
      atomicAdd(&a.dbg[8], (unsigned)((clock64() - tk0) >> 6)); atomicAdd(&a.dbg[9], (unsigned)(tk_wait >> 6));
      atomicAdd(&a.dbg[10], (unsigned)(tk_sort >> 6)); atomicAdd(&a.dbg[11], (unsigned)(tk_search >> 6));
      if (wk.qcnt <= 8) atomicAdd(&a.dbg[12], 1u);
      atomicAdd(&a.dbg[14], (unsigned)(tk_p1 >> 6)); atomicAdd(&a.dbg[15], (unsigned)(tk_p2 >> 6));
    }
    // ---- fit + lazy GNC weight update + outputs (as k_correspond) ----
    if (owner) {
      const int k = t.count();
      double nb[5][3];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        nb[j][0] = ctx.origin[0] + (double)t.x[j]; nb[j][1] = ctx.origin[1] + (double)t.y[j]; nb[j][2] = ctx.origin[2] + (double)t.z[j];
        if (j >= k) nb[j][0] = nb[j][1] = nb[j][2] = 0.0;
      }
      double prim[6];
      const unsigned char flag = fit_neighbours<5>(ctx, c, k, t.d2[0], nb, prim);
      double wv = 1.0;
      if (st->outer != 0) {
        wv = ctx.w[gi];
        const double res = ctx.slot[gi];
        if (res != 0.0) {
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
      if ((flag & kFlagCounted) && ctx.maxnum[c] < ctx.n[c])
        atomicAdd(&ctx.blk_count[buf * ctx.blk_cap + ctx.blk_off[c] + il / kBlk], 1);
    }
  }
}

}  // namespace tloam
