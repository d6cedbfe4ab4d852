// Cluster K-split version of the BLSTM backward recurrence (row a1; semantics in blstm.cu).
//
// dh_{t-1}[B,H] = dz_t[B,4H] . Kh^T: with the flat partition of blstm.cu every CTA owns 8 outputs and must
// stream the whole dz_t (1 MB at cfg-3) every step, in 64-row chunks with a barrier and an 8-way smem
// reduction per chunk -- measured 32 us per step, loads and FFMA not overlapping.  Here
//   * a thread-block CLUSTER of 8 CTAs owns 8*HS hidden units (64 at H=512); CTA r of every cluster multiplies
//     only the K-slice of dz that cluster r of the same direction produces (4 gates x 64 units = 256 columns,
//     128 KB per step instead of 1 MB, fetched as four 32 KB bulk copies through a 2-stage ring) against a
//     resident [256 x 64] slice of Kh^T -- a straight, barrier-free 256-deep FFMA loop with no k-split;
//   * the 8 partial [B x 64] products of a cluster are reduce-scattered through distributed shared memory:
//     warp w of every CTA holds exactly the columns CTA w owns and stores them into CTA w's receive buffer
//     (st.shared::cluster, 32 KB in per CTA), one cluster barrier, then each CTA sums its 8 partials in a fixed
//     order (bit-reproducible) and does the pointwise gate gradients for its own HS units as before;
//   * dz_t is handed between clusters through L2 with one release/acquire counter PER CLUSTER, so a CTA only
//     waits for the producers of its own K-slice, not for the whole direction.
// Only 15 clusters of 8 are co-resident on a B200 (16 needed), so the kernel is also built for clusters of 4
// (K-slice of 512 columns = 256 KB per step, two k-split warp groups per CTA); the host picks the largest that fits.
#include "cl_common.cuh"
#include "blstm_cl.h"
#include <stdlib.h>
#include <string.h>

namespace nabu {
namespace {

// CLS = cluster size (8 or 4).  Per direction there are H/HS = 64 CTAs = 64/CLS clusters.  CTA r of a cluster
// multiplies K-slice r = the dz columns produced by clusters [r*CPS, (r+1)*CPS) of its direction (CPS = 64/CLS/CLS).
// The 8 warps are KH = 8/CLS k-split groups x CLS destination ranks: warp w owns the HS output columns of rank
// w % CLS and the k-range w / CLS of the slice, so its partial goes to receive slot r*KH + w/CLS of that rank.
template <int TBT, int HS, int CLS>
__global__ void __launch_bounds__(CL_THREADS, 1)
blstm_rec_bwd_cluster_kernel(const ClParams p) {
  constexpr int BT = 16 * TBT;             // batch rows (one tile)
  constexpr int NC = CLS * HS;             // hidden units (= output columns) per cluster
  constexpr int KH = CL_WARPS / CLS;       // k-split groups per CTA
  constexpr int CPS = 64 / CLS / CLS;      // producer clusters per K-slice
  constexpr int KW = 32 * HS;              // k rows per warp group per step (= 4H / 8)
  constexpr int RB = 64 / KH;              // rows per ring sub-block; a stage = KH sub-blocks = 64 rows = 32 KB at BT=128
  constexpr int NBLK = KW / RB;            // stages per step
  constexpr int CW = HS >= 2 ? HS / 2 : 1; // output columns per lane
  constexpr int PAIRS = BT * HS;
  constexpr int PP = (PAIRS + CL_THREADS - 1) / CL_THREADS;
  constexpr int SLAB = 4 * NC;             // exchange rows per producer cluster
  extern __shared__ __align__(16) float smem[];
  float* Wl = smem;                                   // [CPS*SLAB][NC]
  float* ring = Wl + CPS * SLAB * NC;                 // [2 stages][KH][RB][BT]
  float* rbuf = ring + 2 * KH * RB * BT;              // [2 parity][8 slots][BT][HS]
  __shared__ __align__(8) uint64_t full_bar[2];

  const int H = p.H, H4 = 4 * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_dir = H / HS;                         // 64 CTAs per direction
  const int dir = blockIdx.x / per_dir;
  const int q = (blockIdx.x % per_dir) / CLS;         // cluster within the direction
  const int r = blockIdx.x % CLS;                     // rank in the cluster == K-slice
  const int j0 = (q * CLS + r) * HS;                  // own hidden units
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  const float* cells = p.cells[dir];
  unsigned* cnt = p.counters + dir * 16;
  float* dzx = p.xchg + (size_t)dir * 2 * H4 * BT;    // [2 parity][cluster][4][NC][BT]
  float* dcb = p.dcbuf + (size_t)dir * BT * H;

  // Wl[kl][n] = Kh[NC*q + n][g*H + NC*(r*CPS + cl) + u]   with kl = (cl*4 + g)*NC + u
  for (int i = tid; i < CPS * SLAB * NC; i += CL_THREADS) {
    const int n = i % NC, kl = i / NC;
    const int cl = kl / SLAB, g = (kl % SLAB) / NC, u = kl % NC;
    Wl[i] = Kh[(size_t)(NC * q + n) * H4 + g * H + NC * (r * CPS + cl) + u];
  }
  if (tid == 0) {
    cb_init(&full_bar[0], 1);
    cb_init(&full_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();                                  // peers' smem exists before anyone stores into it

  const int bg = lane & 15, jj = lane >> 4;
  const int dst_rank = warp % CLS, kh = warp / CLS;
  const int ncol0 = HS * dst_rank + jj * CW;           // first of this lane's output columns
  const uint32_t rbuf_remote = map_to_rank(s_u32(rbuf), (uint32_t)dst_rank);
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned useq = 0;                                   // ring stages consumed so far

  int iter = 0;
  for (int s = p.T - 1; s >= 0; --s, ++iter) {
    const float* dzprev = dzx + (size_t)((iter + 1) & 1) * H4 * BT;
    float* dznext = dzx + (size_t)(iter & 1) * H4 * BT;
    float* rb = rbuf + (size_t)(iter & 1) * 8 * BT * HS;
    CL_STAMP(iter, 0);
    // ---- prefetch pointwise operands -----------------------------------------------------------
    float gt[PP][4], ct[PP], cprev[PP], dyv[PP], dcr[PP];
    int tb[PP];
    bool valid[PP];
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const int pr = tid + k * CL_THREADS;
      const int jl = pr % HS, b = pr / HS;
      valid[k] = false; tb[k] = 0; ct[k] = cprev[k] = dyv[k] = dcr[k] = 0.f;
      gt[k][0] = gt[k][1] = gt[k][2] = gt[k][3] = 0.f;
      if (pr < PAIRS && b < p.B) {
        const int L = p.len[b];
        valid[k] = s < L;
        const int t = valid[k] ? (dir ? L - 1 - s : s) : s;
        tb[k] = t;
        if (valid[k]) {
          const float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
#pragma unroll
          for (int g = 0; g < 4; ++g) gt[k][g] = __ldcg(gp + g * H);
          ct[k] = __ldcg(cells + ((size_t)b * p.T + t) * H + j0 + jl);
          if (s > 0) cprev[k] = __ldcg(cells + ((size_t)b * p.T + (dir ? t + 1 : t - 1)) * H + j0 + jl);
          dyv[k] = __ldcg(p.dy + ((size_t)b * p.yT + t) * 2 * H + dir * H + j0 + jl);
          if (iter > 0) dcr[k] = __ldcg(dcb + (size_t)b * H + j0 + jl);
        }
      }
    }

    if (iter > 0) {
      // ---- partial[b, n] = sum over my K-slice of dz_{s+1}[k][b] * Wl[k][n] ------------------------
      float acc[TBT][CW];
#pragma unroll
      for (int i = 0; i < TBT; ++i)
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[i][c] = 0.f;
      const float* slab = dzprev + (size_t)r * CPS * SLAB * BT;       // rows of my K-slice, contiguous
      // stage `blk`: for every k-split group kh2 the RB rows [kh2*KW + blk*RB, +RB) of the slice
      auto issue = [&](int blk, unsigned seq) {
        uint64_t* bar = &full_bar[seq & 1];
        cb_expect_tx(bar, KH * RB * BT * 4);
#pragma unroll
        for (int kh2 = 0; kh2 < KH; ++kh2)
          cb_bulk(ring + ((size_t)(seq & 1) * KH + kh2) * RB * BT, slab + ((size_t)kh2 * KW + (size_t)blk * RB) * BT,
                  RB * BT * 4, bar);
      };
      if (tid == 0) {
        const unsigned target = (unsigned)CLS * (unsigned)iter;
        for (int c = 0; c < CPS; ++c)
          while (ld_acquire_gpu(cnt + r * CPS + c) < target) { }
        CL_STAMP(iter, 1);
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        issue(0, useq);
        if (NBLK > 1) issue(1, useq + 1);
      }
#pragma unroll 1
      for (int blk = 0; blk < NBLK; ++blk, ++useq) {
        cb_wait(&full_bar[useq & 1], (useq >> 1) & 1);
        if (blk == 0) CL_STAMP(iter, 2);
        const float* hs_ = ring + ((size_t)(useq & 1) * KH + kh) * RB * BT;
        const float* ws_ = Wl + ((size_t)kh * KW + (size_t)blk * RB) * NC + ncol0;
#pragma unroll 4
        for (int kk = 0; kk < RB; ++kk) {
          float w[CW];
          if (CW == 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(ws_ + kk * NC);
            w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[CW - 1] = w4.w;
          } else if (CW == 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(ws_ + kk * NC);
            w[0] = w2.x; w[CW - 1] = w2.y;
          } else {
            w[0] = ws_[kk * NC];
          }
          float hv[TBT];
#pragma unroll
          for (int v = 0; v < TBT / 4; ++v) {
            const float4 t4 = *reinterpret_cast<const float4*>(hs_ + kk * BT + cl_row<TBT>(bg, v * 4));
            hv[v * 4 + 0] = t4.x; hv[v * 4 + 1] = t4.y; hv[v * 4 + 2] = t4.z; hv[v * 4 + 3] = t4.w;
          }
#pragma unroll
          for (int i = 0; i < TBT; ++i)
#pragma unroll
            for (int c = 0; c < CW; ++c) acc[i][c] = fmaf(hv[i], w[c], acc[i][c]);
        }
        if (blk + 2 < NBLK) {
          __syncthreads();                              // every warp is done with this ring stage
          if (tid == 0) issue(blk + 2, useq);           // (useq + 2) & 1 == useq & 1
        }
      }
      CL_STAMP(iter, 3);
      // ---- reduce-scatter: my columns belong to CTA dst_rank, receive slot r*KH + kh -------------------------
      const uint32_t dst = rbuf_remote + (uint32_t)(((size_t)(iter & 1) * 8 + r * KH + kh) * BT * HS) * 4u;
#pragma unroll
      for (int i = 0; i < TBT; ++i) {
        const uint32_t a = dst + (uint32_t)(cl_row<TBT>(bg, i) * HS + jj * CW) * 4u;
        if (CW == 4) st_cluster_v4(a, acc[i][0], acc[i][1], acc[i][2], acc[i][CW - 1]);
        else if (CW == 2) st_cluster_v2(a, acc[i][0], acc[i][CW - 1]);
        else st_cluster_f32(a, acc[i][0]);
      }
      CL_STAMP(iter, 4);
      cluster_sync_all();
      CL_STAMP(iter, 5);
    }

    // ---- pointwise gate gradients for my HS units -------------------------------------------------------
    float* dzmine = dznext + ((size_t)q * SLAB) * BT;                  // my cluster's slab
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const int pr = tid + k * CL_THREADS;
      const int jl = pr % HS, b = pr / HS;
      if (pr < PAIRS && b < p.B) {
        float dh = dyv[k];
        if (iter > 0) {
#pragma unroll
          for (int src = 0; src < 8; ++src) dh += rb[((size_t)src * BT + b) * HS + jl];
        }
        float dz[4] = {0.f, 0.f, 0.f, 0.f};
        float dcn = 0.f;
        if (valid[k]) {
          const float ig = gt[k][0], gg = gt[k][1], fg = gt[k][2], og = gt[k][3];
          const float tc = tanhf(ct[k]);
          const float d_o = dh * tc;
          const float dc = dcr[k] + dh * og * (1.f - tc * tc);
          dz[0] = dc * gg * ig * (1.f - ig);
          dz[1] = dc * ig * (1.f - gg * gg);
          dz[2] = dc * cprev[k] * fg * (1.f - fg);
          dz[3] = d_o * og * (1.f - og);
          dcn = dc * fg;
        }
        const int t = tb[k];
        float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __stcg(gp + g * H, dz[g]);
          __stcg(dzmine + ((size_t)g * NC + r * HS + jl) * BT + b, dz[g]);
          dbacc[g] += dz[g];
        }
        __stcg(dcb + (size_t)b * H + j0 + jl, dcn);
      }
    }
    // ---- publish dz_s of this CTA (per-cluster counter) ---------------------------------------------------
    CL_STAMP(iter, 6);
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    CL_STAMP(iter, 7);
    __syncthreads();
    CL_STAMP(iter, 8);
    if (tid == 0) red_release_gpu_add(cnt + q, 1u);
    CL_STAMP(iter, 9);
  }

  // bias gradient: every thread's pairs share jl = tid % HS; fixed-order sum over threads
  {
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 4; ++g) ring[tid * 4 + g] = dbacc[g];
    __syncthreads();
    if (tid < 4 * HS) {
      const int g = tid / HS, j = tid % HS;
      float sum = 0.f;
      for (int i = j; i < CL_THREADS; i += HS) sum += ring[i * 4 + g];
      p.dbpart[((size_t)dir * 8) * H4 + g * H + j0 + j] = sum;
    }
  }
  cluster_sync_all();                                    // nobody exits while a peer may still store into it
}

template <int TBT, int HS, int CLS>
int launch_cl(const ClParams& p, cudaStream_t stream, bool* launched) {
  constexpr int BT = 16 * TBT, NC = CLS * HS, KH = CL_WARPS / CLS, CPS = 64 / CLS / CLS, RB = 64 / KH;
  const size_t smem = ((size_t)CPS * 4 * NC * NC + 2 * KH * RB * BT + 2 * 8 * BT * HS) * sizeof(float);
  auto* fn = blstm_rec_bwd_cluster_kernel<TBT, HS, CLS>;
  *launched = false;
  if (smem > (size_t)max_smem_optin()) return 0;
  NABU_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (p.H / HS));
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  int nclusters = 0;
  const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg);
  if (getenv("NABU_DEBUG"))
    fprintf(stderr, "[nabu] bwd cluster kernel TBT=%d HS=%d CLS=%d: smem %zu B, max active clusters %d (%s), need %d\n", TBT,
            HS, CLS, smem, nclusters, cudaGetErrorString(oe), (int)cfg.gridDim.x / CLS);
  if (oe != cudaSuccess || nclusters * CLS < (int)cfg.gridDim.x) {
    cudaGetLastError();
    return 0;                                            // not co-resident on this device
  }
  KernelScope ks(CLS == 8 ? "blstm_rec_bwd_cluster8" : "blstm_rec_bwd_cluster4", stream);
  ClParams pt = p;
  pt.trace = trace_buffer();
  NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt));
  trace_dump("bwd", pt.trace, stream);
  *launched = true;
  return 0;
}

// largest co-resident cluster size first (8 clusters of 8 per direction need 16 free 8-SM groups; B200 offers 15)
template <int TBT, int HS>
int launch_any(const ClParams& p, cudaStream_t stream, bool* launched) {
  if (int e = launch_cl<TBT, HS, 8>(p, stream, launched)) return e;
  if (*launched) return 0;
  return launch_cl<TBT, HS, 4>(p, stream, launched);
}


// ---------------------------------------------------------------------------------------------------------
// forward: z_t[B, 4H] = Gx_t + h_{t-1}[B, H] . Kh.  Same K-split, clusters of 4: CTA r of a cluster multiplies the
// h rows [r*H/4, (r+1)*H/4) (64 KB per step at cfg-3 instead of 256 KB) against the resident [H/4 x 16*HS] block of
// Kh that feeds the 4 gates of the cluster's 4*HS units, then the four partial [B x 16*HS] products are
// reduce-scattered over DSMEM (each CTA receives 3 x [B x 4*HS]) and summed in rank order.  Warp w computes the
// columns of destination rank w/2, gates 2*(w%2) + {0,1}; a lane owns 8 batch rows x HS columns (64 accumulators
// at HS = 8: 4 LDS.128 per 64 FFMA).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_cl(float x) { return 1.f / (1.f + expf(-x)); }

template <int TBT, int HS>
__global__ void __launch_bounds__(CL_THREADS, 1)
blstm_rec_fwd_cluster_kernel(const ClParams p) {
  constexpr int CLS = 4;
  constexpr int BT = 16 * TBT;
  constexpr int NC = CLS * HS;             // hidden units per cluster
  constexpr int GC = 4 * NC;               // gate columns per cluster
  constexpr int KS = 64 * HS / CLS;        // h rows per K-slice (= H / CLS)
  constexpr int CPS = 64 / CLS / CLS;      // producer clusters per K-slice
  constexpr int RB = KS < 32 ? KS : 32;    // rows per ring stage
  constexpr int NBLK = KS / RB;
  constexpr int PAIRS = BT * HS;
  constexpr int PP = (PAIRS + CL_THREADS - 1) / CL_THREADS;
  extern __shared__ __align__(16) float smem[];
  float* Wl = smem;                        // [KS][GC]
  float* ring = Wl + KS * GC;              // [2][RB][BT]
  float* rbuf = ring + 2 * RB * BT;        // [2 parity][CLS src][BT][4 gates][HS]
  __shared__ __align__(8) uint64_t full_bar[2];

  const int H = p.H, H4 = 4 * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_dir = H / HS;
  const int dir = blockIdx.x / per_dir;
  const int q = (blockIdx.x % per_dir) / CLS;
  const int r = blockIdx.x % CLS;
  const int j0 = (q * CLS + r) * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  float* cells = const_cast<float*>(p.cells[dir]);
  unsigned* cnt = p.counters + dir * 16;
  float* hx = p.xchg + (size_t)dir * 2 * H * BT;      // [2 parity][H][BT]

  // Wl[k][d*4*HS + g*HS + u] = Kh[r*KS + k][g*H + NC*q + d*HS + u]
  for (int i = tid; i < KS * GC; i += CL_THREADS) {
    const int n = i % GC, k = i / GC;
    const int d = n / (4 * HS), g = (n / HS) % 4, u = n % HS;
    Wl[i] = Kh[(size_t)(r * KS + k) * H4 + g * H + NC * q + d * HS + u];
  }
  if (tid == 0) {
    cb_init(&full_bar[0], 1);
    cb_init(&full_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();

  const int bg = lane & 15, jj = lane >> 4;
  const int dst_rank = warp >> 1, gate = 2 * (warp & 1) + jj;
  const int ncol0 = dst_rank * 4 * HS + gate * HS;
  const uint32_t rbuf_remote = map_to_rank(s_u32(rbuf), (uint32_t)dst_rank);
  unsigned useq = 0;

  for (int s = 0; s < p.T; ++s) {
    const float* hprev = hx + (size_t)((s + 1) & 1) * H * BT;
    float* hnext = hx + (size_t)(s & 1) * H * BT;
    float* rb = rbuf + (size_t)(s & 1) * CLS * BT * 4 * HS;
    CL_STAMP(s, 0);
    // ---- prefetch pointwise operands -------------------------------------------------------------
    float gx[PP][4], cprev[PP];
    int tb[PP];
    bool valid[PP];
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const int pr = tid + k * CL_THREADS;
      const int jl = pr % HS, b = pr / HS;
      valid[k] = false; tb[k] = 0; cprev[k] = 0.f;
      gx[k][0] = gx[k][1] = gx[k][2] = gx[k][3] = 0.f;
      if (pr < PAIRS && b < p.B) {
        const int L = p.len[b];
        valid[k] = s < L;
        const int t = valid[k] ? (dir ? L - 1 - s : s) : s;
        tb[k] = t;
        if (valid[k]) {
          const float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
#pragma unroll
          for (int g = 0; g < 4; ++g) gx[k][g] = __ldcg(gp + g * H);
          if (s > 0) cprev[k] = __ldcg(cells + ((size_t)b * p.T + (dir ? t + 1 : t - 1)) * H + j0 + jl);
        }
      }
    }

    if (s > 0) {
      float acc[TBT][HS];
#pragma unroll
      for (int i = 0; i < TBT; ++i)
#pragma unroll
        for (int c = 0; c < HS; ++c) acc[i][c] = 0.f;
      const float* slab = hprev + (size_t)r * KS * BT;
      auto issue = [&](int blk, unsigned seq) {
        uint64_t* bar = &full_bar[seq & 1];
        cb_expect_tx(bar, RB * BT * 4);
        cb_bulk(ring + (size_t)(seq & 1) * RB * BT, slab + (size_t)blk * RB * BT, RB * BT * 4, bar);
      };
      if (tid == 0) {
        const unsigned target = (unsigned)CLS * (unsigned)s;
        for (int c = 0; c < CPS; ++c)
          while (ld_acquire_gpu(cnt + r * CPS + c) < target) { }
        CL_STAMP(s, 1);
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        issue(0, useq);
        if (NBLK > 1) issue(1, useq + 1);
      }
#pragma unroll 1
      for (int blk = 0; blk < NBLK; ++blk, ++useq) {
        cb_wait(&full_bar[useq & 1], (useq >> 1) & 1);
        if (blk == 0) CL_STAMP(s, 2);
        const float* hs_ = ring + (size_t)(useq & 1) * RB * BT;
        const float* ws_ = Wl + (size_t)blk * RB * GC + ncol0;
#pragma unroll 2
        for (int kk = 0; kk < RB; ++kk) {
          float w[HS];
          if (HS >= 4) {
#pragma unroll
            for (int v = 0; v < HS / 4; ++v) {
              const float4 w4 = *reinterpret_cast<const float4*>(ws_ + kk * GC + v * 4);
              w[v * 4 + 0] = w4.x; w[v * 4 + 1] = w4.y; w[v * 4 + 2] = w4.z; w[v * 4 + 3] = w4.w;
            }
          } else {
            const float2 w2 = *reinterpret_cast<const float2*>(ws_ + kk * GC);
            w[0] = w2.x; w[HS - 1] = w2.y;
          }
          float hv[TBT];
#pragma unroll
          for (int v = 0; v < TBT / 4; ++v) {
            const float4 t4 = *reinterpret_cast<const float4*>(hs_ + kk * BT + cl_row<TBT>(bg, v * 4));
            hv[v * 4 + 0] = t4.x; hv[v * 4 + 1] = t4.y; hv[v * 4 + 2] = t4.z; hv[v * 4 + 3] = t4.w;
          }
#pragma unroll
          for (int i = 0; i < TBT; ++i)
#pragma unroll
            for (int c = 0; c < HS; ++c) acc[i][c] = fmaf(hv[i], w[c], acc[i][c]);
        }
        if (blk + 2 < NBLK) {
          __syncthreads();
          if (tid == 0) issue(blk + 2, useq);
        }
      }
      CL_STAMP(s, 3);
      // ---- reduce-scatter: receive slot r of CTA dst_rank -------------------------------------------------
      const uint32_t dst = rbuf_remote + (uint32_t)(((size_t)(s & 1) * CLS + r) * BT * 4 * HS) * 4u;
#pragma unroll
      for (int i = 0; i < TBT; ++i) {
        const uint32_t a = dst + (uint32_t)((cl_row<TBT>(bg, i) * 4 + gate) * HS) * 4u;
        if (HS >= 4) {
#pragma unroll
          for (int v = 0; v < HS / 4; ++v)
            st_cluster_v4(a + v * 16, acc[i][v * 4 + 0], acc[i][v * 4 + 1], acc[i][v * 4 + 2], acc[i][v * 4 + 3]);
        } else {
          st_cluster_v2(a, acc[i][0], acc[i][HS - 1]);
        }
      }
      CL_STAMP(s, 4);
      cluster_sync_all();
      CL_STAMP(s, 5);
    }

    // ---- pointwise cell update for my HS units ----------------------------------------------------------
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const int pr = tid + k * CL_THREADS;
      const int jl = pr % HS, b = pr / HS;
      if (pr < PAIRS && b < p.B) {
        float z[4] = {gx[k][0], gx[k][1], gx[k][2], gx[k][3]};
        if (s > 0) {
#pragma unroll
          for (int src = 0; src < CLS; ++src)
#pragma unroll
            for (int g = 0; g < 4; ++g) z[g] += rb[(((size_t)src * BT + b) * 4 + g) * HS + jl];
        }
        const float ig = sigmoid_cl(z[0]);
        const float gg = tanhf(z[1]);
        const float fg = sigmoid_cl(z[2] + 1.0f);
        const float og = sigmoid_cl(z[3]);
        const float cn = cprev[k] * fg + ig * gg;
        const float hn = tanhf(cn) * og;
        const int t = tb[k];
        if (valid[k]) {
          float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
          __stcg(gp, ig); __stcg(gp + H, gg); __stcg(gp + 2 * H, fg); __stcg(gp + 3 * H, og);
          __stcg(cells + ((size_t)b * p.T + t) * H + j0 + jl, cn);
        }
        __stcg(p.y + ((size_t)b * p.yT + t) * 2 * H + dir * H + j0 + jl, valid[k] ? hn : 0.f);
        __stcg(hnext + (size_t)(j0 + jl) * BT + b, valid[k] ? hn : 0.f);
      }
    }
    CL_STAMP(s, 6);
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    CL_STAMP(s, 7);
    __syncthreads();
    CL_STAMP(s, 8);
    if (tid == 0) red_release_gpu_add(cnt + q, 1u);
    CL_STAMP(s, 9);
  }
  cluster_sync_all();
}

template <int TBT, int HS>
int launch_fwd_cl(const ClParams& p, cudaStream_t stream, bool* launched) {
  constexpr int CLS = 4, BT = 16 * TBT, KS = 64 * HS / CLS, GC = 16 * HS, RB = KS < 32 ? KS : 32;
  const size_t smem = ((size_t)KS * GC + 2 * RB * BT + 2 * CLS * BT * 4 * HS) * sizeof(float);
  auto* fn = blstm_rec_fwd_cluster_kernel<TBT, HS>;
  *launched = false;
  if (smem > (size_t)max_smem_optin()) return 0;
  NABU_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (p.H / HS));
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  int nclusters = 0;
  const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg);
  if (getenv("NABU_DEBUG"))
    fprintf(stderr, "[nabu] fwd cluster kernel TBT=%d HS=%d: smem %zu B, max active clusters %d (%s), need %d\n", TBT, HS,
            smem, nclusters, cudaGetErrorString(oe), (int)cfg.gridDim.x / CLS);
  if (oe != cudaSuccess || nclusters * CLS < (int)cfg.gridDim.x) {
    cudaGetLastError();
    return 0;
  }
  KernelScope ks("blstm_rec_fwd_cluster4", stream);
  ClParams pt = p;
  pt.trace = trace_buffer();
  NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt));
  trace_dump("fwd", pt.trace, stream);
  *launched = true;
  return 0;
}

}  // namespace

bool blstm_bwd_cluster_eligible(int B, int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NABU_REC_BWD");
    enabled = (e && strcmp(e, "flat") == 0) ? 0 : 1;
  }
  if (!enabled) return false;
  return B <= 128 && B > 0 && (H == 128 || H == 256 || H == 512);
}

int blstm_rec_bwd_cluster(const float* const kernel[2], float* const gates[2], const float* const cells[2],
                          const float* dy, float* dbpart, float* xchg, float* dcbuf, unsigned* counters,
                          const int* len, int B, int T, int yT, int D, int H, cudaStream_t stream, bool* launched) {
  ClParams p = {};
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.dy = dy; p.dbpart = dbpart; p.xchg = xchg; p.dcbuf = dcbuf; p.counters = counters; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  const int hs = H / 64;
  const bool small = B <= 64;
  if (hs == 8) return small ? launch_any<4, 8>(p, stream, launched) : launch_any<8, 8>(p, stream, launched);
  if (hs == 4) return small ? launch_any<4, 4>(p, stream, launched) : launch_any<8, 4>(p, stream, launched);
  return small ? launch_any<4, 2>(p, stream, launched) : launch_any<8, 2>(p, stream, launched);
}

bool blstm_fwd_cluster_eligible(int B, int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NABU_REC_FWD");
    enabled = (e && strcmp(e, "flat") == 0) ? 0 : 1;
  }
  if (!enabled) return false;
  return B <= 128 && B > 0 && (H == 128 || H == 256 || H == 512);
}

int blstm_rec_fwd_cluster(const float* const kernel[2], float* const gates[2], float* const cells[2], float* y,
                          float* xchg, unsigned* counters, const int* len, int B, int T, int yT, int D, int H,
                          cudaStream_t stream, bool* launched) {
  ClParams p = {};
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.y = y; p.xchg = xchg; p.counters = counters; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  const int hs = H / 64;
  const bool small = B <= 64;
  if (hs == 8) return small ? launch_fwd_cl<4, 8>(p, stream, launched) : launch_fwd_cl<8, 8>(p, stream, launched);
  if (hs == 4) return small ? launch_fwd_cl<4, 4>(p, stream, launched) : launch_fwd_cl<8, 4>(p, stream, launched);
  return small ? launch_fwd_cl<4, 2>(p, stream, launched) : launch_fwd_cl<8, 2>(p, stream, launched);
}

}  // namespace nabu
