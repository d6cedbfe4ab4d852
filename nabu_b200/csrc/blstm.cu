// Persistent bidirectional-LSTM recurrence for sm_100a (replaces
// nabu/neuralnetworks/components/layer.py:8-51 = LayerNormBasicLSTMCell(layer_norm=False)
// under bidirectional_dynamic_rnn).
//
// Layer forward  = input projection  Gx[d] = X.Kx[d] + b[d]   (one dense GEMM per direction,
//                  all T at once -- the reference re-does this contraction every time step)
//                + ONE cooperative kernel that walks the T serial steps of BOTH directions:
//                  every CTA owns `HS` hidden units (4*HS gate columns) of one direction, keeps
//                  that slice of the recurrent matrix Kh resident in shared memory for the whole
//                  sequence, and per step multiplies the previous hidden state (streamed through
//                  a 3-stage cp.async ring from a [H][B] exchange buffer in L2) against it.
//                  CTAs of one direction hand h_t to each other through that buffer and a
//                  monotonic release/acquire counter -- no grid-wide barrier, and the two
//                  directions never wait for each other.
// Layer backward = the mirrored cooperative kernel (dz_t exchanged instead of h_t, Kh^T slice
//                  resident) followed by the three batched GEMMs dKx = X^T.dZ, dKh = Hprev^T.dZ,
//                  dX = dZ.Kx^T.
//
// Semantics restated from TF-1.8: gate order i,j,f,o; forget bias +1.0 at
// run time; zero initial state; outputs are 0 and state is frozen for t >= len[b]; the backward
// direction visits t = len[b]-1-s at step s.
#include "common.cuh"
#include "gemm.h"
#include "blstm_tc.h"
#include "blstm_cl.h"
#include <cuda_fp16.h>
#include <algorithm>
#include "nabu_b200.h"
#include <stdlib.h>
#include <string.h>

namespace nabu {
namespace {

constexpr int RNN_THREADS = 256;
constexpr int RNN_WARPS = RNN_THREADS / 32;
constexpr int KC = 64;        // rows of the exchanged operand per pipeline stage
constexpr int STAGES = 3;

// ---- mbarrier + 1-D bulk async copy (TMA without a tensor map): the ring stages are contiguous slabs of the
// exchange buffer, so ONE thread issues ONE instruction per stage instead of 256 threads x 16 cp.async.
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rb_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rb_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rb_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(sm_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sm_u32(dst)), "l"(src), "r"(bytes), "r"(sm_u32(bar)) : "memory");
}
// Stage of the current product: `kc` rows of one batch tile's [R][BT] exchange slab (tile-major layout, so a
// stage is one contiguous slab and ONE thread issues ONE bulk copy for it).
__device__ __forceinline__ void ring_issue(float* dst, const float* src, int kc, int BT, uint64_t* bar, int tid) {
  if (tid == 0) {
    rb_expect_tx(bar, (unsigned)(kc * BT * 4));
    bulk_g2s(dst, src, (unsigned)(kc * BT * 4), bar);
  }
}

struct RecParams {
  const float* kernel[2];   // [(D+H), 4H] per direction
  float* gates[2];          // [B, T, 4H]  fwd: in = Gx, out = activated i,g,f,o ; bwd: in = gates, out = dZ
  float* cells[2];          // [B, T, H]
  float* y;                 // fwd: output [B, yT, 2H]
  const float* dy;          // bwd: [B, yT, 2H]
  float* dbias[2];          // bwd: [4H]
  float* dbpart;            // bwd: per batch-group partials [2 dir][8 grp][4H]
  float* xchg;              // exchange buffer [2 dir][2 parity][tile][R][BT]  (R = H fwd, 4H bwd)
  float* dcbuf;             // bwd: carried dc [2 dir][Bp][H]
  unsigned* counters;       // [2 dir][8 groups], zeroed before launch
  const int* len;           // [B]
  int B, Bp, T, yT, D, H;
  int nsl;                  // CTAs (hidden slices) per direction and batch group
  int ngrp;                 // batch groups: CTA (dir, grp, slice) owns batch tiles grp, grp+ngrp, ...
  int dir0;                 // first direction handled by this launch
  int kc, stages;           // bwd ring: rows per stage, number of stages
  int fstages;              // fwd ring stages (KC rows each)
  int dbg;                  // profiling aid (NABU_REC_DBG): bit0 = skip the ring loads, bit1 = skip the FMAs
};

// Batch rows of a thread inside a tile.  For TBT == 8 the rows are two groups of 4 (bg*4.. and 64+bg*4..)
// so that the 16 lanes of a half-warp read 256 contiguous bytes per LDS.128 (conflict-free); 8 contiguous
// rows per lane would put lanes 0 and 4 on the same banks (2-way conflict on every operand load).
template <int TBT>
__device__ __forceinline__ int tile_row(int bg, int r) {
  return TBT == 8 ? ((r >> 2) * 64 + bg * 4 + (r & 3)) : bg * TBT + r;
}

// Shared-memory carve-up (floats): W slice | ring stages | k-split partials
template <int TBT, int HS>
struct RecCfg {
  static constexpr int BT = 16 * TBT;               // batch rows per tile
  static constexpr int CG = HS / 2;                 // column groups (2 hidden units each)
  static constexpr int KS = RNN_WARPS / CG;         // k-split factor (warps per column group)
  static constexpr int PAIRS = BT * HS;             // (b, j) pairs per tile
  static constexpr int PP = (PAIRS + RNN_THREADS - 1) / RNN_THREADS;
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int TBT, int HS>
__global__ void __launch_bounds__(RNN_THREADS, (TBT <= 4) ? 2 : 1)
blstm_rec_fwd_kernel(const RecParams p) {
  using C = RecCfg<TBT, HS>;
  constexpr int BT = C::BT, CG = C::CG, KS = C::KS, PP = C::PP;
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H4 = 4 * p.H;
  float* Ws = smem;                                  // [H][HS][4]
  float* ring = Ws + (size_t)H * HS * 4;             // [STAGES][KC][BT]
  const int S = p.fstages;
  float* red = ring + (size_t)S * KC * BT;           // [KS][BT][HS][4]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dir = p.dir0 + blockIdx.x / (p.nsl * p.ngrp);
  const int grp = (blockIdx.x / p.nsl) % p.ngrp;
  const int slice = blockIdx.x % p.nsl;
  const int j0 = slice * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  float* cells = p.cells[dir];
  unsigned* counter = p.counters + dir * 8 + grp;
  float* hx = p.xchg + (size_t)dir * 2 * H * p.Bp;   // [2][H][Bp]

  // resident weight slice: Ws[k][jl][g] = Kh[k][g*H + j0 + jl]
  for (int i = tid; i < H * HS * 4; i += RNN_THREADS) {
    const int g = i & 3, jl = (i >> 2) % HS, k = i / (4 * HS);
    Ws[i] = Kh[(size_t)k * H4 + g * H + j0 + jl];
  }
  __syncthreads();

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) rb_init(&full_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned gq = 0;                                   // ring stages consumed so far (same in every thread)
  const int cg = warp % CG, ks = warp / CG;
  const int bg = lane & 15, jj = lane >> 4;
  const int jl_mm = cg * 2 + jj;                     // hidden unit of this thread in the matmul
  const int nchunks = H / KC;                        // host guarantees H % KC == 0
  const int kper = KC / KS;                          // k rows per warp per chunk
  const int ntile = (p.B + BT - 1) / BT;

  for (int s = 0; s < p.T; ++s) {
    const float* hprev = hx + (size_t)((s + 1) & 1) * H * p.Bp;   // written at step s-1
    float* hnext = hx + (size_t)(s & 1) * H * p.Bp;
    for (int tile = grp; tile < ntile; tile += p.ngrp) {
      const int b0 = tile * BT;
      // ---- prefetch the pointwise operands of this thread's (b, j) pairs -----------------------
      float gx[PP][4], cprev[PP];
      int tb[PP];
      bool valid[PP];
#pragma unroll
      for (int q = 0; q < PP; ++q) {
        const int pr = tid + q * RNN_THREADS;
        const int jl = pr % HS, b = b0 + pr / HS;
        valid[q] = false; tb[q] = 0; cprev[q] = 0.f;
        gx[q][0] = gx[q][1] = gx[q][2] = gx[q][3] = 0.f;
        if (pr < C::PAIRS && b < p.B) {
          const int L = p.len[b];
          valid[q] = s < L;
          const int t = valid[q] ? (dir ? L - 1 - s : s) : s;
          tb[q] = t;
          if (valid[q]) {
            const float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
#pragma unroll
            for (int g = 0; g < 4; ++g) gx[q][g] = __ldcg(gp + g * H);
            if (s > 0) cprev[q] = __ldcg(cells + ((size_t)b * p.T + (dir ? t + 1 : t - 1)) * H + j0 + jl);
          }
        }
      }

      // ---- recurrent product  z[b, 4 gates of jl] = sum_k h_{s-1}[k][b] * Ws[k][jl][:] ---------
      float acc[TBT][4];
#pragma unroll
      for (int r = 0; r < TBT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;

      if (s > 0) {
        if (tile == grp) {
          if (tid == 0) {
            const unsigned target = (unsigned)p.nsl * (unsigned)s;
            while (ld_acquire_gpu(counter) < target) { }
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");   // peers' generic stores -> our bulk (async proxy) loads
          }
          __syncthreads();
        }
        const unsigned g0 = gq;
        auto issue = [&](int c) {
          if (c < nchunks)
            ring_issue(ring + (size_t)((g0 + c) % S) * KC * BT, hprev + ((size_t)tile * H + (size_t)c * KC) * BT, KC, BT,
                       &full_bar[(g0 + c) % S], tid);
        };
        for (int c = 0; c < S - 1; ++c) issue(c);
        for (int c = 0; c < nchunks; ++c) {
          rb_wait(&full_bar[(g0 + c) % S], ((g0 + c) / S) & 1);
          __syncthreads();                           // everyone is done with stage c-1 -> its buffer may be refilled
          issue(c + S - 1);
          const float* hs_ = ring + (size_t)((g0 + c) % S) * KC * BT + (size_t)ks * kper * BT;
          const float* ws_ = Ws + ((size_t)(c * KC + ks * kper) * HS + jl_mm) * 4;
#pragma unroll 4
          for (int kk = 0; kk < kper; ++kk) {
            const float4 w = *reinterpret_cast<const float4*>(ws_ + (size_t)kk * HS * 4);
            float hv[TBT];
            if (TBT >= 4) {
#pragma unroll
              for (int v = 0; v < TBT / 4; ++v) {
                const float4 t4 = *reinterpret_cast<const float4*>(hs_ + kk * BT + tile_row<TBT>(bg, v * 4));
                hv[v * 4 + 0] = t4.x; hv[v * 4 + 1] = t4.y; hv[v * 4 + 2] = t4.z; hv[v * 4 + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int r = 0; r < TBT; ++r) hv[r] = hs_[kk * BT + bg * TBT + r];
            }
#pragma unroll
            for (int r = 0; r < TBT; ++r) {
              acc[r][0] = fmaf(hv[r], w.x, acc[r][0]);
              acc[r][1] = fmaf(hv[r], w.y, acc[r][1]);
              acc[r][2] = fmaf(hv[r], w.z, acc[r][2]);
              acc[r][3] = fmaf(hv[r], w.w, acc[r][3]);
            }
          }
        }
        gq = g0 + nchunks;
      }
      // ---- k-split partials -> shared ----------------------------------------------------------
#pragma unroll
      for (int r = 0; r < TBT; ++r) {
        float4* dst = reinterpret_cast<float4*>(red + (((size_t)ks * BT + tile_row<TBT>(bg, r)) * HS + jl_mm) * 4);
        *dst = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      }
      __syncthreads();

      // ---- pointwise cell update ---------------------------------------------------------------
#pragma unroll
      for (int q = 0; q < PP; ++q) {
        const int pr = tid + q * RNN_THREADS;
        const int jl = pr % HS, bl = pr / HS, b = b0 + bl;
        if (pr < C::PAIRS && b < p.B) {
          float z[4] = {gx[q][0], gx[q][1], gx[q][2], gx[q][3]};
#pragma unroll
          for (int k2 = 0; k2 < KS; ++k2) {
            const float4 v = *reinterpret_cast<const float4*>(red + (((size_t)k2 * BT + bl) * HS + jl) * 4);
            z[0] += v.x; z[1] += v.y; z[2] += v.z; z[3] += v.w;
          }
          const float ig = sigmoid_acc(z[0]);
          const float gg = tanhf(z[1]);
          const float fg = sigmoid_acc(z[2] + 1.0f);
          const float og = sigmoid_acc(z[3]);
          const float cn = cprev[q] * fg + ig * gg;
          const float hn = tanhf(cn) * og;
          const int t = tb[q];
          if (valid[q]) {
            float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
            __stcg(gp, ig); __stcg(gp + H, gg); __stcg(gp + 2 * H, fg); __stcg(gp + 3 * H, og);
            __stcg(cells + ((size_t)b * p.T + t) * H + j0 + jl, cn);
          }
          __stcg(p.y + ((size_t)b * p.yT + t) * 2 * H + dir * H + j0 + jl, valid[q] ? hn : 0.f);
          __stcg(hnext + ((size_t)tile * H + j0 + jl) * BT + bl, valid[q] ? hn : 0.f);
        }
      }
      __syncthreads();   // red[] and ring are reused by the next tile / step
    }
    // ---- publish h_s -----------------------------------------------------------------------------
    asm volatile("fence.proxy.async;" ::: "memory");   // our generic stores -> the peers' bulk (async proxy) loads
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_gpu_add(counter, 1u);
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int TBT, int HS>
__global__ void __launch_bounds__(RNN_THREADS, (TBT <= 4) ? 2 : 1)
blstm_rec_bwd_kernel(const RecParams p) {
  using C = RecCfg<TBT, HS>;
  constexpr int BT = C::BT, PP = C::PP;
  constexpr int CW = HS / 2;                         // output columns per thread in the matmul
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H4 = 4 * p.H;
  float* Ws = smem;                                  // [4H][HS]   Ws[k][jl] = Kh[j0+jl][k]
  float* ring = Ws + (size_t)H4 * HS;                // [STAGES][KC][BT]
  const int kc = p.kc;
  float* red = ring + (size_t)p.stages * kc * BT;    // [RNN_WARPS][BT][HS]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dir = p.dir0 + blockIdx.x / (p.nsl * p.ngrp);
  const int grp = (blockIdx.x / p.nsl) % p.ngrp;
  const int slice = blockIdx.x % p.nsl;
  const int j0 = slice * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  const float* cells = p.cells[dir];
  unsigned* counter = p.counters + dir * 8 + grp;
  float* dzx = p.xchg + (size_t)dir * 2 * H4 * p.Bp; // [2][4H][Bp]
  float* dcb = p.dcbuf + (size_t)dir * p.Bp * H;     // [Bp][H]

  for (int i = tid; i < H4 * HS; i += RNN_THREADS) {
    const int jl = i % HS, k = i / HS;
    Ws[i] = Kh[(size_t)(j0 + jl) * H4 + k];
  }
  __syncthreads();

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) rb_init(&full_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned gq = 0;
  const int bg = lane & 15, jj = lane >> 4;
  const int nchunks = H4 / kc;
  const int kper = kc / RNN_WARPS;                   // all warps split every chunk
  const int ntile = (p.B + BT - 1) / BT;
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};

  int iter = 0;
  for (int s = p.T - 1; s >= 0; --s, ++iter) {
    const float* dzprev = dzx + (size_t)((iter + 1) & 1) * H4 * p.Bp;   // published at iter-1 (step s+1)
    float* dznext = dzx + (size_t)(iter & 1) * H4 * p.Bp;
    for (int tile = grp; tile < ntile; tile += p.ngrp) {
      const int b0 = tile * BT;
      // ---- prefetch pointwise operands ---------------------------------------------------------
      float gt[PP][4], ct[PP], cprev[PP], dyv[PP], dcr[PP];
      int tb[PP];
      bool valid[PP];
#pragma unroll
      for (int q = 0; q < PP; ++q) {
        const int pr = tid + q * RNN_THREADS;
        const int jl = pr % HS, b = b0 + pr / HS;
        valid[q] = false; tb[q] = 0; ct[q] = cprev[q] = dyv[q] = dcr[q] = 0.f;
        gt[q][0] = gt[q][1] = gt[q][2] = gt[q][3] = 0.f;
        if (pr < C::PAIRS && b < p.B) {
          const int L = p.len[b];
          valid[q] = s < L;
          const int t = valid[q] ? (dir ? L - 1 - s : s) : s;
          tb[q] = t;
          if (valid[q]) {
            const float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
#pragma unroll
            for (int g = 0; g < 4; ++g) gt[q][g] = __ldcg(gp + g * H);
            ct[q] = __ldcg(cells + ((size_t)b * p.T + t) * H + j0 + jl);
            if (s > 0) cprev[q] = __ldcg(cells + ((size_t)b * p.T + (dir ? t + 1 : t - 1)) * H + j0 + jl);
            dyv[q] = __ldcg(p.dy + ((size_t)b * p.yT + t) * 2 * H + dir * H + j0 + jl);
            if (iter > 0) dcr[q] = __ldcg(dcb + (size_t)b * H + j0 + jl);
          }
        }
      }

      // ---- dh_rec[b, jl] = sum_k dz_{s+1}[k][b] * Kh[j0+jl][k] ---------------------------------
      float acc[TBT][CW];
#pragma unroll
      for (int r = 0; r < TBT; ++r)
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[r][c] = 0.f;

      if (iter > 0) {
        if (tile == grp) {
          if (tid == 0) {
            const unsigned target = (unsigned)p.nsl * (unsigned)iter;
            while (ld_acquire_gpu(counter) < target) { }
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");   // peers' generic stores -> our bulk (async proxy) loads
          }
          __syncthreads();
        }
        const unsigned g0 = gq;
        const int S = p.stages;
        auto issue = [&](int c) {
          if (c < nchunks && !(p.dbg & 1))
            ring_issue(ring + (size_t)((g0 + c) % S) * kc * BT, dzprev + ((size_t)tile * H4 + (size_t)c * kc) * BT, kc, BT,
                       &full_bar[(g0 + c) % S], tid);
        };
        // ring of S buffers with S-1 stages in flight
        for (int c = 0; c < S - 1; ++c) issue(c);
        for (int c = 0; c < nchunks; ++c) {
          if (!(p.dbg & 1)) rb_wait(&full_bar[(g0 + c) % S], ((g0 + c) / S) & 1);
          __syncthreads();                           // everyone is done with stage c-1 -> its buffer may be refilled
          issue(c + S - 1);
          const float* hs_ = ring + (size_t)((g0 + c) % S) * kc * BT + (size_t)warp * kper * BT;
          const float* ws_ = Ws + (size_t)(c * kc + warp * kper) * HS + jj * CW;
          if (p.dbg & 2) continue;
#pragma unroll 4
          for (int kk = 0; kk < kper; ++kk) {
            float w[CW];
            if (CW == 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(ws_ + kk * HS);
              w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[CW - 1] = w4.w;
            } else if (CW == 8) {
              const float4 w4 = *reinterpret_cast<const float4*>(ws_ + kk * HS);
              const float4 w5 = *reinterpret_cast<const float4*>(ws_ + kk * HS + 4);
              w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
              w[CW - 4] = w5.x; w[CW - 3] = w5.y; w[CW - 2] = w5.z; w[CW - 1] = w5.w;
            } else {
#pragma unroll
              for (int c2 = 0; c2 < CW; ++c2) w[c2] = ws_[kk * HS + c2];
            }
            float hv[TBT];
            if (TBT >= 4) {
#pragma unroll
              for (int v = 0; v < TBT / 4; ++v) {
                const float4 t4 = *reinterpret_cast<const float4*>(hs_ + kk * BT + tile_row<TBT>(bg, v * 4));
                hv[v * 4 + 0] = t4.x; hv[v * 4 + 1] = t4.y; hv[v * 4 + 2] = t4.z; hv[v * 4 + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int r = 0; r < TBT; ++r) hv[r] = hs_[kk * BT + bg * TBT + r];
            }
#pragma unroll
            for (int r = 0; r < TBT; ++r)
#pragma unroll
              for (int c2 = 0; c2 < CW; ++c2) acc[r][c2] = fmaf(hv[r], w[c2], acc[r][c2]);
          }
        }
        if (!(p.dbg & 1)) gq = g0 + nchunks;
      }
#pragma unroll
      for (int r = 0; r < TBT; ++r)
#pragma unroll
        for (int c2 = 0; c2 < CW; ++c2)
          red[((size_t)warp * BT + tile_row<TBT>(bg, r)) * HS + jj * CW + c2] = acc[r][c2];
      __syncthreads();

      // ---- pointwise gate gradients ------------------------------------------------------------
#pragma unroll
      for (int q = 0; q < PP; ++q) {
        const int pr = tid + q * RNN_THREADS;
        const int jl = pr % HS, bl = pr / HS, b = b0 + bl;
        if (pr < C::PAIRS && b < p.B) {
          float dh = dyv[q];
#pragma unroll
          for (int w8 = 0; w8 < RNN_WARPS; ++w8) dh += red[((size_t)w8 * BT + bl) * HS + jl];
          float dz[4] = {0.f, 0.f, 0.f, 0.f};
          float dcn = 0.f;
          if (valid[q]) {
            const float ig = gt[q][0], gg = gt[q][1], fg = gt[q][2], og = gt[q][3];
            const float tc = tanhf(ct[q]);
            const float d_o = dh * tc;
            const float dc = dcr[q] + dh * og * (1.f - tc * tc);
            dz[0] = dc * gg * ig * (1.f - ig);
            dz[1] = dc * ig * (1.f - gg * gg);
            dz[2] = dc * cprev[q] * fg * (1.f - fg);
            dz[3] = d_o * og * (1.f - og);
            dcn = dc * fg;
          }
          const int t = tb[q];
          float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            __stcg(gp + g * H, dz[g]);
            __stcg(dznext + ((size_t)tile * H4 + g * H + j0 + jl) * BT + bl, dz[g]);
            dbacc[g] += dz[g];
          }
          __stcg(dcb + (size_t)b * H + j0 + jl, dcn);
        }
      }
      __syncthreads();
    }
    asm volatile("fence.proxy.async;" ::: "memory");   // our generic stores -> the peers' bulk (async proxy) loads
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_gpu_add(counter, 1u);
  }

  // bias gradient: every thread's pairs share jl = tid % HS (256 % HS == 0); fixed-order sum
  {
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 4; ++g) ring[tid * 4 + g] = dbacc[g];
    __syncthreads();
    if (tid < 4 * HS) {
      const int g = tid / HS, j = tid % HS;
      float sum = 0.f;
      for (int i = j; i < RNN_THREADS; i += HS) sum += ring[i * 4 + g];
      p.dbpart[((size_t)dir * 8 + grp) * H4 + g * H + j0 + j] = sum;
    }
  }
}

__global__ void sum_groups_kernel(const float* part, int ngrp, int n, float* out0, float* out1, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n) return;
  const int dir = i / n, k = i % n;
  float* out = dir ? out1 : out0;
  float s = accumulate ? out[k] : 0.f;
  for (int g = 0; g < ngrp; ++g) s += part[((size_t)dir * 8 + g) * n + k];
  out[k] = s;
}
// word 0 = max over the batch of the rows' max |dy|, words 1..127 = 0 (see blstm_rec_bwd_chain: rowmax_ready)
__global__ void batch_max_kernel(const unsigned* rowmax, int B, unsigned* out) {
  __shared__ unsigned m;
  if (threadIdx.x == 0) m = 0u;
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float G = __uint_as_float(rowmax[i]);
    if (G > 0.f && G < 3.0e38f) atomicMax(&m, rowmax[i]);
  }
  __syncthreads();
  if (threadIdx.x < 128) out[threadIdx.x] = threadIdx.x == 0 ? m : 0u;
}
__global__ void rows_absmax_kernel(const float* __restrict__ dy, const int* __restrict__ len, int yT, int W, unsigned* rowmax) {
  const int b = blockIdx.y;
  const size_t n = (size_t)len[b] * W;
  const float* src = dy + (size_t)b * yT * W;
  float m = 0.f;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i + 3 < n; i += (size_t)gridDim.x * blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(rowmax + b, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Plan {
  int tbt, hs, nsl, ngrp, ndir_concurrent, bwd_kc, bwd_stages, fwd_stages, ctas_per_sm;
  size_t smem_fwd, smem_bwd;
};

// split = true: batch tiles of 64 rows, each owned by its own set of CTAs, two CTAs per SM -- while one batch
// group waits on its per-step handshake (about 10 us of fences, polling and barriers) the other one computes.
int make_plan(int B, int H, bool split, Plan* pl) {
  NABU_REQUIRE(H % KC == 0, "blstm: num_units=%d must be a multiple of %d", H, KC);
  pl->tbt = split ? 4 : (B <= 16 ? 1 : B <= 32 ? 2 : B <= 64 ? 4 : 8);
  const int BT = 16 * pl->tbt;
  pl->ngrp = split ? ceil_div(B, BT) : 1;
  pl->ctas_per_sm = split ? 2 : 1;
  if (split && (pl->ngrp < 2 || pl->ngrp > 8)) return 3;
  const int sms = num_sms();
  const size_t cap = split ? (size_t)(233472 - 2048) / 2 : (size_t)max_smem_optin();
  const int cand[4] = {2, 4, 8, 16};
  for (int pass = 0; pass < 2; ++pass) {          // pass 0: both directions concurrently
    const int ndir = pass == 0 ? 2 : 1;
    for (int ci = 0; ci < 4; ++ci) {
      const int hs = cand[ci];
      if (H % hs) continue;
      const int nsl = H / hs;
      if (ndir * nsl * pl->ngrp > sms * pl->ctas_per_sm) continue;
      if (ndir * nsl > sms) continue;               // keep at most one CTA of a batch group per SM
      const int ks = RNN_WARPS / (hs / 2);
      for (int st = STAGES; st >= 2; --st) {
        const size_t fwd = ((size_t)H * hs * 4 + (size_t)st * KC * BT + (size_t)ks * BT * hs * 4) * sizeof(float) + 64;
        const size_t bwd = ((size_t)4 * H * hs + (size_t)st * KC * BT + (size_t)RNN_WARPS * BT * hs) * sizeof(float) + 64;
        if (fwd > cap || bwd > cap) continue;
        pl->bwd_kc = KC; pl->bwd_stages = st; pl->fwd_stages = st;
        pl->hs = hs; pl->nsl = nsl; pl->ndir_concurrent = ndir; pl->smem_fwd = fwd; pl->smem_bwd = bwd;
        return 0;
      }
    }
  }
  if (split) return 3;
  set_error("blstm: num_units=%d does not fit the persistent kernel (needs H/hs <= %d CTAs and the Kh slice in %zu B shared memory)",
            H, sms, cap);
  return 2;
}

template <int TBT, int HS>
int launch_rec(bool backward, const RecParams& rp, const Plan& pl, cudaStream_t stream) {
  const void* fn = backward ? (const void*)blstm_rec_bwd_kernel<TBT, HS> : (const void*)blstm_rec_fwd_kernel<TBT, HS>;
  const size_t smem = backward ? pl.smem_bwd : pl.smem_fwd;
  NABU_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = pl.nsl * pl.ngrp * pl.ndir_concurrent;
  int per_sm = 0;
  NABU_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, RNN_THREADS, smem));
  if (per_sm * num_sms() < grid) return 3;          // not co-resident: the caller falls back to the unsplit plan
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("NABU_REC_DBG"); dbg = e ? atoi(e) : 0; }
  for (int d0 = 0; d0 < 2; d0 += pl.ndir_concurrent) {
    RecParams q = rp;
    q.dir0 = d0;
    q.dbg = dbg;
    q.kc = pl.bwd_kc; q.stages = pl.bwd_stages; q.fstages = pl.fwd_stages;
    q.nsl = pl.nsl; q.ngrp = pl.ngrp;
    void* args[] = {(void*)&q};
    KernelScope ks(backward ? "blstm_rec_bwd" : "blstm_rec_fwd", stream);
    NABU_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(RNN_THREADS), args, smem, stream));
  }
  return 0;
}

template <int TBT>
int dispatch_hs(bool backward, const RecParams& rp, const Plan& pl, cudaStream_t stream) {
  switch (pl.hs) {
    case 2: return launch_rec<TBT, 2>(backward, rp, pl, stream);
    case 4: return launch_rec<TBT, 4>(backward, rp, pl, stream);
    case 8: return launch_rec<TBT, 8>(backward, rp, pl, stream);
    case 16: return launch_rec<TBT, 16>(backward, rp, pl, stream);
  }
  set_error("blstm: bad hs");
  return 2;
}

int dispatch(bool backward, const RecParams& rp, const Plan& pl, cudaStream_t stream) {
  switch (pl.tbt) {
    case 1: return dispatch_hs<1>(backward, rp, pl, stream);
    case 2: return dispatch_hs<2>(backward, rp, pl, stream);
    case 4: return dispatch_hs<4>(backward, rp, pl, stream);
    case 8: return dispatch_hs<8>(backward, rp, pl, stream);
  }
  set_error("blstm: bad tbt");
  return 2;
}

// Launch the recurrence: batch-group split (two CTAs per SM) when the batch is large enough and it is co-resident,
// else one CTA per (direction, slice).  NABU_REC_SPLIT=1 enables the split.
int run_recurrence(bool backward, RecParams rp, int B, int H, cudaStream_t stream, int* ngrp_used = nullptr) {
  static int use_split = -1;
  if (use_split < 0) {
    const char* e = getenv("NABU_REC_SPLIT");
    use_split = (e && strcmp(e, "1") == 0) ? 1 : 0;     // opt-in: measured no gain at cfg-3 (547 vs 550 ms/step)
  }
  Plan pl;
  if (use_split && B > 64 && make_plan(B, H, true, &pl) == 0) {
    rp.Bp = pl.ngrp * 16 * pl.tbt;
    const int e = dispatch(backward, rp, pl, stream);
    if (ngrp_used) *ngrp_used = pl.ngrp;
    if (e != 3) return e;
  }
  if (int e = make_plan(B, H, false, &pl)) return e;
  rp.Bp = ceil_div(B, 16 * pl.tbt) * 16 * pl.tbt;
  if (ngrp_used) *ngrp_used = 1;
  const int e = dispatch(backward, rp, pl, stream);
  NABU_REQUIRE(e != 3, "blstm: the persistent kernel is not co-resident on this device");
  return e;
}

// workspace layout: [counters 256 B | per-row max |dy| 512 B | pad][exchange 2*2*4H*Bp floats][dcbuf 2*Bp*H floats]
// [gemm scratch][fp16 split planes P1 (layer input / output side), P2, P3 (gate side), weights, scales]
constexpr int YT_SLACK = 16;    // yT - T the split planes are sized for (pyramid padding)
struct Ws {
  unsigned* rowmax_all; unsigned* rowmax_tile;     // B > 128: every row's max |dy|, and the 128-word image handed to the tiles
  unsigned* counters; unsigned* rowmax; float* xchg; float* dcbuf; float* dbpart; float* gemm; size_t gemm_bytes;
  void *p1h, *p1l, *p2h, *p2l, *p3h, *p3l, *wh[2], *wl[2];
  float *row1, *row2, *glob;      // row scales of P1 / P2|P3 rows; 16 global scalars (inverse scales + max bits)
  size_t total;
};
Ws carve(void* base, int H, int B, int T, int D) {
  const int Bp = ceil_div(B, 128) * 128;
  Ws w;
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) { void* p = b + off; off += align_up(bytes, 256); return p; };
  w.counters = (unsigned*)take(1024); w.rowmax = w.counters + 64;
  w.rowmax_all = (unsigned*)take((size_t)Bp * 4); w.rowmax_tile = (unsigned*)take(512);
  {
    size_t xf = (size_t)2 * 2 * 4 * H * Bp;
    if (xf < (size_t)4 * 128 * H) xf = (size_t)4 * 128 * H;      // the tcgen05 path exchanges [2][2][128][H]
    w.xchg = (float*)take(xf * sizeof(float));
  }
  w.dcbuf = (float*)take((size_t)2 * Bp * H * sizeof(float));
  w.dbpart = (float*)take((size_t)2 * 8 * 4 * H * sizeof(float));
  w.gemm_bytes = sgemm_workspace_bytes();
  w.gemm = (float*)take(w.gemm_bytes);
  const size_t D8 = align_up(D, 8), H4 = (size_t)4 * H;
  const size_t e1 = std::max((size_t)B * T * D8, (size_t)B * (T + YT_SLACK) * 2 * H);
  const size_t e2 = (size_t)B * T * H4;
  w.p1h = take(e1 * 2); w.p1l = take(e1 * 2);
  w.p2h = take(e2 * 2); w.p2l = take(e2 * 2);
  w.p3h = take(e2 * 2); w.p3l = take(e2 * 2);
  for (int d = 0; d < 2; ++d) { w.wh[d] = take((size_t)(D + H) * H4 * 2); w.wl[d] = take((size_t)(D + H) * H4 * 2); }
  w.row1 = (float*)take((size_t)B * T * 4); w.row2 = (float*)take((size_t)B * T * 4);
  w.glob = (float*)take(256);
  w.total = off;
  return w;
}

// scratch of the deferred weight gradients (library-owned, see Overlap): gemm scratch, scales, planes P1..P3
Ws carve_side(void* base, int H, int B, int T, int D) {
  Ws w = {};
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) { void* p = b + off; off += align_up(bytes, 256); return p; };
  w.gemm_bytes = sgemm_workspace_bytes();
  w.gemm = (float*)take(w.gemm_bytes);
  w.glob = (float*)take(256);
  const size_t D8 = align_up(D, 8), H4 = (size_t)4 * H;
  const size_t e1 = std::max((size_t)B * T * D8, (size_t)B * (T + YT_SLACK) * 2 * H);
  const size_t e2 = (size_t)B * T * H4;
  w.p1h = take(e1 * 2); w.p1l = take(e1 * 2);
  w.p2h = take(e2 * 2); w.p2l = take(e2 * 2);
  w.p3h = take(e2 * 2); w.p3l = take(e2 * 2);
  w.total = off;
  return w;
}

// ---- operand planes that travel with the activations (include/nabu_b200.h: nabu_blstm_*_planes) ----------------------
// planes of an [R, C] fp32 matrix: hi at the base, lo plane_half(R * C) bytes further
size_t plane_half(size_t elems) { return align_up(elems * 2, 256); }

__device__ float g_inv_y_scale = 1.f / Y_PLANE_SCALE;
const float* inv_y_scale_ptr() {
  static float* p = nullptr;
  if (!p && cudaGetSymbolAddress((void**)&p, g_inv_y_scale) != cudaSuccess) p = nullptr;
  return p;
}

// y planes for the recurrence kernels that do not write them themselves: split of y * 32 (8 values per thread)
__global__ void __launch_bounds__(256) split_fixed_kernel(const float* __restrict__ src, size_t n8, __half* __restrict__ hi,
                                                          __half* __restrict__ lo, float S) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    union { __half h[8]; uint4 u; } ph, pl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __half h = __float2half_rn(v[j] * S);
      ph.h[j] = h;
      pl.h[j] = __float2half_rn((v[j] * S - __half2float(h)) * 2048.f);
    }
    reinterpret_cast<uint4*>(hi)[i] = ph.u;
    reinterpret_cast<uint4*>(lo)[i] = pl.u;
  }
}

// The backward recurrence writes the dZ planes with the gate columns of a direction in the order [unit quad][gate][4 units]
// (blstm_cl_bwd8.cu: a thread's 16 values are then 32 contiguous bytes): column g*H + j sits at zperm(g, j).
__host__ __device__ inline int zperm(int g, int j) { return (j >> 2) * 16 + g * 4 + (j & 3); }

// dK[r][g*H + j] = tmp[r][zperm(g, j)]  (the weight-gradient GEMMs contract against the permuted planes)
__global__ void __launch_bounds__(256) unpermute_cols_kernel(const float* __restrict__ tmp, int R, int H, float* __restrict__ out) {
  const int H4 = 4 * H;
  const size_t n = (size_t)R * H4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % H4);
    const size_t r = i / H4;
    out[i] = tmp[r * H4 + zperm(c / H, c % H)];
  }
}
// [Kx_fw | Kx_bw] with the same column order as the dZ planes: out[r][d*4H + zperm(g, j)] = kern[d][r][g*H + j], r < D
__global__ void __launch_bounds__(256) permute_kx_kernel(const float* __restrict__ k0, const float* __restrict__ k1, int D, int H,
                                                         float* __restrict__ out) {
  const int H4 = 4 * H;
  const size_t n = (size_t)D * 2 * H4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % (2 * H4));
    const size_t r = i / (2 * H4);
    const int d = cc / H4, c = cc % H4;
    out[r * 2 * H4 + (size_t)d * H4 + zperm(c / H, c % H)] = (d ? k1 : k0)[r * H4 + c];
  }
}

// Library-owned scratch of the planes path: two sets of dZ planes [B*T, 8H] (hi | lo) + their 1/scale -- the deferred weight
// gradients of layer l read set k on the side stream while layer l-1's recurrence writes set k^1 --, the planes of a
// layer input that arrived without planes (layer 0's features), the weights' planes for dX, split-K scratch.
struct SideZ {
  void *zh[2], *zl[2]; float* zglob[2];
  void *xh, *xl; float* xglob;          // xglob: [0], [1] 1/scale of x, y; [8], [9] max bits (split_global's scratch)
  void *yh, *yl;                        // planes of a y that arrived without planes (the plain entry points)
  void *kh, *kl; float* kglob;          // [Kx_fw | Kx_bw] planes [D, 8H] for dX (columns in the planes' order)
  float* kperm;                         // fp32 [D, 8H]: the permuted weights before their split (caller's stream)
  float* dktmp[2];                      // fp32 [(D+H), 4H] per direction: weight gradients in the planes' column order
  float* gemm; size_t gemm_bytes;
  size_t total;
};
// The two dZ sets sit at FIXED places of the scratch (set k at k * set_cap), whatever the shape of the call: a layer's
// deferred GEMMs still read set k while the next layer -- of another length or width in a pyramidal encoder -- writes
// set k^1 and lays out its own misc region.  The capacities only grow, and growing reallocates behind a device sync.
struct SideCap { size_t set_cap, misc_cap; };
SideCap& side_cap() {
  static SideCap c = {0, 0};
  return c;
}
size_t sidez_set_bytes(int H, int B, int T) { return 2 * align_up((size_t)B * T * 8 * H * 2, 256) + 256; }
SideZ carve_sidez(void* base, size_t set_cap, int H, int B, int T, int D) {
  SideZ w = {};
  char* b = (char*)base;
  const size_t H4 = (size_t)4 * H, D8 = align_up(D, 8);
  const size_t plane = align_up((size_t)B * T * 2 * H4 * 2, 256);
  for (int k = 0; k < 2; ++k) {
    char* s = b + (size_t)k * set_cap;
    w.zh[k] = s; w.zl[k] = s + plane; w.zglob[k] = (float*)(s + 2 * plane);
  }
  size_t off = 2 * set_cap;
  auto take = [&](size_t bytes) { void* p = b + off; off += align_up(bytes, 256); return p; };
  w.xh = take((size_t)B * T * D8 * 2); w.xl = take((size_t)B * T * D8 * 2);
  w.xglob = (float*)take(256);
  w.yh = take((size_t)B * (T + YT_SLACK) * 2 * H * 2); w.yl = take((size_t)B * (T + YT_SLACK) * 2 * H * 2);
  w.kh = take((size_t)D * 2 * H4 * 2); w.kl = take((size_t)D * 2 * H4 * 2);
  w.kglob = (float*)take(256);
  w.kperm = (float*)take((size_t)D * 2 * H4 * 4);
  for (int d = 0; d < 2; ++d) w.dktmp[d] = (float*)take((size_t)(D + H) * H4 * 4);
  w.gemm_bytes = sgemm_workspace_bytes();
  w.gemm = (float*)take(w.gemm_bytes);
  w.total = off;
  return w;
}
struct ZState { unsigned calls; cudaEvent_t done[2]; bool recorded[2]; cudaEvent_t dx_done[2]; bool dx_recorded[2]; };
ZState& zstate() {
  static ZState z = {};
  return z;
}

// fp16-split tensor-core GEMMs (gemm_h2.cu) for the layer-sized contractions; NABU_GEMM=tf32|simt keeps gemm().
bool use_h2(int B, int T, int D, int H, int yT) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NABU_GEMM");
    enabled = (e && (strcmp(e, "tf32") == 0 || strcmp(e, "simt") == 0)) ? 0 : 1;
  }
  return enabled && yT - T <= YT_SLACK && D % 4 == 0 && H % 8 == 0 && gemm_h2_eligible(GEMM_NN, B * T, 4 * H, D) &&
         gemm_h2_eligible(GEMM_TN, D, 4 * H, B * T);
}

}  // namespace
}  // namespace nabu

using namespace nabu;

extern "C" size_t nabu_blstm_workspace_bytes(int B, int T, int D, int H) {
  Plan pl;
  if (make_plan(B, H, false, &pl)) return 0;
  return carve(nullptr, H, B, T, D).total;
}

extern "C" size_t nabu_blstm_planes_bytes(int B, int yT, int H) {
  if (B <= 0 || yT <= 0 || H <= 0 || H % 4) return 0;
  return 2 * plane_half((size_t)B * yT * 2 * H);
}

extern "C" int nabu_blstm_fwd(const float* x, const int* len, int B, int T, int D, int H,
                              const float* kernel_fw, const float* bias_fw, const float* kernel_bw,
                              const float* bias_bw, float* y, int yT, float* gates, float* cells,
                              void* workspace, size_t ws_bytes, void* stream_) {
  return nabu_blstm_fwd_planes(x, nullptr, len, B, T, D, H, kernel_fw, bias_fw, kernel_bw, bias_bw, y, nullptr, yT, gates, cells,
                               workspace, ws_bytes, stream_);
}

extern "C" int nabu_blstm_fwd_planes(const float* x, const void* x_planes, const int* len, int B, int T, int D, int H,
                                     const float* kernel_fw, const float* bias_fw, const float* kernel_bw,
                                     const float* bias_bw, float* y, void* y_planes, int yT, float* gates, float* cells,
                                     void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NABU_REQUIRE(!x_planes || D % 8 == 0, "blstm_fwd: input planes need D %% 8 == 0 (D=%d)", D);
  NABU_REQUIRE(!y_planes || H % 4 == 0, "blstm_fwd: output planes need num_units %% 4 == 0 (H=%d)", H);
  void* yh = y_planes;
  void* yl = y_planes ? (char*)y_planes + plane_half((size_t)B * yT * 2 * H) : nullptr;
  NABU_REQUIRE(B > 0 && T > 0 && D > 0 && H > 0 && yT >= T, "blstm_fwd: bad shape B=%d T=%d D=%d H=%d yT=%d", B, T, D, H, yT);
  Ws w = carve(workspace, H, B, T, D);
  NABU_REQUIRE(ws_bytes >= w.total, "blstm_fwd: workspace %zu < %zu bytes", ws_bytes, w.total);
  if (int e = overlap_join(stream)) return e;          // deferred weight gradients of the previous step read the parameters' neighbours
  const int H4 = 4 * H;
  const float* kern[2] = {kernel_fw, kernel_bw};
  const float* bias[2] = {bias_fw, bias_bw};
  float* g[2] = {gates, gates + (size_t)B * T * H4};
  float* c[2] = {cells, cells + (size_t)B * T * H};
  // input projection for all T at once: Gx = X . Kx + b
  if (use_h2(B, T, D, H, yT)) {
    const int D8 = (int)align_up(D, 8);
    H2Operand xa = {w.p1h, w.p1l, D8, w.row1, nullptr};
    if (x_planes) {        // the producer of x already wrote its operand planes (scale 32): no pass over x at all
      xa.hi = x_planes; xa.lo = (const char*)x_planes + plane_half((size_t)B * T * D); xa.ld = D;
      xa.row_inv = nullptr; xa.glob_inv = inv_y_scale_ptr();
      NABU_REQUIRE(xa.glob_inv != nullptr, "blstm_fwd: device constant unavailable");
    } else if (int e = split_rows(x, D, B * T, D, w.p1h, w.p1l, D8, w.row1, stream)) {
      return e;
    }
    for (int d = 0; d < 2; ++d) {
      if (int e = split_global(kern[d], H4, D, H4, w.wh[d], w.wl[d], H4, (unsigned*)(w.glob + 8 + d), w.glob + d, stream))
        return e;
      H2Operand kb = {w.wh[d], w.wl[d], H4, nullptr, w.glob + d};
      if (int e = gemm_h2(GEMM_NN, B * T, H4, D, 1.f, xa, kb, 0.f, g[d], H4, bias[d], nullptr, nullptr, 0, stream)) return e;
    }
  } else {
    for (int d = 0; d < 2; ++d)
      if (int e = gemm(GEMM_NN, B * T, H4, D, 1.f, x, D, kern[d], H4, 0.f, g[d], H4, bias[d], nullptr, nullptr, 0, stream))
        return e;
  }
  NABU_CHECK_CUDA(cudaMemsetAsync(w.counters, 0, 256, stream));
  if (yT > T) {
    NABU_CHECK_CUDA(cudaMemset2DAsync(y + (size_t)T * 2 * H, (size_t)yT * 2 * H * sizeof(float), 0,
                                      (size_t)(yT - T) * 2 * H * sizeof(float), B, stream));
    if (y_planes)
      for (void* pl : {yh, yl})
        NABU_CHECK_CUDA(cudaMemset2DAsync((char*)pl + (size_t)T * 2 * H * 2, (size_t)yT * 2 * H * 2, 0,
                                          (size_t)(yT - T) * 2 * H * 2, B, stream));
  }
  // y planes for the kernels that do not write them themselves (everything but the tcgen05 cluster kernel)
  auto planes_after = [&]() -> int {
    if (!y_planes) return 0;
    const size_t n8 = (size_t)B * yT * 2 * H / 8;
    KernelScope ks("split_fixed", stream);
    split_fixed_kernel<<<(unsigned)std::min<size_t>((n8 + 255) / 256, (size_t)num_sms() * 8), 256, 0, stream>>>(
        y, n8, (__half*)yh, (__half*)yl, Y_PLANE_SCALE);
    NABU_CHECK_LAUNCH();
    return 0;
  };
  // The tcgen05 recurrence (blstm_tc.cu) is opt-in (NABU_REC=tc): measured 20.8 us/step at cfg-3, the FFMA kernel
  // with the batch-group split is faster; see DESIGN.md section 6.
  static int use_tc = -1;
  if (use_tc < 0) {
    const char* e = getenv("NABU_REC");
    use_tc = (e && strcmp(e, "tc") == 0) ? 1 : 0;
  }
  BlstmTcPlan tpl;
  if (use_tc && blstm_tc_plan(B, H, &tpl)) {
    NABU_CHECK_CUDA(cudaMemsetAsync(w.xchg, 0, (size_t)4 * 128 * H * sizeof(float), stream));
    if (int e = blstm_rec_fwd_tc(tpl, kern, g, c, y, w.xchg, w.counters, len, B, T, yT, D, H, stream)) return e;
    return planes_after();
  }
  NABU_CHECK_CUDA(cudaMemsetAsync(w.xchg, 0, (size_t)2 * 2 * H * (ceil_div(B, 128) * 128) * sizeof(float), stream));
  // the tcgen05 recurrences: the chain kernel (blstm_cl_fwdc.cu) for up to 128 rows, the 128-row kernel (blstm_cl_tc.cu)
  // when it is forced or the chain kernel cannot be placed; more than 128 rows tile by tile, one tile after the other
  auto rec_tile = [&](int b0, int Bt, bool* launched) -> int {
    float* gt[2] = {g[0] + (size_t)b0 * T * H4, g[1] + (size_t)b0 * T * H4};
    float* ct[2] = {c[0] + (size_t)b0 * T * H, c[1] + (size_t)b0 * T * H};
    void* th = yh ? (char*)yh + (size_t)b0 * yT * 2 * H * 2 : nullptr;
    void* tl = yl ? (char*)yl + (size_t)b0 * yT * 2 * H * 2 : nullptr;
    *launched = false;
    if (blstm_fwd_chain_eligible(Bt, H))
      if (int e = blstm_rec_fwd_chain(kern, gt, ct, y + (size_t)b0 * yT * 2 * H, w.xchg, len + b0, Bt, T, yT, D, H, stream, launched, th, tl))
        return e;
    if (!*launched && blstm_fwd_cluster_tc_eligible(Bt, H))
      if (int e = blstm_rec_fwd_cluster_tc(kern, gt, ct, y + (size_t)b0 * yT * 2 * H, w.xchg, w.counters, len + b0, Bt, T, yT, D, H,
                                           stream, launched, th, tl))
        return e;
    return 0;
  };
  if (blstm_fwd_chain_eligible(std::min(B, 128), H) || blstm_fwd_cluster_tc_eligible(std::min(B, 128), H)) {
    bool any = false;
    for (int b0 = 0; b0 < B; b0 += 128) {
      const int Bt = std::min(128, B - b0);
      if (b0 > 0) {
        NABU_CHECK_CUDA(cudaMemsetAsync(w.counters, 0, 256, stream));
        NABU_CHECK_CUDA(cudaMemsetAsync(w.xchg, 0, (size_t)2 * 2 * H * 128 * sizeof(float), stream));
      }
      bool launched = false;
      if (int e = rec_tile(b0, Bt, &launched)) return e;
      if (!launched) {
        NABU_REQUIRE(b0 == 0, "blstm_fwd: the tcgen05 recurrence stopped being launchable between batch tiles");
        break;
      }
      any = true;
    }
    if (any) return 0;
  }
  {
    char key[96];
    snprintf(key, sizeof(key), "fwd B=%d H=%d", B, H);
    warn_once(key, "blstm forward recurrence B=%d num_units=%d is not on the tcgen05 cluster kernels (they exist for num_units 256, "
              "512 and 1024; the Python engine pads other widths up to them): falling back to the FFMA kernels", B, H);
  }
  if (blstm_fwd_cluster_eligible(B, H)) {
    bool launched = false;
    if (int e = blstm_rec_fwd_cluster(kern, g, c, y, w.xchg, w.counters, len, B, T, yT, D, H, stream, &launched)) return e;
    if (launched) return planes_after();
  }
  RecParams rp = {};
  rp.kernel[0] = kern[0]; rp.kernel[1] = kern[1];
  rp.gates[0] = g[0]; rp.gates[1] = g[1];
  rp.cells[0] = c[0]; rp.cells[1] = c[1];
  rp.y = y; rp.xchg = w.xchg; rp.counters = w.counters; rp.len = len;
  rp.B = B; rp.T = T; rp.yT = yT; rp.D = D; rp.H = H;
  if (int e = run_recurrence(false, rp, B, H, stream)) return e;
  return planes_after();
}

// One-shot hand-over of max |dx| between the backward calls of consecutive layers (nabu_blstm_bwd_hints).
namespace {
struct BwdHints { unsigned* dx_out; const unsigned* dy_in; };
thread_local BwdHints t_bwd_hints = {nullptr, nullptr};
// fills the promised max |dx| on every path that did not get it from the dX contraction's epilogue
struct DxMaxGuard {
  unsigned* out; const float* dx; size_t rows, cols; cudaStream_t stream; bool filled;
  ~DxMaxGuard() {
    if (out && dx && !filled) absmax_accumulate(dx, (int)cols, rows, cols, out, stream);
  }
};
}  // namespace

extern "C" int nabu_blstm_bwd_hints(unsigned* dx_absmax_out, const unsigned* dy_absmax_in) {
  t_bwd_hints.dx_out = dx_absmax_out;
  t_bwd_hints.dy_in = dy_absmax_in;
  return 0;
}

extern "C" int nabu_blstm_bwd(const float* x, const int* len, int B, int T, int D, int H,
                              const float* kernel_fw, const float* kernel_bw, const float* y, int yT,
                              float* gates, const float* cells, const float* dy, float* dx,
                              float* dkernel_fw, float* dbias_fw, float* dkernel_bw, float* dbias_bw,
                              void* workspace, size_t ws_bytes, void* stream_) {
  return nabu_blstm_bwd_planes(x, nullptr, len, B, T, D, H, kernel_fw, kernel_bw, y, nullptr, yT, gates, cells, dy, dx,
                               dkernel_fw, dbias_fw, dkernel_bw, dbias_bw, workspace, ws_bytes, stream_);
}

extern "C" int nabu_blstm_bwd_planes(const float* x, const void* x_planes, const int* len, int B, int T, int D, int H,
                                     const float* kernel_fw, const float* kernel_bw, const float* y, const void* y_planes,
                                     int yT, float* gates, const float* cells, const float* dy, float* dx,
                                     float* dkernel_fw, float* dbias_fw, float* dkernel_bw, float* dbias_bw,
                                     void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const BwdHints hints = t_bwd_hints;                   // consumed by this call, whatever its outcome
  t_bwd_hints = BwdHints{nullptr, nullptr};
  NABU_REQUIRE(!x_planes || D % 8 == 0, "blstm_bwd: input planes need D %% 8 == 0 (D=%d)", D);
  NABU_REQUIRE(B > 0 && T > 0 && D > 0 && H > 0 && yT >= T, "blstm_bwd: bad shape");
  if (hints.dx_out && dx) NABU_CHECK_CUDA(cudaMemsetAsync(hints.dx_out, 0, 512, stream));
  DxMaxGuard dxmax = {hints.dx_out, dx, (size_t)B * T, (size_t)D, stream, false};
  Ws w = carve(workspace, H, B, T, D);
  NABU_REQUIRE(ws_bytes >= w.total, "blstm_bwd: workspace %zu < %zu bytes", ws_bytes, w.total);
  const int H4 = 4 * H;
  const float* kern[2] = {kernel_fw, kernel_bw};
  float* dkern[2] = {dkernel_fw, dkernel_bw};
  float* g[2] = {gates, gates + (size_t)B * T * H4};
  const float* c[2] = {cells, cells + (size_t)B * T * H};
  NABU_CHECK_CUDA(cudaMemsetAsync(w.counters, 0, 256, stream));
  RecParams rp = {};
  rp.kernel[0] = kern[0]; rp.kernel[1] = kern[1];
  rp.gates[0] = g[0]; rp.gates[1] = g[1];
  rp.cells[0] = (float*)c[0]; rp.cells[1] = (float*)c[1];
  rp.dy = dy; rp.dbias[0] = dbias_fw; rp.dbias[1] = dbias_bw;
  rp.xchg = w.xchg; rp.dcbuf = w.dcbuf; rp.dbpart = w.dbpart; rp.counters = w.counters; rp.len = len;
  rp.B = B; rp.T = T; rp.yT = yT; rp.D = D; rp.H = H;
  int ngrp = 1;
  bool launched = false;
  // Deferred weight gradients (nabu_set_overlap): the cluster-of-8 recurrence leaves 84 of the 148 SMs idle, so the
  // weight-gradient GEMMs of THIS layer run on a side stream while the caller goes on to the next layer's recurrence,
  // which is issued on a high-priority stream so that its clusters take SMs as the GEMM's CTAs retire.
  Overlap& ov = overlap();
  // a batch of more than 128 rows runs the tcgen05 recurrence tile by tile (128 rows each, one after the other: two
  // tiles' clusters are not co-resident)
  // (num_units = 1024: tiles of 32 rows, the chain kernel's limit there -- blstm_cl_bwd8c.cu)
  const int TILE_B = H == 1024 ? 32 : 128;
  const int ntile = ceil_div(B, TILE_B);
  const bool tc_ok = H == 1024 ? blstm_bwd_chain_eligible(std::min(B, TILE_B), H)
                               : blstm_bwd_cluster8_eligible(std::min(B, 128), H) && (ntile == 1 || blstm_bwd_chain_eligible(128, H));
  const bool defer = ov.on && tc_ok && use_h2(B, T, D, H, yT);
  // dZ as operand planes straight from the recurrence (NABU_ZPLANES=0 keeps the fp32 dZ + split passes): the planes live
  // in library-owned scratch, two sets used alternately (see SideZ)
  static int zplanes_on = -1;
  if (zplanes_on < 0) zplanes_on = (getenv("NABU_ZPLANES") && atoi(getenv("NABU_ZPLANES")) == 0) ? 0 : 1;
  const bool zplanes = zplanes_on && tc_ok && use_h2(B, T, D, H, yT) &&
                       gemm_h2_eligible(GEMM_NT, B * T, D, 2 * H4) && inv_y_scale_ptr() != nullptr;
  SideZ sz = {};
  int zk = 0;
  if (zplanes) {
    if (int e = overlap_init()) return e;
    SideCap& cap = side_cap();
    const size_t set_need = sidez_set_bytes(H, B, T);
    const size_t misc_need = carve_sidez(nullptr, 0, H, B, T, D).total;
    if (set_need > cap.set_cap || misc_need > cap.misc_cap) {
      NABU_CHECK_CUDA(cudaDeviceSynchronize());         // the sets move: nothing may be in flight
      cap.set_cap = std::max(cap.set_cap, set_need);
      cap.misc_cap = std::max(cap.misc_cap, misc_need);
      zstate().recorded[0] = zstate().recorded[1] = zstate().dx_recorded[0] = zstate().dx_recorded[1] = false;
    }
    if (int e = overlap_workspace(2 * cap.set_cap + cap.misc_cap)) return e;
    sz = carve_sidez(ov.ws, cap.set_cap, H, B, T, D);
    ZState& zs = zstate();
    zk = (int)(zs.calls++ & 1u);
    if (!zs.done[0]) {
      for (int k = 0; k < 2; ++k) {
        NABU_CHECK_CUDA(cudaEventCreateWithFlags(&zs.done[k], cudaEventDisableTiming));
        NABU_CHECK_CUDA(cudaEventCreateWithFlags(&zs.dx_done[k], cudaEventDisableTiming));
      }
    }
    // set zk was last read by the GEMMs of two calls ago (side stream and the caller's stream of that call)
    if (zs.recorded[zk]) NABU_CHECK_CUDA(cudaStreamWaitEvent(stream, zs.done[zk], 0));
    if (zs.dx_recorded[zk]) NABU_CHECK_CUDA(cudaStreamWaitEvent(stream, zs.dx_done[zk], 0));
  }
  if (tc_ok) {
    const float* cc[2] = {c[0], c[1]};
    NABU_CHECK_CUDA(cudaMemsetAsync(w.counters, 0, 1024, stream));
    NABU_CHECK_CUDA(cudaMemsetAsync(w.xchg, 0, (size_t)2 * 2 * H4 * 128 * sizeof(float), stream));
    cudaStream_t rs = stream;
    if (defer) {
      NABU_CHECK_CUDA(cudaEventRecord(ov.ev_pre, stream));
      NABU_CHECK_CUDA(cudaStreamWaitEvent(ov.hp, ov.ev_pre, 0));
      rs = ov.hp;
      // The concurrent GEMMs stream gigabytes through L2; keep the 4 MB exchange buffer of the recurrence resident
      // (persisting access-policy window on the recurrence's stream) so that its polls do not go to DRAM.
      static int l2_pinned = -1;
      if (l2_pinned < 0) {
        l2_pinned = (getenv("NABU_L2_PIN") && atoi(getenv("NABU_L2_PIN")) == 0) ? 0 : 1;
        if (l2_pinned && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)16 << 20) != cudaSuccess) {
          cudaGetLastError();
          l2_pinned = 0;
        }
      }
      if (l2_pinned) {
        cudaStreamAttrValue av = {};
        av.accessPolicyWindow.base_ptr = w.xchg;
        av.accessPolicyWindow.num_bytes = (size_t)2 * 2 * H4 * 128 * sizeof(float);
        av.accessPolicyWindow.hitRatio = 1.f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(ov.hp, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
      }
    }
    ov.in_defer = defer;
    int re = 0;
    if (ntile == 1) {
      if (blstm_bwd_chain_eligible(B, H)) {
        // max |dy| handed over by the caller (the layer above's dX epilogue wrote it): no pass over dy.  The image is
        // 128 words, word 0 = the maximum, the others 0 -- what the kernel expects of its per-row array.
        const bool have_max = hints.dy_in != nullptr && B <= 128;
        if (have_max) NABU_CHECK_CUDA(cudaMemcpyAsync(w.rowmax, hints.dy_in, 512, cudaMemcpyDeviceToDevice, rs));
        re = blstm_rec_bwd_chain(kern, g, cc, dy, w.dbpart, w.xchg, w.rowmax, len, B, T, yT, D, H, rs, &launched, &ngrp,
                                 zplanes ? sz.zh[zk] : nullptr, zplanes ? sz.zl[zk] : nullptr, zplanes ? sz.zglob[zk] : nullptr,
                                 have_max);
        if (!launched && have_max) NABU_CHECK_CUDA(cudaMemsetAsync(w.rowmax, 0, 512, rs));   // the fallback computes its own
      }
      if (!re && !launched && blstm_bwd_cluster8_eligible(B, H)) {
        ngrp = 1;
        re = blstm_rec_bwd_cluster8(kern, g, cc, dy, w.dbpart, w.xchg, w.rowmax, len, B, T, yT, D, H, rs, &launched,
                                    zplanes ? sz.zh[zk] : nullptr, zplanes ? sz.zl[zk] : nullptr,
                                    zplanes ? sz.zglob[zk] : nullptr);
      }
    } else {
      // one scale for the whole batch: every row's max |dy| first, then its maximum in word 0 of the tiles' 128-word image
      NABU_CHECK_CUDA(cudaMemsetAsync(w.rowmax_all, 0, (size_t)B * 4, rs));
      {
        KernelScope ks("row_absmax", rs);
        rows_absmax_kernel<<<dim3(32, B), 256, 0, rs>>>(dy, len, yT, 2 * H, w.rowmax_all);
        NABU_CHECK_LAUNCH();
        batch_max_kernel<<<1, 256, 0, rs>>>(w.rowmax_all, B, w.rowmax_tile);
        NABU_CHECK_LAUNCH();
      }
      for (int tb = 0; tb < ntile && !re; ++tb) {
        const int b0 = tb * TILE_B, Bt = std::min(TILE_B, B - b0);
        if (tb > 0) {
          NABU_CHECK_CUDA(cudaMemsetAsync(w.counters, 0, 1024, rs));
          NABU_CHECK_CUDA(cudaMemsetAsync(w.xchg, 0, (size_t)2 * 2 * H4 * 128 * sizeof(float), rs));
        }
        float* gt[2] = {g[0] + (size_t)b0 * T * H4, g[1] + (size_t)b0 * T * H4};
        const float* ct[2] = {c[0] + (size_t)b0 * T * H, c[1] + (size_t)b0 * T * H};
        bool l1 = false;
        int slots = 1;
        re = blstm_rec_bwd_chain(kern, gt, ct, dy + (size_t)b0 * yT * 2 * H, w.dbpart, w.xchg, w.rowmax_tile, len + b0, Bt, T, yT,
                                 D, H, rs, &l1, &slots,
                                 zplanes ? (char*)sz.zh[zk] + (size_t)b0 * T * 2 * H4 * 2 : nullptr,
                                 zplanes ? (char*)sz.zl[zk] + (size_t)b0 * T * 2 * H4 * 2 : nullptr,
                                 zplanes ? sz.zglob[zk] : nullptr, true);
        if (re) break;
        if (!l1) {
          NABU_REQUIRE(tb == 0, "blstm_bwd: the tcgen05 recurrence stopped being launchable between batch tiles");
          break;
        }
        launched = true;
        KernelScope ks("sum_groups", rs);
        sum_groups_kernel<<<ceil_div(2 * H4, 256), 256, 0, rs>>>(w.dbpart, slots, H4, dbias_fw, dbias_bw, tb > 0);
        NABU_CHECK_LAUNCH();
      }
      ngrp = launched ? 0 : 1;                         // tiles done: the bias gradients are complete
    }
    ov.in_defer = false;
    if (re) return re;
    if (defer) {
      NABU_CHECK_CUDA(cudaEventRecord(ov.ev_rec, ov.hp));
      NABU_CHECK_CUDA(cudaStreamWaitEvent(stream, ov.ev_rec, 0));
    }
  }
  if (!launched) {
    char key[96];
    snprintf(key, sizeof(key), "bwd B=%d H=%d", B, H);
    warn_once(key, "blstm backward recurrence B=%d num_units=%d is not on the TMEM-resident tcgen05 kernels (they exist for "
              "num_units 256, 512 and 1024; the Python engine pads other widths up to them): falling back", B, H);
  }
  if (!launched && blstm_bwd_cluster_tc_eligible(B, H)) {
    const float* cc[2] = {c[0], c[1]};
    NABU_CHECK_CUDA(cudaMemsetAsync(w.counters, 0, 1024, stream));
    NABU_CHECK_CUDA(cudaMemsetAsync(w.xchg, 0, (size_t)2 * 2 * H4 * 128 * sizeof(float), stream));
    if (int e = blstm_rec_bwd_cluster_tc(kern, g, cc, dy, w.dbpart, w.xchg, w.dcbuf, w.counters, w.rowmax, len, B, T, yT, D,
                                         H, stream, &launched))
      return e;
  }
  if (!launched && blstm_bwd_cluster_eligible(B, H)) {
    const float* cc[2] = {c[0], c[1]};
    if (int e = blstm_rec_bwd_cluster(kern, g, cc, dy, w.dbpart, w.xchg, w.dcbuf, w.counters, len, B, T, yT, D, H, stream,
                                      &launched))
      return e;
  }
  if (!launched)
    if (int e = run_recurrence(true, rp, B, H, stream, &ngrp)) return e;
  if (ngrp > 0 || !launched) {
    KernelScope ks("sum_groups", stream);
    sum_groups_kernel<<<ceil_div(2 * H4, 256), 256, 0, stream>>>(w.dbpart, std::max(ngrp, 1), H4, dbias_fw, dbias_bw, 0);
    NABU_CHECK_LAUNCH();
  }
  if (zplanes && launched) {
    // ---- every contraction reads the planes the recurrences wrote: no split / absmax pass over x, y or dZ ------------
    const float* inv_y = inv_y_scale_ptr();
    const int D8 = (int)align_up(D, 8);
    ZState& zs = zstate();
    H2Operand zb[2];
    for (int d = 0; d < 2; ++d)
      zb[d] = {(const __half*)sz.zh[zk] + (size_t)d * H4, (const __half*)sz.zl[zk] + (size_t)d * H4, 2 * H4, nullptr, sz.zglob[zk]};
    auto weight_grads = [&](cudaStream_t ws) -> int {
      H2Operand xa;
      if (x_planes) {
        xa = {x_planes, (const char*)x_planes + plane_half((size_t)B * T * D), D, nullptr, inv_y};
      } else {
        if (int e = split_global(x, D, B * T, D, sz.xh, sz.xl, D8, (unsigned*)(sz.xglob + 8), sz.xglob, ws)) return e;
        xa = {sz.xh, sz.xl, D8, nullptr, sz.xglob};
      }
      for (int d = 0; d < 2; ++d)
        if (int e = gemm_h2(GEMM_TN, D, H4, B * T, 1.f, xa, zb[d], 0.f, sz.dktmp[d], H4, nullptr, nullptr, sz.gemm, sz.gemm_bytes, ws))
          return e;
      const void* yh = y_planes;
      const void* yl = y_planes ? (const char*)y_planes + plane_half((size_t)B * yT * 2 * H) : nullptr;
      const float* yinv = inv_y;
      if (T > 1 && !y_planes) {          // y without planes (a caller of the plain entry point): split it here
        if (int e = split_global(y, 2 * H, B * yT, 2 * H, sz.yh, sz.yl, 2 * H, (unsigned*)(sz.xglob + 9), sz.xglob + 1, ws)) return e;
        yh = sz.yh; yl = sz.yl; yinv = sz.xglob + 1;
      }
      for (int d = 0; d < 2; ++d) {
        float* dKh = sz.dktmp[d] + (size_t)D * H4;
        if (T > 1) {
          GemmSeg seg;
          seg.seg = T - 1; seg.segA = yT; seg.segB = T;
          seg.offA = d == 0 ? 0 : 1; seg.offB = d == 0 ? 1 : 0;
          H2Operand ya = {(const __half*)yh + d * H, (const __half*)yl + d * H, 2 * H, nullptr, yinv};
          if (int e = gemm_h2(GEMM_TN, H, H4, B * (T - 1), 1.f, ya, zb[d], 0.f, dKh, H4, nullptr, &seg, sz.gemm, sz.gemm_bytes, ws))
            return e;
        } else {
          NABU_CHECK_CUDA(cudaMemsetAsync(dKh, 0, (size_t)H * H4 * sizeof(float), ws));
        }
        KernelScope ks("unpermute_cols", ws);
        unpermute_cols_kernel<<<num_sms() * 4, 256, 0, ws>>>(sz.dktmp[d], D + H, H, dkern[d]);
        NABU_CHECK_LAUNCH();
      }
      return 0;
    };
    if (dx) {
      // dX = [dZ_fw | dZ_bw] . [Kx_fw | Kx_bw]^T as ONE contraction over K = 8H.  It is on the critical path (the layer
      // below waits for it), so it runs BEFORE the deferred weight gradients are released: sharing the tensor cores
      // with them cost it 6 ms per step.
      {
        KernelScope ks("permute_kx", stream);
        permute_kx_kernel<<<num_sms() * 4, 256, 0, stream>>>(kern[0], kern[1], D, H, sz.kperm);
        NABU_CHECK_LAUNCH();
      }
      if (int e = split_global(sz.kperm, 2 * H4, D, 2 * H4, sz.kh, sz.kl, 2 * H4, (unsigned*)(sz.kglob + 8), sz.kglob, stream)) return e;
      H2Operand za = {sz.zh[zk], sz.zl[zk], 2 * H4, nullptr, sz.zglob[zk]};
      H2Operand kb = {sz.kh, sz.kl, 2 * H4, nullptr, sz.kglob};
      if (int e = gemm_h2(GEMM_NT, B * T, D, 2 * H4, 1.f, za, kb, 0.f, dx, D, nullptr, nullptr, nullptr, 0, stream, hints.dx_out)) return e;
      dxmax.filled = true;
    }
    NABU_CHECK_CUDA(cudaEventRecord(zs.dx_done[zk], stream));
    zs.dx_recorded[zk] = true;
    if (defer) {
      NABU_CHECK_CUDA(cudaStreamWaitEvent(ov.side, zs.dx_done[zk], 0));
      if (int e = weight_grads(ov.side)) return e;
      NABU_CHECK_CUDA(cudaEventRecord(ov.ev_done, ov.side));
      NABU_CHECK_CUDA(cudaEventRecord(zs.done[zk], ov.side));
      zs.recorded[zk] = true;
      ov.pending = true;
    } else {
      if (int e = weight_grads(stream)) return e;
      zs.recorded[zk] = false;
    }
    return 0;
  }
  // gates[] now hold dZ (zero for t >= len)
  if (use_h2(B, T, D, H, yT)) {
    // One split per operand and use: x, y and dZ with a global scale for the contractions over the B*T rows (dKx, dKh),
    // dZ again with per-row scales and the weights with a global scale for dX.
    const int D8 = (int)align_up(D, 8);
    void* zh[2] = {w.p2h, w.p3h};
    void* zl[2] = {w.p2l, w.p3l};
    // weight gradients: on `ws` with the planes of `pw` (the caller's workspace, or the library's side scratch)
    auto weight_grads = [&](cudaStream_t ws, const Ws& pw) -> int {
      unsigned* mx = (unsigned*)(pw.glob + 8);
      void* qh[2] = {pw.p2h, pw.p3h};
      void* ql[2] = {pw.p2l, pw.p3l};
      if (int e = split_global(x, D, B * T, D, pw.p1h, pw.p1l, D8, mx + 0, pw.glob + 0, ws)) return e;
      H2Operand xa = {pw.p1h, pw.p1l, D8, nullptr, pw.glob + 0};
      for (int d = 0; d < 2; ++d) {
        if (int e = split_global(g[d], H4, B * T, H4, qh[d], ql[d], H4, mx + 1 + d, pw.glob + 1 + d, ws)) return e;
        H2Operand zb = {qh[d], ql[d], H4, nullptr, pw.glob + 1 + d};
        if (int e = gemm_h2(GEMM_TN, D, H4, B * T, 1.f, xa, zb, 0.f, dkern[d], H4, nullptr, nullptr, pw.gemm, pw.gemm_bytes, ws))
          return e;
      }
      if (T > 1) {
        if (int e = split_global(y, 2 * H, B * yT, 2 * H, pw.p1h, pw.p1l, 2 * H, mx + 3, pw.glob + 3, ws)) return e;
      }
      for (int d = 0; d < 2; ++d) {
        float* dKh = dkern[d] + (size_t)D * H4;
        if (T > 1) {
          GemmSeg seg;
          seg.seg = T - 1; seg.segA = yT; seg.segB = T;
          seg.offA = d == 0 ? 0 : 1; seg.offB = d == 0 ? 1 : 0;
          H2Operand ya = {(const __half*)pw.p1h + d * H, (const __half*)pw.p1l + d * H, 2 * H, nullptr, pw.glob + 3};
          H2Operand zb = {qh[d], ql[d], H4, nullptr, pw.glob + 1 + d};
          if (int e = gemm_h2(GEMM_TN, H, H4, B * (T - 1), 1.f, ya, zb, 0.f, dKh, H4, nullptr, &seg, pw.gemm, pw.gemm_bytes, ws))
            return e;
        } else {
          NABU_CHECK_CUDA(cudaMemsetAsync(dKh, 0, (size_t)H * H4 * sizeof(float), ws));
        }
      }
      return 0;
    };
    if (defer && launched) {
      const Ws side_probe = carve_side(nullptr, H, B, T, D);
      if (int e = overlap_workspace(side_probe.total)) return e;
      const Ws sw = carve_side(ov.ws, H, B, T, D);
      NABU_CHECK_CUDA(cudaStreamWaitEvent(ov.side, ov.ev_rec, 0));
      if (int e = weight_grads(ov.side, sw)) return e;
      NABU_CHECK_CUDA(cudaEventRecord(ov.ev_done, ov.side));
      ov.pending = true;
    } else {
      if (int e = weight_grads(stream, w)) return e;
    }
    unsigned* mx = (unsigned*)(w.glob + 8);
    if (dx) {
      // dX = [dZ_fw | dZ_bw] . [Kx_fw | Kx_bw]^T as ONE contraction over K = 8H (a second GEMM accumulating into dX
      // re-reads and re-writes C: 3.3 ms instead of 1.9 ms at cfg-3).  The planes P2h|P2l and P3h|P3l are adjacent in
      // the workspace and hold the [B*T, 8H] hi and lo operands, the weight planes likewise.
      const bool joint = (char*)w.p2l == (char*)w.p2h + (size_t)B * T * H4 * 2 && (char*)w.p3l == (char*)w.p3h + (size_t)B * T * H4 * 2 &&
                         (char*)w.wl[0] >= (char*)w.wh[0] && (char*)w.wl[1] >= (char*)w.wh[1] &&
                         (size_t)((char*)w.wh[1] - (char*)w.wh[0]) >= (size_t)D * 2 * H4 * 2 && gemm_h2_eligible(GEMM_NT, B * T, D, 2 * H4);
      if (joint) {
        if (int e = split_rows_pair(g[0], g[1], H4, B * T, H4, w.p2h, w.p3h, w.row2, stream)) return e;
        if (int e = split_global_pair(kern[0], kern[1], H4, D, H4, w.wh[0], w.wh[1], mx + 4, w.glob + 4, stream)) return e;
        H2Operand za = {w.p2h, w.p3h, 2 * H4, w.row2, nullptr};
        H2Operand kb = {w.wh[0], w.wh[1], 2 * H4, nullptr, w.glob + 4};
        if (int e = gemm_h2(GEMM_NT, B * T, D, 2 * H4, 1.f, za, kb, 0.f, dx, D, nullptr, nullptr, nullptr, 0, stream)) return e;
      } else {
        for (int d = 0; d < 2; ++d) {
          if (int e = split_rows(g[d], H4, B * T, H4, zh[d], zl[d], H4, w.row2, stream)) return e;
          if (int e = split_global(kern[d], H4, D, H4, w.wh[d], w.wl[d], H4, mx + 4 + d, w.glob + 4 + d, stream)) return e;
          H2Operand za = {zh[d], zl[d], H4, w.row2, nullptr};
          H2Operand kb = {w.wh[d], w.wl[d], H4, nullptr, w.glob + 4 + d};
          if (int e = gemm_h2(GEMM_NT, B * T, D, H4, 1.f, za, kb, d == 0 ? 0.f : 1.f, dx, D, nullptr, nullptr, nullptr, 0, stream))
            return e;
        }
      }
    }
    return 0;
  }
  for (int d = 0; d < 2; ++d) {
    // dKx = X^T . dZ
    if (int e = gemm(GEMM_TN, D, H4, B * T, 1.f, x, D, g[d], H4, 0.f, dkern[d], H4, nullptr, nullptr, w.gemm,
                      w.gemm_bytes, stream))
      return e;
    // dKh = Hprev^T . dZ ; fw: Hprev[b,t] = y[b,t-1,:H] ; bw: Hprev[b,t] = y[b,t+1,H:]
    float* dKh = dkern[d] + (size_t)D * H4;
    if (T > 1) {
      GemmSeg seg;
      seg.seg = T - 1; seg.segA = yT; seg.segB = T;
      seg.offA = d == 0 ? 0 : 1; seg.offB = d == 0 ? 1 : 0;
      if (int e = gemm(GEMM_TN, H, H4, B * (T - 1), 1.f, y + d * H, 2 * H, g[d], H4, 0.f, dKh, H4, nullptr, &seg,
                        w.gemm, w.gemm_bytes, stream))
        return e;
    } else {
      NABU_CHECK_CUDA(cudaMemsetAsync(dKh, 0, (size_t)H * H4 * sizeof(float), stream));
    }
    // dX (+)= dZ . Kx^T
    if (dx)
      if (int e = gemm(GEMM_NT, B * T, D, H4, 1.f, g[d], H4, kern[d], H4, d == 0 ? 0.f : 1.f, dx, D, nullptr, nullptr,
                        nullptr, 0, stream))
        return e;
  }
  return 0;
}
