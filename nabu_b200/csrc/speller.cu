// Teacher-forced Speller forward / backward over the whole target sequence (SURVEY.md section 8
// rows a6-a8; replaces models/ed_decoders/rnn_decoder.py:40-82 = dynamic_decode(BasicDecoder(
// ScheduledEmbeddingTrainingHelper(sample_prob=0)), impute_finished=True) around speller.py's cell).
//
// Forward:  values = memory*mask, keys = values.Wm (one GEMM), then U steps of
//           {dec_lstm_step x layers, dec_attn_step}.  Everything the backward needs is written into the
//           caller's `saved` buffer as [U+1] state slots (slot 0 = zero state, slot u+1 = after step u).
// Backward: U reversed steps of {dec_attn_bwd_step, (dec_lstm_bwd_pointwise, dec_matmul_t) x layers}
//           carrying d(state) between steps, then every weight gradient as ONE batched GEMM over all
//           (step, row) pairs (the per-step kernels only produce dz / dq / per-row partials).
// Rows whose target is shorter than U are frozen after their last step (state copied through, zero
// logits), exactly like impute_finished=True; their gradients are zero.
#include <cooperative_groups.h>
#include "speller_kernels.cuh"
#include "speller_api.h"
#include "gemm.h"
#include "nabu_b200.h"

namespace nabu {
namespace dec {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void mask_memory_kernel(const float* mem, const int* len, int Tm, int E, float* out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long row = i / E;
  const int b = row / Tm, t = row % Tm;
  out[i] = (t < len[b]) ? mem[i] : 0.f;
}

// ids_in[u][r] = (u == 0) ? V-1 : targets[r][u-1]     (rnn_decoder.py:46-47)
// WindowedAttention.initial_alignments (attention.py:344-351): all mass on frame 0
__global__ void onehot0_kernel(float* align, int R, int Tm) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) align[(size_t)r * Tm] = 1.f;
}

// one thread per decoder row: inverse-CDF draw from softmax(logits) with the counter generator (dec_uniform)
__global__ void dec_sample_ids_kernel(const float* logits, long row_stride, int V, int* ids_next, int R, float prob,
                                      unsigned seed, int u) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (!(dec_uniform(seed, 0x40000000u + (uint32_t)u, (uint32_t)r, 0u) < prob)) return;
  const float ub = dec_uniform(seed, 0x40000000u + (uint32_t)u, (uint32_t)r, 1u);
  const float* lg = logits + r * row_stride;
  float mx = lg[0];
  for (int k = 1; k < V; ++k) mx = fmaxf(mx, lg[k]);
  float total = 0.f;
  for (int k = 0; k < V; ++k) total += expf(lg[k] - mx);
  const float thr = ub * total;
  float c = 0.f;
  int id = 0;
  for (int k = 0; k < V; ++k) {
    c += expf(lg[k] - mx);
    if (c < thr) ++id;
  }
  ids_next[r] = min(id, V - 1);
}

__global__ void build_ids_kernel(const int* targets, int ldt, int R, int U, int V, int* ids_in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U * R) return;
  const int u = i / R, r = i % R;
  ids_in[i] = (u == 0) ? V - 1 : targets[(size_t)r * ldt + u - 1];
}

// out[i] = sum_r part[r][i]   (fixed order)
__global__ void reduce_rows_kernel(const float* part, int R, long n, float* out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += part[(size_t)r * n + i];
  out[i] = s;
}

// One-hot rows of the layer-0 kernel: dK0[y][n] = sum over (u, r) with ids_in[u][r] == y of dz0[u][r][n].
// grid = (ceil(4H/256), V); fixed summation order.
__global__ void embedding_grad_kernel(const int* ids_in, const float* dz, int UR, int H4, float* dK0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (n >= H4) return;
  float s = 0.f;
  for (int i = 0; i < UR; ++i)
    if (ids_in[i] == y) s += dz[(size_t)i * H4 + n];
  dK0[(size_t)y * H4 + n] = s;
}

// ------------------------------------------------------------------------------------------------
// backward of the LSTM pointwise part for one layer / step.  One thread per (r, j).
// ------------------------------------------------------------------------------------------------
__global__ void dec_lstm_bwd_pointwise_kernel(const LstmBwdPw p, const float* dh_above, int R, int H) {
  chain_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * H) return;
  lstm_bwd_unit(p, R, H, i / H, i % H, dh_above[i]);
}

// ------------------------------------------------------------------------------------------------
// backward of dec_attn_step: one CTA (256 threads) per row.  NA = ceil(A/256).
// ------------------------------------------------------------------------------------------------
struct AttnBwdArgs {
  int R, Tm, E, H, A, V, F, ksz, U, u;
  const float* dlogits; long dl_row_stride;          // dlogits + r*stride : V values
  const float* outin; long outin_row_stride;         // [h_top, ctx] of this step
  const float* alpha; const float* alpha_prev;       // [R][Tm]
  const float* q; const float* cf;                   // [R][A], [R][Tm][F]
  const float* Wq; const float* Wc; const float* Wd; const float* v; const float* Wo;
  const float* keys; const float* values; const int* mem_len;
  const float* dctx_carry;                           // [R][E]
  float* dalign_carry;                               // [R][Tm] in/out
  float* dh_above;                                   // [R][H] out
  float* dq_save;                                    // [R][A] out
  float* dkeys; float* dvalues;                      // [R][Tm][A], [R][Tm][E] accumulate
  float* dv_part; float* dWd_part; float* dWc_part;  // [R][A], [R][F][A], [R][ksz][F] accumulate
  const int* tlen;
  int prob;                                          // probability_fn: 0 softmax, 1 normalized_sigmoid, 2 sigmoid
  const float* asum;                                 // [R] sum of sigmoids of this step (prob == 1)
  LstmBwdPw pw;                                      // pw.gates != nullptr: dh_above goes straight through the top LSTM layer's cell backward
  int ablate;                                        // NABU_ATTN_ABLATE (profiling only, results are then wrong): 1 skip dcf, 2 skip the conv backward, 4 skip the score backward
};

constexpr int TT = 16;     // memory positions per dpre tile
constexpr int MAXF = 16;   // max numfilt held in registers

// 512 threads per CTA.  TSPLIT = 2 (A <= 256): thread (unit c = tid % 256, half th = tid / 256) takes the tile's positions
// of parity th in the score backward; TSPLIT = 1 (A <= 512): one unit per thread.
// CS = CTAs per decoder row (a thread-block cluster, round 2): with one CTA per row a batch of 64 rows runs on 64 of the
// 148 SMs at 16 warps each and the kernel is bound by instruction issue and load latency.  The memory positions of a row
// are cut into CS contiguous ranges (multiples of TT); a CTA runs phases B-D on its range; the softmax dot product, the
// per-unit sums dq / dv / dWd and the rows of dcf are exchanged through distributed shared memory (two cluster barriers);
// the taus of the location backward, the taps of dWc and the output units of phase F are cut CS ways as well.
template <int TSPLIT, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(512) dec_attn_bwd_step_kernel(const AttnBwdArgs a) {
  constexpr int NT = 512, NW = NT / 32, NA = 1;
  extern __shared__ __align__(16) float sm[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cr = CS > 1 ? (int)cluster.block_rank() : 0;
  const int r = blockIdx.x / CS;
  const int Tm = a.Tm, E = a.E, H = a.H, A = a.A, V = a.V, F = a.F, ksz = a.ksz;
  const int padl = (ksz - 1) / 2;
  auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };   // every piece starts on a 16-byte boundary (attn_bwd_smem)
  float* dl = sm;                          // [V]
  float* oin = dl + r4(V);                 // [H+E]
  float* dctx = oin + r4(H + E);           // [E]
  float* dquery = dctx + r4(E);            // [H]
  float* al = dquery + r4(H);              // [Tm]
  float* ap = al + r4(Tm);                 // [Tm + ksz + 4] zero-padded alpha_prev
  float* dal = ap + r4(Tm + ksz + 4);      // [Tm] dalpha, then de
  float* qs = dal + r4(Tm);                // [A]
  float* dqs = qs + r4(A);                 // [A]
  float* red = dqs + r4(A);                // [32]
  float* cf = red + 32;                    // [Tm][F]
  float* dcf = cf + r4((size_t)Tm * F);    // [Tm][F]
  float* wd = dcf + r4((size_t)Tm * F);    // [F][A]
  float* wc = wd + r4((size_t)F * A);      // [ksz][F]
  float* dpre = wc + r4((size_t)ksz * F);  // [TT][A]
  float* xdot = dpre + (size_t)TT * A;     // [CS] partial softmax dots, slot = source CTA           (CS > 1)
  float* xch = xdot + 4;                   // [CS][2 + F][A] partial dq, dv, dWd, slot = source CTA   (CS > 1)

  if (!(a.u < a.tlen[r])) {                // the whole cluster leaves: no barrier has been touched yet
    chain_wait();
    if (cr == 0) {
      if (a.pw.gates) for (int i = tid; i < H; i += NT) lstm_bwd_unit(a.pw, a.R, H, r, i, 0.f);      // dz = 0 for this row
      else for (int i = tid; i < H; i += NT) a.dh_above[(size_t)r * H + i] = 0.f;
      for (int i = tid; i < A; i += NT) a.dq_save[(size_t)r * A + i] = 0.f;
    }
    return;
  }
  // every CTA of the cluster is running before the first store into a peer's shared memory: arrive here, wait in phase C
  if (CS > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  const int len = min(a.mem_len[r], Tm);
  // this CTA's memory positions [tb, te), output taus [ub, ue)
  const int per = CS > 1 ? ((len + CS * TT - 1) / (CS * TT)) * TT : len;
  const int tb = min(len, cr * per), te = min(len, tb + per);
  const int uper = (Tm + CS - 1) / CS, ub = min(Tm, cr * uper), ue = min(Tm, ub + uper);
  for (int i = tid; i < V; i += NT) dl[i] = a.dlogits[r * a.dl_row_stride + i];
  for (int i = tid; i < H + E; i += NT) oin[i] = a.outin[r * a.outin_row_stride + i];
  for (int i = tid; i < Tm; i += NT) al[i] = a.alpha[(size_t)r * Tm + i];
  for (int i = tid; i < Tm + ksz + 4; i += NT) {
    const int t = i - padl;
    ap[i] = (F > 0 && t >= 0 && t < Tm) ? a.alpha_prev[(size_t)r * Tm + t] : 0.f;
  }
  for (int i = tid; i < A; i += NT) qs[i] = a.q[(size_t)r * A + i];
  for (int i = tb * F + tid; i < te * F; i += NT) cf[i] = a.cf[(size_t)r * Tm * F + i];
  for (int i = tid; i < F * A; i += NT) wd[i] = a.Wd[i];
  for (int i = tid; i < ksz * F; i += NT) wc[i] = a.Wc[i];
  __syncthreads();

  // phase A: d[h_top, ctx] = dlogits . Wo^T ; dctx += carry   (every CTA of the row computes all of it: it is small)
  for (int k = tid; k < H + E; k += NT) {
    float s = 0.f;
    for (int vv = 0; vv < V; ++vv) s = fmaf(dl[vv], a.Wo[(size_t)k * V + vv], s);
    if (k < H) dquery[k] = s;
    else dctx[k - H] = s;
  }
  // everything above read the forward's saved tensors and weights only and ran next to the previous kernel of the chain
  chain_wait();
  __syncthreads();
  for (int i = tid; i < E; i += NT) dctx[i] += a.dctx_carry[(size_t)r * E + i];
  __syncthreads();
  // phase B: dalpha[t] = dctx . values[t] + carry ; dvalues[t] += alpha[t] * dctx
  const float* values = a.values + (size_t)r * Tm * E;
  float* dvalues = a.dvalues + (size_t)r * Tm * E;
  const bool vecE = (E & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.values) | reinterpret_cast<uintptr_t>(a.dvalues)) & 15) == 0;
  for (int t = tb + warp; t < te; t += NW) {
    float s = 0.f;
    const float at = al[t];
    if (vecE && E <= 512) {
      // all of this position's loads (values and the dvalues accumulator) are issued before the first use: the old
      // scalar read-modify-write loop was a chain of E/32 dependent L2 round trips per position
      const float4* vp = reinterpret_cast<const float4*>(values + (size_t)t * E);
      float4* dp = reinterpret_cast<float4*>(dvalues + (size_t)t * E);
      float4 xv[4], xd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i4 = lane + 32 * j;
        if (i4 * 4 < E) { xv[j] = __ldg(vp + i4); xd[j] = dp[i4]; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i4 = lane + 32 * j;
        if (i4 * 4 < E) {
          const float4 dc = *reinterpret_cast<const float4*>(dctx + i4 * 4);
          s += (dc.x * xv[j].x + dc.y * xv[j].y) + (dc.z * xv[j].z + dc.w * xv[j].w);
          xd[j].x = fmaf(at, dc.x, xd[j].x); xd[j].y = fmaf(at, dc.y, xd[j].y);
          xd[j].z = fmaf(at, dc.z, xd[j].z); xd[j].w = fmaf(at, dc.w, xd[j].w);
          dp[i4] = xd[j];
        }
      }
    } else {
      for (int i = lane; i < E; i += 32) {
        s = fmaf(dctx[i], values[(size_t)t * E + i], s);
        dvalues[(size_t)t * E + i] += at * dctx[i];
      }
    }
    s = warp_sum(s);
    if (lane == 0) dal[t] = s + a.dalign_carry[(size_t)r * Tm + t];
  }
  __syncthreads();
  // phase C: probability function backward (components/attention.py:9-13, 41-55), positions [tb, te)
  //   softmax:             de = alpha * (dalpha - sum alpha*dalpha)
  //   normalized_sigmoid:  alpha = s / S  ->  de = (dalpha - sum alpha*dalpha) / S * s * (1 - s),  s = alpha * S
  //   sigmoid:             de = dalpha * alpha * (1 - alpha)
  if (a.prob == 2) {
    for (int t = tb + tid; t < te; t += NT) dal[t] = dal[t] * al[t] * (1.f - al[t]);
    if (CS > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  } else {
    float part = 0.f;
    for (int t = tb + tid; t < te; t += NT) part += al[t] * dal[t];
    float dot = block_reduce(part, red, false);
    if (CS > 1) {
      asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
      if (tid < CS) *cluster.map_shared_rank(xdot + cr, tid) = dot;      // my partial into slot cr of every CTA
      cluster.sync();
      dot = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) dot += xdot[c];
    }
    if (a.prob == 0) {
      for (int t = tb + tid; t < te; t += NT) dal[t] = al[t] * (dal[t] - dot);
    } else {
      const float S = a.asum[r], iS = 1.f / S;
      for (int t = tb + tid; t < te; t += NT) {
        const float sg = al[t] * S;
        dal[t] = (dal[t] - dot) * iS * sg * (1.f - sg);
      }
    }
  }
  __syncthreads();
  // phase D: score backward, tiles of TT memory positions; thread owns attention units tid + 256*i
  const float* keys = a.keys + (size_t)r * Tm * A;
  float* dkeys = a.dkeys + (size_t)r * Tm * A;
  float dq[NA], dv[NA], dWd[NA][MAXF], vreg[NA], wdreg[NA][MAXF];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    dq[i] = 0.f; dv[i] = 0.f;
    const int c = TSPLIT == 2 ? (tid & 255) : tid;
    vreg[i] = c < A ? a.v[c] : 0.f;
#pragma unroll
    for (int f = 0; f < MAXF; ++f) {
      dWd[i][f] = 0.f;
      wdreg[i][f] = (f < F && c < A) ? wd[f * A + c] : 0.f;   // this thread's column of Wd: constant over the positions
    }
  }
  for (int t0 = tb; t0 < te; t0 += TT) {
    const int nt = min(TT, te - t0);
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int c = TSPLIT == 2 ? (tid & 255) : tid;
      const int th = TSPLIT == 2 ? (tid >> 8) : 0;
      if (c < A && !(a.ablate & 4)) {
        // the tile's keys and dkeys accumulators are fetched up front (2*TT independent loads in flight per thread)
        float kv[TT / TSPLIT], dk[TT / TSPLIT];
#pragma unroll
        for (int j = 0; j < TT / TSPLIT; ++j) {
          const int tt = j * TSPLIT + th;
          kv[j] = tt < nt ? __ldg(keys + (size_t)(t0 + tt) * A + c) : 0.f;
          dk[j] = tt < nt ? dkeys[(size_t)(t0 + tt) * A + c] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < TT / TSPLIT; ++j) {
          const int tt = j * TSPLIT + th;
          if (tt < nt) {
            const int t = t0 + tt;
            float cfr[MAXF];                         // location features of this position: one broadcast read each
#pragma unroll
            for (int f = 0; f < MAXF; ++f) cfr[f] = f < F ? cf[t * F + f] : 0.f;
            float pre = qs[c] + kv[j];
#pragma unroll
            for (int f = 0; f < MAXF; ++f) pre = fmaf(cfr[f], wdreg[i][f], pre);
            const float s = tanh_fast(pre);
            const float de = dal[t];
            const float dp = de * vreg[i] * (1.f - s * s);
            dq[i] += dp;
            dv[i] = fmaf(de, s, dv[i]);
#pragma unroll
            for (int f = 0; f < MAXF; ++f) dWd[i][f] = fmaf(cfr[f], dp, dWd[i][f]);
            dkeys[(size_t)t * A + c] = dk[j] + dp;
            dpre[tt * A + c] = dp;
          }
        }
      }
    }
    __syncthreads();
    // dcf[t][f] = sum_c dpre[t][c] * Wd[f][c]: one warp per memory position, lanes stride the attention units (both
    // operands are then read at consecutive addresses; the (t, f)-per-thread mapping hit one bank with 10 lanes).
    // The row goes to every CTA of the cluster (phase E needs the neighbours' rows within the filter's reach).
    if (!(a.ablate & 1))
    for (int tt = warp; tt < nt; tt += NW) {
      float acc0[MAXF];
#pragma unroll
      for (int f = 0; f < MAXF; ++f) acc0[f] = 0.f;
      for (int c = lane; c < A; c += 32) {
        const float d0 = dpre[tt * A + c];
#pragma unroll
        for (int f = 0; f < MAXF; ++f)
          if (f < F) acc0[f] = fmaf(d0, wd[f * A + c], acc0[f]);
      }
#pragma unroll
      for (int f = 0; f < MAXF; ++f)
        if (f < F) {
          const float s0 = warp_sum(acc0[f]);
          if (CS == 1) {
            if (lane == 0) dcf[(t0 + tt) * F + f] = s0;
          } else if (lane < CS) {
            *cluster.map_shared_rank(dcf + (t0 + tt) * F + f, lane) = s0;
          }
        }
    }
    __syncthreads();
  }
  for (int i = tid; i < (Tm - len) * F; i += NT) dcf[len * F + i] = 0.f;
  {
    const int c = TSPLIT == 2 ? (tid & 255) : tid;
    const int th = TSPLIT == 2 ? (tid >> 8) : 0;
    // the upper half hands its partial sums to the lower half through the dpre scratch ([2 + F][A] <= [TT][A] floats)
    if (TSPLIT == 2 && th == 1 && c < A) {
      dpre[c] = dq[0];
      dpre[A + c] = dv[0];
#pragma unroll
      for (int f = 0; f < MAXF; ++f)
        if (f < F) dpre[(2 + f) * A + c] = dWd[0][f];
    }
    __syncthreads();
    float q0 = dq[0], v0 = dv[0];
    if (th == 0 && c < A) {
      if (TSPLIT == 2) {
        q0 += dpre[c]; v0 += dpre[A + c];
#pragma unroll
        for (int f = 0; f < MAXF; ++f)
          if (f < F) dWd[0][f] += dpre[(2 + f) * A + c];
      }
      if (CS > 1) {                          // this CTA's sums into slot cr of every other CTA
        const size_t slot = (size_t)cr * (2 + F) * A;
        for (int pc = 0; pc < CS; ++pc) {
          if (pc == cr) continue;
          float* px = cluster.map_shared_rank(xch, pc) + slot;
          px[c] = q0;
          px[A + c] = v0;
#pragma unroll
          for (int f = 0; f < MAXF; ++f)
            if (f < F) px[(2 + f) * A + c] = dWd[0][f];
        }
      }
    }
    if (CS > 1) cluster.sync();              // the sums and every CTA's rows of dcf have arrived; no remote access after this
    if (th == 0 && c < A) {
      if (CS > 1) {
        // summed in CTA order whatever the CTA, so that every CTA of the row holds the same bits of dq
        float qs_ = 0.f, vs_ = 0.f, ws_[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) ws_[f] = 0.f;
        for (int pc = 0; pc < CS; ++pc) {
          const float* px = xch + (size_t)pc * (2 + F) * A;
          qs_ += pc == cr ? q0 : px[c];
          vs_ += pc == cr ? v0 : px[A + c];
#pragma unroll
          for (int f = 0; f < MAXF; ++f)
            if (f < F) ws_[f] += pc == cr ? dWd[0][f] : px[(2 + f) * A + c];
        }
        q0 = qs_; v0 = vs_;
#pragma unroll
        for (int f = 0; f < MAXF; ++f) dWd[0][f] = ws_[f];
      }
      dqs[c] = q0;
      if (cr == 0) {
        a.dq_save[(size_t)r * A + c] = q0;
        a.dv_part[(size_t)r * A + c] += v0;
      }
#pragma unroll
      for (int f = 0; f < MAXF; ++f)
        if (f < F && f % CS == cr) a.dWd_part[((size_t)r * F + f) * A + c] += dWd[0][f];
    }
  }
  __syncthreads();
  // phase E: location-feature backward, taus [ub, ue) and every CS-th tap product of dWc
  if (F > 0 && !(a.ablate & 2)) {
    // dalign_prev[tau] = sum_{k,f} dcf[tau - k + padl][f] * Wc[k][f]
    // (the taps are split over the groups of 128 threads; partial sums meet in the dpre scratch)
    {
      // tap groups of GS threads, GS the power of two that covers this CTA's taus (64 threads x 8 groups for 63 taus)
      int GS = 32;
      while (GS < ue - ub && GS < NT) GS <<= 1;
      const int NG = NT / GS;
      const bool split2 = NG * Tm <= TT * A;            // room for every group's partial sums in the scratch
      const int half = split2 ? tid / GS : 0, kh = split2 ? (ksz + NG - 1) / NG : ksz;
      const int k0 = half * kh, k1 = min(ksz, k0 + kh);
      for (int tau = ub + (split2 ? (tid & (GS - 1)) : tid); tau < ue; tau += split2 ? GS : NT) {
        float s = 0.f;
        for (int k = k0; k < k1; ++k) {
          const int t = tau - k + padl;
          if (t >= 0 && t < Tm) {
            if ((F & 1) == 0) {
              const float2* dr = reinterpret_cast<const float2*>(dcf + t * F);
              const float2* wr = reinterpret_cast<const float2*>(wc + k * F);
              for (int f2 = 0; f2 < F / 2; ++f2) {
                const float2 x = dr[f2], y = wr[f2];
                s = fmaf(x.x, y.x, s);
                s = fmaf(x.y, y.y, s);
              }
            } else {
              for (int f = 0; f < F; ++f) s = fmaf(dcf[t * F + f], wc[k * F + f], s);
            }
          }
        }
        if (split2) dpre[half * Tm + tau] = s;
        else a.dalign_carry[(size_t)r * Tm + tau] = s;
      }
      __syncthreads();
      if (split2)
        for (int tau = ub + tid; tau < ue; tau += NT) {
          float s = 0.f;
          for (int gq = 0; gq < NG; ++gq) s += dpre[gq * Tm + tau];
          a.dalign_carry[(size_t)r * Tm + tau] = s;
        }
    }
    // dWc[k][f] += sum_t alpha_prev[t + k - padl] * dcf[t][f]
    // a thread takes 4 consecutive taps of one filter (a sliding window over alpha_prev in registers: two reads per
    // 4 FMAs); the quads are dealt out to the CTAs of the row in turn
    for (int i = cr + CS * tid; i < ((ksz + 3) / 4) * F; i += CS * NT) {
      const int f = i % F, kq = (i / F) * 4;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      float a0 = ap[kq], a1 = ap[kq + 1], a2 = ap[kq + 2];
#pragma unroll 4
      for (int t = 0; t < Tm; ++t) {
        const float a3 = ap[t + kq + 3], dd = dcf[t * F + f];
        s0 = fmaf(a0, dd, s0); s1 = fmaf(a1, dd, s1); s2 = fmaf(a2, dd, s2); s3 = fmaf(a3, dd, s3);
        a0 = a1; a1 = a2; a2 = a3;
      }
      float* out = a.dWc_part + (size_t)r * ksz * F;
      out[kq * F + f] += s0;
      if (kq + 1 < ksz) out[(kq + 1) * F + f] += s1;
      if (kq + 2 < ksz) out[(kq + 2) * F + f] += s2;
      if (kq + 3 < ksz) out[(kq + 3) * F + f] += s3;
    }
  } else {
    for (int tau = ub + tid; tau < ue; tau += NT) a.dalign_carry[(size_t)r * Tm + tau] = 0.f;
  }
  // phase F: dh_top = dquery_part + dq . Wq^T   (warp per output unit, lanes over A; the CTAs take turns over the groups)
  for (int k = warp + cr * 4 * NW; k < H; k += CS * 4 * NW) {      // 4 output units per warp iteration: independent load chains
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < A; c += 32) {
      const float dqc = dqs[c];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + NW * j < H) s[j] = fmaf(dqc, __ldg(a.Wq + (size_t)(k + NW * j) * A + c), s[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float t = warp_sum(s[j]);
      if (lane == 0 && k + NW * j < H) {
        if (a.pw.gates) dquery[k + NW * j] += t;             // finished below, one thread per unit
        else a.dh_above[(size_t)r * H + k + NW * j] = dquery[k + NW * j] + t;
      }
    }
  }
  if (a.pw.gates) {
    // the top LSTM layer's cell backward on this CTA's units (group k / (4 NW) belongs to CTA group % CS)
    __syncthreads();
    for (int k = tid; k < H; k += NT)
      if ((k / (4 * NW)) % CS == cr) lstm_bwd_unit(a.pw, a.R, H, r, k, dquery[k]);
  }
}

inline size_t attn_bwd_smem(int Tm, int E, int H, int A, int V, int F, int ksz, int cs = 4) {
  auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  return (r4(V) + r4(H + E) + r4(E) + r4(H) + r4(Tm) + r4(Tm + ksz + 4) + r4(Tm) + r4(A) + r4(A) + 32 + 2 * r4((size_t)Tm * F) +
          r4((size_t)F * A) + r4((size_t)ksz * F) + (size_t)TT * A + 4 + (cs > 1 ? (size_t)cs * (2 + F) * A : 0)) * sizeof(float);
}

// CTAs per decoder row: as many as keep the grid within one wave of the 148 SMs (NABU_ATTN_BWD_CLUSTER=1|2|4 forces)
inline int attn_bwd_cluster(int rows) {
  static int forced = -1;
  if (forced < 0) forced = getenv("NABU_ATTN_BWD_CLUSTER") ? atoi(getenv("NABU_ATTN_BWD_CLUSTER")) : 0;
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  return rows * 4 <= 160 ? 4 : rows * 2 <= 160 ? 2 : 1;
}

// the kernel for (A, CTAs per row); its dynamic shared memory limit raised once per variant
inline int attn_bwd_launch(const AttnBwdArgs& a, int rows, cudaStream_t stream) {
  const int cs = attn_bwd_cluster(rows);
  const size_t smem = attn_bwd_smem(a.Tm, a.E, a.H, a.A, a.V, a.F, a.ksz, cs);
  NABU_REQUIRE(smem <= (size_t)max_smem_optin(), "attention backward: memory too long for the kernel (Tm=%d)", a.Tm);
  void (*fn)(const AttnBwdArgs) =
      a.A <= 256 ? (cs == 4 ? dec_attn_bwd_step_kernel<2, 4> : cs == 2 ? dec_attn_bwd_step_kernel<2, 2> : dec_attn_bwd_step_kernel<2, 1>)
                 : (cs == 4 ? dec_attn_bwd_step_kernel<1, 4> : cs == 2 ? dec_attn_bwd_step_kernel<1, 2> : dec_attn_bwd_step_kernel<1, 1>);
  if (smem > 48 * 1024) NABU_CHECK_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NABU_CHECK_CUDA(chain_launch(fn, dim3(rows * cs), dim3(512), smem, stream, a));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// buffer carving
// ------------------------------------------------------------------------------------------------
struct Saved {        // written by the forward, read by the backward
  float* values; float* keys;                 // [B][Tm][E], [B][Tm][A]
  int* ids_in;                                // [U][B]
  float* hT[4]; float* h[4]; float* c[4];     // [(U+1)][H][B], [(U+1)][B][H], [(U+1)][B][H]
  float* gates[4];                            // [U][B][4H]
  float* ctx; float* ctxT; float* align;      // [(U+1)][B][E], [(U+1)][E][B], [(U+1)][B][Tm]
  float* q; float* cf; float* outin;          // [U][B][A], [U][B][Tm][F], [B][U][H+E]
  float* asum;                                // [U][B] sum of sigmoids (probability_fn = normalized_sigmoid)
  float* out[4]; float* outT[4];              // dropout only: cell outputs [(U+1)][B][H] (slot u+1 = step u), [H][B] scratch
  float* wf[4]; float* wb[4];                 // the cells' weight slices as the step / transposed-matmul kernels stage them
  size_t total;
};

Saved carve_saved(void* base, const nabu_speller_desc_t& d) {
  Saved s;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloats) { float* q = (float*)(p + off); off += align_up(nfloats * sizeof(float), 256); return q; };
  const size_t B = d.B, Tm = d.Tm, E = d.E, H = d.H, A = d.A, U = d.U, F = d.attention == 1 ? d.numfilt : 0;
  s.values = take(B * Tm * E);
  s.keys = take(B * Tm * A);
  s.ids_in = (int*)take(U * B);
  for (int l = 0; l < d.num_layers; ++l) {
    s.hT[l] = take((U + 1) * H * B);
    s.h[l] = take((U + 1) * B * H);
    s.c[l] = take((U + 1) * B * H);
    s.gates[l] = take(U * B * 4 * H);
  }
  s.ctx = take((U + 1) * B * E);
  s.ctxT = take((U + 1) * E * B);
  s.align = take((U + 1) * B * Tm);
  s.q = take(U * B * A);
  s.cf = take(U * B * Tm * (F ? F : 1));
  s.outin = take(B * U * (H + E));
  s.asum = take(U * B);
  for (int l = 0; l < d.num_layers; ++l) {
    s.wf[l] = take(relayout_fwd_floats(d, l));
    s.wb[l] = take((l == 0 ? E + H : 2 * H) * 4 * H);
  }
  const bool drop = d.dropout_keep > 0.f && d.dropout_keep < 1.f;
  for (int l = 0; l < 4; ++l) {
    s.out[l] = (drop && l < d.num_layers) ? take((U + 1) * B * H) : nullptr;
    s.outT[l] = (drop && l < d.num_layers) ? take(H * B) : nullptr;
  }
  s.total = off;
  return s;
}

struct Work {         // backward scratch
  float* dh_carry[4]; float* dc_carry[4];     // [B][H]
  float* dh_above; float* dzT[4];             // [B][H], per layer [4H][B] (the fused cell backward of layer l-1 writes its dz while layer l's is still being read)
  float* dctx_carry; float* dalign_carry;     // [B][E], [B][Tm]
  float* dq; float* dkeys; float* dvalues;    // [U][B][A], [B][Tm][A], [B][Tm][E]
  float* dv_part; float* dWd_part; float* dWc_part;
  float* gemm; size_t gemm_bytes;
  size_t zero_bytes;                          // leading region that must be zeroed
  size_t total;
};

Work carve_work(void* base, const nabu_speller_desc_t& d) {
  Work w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloats) { float* q = (float*)(p + off); off += align_up(nfloats * sizeof(float), 256); return q; };
  const size_t B = d.B, Tm = d.Tm, E = d.E, H = d.H, A = d.A, U = d.U;
  const size_t F = d.attention == 1 ? d.numfilt : 0, ksz = d.attention == 1 ? d.filtersize : 1;
  for (int l = 0; l < d.num_layers; ++l) { w.dh_carry[l] = take(B * H); w.dc_carry[l] = take(B * H); }
  w.dctx_carry = take(B * E);
  w.dalign_carry = take(B * Tm);
  w.dkeys = take(B * Tm * A);
  w.dvalues = take(B * Tm * E);
  w.dv_part = take(B * A);
  w.dWd_part = take(B * (F ? F : 1) * A);
  w.dWc_part = take(B * ksz * (F ? F : 1));
  w.zero_bytes = off;
  w.dh_above = take(B * H);
  for (int l = 0; l < d.num_layers; ++l) w.dzT[l] = take(4 * H * B);
  w.dq = take(U * B * A);
  w.gemm = (float*)(p + off); w.gemm_bytes = sgemm_workspace_bytes(); off += w.gemm_bytes;
  w.total = off;
  return w;
}

int init_window_alignments(float* align, int R, int Tm, cudaStream_t stream) {
  KernelScope ks("onehot0", stream);
  onehot0_kernel<<<ceil_div(R, 256), 256, 0, stream>>>(align, R, Tm);
  NABU_CHECK_LAUNCH();
  return 0;
}

int check_desc(const nabu_speller_desc_t& d) {
  NABU_REQUIRE(d.B > 0 && d.Tm > 0 && d.E > 0 && d.V > 1 && d.H > 0 && d.U > 0, "speller: bad shape");
  NABU_REQUIRE(d.num_layers >= 1 && d.num_layers <= 4, "speller: num_layers=%d not in 1..4", d.num_layers);
  NABU_REQUIRE(d.H % 8 == 0 && d.E % 8 == 0, "speller: num_units=%d and memory dim=%d must be multiples of 8", d.H, d.E);
  NABU_REQUIRE(d.A > 0 && d.A <= 512, "speller: attention units=%d not in 1..512", d.A);
  NABU_REQUIRE(d.attention >= 0 && d.attention <= 2, "speller: attention %d not in 0..2", d.attention);
  if (d.attention == 2)
    NABU_REQUIRE(d.numfilt >= 0 && d.filtersize >= 1 && d.Tm <= 2048,
                 "speller: windowed attention needs left_window_width >= 0, right_window_width >= 1 (the reference slices "
                 "[:, :-right]) and Tm <= 2048 (left=%d right=%d Tm=%d)", d.numfilt, d.filtersize, d.Tm);
  NABU_REQUIRE(d.probability_fn >= 0 && d.probability_fn <= 2, "speller: probability_fn=%d not in 0..2", d.probability_fn);
  if (d.attention == 1)
    NABU_REQUIRE(d.numfilt >= 1 && d.numfilt <= MAXF && d.filtersize >= 1, "speller: numfilt=%d (max %d), filtersize=%d",
                 d.numfilt, MAXF, d.filtersize);
  return 0;
}

size_t relayout_fwd_floats(const nabu_speller_desc_t& d, int l) {
  return (size_t)(l == 0 ? d.E + d.H : 2 * d.H) * 4 * d.H;
}
int relayout_fwd(const nabu_speller_desc_t& d, const nabu_speller_params_t& p, float* const* out, cudaStream_t stream) {
  for (int l = 0; l < d.num_layers; ++l) {
    const long n = (long)relayout_fwd_floats(d, l);
    KernelScope ks("dec_relayout", stream);
    if (l == 0) dec_relayout_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p.cell_kernel[0], out[0], d.H, d.E, d.V, d.H, d.V + d.E);
    else dec_relayout_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p.cell_kernel[l], out[l], d.H, d.H, 0, d.H, d.H);
    NABU_CHECK_LAUNCH();
  }
  return 0;
}

// Launch one full decoder step (all LSTM layers + attention/projection).  Shared with las_beam.cu.
int launch_step(const nabu_speller_desc_t& d, const nabu_speller_params_t& p, int R, int rows_per_mem,
                const int* ids, const float* keys, const float* values, const int* mem_len,
                float* const* hT_prev, float* const* h_prev, float* const* c_prev, const float* ctx_prev,
                const float* ctxT_prev, const float* align_prev,
                float* const* hT_new, float* const* h_new, float* const* c_new, float* ctx_new, float* ctxT_new,
                float* align_new, float* const* gates_out, float* logits, long logits_row_stride,
                float temperature, float* q_save, float* cf_save, float* outin_save, long outin_row_stride,
                const int* tlen, int u, const int* done, cudaStream_t stream, float* asum_save,
                float* const* out_new, float* const* outT_new, float keep, unsigned seed, const float* const* cell_relayout) {
  const int H = d.H, E = d.E, V = d.V;
  for (int l = 0; l < d.num_layers; ++l) {
    LstmStepArgs a = {};
    if (l == 0) { a.inT0 = ctxT_prev; a.K0 = E; a.w0 = V; a.inT1 = hT_prev[0]; a.K1 = H; a.w1 = V + E; a.ids = ids; }
    else { a.inT0 = outT_new ? outT_new[l - 1] : hT_new[l - 1]; a.K0 = H; a.w0 = 0; a.inT1 = hT_prev[l]; a.K1 = H; a.w1 = H; a.ids = nullptr; }
    a.W = p.cell_kernel[l]; a.bias = p.cell_bias[l]; a.H = H; a.R = R;
    a.Wr = cell_relayout ? cell_relayout[l] : nullptr;
    a.c_prev = c_prev[l]; a.h_prev = h_prev[l];
    a.c_new = c_new[l]; a.h_new = h_new[l]; a.hT_new = hT_new[l];
    a.gates_out = gates_out ? gates_out[l] : nullptr;
    a.tlen = tlen; a.u = u; a.done = done;
    a.out_new = out_new ? out_new[l] : nullptr; a.outT_new = outT_new ? outT_new[l] : nullptr;
    a.keep = keep; a.seed = seed; a.layer = l;
    const size_t smem = ((size_t)(a.K0 + a.K1) * 8 + SK_KSPLIT * ROWS * 8) * sizeof(float) + dk_stage_bytes();
    NABU_REQUIRE(smem <= (size_t)max_smem_optin(), "speller: LSTM input too wide for the step kernel");
    if (smem > 48 * 1024)
      NABU_CHECK_CUDA(cudaFuncSetAttribute(dec_lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelScope ks("dec_lstm_step", stream);
    NABU_CHECK_CUDA(chain_launch(dec_lstm_step_kernel, dim3(H / 2, ceil_div(R, ROWS)), dim3(SK_THREADS), smem, stream, a));
  }
  AttnStepArgs a = {};
  a.R = R; a.Tm = d.Tm; a.E = E; a.H = H; a.A = d.A; a.V = V;
  a.F = d.attention == 1 ? d.numfilt : 0; a.ksz = d.attention == 1 ? d.filtersize : 1;
  a.rows_per_mem = rows_per_mem;
  a.h_top = out_new ? out_new[d.num_layers - 1] : h_new[d.num_layers - 1];
  a.Wq = p.query_kernel; a.Wc = p.conv_kernel; a.Wd = p.conv_dense_kernel; a.v = p.attention_v;
  a.Wo = p.out_kernel; a.bo = p.out_bias;
  a.keys = keys; a.values = values; a.mem_len = mem_len;
  a.align_prev = align_prev; a.ctx_prev = ctx_prev;
  a.align_new = align_new; a.ctx_new = ctx_new; a.ctxT_new = ctxT_new;
  a.logits = logits; a.logits_row_stride = logits_row_stride; a.temperature = temperature;
  a.q_save = q_save; a.cf_save = cf_save; a.outin_save = outin_save; a.outin_row_stride = outin_row_stride;
  a.tlen = tlen; a.u = u; a.done = done;
  a.prob = d.probability_fn; a.asum_save = asum_save;
  a.win_left = d.attention == 2 ? d.numfilt : -1; a.win_right = d.filtersize;
  NABU_REQUIRE(attn_step_smem(d.Tm, E, H, d.A, a.F, a.ksz, attn_step_cluster(R, a.rows_per_mem)) <= (size_t)max_smem_optin(), "speller: memory too long for the attention step kernel (Tm=%d)", d.Tm);
  KernelScope ks("dec_attn_step", stream);
  NABU_CHECK_CUDA(attn_step_launch(a, R, nullptr, stream));
  return 0;
}

// values = memory*mask ; keys = values.Wm
int prepare_memory(const nabu_speller_desc_t& d, const nabu_speller_params_t& p, const float* memory,
                   const int* mem_len, float* values, float* keys, cudaStream_t stream) {
  const long n = (long)d.B * d.Tm * d.E;
  {
    KernelScope ks("mask_memory", stream);
    mask_memory_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(memory, mem_len, d.Tm, d.E, values, n);
    NABU_CHECK_LAUNCH();
  }
  return gemm(GEMM_NN, d.B * d.Tm, d.A, d.E, 1.f, values, d.E, p.memory_kernel, d.A, 0.f, keys, d.A, nullptr, nullptr,
               nullptr, 0, stream);
}

}  // namespace dec
}  // namespace nabu

using namespace nabu;
using namespace nabu::dec;

extern "C" size_t nabu_speller_saved_bytes(const nabu_speller_desc_t* d) {
  if (check_desc(*d)) return 0;
  return carve_saved(nullptr, *d).total;
}
extern "C" size_t nabu_speller_workspace_bytes(const nabu_speller_desc_t* d) {
  if (check_desc(*d)) return 0;
  return carve_work(nullptr, *d).total;
}

extern "C" int nabu_speller_fwd(const nabu_speller_desc_t* dp, const nabu_speller_params_t* p, const float* memory,
                                const int* mem_len, const int* targets, int ldt, const int* target_len,
                                float* logits, void* saved, void* workspace, size_t ws_bytes, void* stream_) {
  (void)workspace; (void)ws_bytes;
  cudaStream_t stream = (cudaStream_t)stream_;
  const nabu_speller_desc_t& d = *dp;
  if (int e = check_desc(d)) return e;
  NABU_REQUIRE(ldt >= d.U, "speller_fwd: targets row stride %d < U=%d", ldt, d.U);
  Saved s = carve_saved(saved, d);
  const size_t B = d.B, Tm = d.Tm, E = d.E, H = d.H, U = d.U;
  if (int e = prepare_memory(d, *p, memory, mem_len, s.values, s.keys, stream)) return e;
  if (int e = relayout_fwd(d, *p, s.wf, stream)) return e;
  {
    KernelScope ks("build_ids", stream);
    build_ids_kernel<<<ceil_div(d.U * d.B, 256), 256, 0, stream>>>(targets, ldt, d.B, d.U, d.V, s.ids_in);
    NABU_CHECK_LAUNCH();
  }
  // zero state = slot 0 of every state array
  for (int l = 0; l < d.num_layers; ++l) {
    NABU_CHECK_CUDA(cudaMemsetAsync(s.hT[l], 0, H * B * sizeof(float), stream));
    NABU_CHECK_CUDA(cudaMemsetAsync(s.h[l], 0, B * H * sizeof(float), stream));
    NABU_CHECK_CUDA(cudaMemsetAsync(s.c[l], 0, B * H * sizeof(float), stream));
  }
  NABU_CHECK_CUDA(cudaMemsetAsync(s.ctx, 0, B * E * sizeof(float), stream));
  NABU_CHECK_CUDA(cudaMemsetAsync(s.ctxT, 0, E * B * sizeof(float), stream));
  NABU_CHECK_CUDA(cudaMemsetAsync(s.align, 0, B * Tm * sizeof(float), stream));
  if (d.attention == 2)
    if (int e = init_window_alignments(s.align, (int)B, (int)Tm, stream)) return e;
  const size_t F = d.attention == 1 ? d.numfilt : 0;
  const bool drop = d.dropout_keep > 0.f && d.dropout_keep < 1.f;
  for (int u = 0; u < d.U; ++u) {
    float *hTp[4], *hp[4], *cp[4], *hTn[4], *hn[4], *cn[4], *go[4], *on[4], *oTn[4];
    for (int l = 0; l < d.num_layers; ++l) {
      hTp[l] = s.hT[l] + (size_t)u * H * B; hTn[l] = s.hT[l] + (size_t)(u + 1) * H * B;
      hp[l] = s.h[l] + (size_t)u * B * H; hn[l] = s.h[l] + (size_t)(u + 1) * B * H;
      cp[l] = s.c[l] + (size_t)u * B * H; cn[l] = s.c[l] + (size_t)(u + 1) * B * H;
      go[l] = s.gates[l] + (size_t)u * B * 4 * H;
      on[l] = drop ? s.out[l] + (size_t)(u + 1) * B * H : nullptr; oTn[l] = s.outT[l];
    }
    if (int e = launch_step(d, *p, d.B, 1, s.ids_in + (size_t)u * B, s.keys, s.values, mem_len, hTp, hp, cp,
                            s.ctx + (size_t)u * B * E, s.ctxT + (size_t)u * E * B, s.align + (size_t)u * B * Tm,
                            hTn, hn, cn, s.ctx + (size_t)(u + 1) * B * E, s.ctxT + (size_t)(u + 1) * E * B,
                            s.align + (size_t)(u + 1) * B * Tm, go, logits + (size_t)u * d.V, (long)U * d.V, 1.f,
                            s.q + (size_t)u * B * d.A, F ? s.cf + (size_t)u * B * Tm * F : nullptr,
                            s.outin + (size_t)u * (H + E), (long)U * (H + E), target_len, u, nullptr, stream,
                            s.asum + (size_t)u * B, drop ? on : nullptr, drop ? oTn : nullptr, d.dropout_keep, d.seed, s.wf))
      return e;
    // ScheduledEmbeddingTrainingHelper (rnn_decoder.py:59-64): with probability sample_prob the NEXT input token is a
    // draw from Categorical(logits of this step) instead of the teacher's
    if (d.sample_prob > 0.f && u + 1 < d.U) {
      KernelScope ks("dec_sample_ids", stream);
      dec_sample_ids_kernel<<<ceil_div(d.B, 128), 128, 0, stream>>>(logits + (size_t)u * d.V, (long)U * d.V, d.V,
                                                                  s.ids_in + (size_t)(u + 1) * B, d.B, d.sample_prob, d.seed, u);
      NABU_CHECK_LAUNCH();
    }
  }
  return 0;
}

extern "C" int nabu_speller_bwd(const nabu_speller_desc_t* dp, const nabu_speller_params_t* p, const float* memory,
                                const int* mem_len, const int* targets, int ldt, const int* target_len,
                                const float* dlogits, void* saved, float* dmemory,
                                const nabu_speller_params_t* g, void* workspace, size_t ws_bytes, void* stream_) {
  (void)memory; (void)targets; (void)ldt;
  cudaStream_t stream = (cudaStream_t)stream_;
  const nabu_speller_desc_t& d = *dp;
  if (int e = check_desc(d)) return e;
  Saved s = carve_saved(saved, d);
  Work w = carve_work(workspace, d);
  NABU_REQUIRE(ws_bytes >= w.total, "speller_bwd: workspace %zu < %zu bytes", ws_bytes, w.total);
  const int B = d.B, Tm = d.Tm, E = d.E, H = d.H, A = d.A, V = d.V, U = d.U, NL = d.num_layers;
  const int F = d.attention == 1 ? d.numfilt : 0, ksz = d.attention == 1 ? d.filtersize : 1;
  const int H4 = 4 * H;
  NABU_CHECK_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, stream));

  const size_t smem_mm = ((size_t)H4 * 8 + MT_KSPLIT * ROWS * 8) * sizeof(float) + dk_stage_bytes();
  NABU_REQUIRE(smem_mm <= (size_t)max_smem_optin(), "speller_bwd: num_units too large");
  if (smem_mm > 48 * 1024)
    NABU_CHECK_CUDA(cudaFuncSetAttribute(dec_matmul_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mm));

  for (int l = 0; l < NL; ++l) {
    const int N = l == 0 ? E + H : 2 * H, row0 = l == 0 ? V : 0;
    KernelScope ks("dec_relayout", stream);
    dec_relayout_bwd_kernel<<<(unsigned)(((long)N * H4 + 255) / 256), 256, 0, stream>>>(p->cell_kernel[l], s.wb[l], H4, H4, row0, N);
    NABU_CHECK_LAUNCH();
  }
  // NABU_DEC_FUSE=0: the cell backward as a kernel of its own (round 1) instead of inside the kernels that produce its input
  static const bool fuse = !(getenv("NABU_DEC_FUSE") && atoi(getenv("NABU_DEC_FUSE")) == 0);
  for (int u = U - 1; u >= 0; --u) {
    auto pw_of = [&](int l) {
      LstmBwdPw q = {};
      q.gates = s.gates[l] + (size_t)u * B * H4; q.c_new = s.c[l] + (size_t)(u + 1) * B * H; q.c_prev = s.c[l] + (size_t)u * B * H;
      q.dh_carry = w.dh_carry[l]; q.dc_carry = w.dc_carry[l]; q.dzT = w.dzT[l]; q.tlen = target_len; q.u = u;
      q.keep = d.dropout_keep; q.seed = d.seed; q.layer = l;
      return q;
    };
    AttnBwdArgs a = {};
    if (fuse) a.pw = pw_of(NL - 1);
    a.R = B; a.Tm = Tm; a.E = E; a.H = H; a.A = A; a.V = V; a.F = F; a.ksz = ksz; a.U = U; a.u = u;
    a.dlogits = dlogits + (size_t)u * V; a.dl_row_stride = (long)U * V;
    a.outin = s.outin + (size_t)u * (H + E); a.outin_row_stride = (long)U * (H + E);
    a.alpha = s.align + (size_t)(u + 1) * B * Tm; a.alpha_prev = s.align + (size_t)u * B * Tm;
    a.q = s.q + (size_t)u * B * A; a.cf = F ? s.cf + (size_t)u * B * Tm * F : nullptr;
    a.Wq = p->query_kernel; a.Wc = p->conv_kernel; a.Wd = p->conv_dense_kernel; a.v = p->attention_v; a.Wo = p->out_kernel;
    a.keys = s.keys; a.values = s.values; a.mem_len = mem_len;
    a.dctx_carry = w.dctx_carry; a.dalign_carry = w.dalign_carry; a.dh_above = w.dh_above;
    a.dq_save = w.dq + (size_t)u * B * A; a.dkeys = w.dkeys; a.dvalues = w.dvalues;
    a.dv_part = w.dv_part; a.dWd_part = w.dWd_part; a.dWc_part = w.dWc_part; a.tlen = target_len;
    {
      a.ablate = getenv("NABU_ATTN_ABLATE") ? atoi(getenv("NABU_ATTN_ABLATE")) : 0;
      a.prob = d.probability_fn; a.asum = s.asum + (size_t)u * B;
      KernelScope ks("dec_attn_bwd_step", stream);
      if (int e = attn_bwd_launch(a, B, stream)) return e;
    }
    for (int l = NL - 1; l >= 0; --l) {
      if (!fuse) {
        KernelScope ks("dec_lstm_bwd_pointwise", stream);
        NABU_CHECK_CUDA(chain_launch(dec_lstm_bwd_pointwise_kernel, dim3(ceil_div(B * H, 256)), dim3(256), 0, stream,
                                     pw_of(l), (const float*)w.dh_above, B, H));
      }
      MatmulTArgs m = {};
      m.xT = w.dzT[l]; m.K = H4; m.R = B; m.W = p->cell_kernel[l]; m.ldw = H4; m.Wr = s.wb[l];
      if (l > 0) { m.row0 = 0; m.N = 2 * H; m.N0 = H; m.out0 = w.dh_above; m.ld0 = H; m.out1 = w.dh_carry[l]; m.ld1 = H; }
      else { m.row0 = V; m.N = E + H; m.N0 = E; m.out0 = w.dctx_carry; m.ld0 = E; m.out1 = w.dh_carry[0]; m.ld1 = H; }
      if (fuse && l > 0) m.pw = pw_of(l - 1);               // d(output of layer l-1) goes through its cell backward at once
      KernelScope ks("dec_matmul_t", stream);
      NABU_CHECK_CUDA(chain_launch(dec_matmul_t_kernel, dim3(ceil_div(m.N, 8), ceil_div(B, ROWS)), dim3(MT_THREADS), smem_mm, stream, m));
    }
  }
  // ---- batched weight gradients over all (step, row) pairs --------------------------------------
  const int UB = U * B;
  const bool drop = d.dropout_keep > 0.f && d.dropout_keep < 1.f;
  for (int l = 0; l < NL; ++l) {
    float* dz = s.gates[l];                                   // [U][B][4H], now dz
    float* dK = g->cell_kernel[l];
    if (l == 0) {
      {
        KernelScope ks("embedding_grad", stream);
        embedding_grad_kernel<<<dim3(ceil_div(H4, 256), V), 256, 0, stream>>>(s.ids_in, dz, UB, H4, dK);
        NABU_CHECK_LAUNCH();
      }
      // context rows: input of step u is ctx slot u
      if (int e = gemm(GEMM_TN, E, H4, UB, 1.f, s.ctx, E, dz, H4, 0.f, dK + (size_t)V * H4, H4, nullptr, nullptr, w.gemm,
                        w.gemm_bytes, stream)) return e;
      if (int e = gemm(GEMM_TN, H, H4, UB, 1.f, s.h[0], H, dz, H4, 0.f, dK + (size_t)(V + E) * H4, H4, nullptr, nullptr,
                        w.gemm, w.gemm_bytes, stream)) return e;
    } else {
      // input rows: h of the layer below AFTER step u = slot u+1 ; recurrent rows: own h slot u
      if (int e = gemm(GEMM_TN, H, H4, UB, 1.f, (drop ? s.out[l - 1] : s.h[l - 1]) + (size_t)B * H, H, dz, H4, 0.f, dK, H4, nullptr, nullptr,
                        w.gemm, w.gemm_bytes, stream)) return e;
      if (int e = gemm(GEMM_TN, H, H4, UB, 1.f, s.h[l], H, dz, H4, 0.f, dK + (size_t)H * H4, H4, nullptr, nullptr,
                        w.gemm, w.gemm_bytes, stream)) return e;
    }
    if (int e = colsum(dz, UB, H4, H4, g->cell_bias[l], stream)) return e;
  }
  // output projection: rows ordered (b, u) on both sides
  if (int e = gemm(GEMM_TN, H + E, V, B * U, 1.f, s.outin, H + E, dlogits, V, 0.f, g->out_kernel, V, nullptr, nullptr,
                    w.gemm, w.gemm_bytes, stream)) return e;
  if (int e = colsum(dlogits, B * U, V, V, g->out_bias, stream)) return e;
  // query layer: h_top after step u (slot u+1) against dq[u]
  if (int e = gemm(GEMM_TN, H, A, UB, 1.f, (drop ? s.out[NL - 1] : s.h[NL - 1]) + (size_t)B * H, H, w.dq, A, 0.f, g->query_kernel, A, nullptr,
                    nullptr, w.gemm, w.gemm_bytes, stream)) return e;
  // memory layer and the memory itself
  if (int e = gemm(GEMM_TN, E, A, B * Tm, 1.f, s.values, E, w.dkeys, A, 0.f, g->memory_kernel, A, nullptr, nullptr,
                    w.gemm, w.gemm_bytes, stream)) return e;
  if (int e = gemm(GEMM_NT, B * Tm, E, A, 1.f, w.dkeys, A, p->memory_kernel, A, 1.f, w.dvalues, E, nullptr, nullptr,
                    nullptr, 0, stream)) return e;
  if (dmemory) {
    const long n = (long)B * Tm * E;
    KernelScope ks("mask_memory", stream);
    mask_memory_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w.dvalues, mem_len, Tm, E, dmemory, n);
    NABU_CHECK_LAUNCH();
  }
  {
    KernelScope ks("reduce_rows", stream);
    reduce_rows_kernel<<<ceil_div(A, 256), 256, 0, stream>>>(w.dv_part, B, A, g->attention_v);
    NABU_CHECK_LAUNCH();
  }
  if (F > 0) {
    KernelScope ks("reduce_rows", stream);
    reduce_rows_kernel<<<ceil_div(F * A, 256), 256, 0, stream>>>(w.dWd_part, B, (long)F * A, g->conv_dense_kernel);
    NABU_CHECK_LAUNCH();
    reduce_rows_kernel<<<ceil_div(ksz * F, 256), 256, 0, stream>>>(w.dWc_part, B, (long)ksz * F, g->conv_kernel);
    NABU_CHECK_LAUNCH();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Row a8 on its own (SURVEY.md section 8b export list): the attention mechanism as three entry points, for a caller that
// steps the mechanism itself the way components/beam_search_decoder.py:176 steps the AttentionWrapper's cell.
// The kernels are the fused ones above; the output projection they also contain runs on scratch / zero operands here.
// ---------------------------------------------------------------------------------------------------------------------
namespace nabu {
namespace dec {
struct AttnWs {
  float *dlogits, *dv_part, *dWd_part, *dWc_part;     // zeroed per call
  float *logits, *ctxT, *outin, *dq;
  int* tlen;
  float* gemm; size_t gemm_bytes; size_t zero_bytes; size_t total;
};
AttnWs carve_attn(void* base, const nabu_speller_desc_t& d, int R) {
  AttnWs w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloats) { float* q = (float*)(p + off); off += align_up(nfloats * sizeof(float), 256); return q; };
  const size_t F = d.attention == 1 ? d.numfilt : 0, ksz = d.attention == 1 ? d.filtersize : 1;
  w.dlogits = take((size_t)R * d.V);
  w.dv_part = take((size_t)R * d.A);
  w.dWd_part = take((size_t)R * (F ? F : 1) * d.A);
  w.dWc_part = take((size_t)R * ksz * (F ? F : 1));
  w.outin = take((size_t)R * (d.H + d.E));
  w.zero_bytes = off;
  w.logits = take((size_t)R * d.V);
  w.ctxT = take((size_t)R * d.E);
  w.dq = take((size_t)R * d.A);
  w.tlen = (int*)take((size_t)R);
  w.gemm = (float*)(p + off); w.gemm_bytes = sgemm_workspace_bytes(); off += w.gemm_bytes;
  w.total = off;
  return w;
}
}  // namespace dec
}  // namespace nabu

extern "C" size_t nabu_attn_workspace_bytes(const nabu_speller_desc_t* d, int R) {
  if (check_desc(*d) || R <= 0) return 0;
  return carve_attn(nullptr, *d, R).total;
}

extern "C" int nabu_attn_keys(const nabu_speller_desc_t* d, const nabu_speller_params_t* p, const float* memory,
                              const int* mem_len, float* values, float* keys, void* stream) {
  if (int e = check_desc(*d)) return e;
  return prepare_memory(*d, *p, memory, mem_len, values, keys, (cudaStream_t)stream);
}

extern "C" int nabu_attn_step_fwd(const nabu_speller_desc_t* dp, const nabu_speller_params_t* p, const float* query, int R,
                                  int rows_per_mem, const float* keys, const float* values, const int* mem_len,
                                  const float* align_prev, float* align_new, float* context, float* q_save, float* cf_save,
                                  float* asum_save, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const nabu_speller_desc_t& d = *dp;
  if (int e = check_desc(d)) return e;
  NABU_REQUIRE(R > 0 && rows_per_mem > 0 && R == d.B * rows_per_mem, "attn_step_fwd: R=%d must be B=%d x rows_per_mem=%d", R, d.B,
               rows_per_mem);
  AttnWs w = carve_attn(workspace, d, R);
  NABU_REQUIRE(ws_bytes >= w.total, "attn_step_fwd: workspace %zu < %zu bytes", ws_bytes, w.total);
  AttnStepArgs a = {};
  a.R = R; a.Tm = d.Tm; a.E = d.E; a.H = d.H; a.A = d.A; a.V = d.V;
  a.F = d.attention == 1 ? d.numfilt : 0; a.ksz = d.attention == 1 ? d.filtersize : 1;
  a.rows_per_mem = rows_per_mem;
  a.h_top = query;
  a.Wq = p->query_kernel; a.Wc = p->conv_kernel; a.Wd = p->conv_dense_kernel; a.v = p->attention_v;
  a.Wo = p->out_kernel; a.bo = p->out_bias;
  a.keys = keys; a.values = values; a.mem_len = mem_len;
  a.align_prev = align_prev; a.ctx_prev = context;
  a.align_new = align_new; a.ctx_new = context; a.ctxT_new = w.ctxT;
  a.logits = w.logits; a.logits_row_stride = d.V; a.temperature = 1.f;
  a.q_save = q_save; a.cf_save = a.F ? cf_save : nullptr; a.outin_save = nullptr; a.outin_row_stride = 0;
  a.tlen = nullptr; a.u = 0; a.done = nullptr;
  a.prob = d.probability_fn; a.asum_save = asum_save;
  a.win_left = d.attention == 2 ? d.numfilt : -1; a.win_right = d.filtersize;
  NABU_REQUIRE(attn_step_smem(d.Tm, d.E, d.H, d.A, a.F, a.ksz, attn_step_cluster(R, a.rows_per_mem)) <= (size_t)max_smem_optin(), "attn_step_fwd: memory too long for the attention step kernel (Tm=%d)", d.Tm);
  KernelScope ks("dec_attn_step", stream);
  NABU_CHECK_CUDA(attn_step_launch(a, R, nullptr, stream));
  return 0;
}

extern "C" int nabu_attn_step_bwd(const nabu_speller_desc_t* dp, const nabu_speller_params_t* p, const float* query, int R,
                                  const float* keys, const float* values, const int* mem_len, const float* align_prev,
                                  const float* align_new, const float* q_save, const float* cf_save, const float* asum_save,
                                  const float* dalign_new, const float* dcontext, float* dquery, float* dalign_prev,
                                  float* dkeys, float* dvalues, const nabu_speller_params_t* g, void* workspace,
                                  size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const nabu_speller_desc_t& d = *dp;
  if (int e = check_desc(d)) return e;
  NABU_REQUIRE(R == d.B, "attn_step_bwd: one decoder row per memory row (R=%d, B=%d)", R, d.B);
  AttnWs w = carve_attn(workspace, d, R);
  NABU_REQUIRE(ws_bytes >= w.total, "attn_step_bwd: workspace %zu < %zu bytes", ws_bytes, w.total);
  const int Tm = d.Tm, E = d.E, H = d.H, A = d.A, V = d.V;
  const int F = d.attention == 1 ? d.numfilt : 0, ksz = d.attention == 1 ? d.filtersize : 1;
  NABU_CHECK_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, stream));
  NABU_CHECK_CUDA(cudaMemsetAsync(w.tlen, 1, (size_t)R * sizeof(int), stream));       // 0x01010101 > u = 0: every row active
  // d(alignments) arriving from the consumer of align_new (the next step's location features); the kernel adds its own
  // contribution through the context and leaves d(align_prev) in the same buffer
  if (dalign_new) NABU_CHECK_CUDA(cudaMemcpyAsync(dalign_prev, dalign_new, (size_t)R * Tm * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  else NABU_CHECK_CUDA(cudaMemsetAsync(dalign_prev, 0, (size_t)R * Tm * sizeof(float), stream));
  AttnBwdArgs a = {};
  a.R = R; a.Tm = Tm; a.E = E; a.H = H; a.A = A; a.V = V; a.F = F; a.ksz = ksz; a.U = 1; a.u = 0;
  a.dlogits = w.dlogits; a.dl_row_stride = V;                   // zeros: the projection's backward contributes nothing
  a.outin = w.outin; a.outin_row_stride = H + E;
  a.alpha = align_new; a.alpha_prev = align_prev;
  a.q = q_save; a.cf = F ? cf_save : nullptr;
  a.Wq = p->query_kernel; a.Wc = p->conv_kernel; a.Wd = p->conv_dense_kernel; a.v = p->attention_v; a.Wo = p->out_kernel;
  a.keys = keys; a.values = values; a.mem_len = mem_len;
  a.dctx_carry = dcontext; a.dalign_carry = dalign_prev; a.dh_above = dquery;
  a.dq_save = w.dq; a.dkeys = dkeys; a.dvalues = dvalues;
  a.dv_part = w.dv_part; a.dWd_part = w.dWd_part; a.dWc_part = w.dWc_part;
  a.tlen = w.tlen; a.prob = d.probability_fn; a.asum = asum_save;
  {
    KernelScope ks("dec_attn_bwd_step", stream);
    if (int e = attn_bwd_launch(a, R, stream)) return e;
  }
  // this step's parameter gradients (overwritten, not accumulated): query layer, attention vector, location layers
  if (int e = gemm(GEMM_TN, H, A, R, 1.f, query, H, w.dq, A, 0.f, g->query_kernel, A, nullptr, nullptr, w.gemm, w.gemm_bytes, stream))
    return e;
  {
    KernelScope ks("reduce_rows", stream);
    reduce_rows_kernel<<<ceil_div(A, 256), 256, 0, stream>>>(w.dv_part, R, A, g->attention_v);
    NABU_CHECK_LAUNCH();
  }
  if (F > 0) {
    KernelScope ks("reduce_rows", stream);
    reduce_rows_kernel<<<ceil_div(F * A, 256), 256, 0, stream>>>(w.dWd_part, R, (long)F * A, g->conv_dense_kernel);
    NABU_CHECK_LAUNCH();
    reduce_rows_kernel<<<ceil_div(ksz * F, 256), 256, 0, stream>>>(w.dWc_part, R, (long)ksz * F, g->conv_kernel);
    NABU_CHECK_LAUNCH();
  }
  return 0;
}
