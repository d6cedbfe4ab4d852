// Device kernels of the attention decoder ("Speller"), shared by
// the teacher-forced training path (speller.cu) and the LAS beam search (las_beam.cu).
//
// One decoder step = TF's AttentionProjectionWrapper(AttentionWrapper(MultiRNNCell(LSTMCell x N)))
// (reference: models/ed_decoders/speller.py:29-69, components/rnn_cell.py:145-155,
// components/attention.py:142-240; TF semantics in SURVEY appendix B3-B5):
//   dec_lstm_step   x N  : z = onehot-row gather + [input, h_prev].K + b -> i,j,f,o -> c', h'
//   dec_attn_step        : q = Wq.h_top ; location features conv1d(alpha_prev)->dense ; score
//                          e = v.tanh(q + keys + f) ; mask ; softmax ; context = alpha.values ;
//                          logits = [h_top, context].Wo + bo      -- one CTA per decoder row
// The decoder state is kept in two layouts: row-major [R][*] for the pointwise consumers and
// transposed [*][R] for the "skinny" matmuls (R = batch rows is small, so threads run along R and
// the weight slice of a CTA sits in shared memory).
#pragma once
#include "common.cuh"
#include <stdint.h>
#include <math_constants.h>
#include <cooperative_groups.h>
#include <stdlib.h>
#include <utility>

namespace nabu {
namespace dec {

// Counter-based generator of the decoder's stochastic parts (DropoutWrapper masks, scheduled sampling): three rounds of
// the lowbias32 integer finaliser over (seed, a, b, c).  Stateless, so the backward regenerates the forward's masks; the
// oracle restates it (oracle/nabu_oracle.py: rng_u32).  Keys: dropout a = layer*65536 + step, b = row, c = unit;
// sampling a = 0x40000000 + step, b = row, c = 0 (Bernoulli) / 1 (categorical).
__host__ __device__ inline uint32_t dec_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__host__ __device__ inline uint32_t dec_rng_u32(uint32_t seed, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t x = dec_mix((seed ^ 0x9E3779B9u) + a);
  x = dec_mix(x + b * 0x85EBCA6Bu);
  x = dec_mix(x + c * 0xC2B2AE35u);
  return x;
}
__host__ __device__ inline float dec_uniform(uint32_t seed, uint32_t a, uint32_t b, uint32_t c) {
  return (float)(dec_rng_u32(seed, a, b, c) >> 8) * (1.f / 16777216.f);       // [0, 1), 24 bits
}

constexpr int ROWS = 64;       // decoder rows per CTA tile in the skinny matmuls
constexpr int SK_THREADS = 512;
constexpr int SK_KSPLIT = SK_THREADS / ROWS;   // k-split of the LSTM step (8 partial sums per output)
constexpr int MT_THREADS = 256;                // dec_matmul_t: measured slower with 512 threads (8.0 vs 5.6 ms per LAS step)
constexpr int MT_KSPLIT = MT_THREADS / ROWS;

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch of the decoder's chain of small kernels (round 2).  A decoder step is a chain of
// kernels of 10-80 us, each of which begins by staging weights (or, for the attention step, by the location features,
// which depend on the PREVIOUS step's alignments only).  Launched with the programmatic-stream-serialization attribute,
// kernel N+1 starts as soon as every CTA of kernel N has passed chain_wait(), runs its prologue next to N's body and
// blocks in its own chain_wait() until N has completed and flushed.  Rule for every kernel of the chain: before
// chain_wait() it reads only what was complete two kernels ago (weights, saved forward tensors, the previous step's
// alignments - kernel N has itself waited for N-1) and writes nothing but its own shared memory.  NABU_PDL=0 launches
// the same kernels without the attribute (chain_wait() is then a no-op).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void chain_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
inline bool chain_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("NABU_PDL") ? atoi(getenv("NABU_PDL")) : 1;
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t chain_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = chain_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------------
// Bulk-copy staging of the transposed activations of the skinny matmuls (round 2).  Warp 0 streams chunks of XC
// consecutive k-rows of a [K][R] operand into a double buffer with cp.async.bulk (completion on an mbarrier) while the
// block multiplies the previous chunk, instead of every thread chasing its own chain of L2 round trips (K / KSPLIT
// dependent-latency loads per thread).  Used when the launch has one row tile (R <= ROWS, training): a chunk is then one
// contiguous copy.
// ------------------------------------------------------------------------------------------------
constexpr int XC = 128;          // k-rows per chunk
__device__ __forceinline__ uint32_t dk_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dk_bar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dk_s32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void dk_bar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "DK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DK_DONE;\n"
      "bra DK_WAIT;\n"
      "DK_DONE:\n"
      "}" ::"r"(dk_s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void dk_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer's last generic reads are ordered before the copy
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dk_s32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dk_s32(dst)), "l"(src), "r"(bytes), "r"(dk_s32(bar)) : "memory");
}
// acc[0..7] += sum_k x[k][r0 + rl] * Ws[(wrow0 + k)][0..7] over the k of this thread's split, operand [K][R] streamed in
// chunks of XC k-rows, one copy per chunk (the block's rows are all R rows, so a chunk is contiguous).  Called by all
// threads of the block; `issued` / `consumed` carry the double buffer's position from one operand to the next.
template <int NTHREADS>
__device__ __forceinline__ void dk_stream_fma(const float* __restrict__ xT, int K, int R, int r0, const float* Ws, int wrow0,
                                              float* xs, uint64_t* bar, unsigned& issued, unsigned& consumed, int rl, int ks,
                                              bool row_ok, float (&acc)[8]) {
  constexpr int KSPLIT = NTHREADS / ROWS;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nch = (K + XC - 1) / XC;
  if (nch == 0) return;
  const int rs = R;                                      // row stride of a staged chunk (one row tile: the chunk is contiguous)
  (void)r0;
  auto issue = [&](int ch) {                             // warp 0, all lanes
    const int kk = min(XC, K - ch * XC);
    const unsigned b = issued & 1;
    float* dst = xs + (size_t)b * XC * ROWS;
    if (lane == 0) dk_bulk_load(dst, xT + (size_t)ch * XC * R, (unsigned)(kk * R * sizeof(float)), &bar[b]);
    ++issued;
  };
  // (the caller guarantees that both buffers are free on entry)
  if (tid < 32) { issue(0); if (nch > 1) issue(1); }
  for (int ch = 0; ch < nch; ++ch) {
    const unsigned b = consumed & 1;
    dk_bar_wait(&bar[b], (consumed >> 1) & 1);
    ++consumed;
    const float* xc = xs + (size_t)b * XC * ROWS;
    const int kk = min(XC, K - ch * XC);
    if (row_ok) {
#pragma unroll 8
      for (int k = ks; k < kk; k += KSPLIT) {
        const float x = xc[k * rs + rl];
        const float* wp = Ws + (size_t)(wrow0 + ch * XC + k) * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wp);
        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
        acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
        acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
        acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
        acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
      }
    }
    __syncthreads();                                     // the buffer is free again
    if (tid < 32 && ch + 2 < nch) issue(ch + 2);
  }
}
// Several row tiles (beam search, R = B * W rows): a chunk is XC separate 256-byte segments; staging them with one bulk
// copy each was measured and is SLOWER than the per-thread loads (LAS beam-16 decode 24.9 -> 38.2 ms), so those launches
// keep the register path.
__device__ __forceinline__ bool dk_bulk_ok(const void* p0, const void* p1, int R) {
  return gridDim.y == 1 && (R & 3) == 0 && ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1)) & 15) == 0;
}
inline size_t dk_stage_bytes() { return (size_t)2 * XC * ROWS * sizeof(float); }

// ------------------------------------------------------------------------------------------------
// Weight slices of the skinny matmuls laid out the way their CTAs stage them (once per forward / backward / beam search
// instead of a strided gather in every one of the U steps):
//   forward:  Wr[slice][k][c] = W[wrow(k)][(c >> 1) * H + 2 * slice + (c & 1)]   slice < H/2, k over both input segments
//   backward: Wb[slice][k][c] = W[row0 + 8 * slice + c][k]                       slice < N/8, k < K
// ------------------------------------------------------------------------------------------------
__global__ void dec_relayout_fwd_kernel(const float* W, float* Wr, int H, int K0, int w0, int K1, int w1) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  const int Ktot = K0 + K1;
  if (i >= (long)(H / 2) * Ktot * 8) return;
  const int c = (int)(i & 7);
  const long t = i >> 3;
  const int k = (int)(t % Ktot), slice = (int)(t / Ktot);
  const int wrow = k < K0 ? w0 + k : w1 + (k - K0);
  Wr[i] = W[(size_t)wrow * 4 * H + (c >> 1) * H + 2 * slice + (c & 1)];
}
__global__ void dec_relayout_bwd_kernel(const float* W, float* Wb, int K, int ldw, int row0, int N) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  if (i >= (long)N * K) return;
  const int k = (int)(i % K), n = (int)(i / K);          // k fastest: coalesced reads of W's rows
  Wb[((size_t)(n >> 3) * K + k) * 8 + (n & 7)] = W[(size_t)(row0 + n) * ldw + k];
}

// ------------------------------------------------------------------------------------------------
// LSTM cell step.  grid = (H/2, ceil(R/ROWS)); CTA (slice, tile) owns hidden units 2*slice,
// 2*slice+1 (8 gate columns) for ROWS rows; SK_KSPLIT-way k-split over the SK_THREADS threads.
// ------------------------------------------------------------------------------------------------
struct LstmStepArgs {
  const float* inT0; int K0; int w0;      // transposed input segment 0 [K0][R], first weight row w0
  const float* inT1; int K1; int w1;      // transposed input segment 1 (this layer's h_prev) [K1][R]
  const int* ids;                         // optional one-hot ids [R] (weight rows 0..V-1), or nullptr
  const float* W; const float* bias;      // [(rows), 4H], [4H]
  const float* Wr;                        // optional [H/2][K0 + K1][8]: the CTAs' weight slices laid out contiguously (relayout_fwd)
  int H, R;
  const float* c_prev;                    // [R][H]
  const float* h_prev;                    // [R][H] row-major (copy-through for finished rows)
  float* c_new; float* h_new; float* hT_new;   // [R][H], [R][H], [H][R]
  float* gates_out;                       // [R][4H] activated i,g,f,o or nullptr
  const int* tlen; int u;                 // row r is active iff tlen == nullptr || u < tlen[r]
  const int* done;                        // optional device flag: non-zero -> the whole launch is a no-op
  // DropoutWrapper(output_keep_prob = keep) (speller.py:37-41): the cell OUTPUT out = h * mask / keep feeds the next
  // layer and the attention query, the state keeps h.  out_new == nullptr: no dropout (consumers read h).
  float* out_new; float* outT_new;        // [R][H], [H][R]
  float keep; unsigned seed; int layer;
};

__global__ void __launch_bounds__(SK_THREADS) dec_lstm_step_kernel(const LstmStepArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int Ktot = a.K0 + a.K1;
  float* Ws = sm;                          // [Ktot][8]
  float* red = sm + (size_t)Ktot * 8;      // [SK_KSPLIT][ROWS][8]
  const int tid = threadIdx.x;
  const int H = a.H, H4 = 4 * a.H, R = a.R;
  const int j0 = blockIdx.x * 2;
  const int r0 = blockIdx.y * ROWS;
  __shared__ uint64_t bars[3];
  if (tid == 0) { dk_bar_init(&bars[0]); dk_bar_init(&bars[1]); dk_bar_init(&bars[2]); }
  if (a.Wr) {                              // the slice is one contiguous piece: one bulk copy
    if (tid == 0) dk_bulk_load(Ws, a.Wr + (size_t)blockIdx.x * Ktot * 8, (unsigned)(Ktot * 8 * sizeof(float)), &bars[2]);
  } else {
    for (int i = tid; i < Ktot * 8; i += SK_THREADS) {
      const int c = i & 7, k = i >> 3;
      const int wrow = k < a.K0 ? a.w0 + k : a.w1 + (k - a.K0);
      Ws[i] = a.W[(size_t)wrow * H4 + (c >> 1) * H + j0 + (c & 1)];
    }
  }
  __syncthreads();                         // (the barriers are initialised for every thread)
  chain_wait();                            // the weight slice was staged next to the previous kernel
  if (a.Wr) dk_bar_wait(&bars[2], 0);
  if (a.done && *a.done) return;
  __syncthreads();
  const int rl = tid % ROWS, ks = tid / ROWS;
  const int r = r0 + rl;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (dk_bulk_ok(a.inT0, a.inT1, R)) {
    float* xs = red + SK_KSPLIT * ROWS * 8;            // [2][XC][R]
    unsigned issued = 0, consumed = 0;
    dk_stream_fma<SK_THREADS>(a.inT0, a.K0, R, r0, Ws, 0, xs, bars, issued, consumed, rl, ks, r < R, acc);
    dk_stream_fma<SK_THREADS>(a.inT1, a.K1, R, r0, Ws, a.K0, xs, bars, issued, consumed, rl, ks, r < R, acc);
  } else if (r < R) {
#pragma unroll 16
    for (int k = ks; k < a.K0; k += SK_KSPLIT) {
      const float x = __ldcg(a.inT0 + (size_t)k * R + r);
      const float4 w0 = *reinterpret_cast<const float4*>(Ws + k * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(Ws + k * 8 + 4);
      acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
      acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
      acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
      acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
    }
#pragma unroll 16
    for (int k = ks; k < a.K1; k += SK_KSPLIT) {
      const float x = __ldcg(a.inT1 + (size_t)k * R + r);
      const float* wp = Ws + (a.K0 + k) * 8;
      const float4 w0 = *reinterpret_cast<const float4*>(wp);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
      acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
      acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
      acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
      acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) red[(ks * ROWS + rl) * 8 + c] = acc[c];
  __syncthreads();
  if (tid < 2 * ROWS) {
    const int rl2 = tid % ROWS, jl = tid / ROWS;
    const int r2 = r0 + rl2, j = j0 + jl;
    if (r2 < R) {
      float z[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float s = a.bias[g * H + j];
#pragma unroll
        for (int k2 = 0; k2 < SK_KSPLIT; ++k2) s += red[(k2 * ROWS + rl2) * 8 + g * 2 + jl];
        if (a.ids) s += a.W[(size_t)a.ids[r2] * H4 + g * H + j];
        z[g] = s;
      }
      const bool active = (a.tlen == nullptr) || (a.u < a.tlen[r2]);
      const float cp = a.c_prev[(size_t)r2 * H + j];
      const float ig = sigmoid_acc(z[0]), gg = tanhf(z[1]), fg = sigmoid_acc(z[2] + 1.0f), og = sigmoid_acc(z[3]);
      float cn = cp * fg + ig * gg;
      float hn = tanhf(cn) * og;
      if (!active) { cn = cp; hn = a.h_prev[(size_t)r2 * H + j]; }
      a.c_new[(size_t)r2 * H + j] = cn;
      a.h_new[(size_t)r2 * H + j] = hn;
      a.hT_new[(size_t)j * R + r2] = hn;
      if (a.out_new) {
        float o = hn;
        if (active) o = dec_uniform(a.seed, (uint32_t)(a.layer * 65536 + a.u), (uint32_t)r2, (uint32_t)j) < a.keep ? hn / a.keep : 0.f;
        a.out_new[(size_t)r2 * H + j] = o;
        a.outT_new[(size_t)j * R + r2] = o;
      }
      if (a.gates_out) {
        float* gp = a.gates_out + (size_t)r2 * H4 + j;
        gp[0] = ig; gp[H] = gg; gp[2 * H] = fg; gp[3 * H] = og;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Attention + output projection step: one CTA (256 threads) per decoder row.
// ------------------------------------------------------------------------------------------------
struct AttnStepArgs {
  int R, Tm, E, H, A, V, F, ksz;          // F = 0 -> vanilla Bahdanau (no location features)
  int rows_per_mem;                       // decoder rows sharing one memory row (beam width; 1 in training)
  const float* h_top;                     // [R][H] query (new top-layer output)
  const float* Wq; const float* Wc; const float* Wd; const float* v;
  const float* Wo; const float* bo;
  const float* keys; const float* values; // [Rm][Tm][A], [Rm][Tm][E]
  const int* mem_len;                     // [Rm]
  const float* align_prev; const float* ctx_prev;    // [R][Tm], [R][E]
  float* align_new; float* ctx_new; float* ctxT_new; // [R][Tm], [R][E], [E][R]
  float* logits; long logits_row_stride;  // logits + r*stride : V values
  float temperature;                      // logits divided by this (1 in training)
  float* q_save; float* cf_save; float* outin_save; long outin_row_stride;   // optional
  const int* tlen; int u;
  const int* done;
  int prob;                               // probability_fn: 0 softmax, 1 normalized_sigmoid, 2 sigmoid (attention.py:9-13)
  float* asum_save;                       // optional [R]: sum of sigmoids (prob == 1), for the backward
  int win_left, win_right;                // WindowedAttention (attention.py:294-396): window widths, win_left < 0 = off
};

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = is_max ? -CUDART_INF_F : 0.f;
  for (int w = 0; w < nw; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// CS = CTAs per decoder row (a thread-block cluster; round 2, training batches whose rows alone cannot fill the 148 SMs):
// the memory positions are cut into CS ranges; a CTA computes the location features, scores, probabilities and the
// partial context of its range; the softmax statistics (maximum and sum of its range, combined as in an online
// softmax) and the partial contexts are exchanged through distributed shared memory (two cluster barriers); the query
// projection is computed by every CTA, the output projection by CTA 0.  CS = 1 is the round-1 kernel, bit for bit (the
// beam search, whose ids are held bit-exact, always runs it).
template <int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(512) dec_attn_step_kernel(const AttnStepArgs a) {
  constexpr int NT = 512, NW = NT / 32;                // 512 threads per CTA
  extern __shared__ __align__(16) float sm[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cr = CS > 1 ? (int)cluster.block_rank() : 0;
  const int r = blockIdx.x / CS;
  const int mrow = r / a.rows_per_mem;
  const int Tm = a.Tm, E = a.E, H = a.H, A = a.A, V = a.V, F = a.F, ksz = a.ksz;
  const int padl = (ksz - 1) / 2;
  auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };   // every piece starts on a 16-byte boundary (attn_step_smem)
  float* query = sm;                       // [H]
  float* q = query + r4(H);                // [A]
  float* ap = q + r4(A);                   // [Tm + ksz + 4] zero padded alpha_prev
  float* e = ap + r4(Tm + ksz + 4);        // [Tm]
  float* ctx = e + r4(Tm);                 // [E]
  float* red = ctx + r4(E);                // [32]
  float* part = red + 32;                  // [NW][32] projection partials
  float* cf = part + NW * 32;              // [Tm][F]
  float* wd = cf + r4((size_t)Tm * F);     // [F][A]
  float* wc = wd + r4((size_t)F * A);      // [ksz][F]
  float* vs = wc + r4((size_t)ksz * F);    // [A] attention vector
  float* cpart = vs + r4(A);               // [NT * 4] context partials of the t-splits
  float* xstat = cpart + NT * 4;           // [CS][4] (maximum, sum) of every CTA's range          (CS > 1)
  float* xctx = xstat + 4 * 4;             // [CS][E] partial contexts, slot = source CTA           (CS > 1)
  // this CTA's memory positions [tb, te)
  const int per = (Tm + CS - 1) / CS, tb = min(Tm, cr * per), te = min(Tm, tb + per);

  const bool active = (a.tlen == nullptr) || (a.u < a.tlen[r]);
  if (active) {
    // prologue, next to the LSTM step that produces this step's query (see chain_wait): weights, the previous
    // alignments (complete since the previous step) and phase 1, the location features
    //   cf[t][f] = sum_k alpha_prev[t + k - padl] * Wc[k][f]
    for (int i = tid; i < Tm + ksz + 4; i += NT) {
      const int t = i - padl;
      ap[i] = ((F > 0 || a.win_left >= 0) && t >= 0 && t < Tm) ? __ldcg(a.align_prev + (size_t)r * Tm + t) : 0.f;
    }
    for (int i = tid; i < F * A; i += NT) wd[i] = a.Wd[i];
    for (int i = tid; i < ksz * F; i += NT) wc[i] = a.Wc[i];
    for (int i = tid; i < A; i += NT) vs[i] = a.v[i];
    __syncthreads();
    // a thread takes 4 consecutive positions of one filter: per tap one weight and one new alignment are read for 4 FMAs
    // (a sliding window in registers) instead of two reads per FMA
    for (int i = tid; i < ((te - tb + 3) / 4) * F; i += NT) {
      const int f = i % F, t0 = tb + (i / F) * 4;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      float a0 = ap[t0], a1 = ap[t0 + 1], a2 = ap[t0 + 2];
#pragma unroll 4
      for (int k = 0; k < ksz; ++k) {
        const float a3 = ap[t0 + k + 3], w = wc[k * F + f];
        s0 = fmaf(a0, w, s0); s1 = fmaf(a1, w, s1); s2 = fmaf(a2, w, s2); s3 = fmaf(a3, w, s3);
        a0 = a1; a1 = a2; a2 = a3;
      }
      cf[t0 * F + f] = s0;
      if (t0 + 1 < te) cf[(t0 + 1) * F + f] = s1;
      if (t0 + 2 < te) cf[(t0 + 2) * F + f] = s2;
      if (t0 + 3 < te) cf[(t0 + 3) * F + f] = s3;
    }
    // every CTA of the cluster is running before the first store into a peer's shared memory: arrive here, wait in phase 3
    if (CS > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  }
  chain_wait();
  if (a.done && *a.done) return;
  if (!active && cr != 0) return;
  if (!active) {                           // finished row: copy the state through, emit zeros
    for (int t = tid; t < Tm; t += NT) a.align_new[(size_t)r * Tm + t] = a.align_prev[(size_t)r * Tm + t];
    for (int i = tid; i < E; i += NT) {
      const float c = a.ctx_prev[(size_t)r * E + i];
      a.ctx_new[(size_t)r * E + i] = c;
      a.ctxT_new[(size_t)i * a.R + r] = c;
    }
    for (int k = tid; k < V; k += NT) a.logits[r * a.logits_row_stride + k] = 0.f;
    if (a.q_save) for (int i = tid; i < A; i += NT) a.q_save[(size_t)r * A + i] = 0.f;
    if (a.cf_save) for (int i = tid; i < Tm * F; i += NT) a.cf_save[(size_t)r * Tm * F + i] = 0.f;
    if (a.outin_save) for (int i = tid; i < H + E; i += NT) a.outin_save[r * a.outin_row_stride + i] = 0.f;
    return;
  }
  const int len = min(a.mem_len[mrow], Tm);

  __syncthreads();                         // the location features were written by other threads (4 positions each)
  for (int i = tid; i < H; i += NT) query[i] = a.h_top[(size_t)r * H + i];
  if (a.cf_save) for (int i = tb * F + tid; i < te * F; i += NT) a.cf_save[(size_t)r * Tm * F + i] = cf[i];
  __syncthreads();

  // phase 0: q = query . Wq
  for (int c = tid; c < A; c += NT) {
    // four independent chains, 16 weight loads in flight per thread (the loop is bound by L2 latency, not by FMAs)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
#pragma unroll 4
    for (; k + 3 < H; k += 4) {
      s0 = fmaf(query[k + 0], __ldg(a.Wq + (size_t)(k + 0) * A + c), s0);
      s1 = fmaf(query[k + 1], __ldg(a.Wq + (size_t)(k + 1) * A + c), s1);
      s2 = fmaf(query[k + 2], __ldg(a.Wq + (size_t)(k + 2) * A + c), s2);
      s3 = fmaf(query[k + 3], __ldg(a.Wq + (size_t)(k + 3) * A + c), s3);
    }
    for (; k < H; ++k) s0 = fmaf(query[k], __ldg(a.Wq + (size_t)k * A + c), s0);
    const float s = (s0 + s1) + (s2 + s3);
    q[c] = s;
    if (a.q_save && cr == 0) a.q_save[(size_t)r * A + c] = s;
  }
  __syncthreads();
  // phase 2: scores e[t] = v . tanh(q + keys[t] + cf[t] . Wd), one warp per memory position, 128-bit loads of the keys
  // (two positions in flight per warp), tanh from MUFU
  const float* keys = a.keys + (size_t)mrow * Tm * A;
  const bool vec4 = ((A | E) & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.keys) | reinterpret_cast<uintptr_t>(a.values)) & 15) == 0;
  if (vec4) {
    for (int t = tb + warp; t < te; t += 2 * NW) {
      const int t1 = t + NW;
      float sa = 0.f, sb = 0.f;
      for (int c = lane * 4; c < A; c += 128) {
        float4 ka = make_float4(0.f, 0.f, 0.f, 0.f), kb = ka;
        if (t < len) ka = __ldg(reinterpret_cast<const float4*>(keys + (size_t)t * A + c));
        if (t1 < len && t1 < te) kb = __ldg(reinterpret_cast<const float4*>(keys + (size_t)t1 * A + c));
        const float4 qq = *reinterpret_cast<const float4*>(q + c);
        const float4 vv = *reinterpret_cast<const float4*>(vs + c);
        float pa[4] = {qq.x + ka.x, qq.y + ka.y, qq.z + ka.z, qq.w + ka.w};
        float pb[4] = {qq.x + kb.x, qq.y + kb.y, qq.z + kb.z, qq.w + kb.w};
        for (int f = 0; f < F; ++f) {
          const float4 w4 = *reinterpret_cast<const float4*>(wd + f * A + c);
          const float ca = cf[t * F + f], cb = t1 < te ? cf[t1 * F + f] : 0.f;
          pa[0] = fmaf(ca, w4.x, pa[0]); pa[1] = fmaf(ca, w4.y, pa[1]); pa[2] = fmaf(ca, w4.z, pa[2]); pa[3] = fmaf(ca, w4.w, pa[3]);
          pb[0] = fmaf(cb, w4.x, pb[0]); pb[1] = fmaf(cb, w4.y, pb[1]); pb[2] = fmaf(cb, w4.z, pb[2]); pb[3] = fmaf(cb, w4.w, pb[3]);
        }
        sa += vv.x * tanh_fast(pa[0]) + vv.y * tanh_fast(pa[1]) + vv.z * tanh_fast(pa[2]) + vv.w * tanh_fast(pa[3]);
        sb += vv.x * tanh_fast(pb[0]) + vv.y * tanh_fast(pb[1]) + vv.z * tanh_fast(pb[2]) + vv.w * tanh_fast(pb[3]);
      }
      sa = warp_sum(sa);
      sb = warp_sum(sb);
      if (lane == 0) {
        e[t] = (t < len) ? sa : -CUDART_INF_F;
        if (t1 < te) e[t1] = (t1 < len) ? sb : -CUDART_INF_F;
      }
    }
  } else {
    for (int t = tb + warp; t < te; t += NW) {
      float s = 0.f;
      if (t < len) {
        for (int c = lane; c < A; c += 32) {
          float pre = q[c] + keys[(size_t)t * A + c];
          for (int f = 0; f < F; ++f) pre = fmaf(cf[t * F + f], wd[f * A + c], pre);
          s = fmaf(vs[c], tanh_fast(pre), s);
        }
        s = warp_sum(s);
      }
      if (lane == 0) e[t] = (t < len) ? s : -CUDART_INF_F;
    }
  }
  __syncthreads();
  // WindowedAttention: keep the scores inside the window around the previous alignment's median frame only
  // (attention.py:372-386): half_step = cumsum(alpha_prev) > 0.5 (sequential fp32 sum, as a CPU tf.cumsum), window =
  // half_step shifted left by win_left + 1 (true shifted in) XOR shifted right by win_right (false shifted in).
  if (a.win_left >= 0) {
    for (int t = tid; t < Tm; t += NT) {
      float c = 0.f;
      for (int j = 0; j <= t; ++j) c += ap[j];         // padl = 0 here: ap[j] = alpha_prev[j]
      cpart[t] = c > 0.5f ? 1.f : 0.f;
    }
    __syncthreads();
    for (int t = tb + tid; t < te; t += NT) {
      const bool sl = t + a.win_left + 1 < Tm ? cpart[t + a.win_left + 1] != 0.f : true;
      const bool sr = t - a.win_right >= 0 ? cpart[t - a.win_right] != 0.f : false;
      if (!(sl != sr)) e[t] = -CUDART_INF_F;
    }
    __syncthreads();
  }
  // phase 3: alignments from the masked scores: softmax, or sigmoid / normalised sigmoid (components/attention.py:41-55;
  // tf.sigmoid(-inf) = 0 on the masked positions)
  if (CS > 1) {
    // the cluster form: statistics of this CTA's range, one exchange, then the global normalisation
    float mx = -CUDART_INF_F, sum = 0.f;
    if (a.prob == 0) {
      for (int t = tb + tid; t < te; t += NT) mx = fmaxf(mx, e[t]);
      mx = block_reduce(mx, red, true);
      for (int t = tb + tid; t < te; t += NT) {
        const float p = (t < len && e[t] > -CUDART_INF_F) ? expf(e[t] - mx) : 0.f;
        e[t] = p;
        sum += p;
      }
    } else {
      for (int t = tb + tid; t < te; t += NT) {
        const float p = (t < len) ? 1.f / (1.f + expf(-e[t])) : 0.f;
        e[t] = p;
        sum += p;
      }
    }
    sum = block_reduce(sum, red, false);
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (tid < CS) {
      float* px = cluster.map_shared_rank(xstat, tid) + cr * 4;
      px[0] = mx; px[1] = sum;
    }
    cluster.sync();
    float scale = 1.f;
    if (a.prob == 0) {
      float M = -CUDART_INF_F, total = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) M = fmaxf(M, xstat[c * 4]);
#pragma unroll
      for (int c = 0; c < CS; ++c) total += xstat[c * 4 + 1] > 0.f ? xstat[c * 4 + 1] * expf(xstat[c * 4] - M) : 0.f;
      scale = sum > 0.f ? expf(mx - M) / total : 0.f;
    } else if (a.prob == 1) {
      float total = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) total += xstat[c * 4 + 1];
      scale = 1.f / total;
      if (a.asum_save && tid == 0 && cr == 0) a.asum_save[r] = total;
    }
    for (int t = tb + tid; t < te; t += NT) {
      const float p = e[t] * scale;
      e[t] = p;
      a.align_new[(size_t)r * Tm + t] = p;
    }
  } else
  if (a.prob == 0) {
    float mx = -CUDART_INF_F;
    for (int t = tid; t < Tm; t += NT) mx = fmaxf(mx, e[t]);
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int t = tid; t < Tm; t += NT) {
      const float p = (t < len) ? expf(e[t] - mx) : 0.f;
      e[t] = p;
      sum += p;
    }
    sum = block_reduce(sum, red, false);
    const float inv = 1.f / sum;
    for (int t = tid; t < Tm; t += NT) {
      const float p = e[t] * inv;
      e[t] = p;
      a.align_new[(size_t)r * Tm + t] = p;
    }
  } else {
    float sum = 0.f;
    for (int t = tid; t < Tm; t += NT) {
      const float p = (t < len) ? 1.f / (1.f + expf(-e[t])) : 0.f;
      e[t] = p;
      sum += p;
    }
    float inv = 1.f;
    if (a.prob == 1) {
      sum = block_reduce(sum, red, false);
      inv = 1.f / sum;
      if (a.asum_save && tid == 0) a.asum_save[r] = sum;
    } else {
      __syncthreads();
    }
    for (int t = tid; t < Tm; t += NT) {
      const float p = e[t] * inv;
      e[t] = p;
      a.align_new[(size_t)r * Tm + t] = p;
    }
  }
  __syncthreads();
  // phase 4: context = alpha . values.  128-bit loads; the memory positions are split over 256 / (E/4) thread groups and
  // every thread keeps 4 loads in flight (the old one-column-per-thread loop was a chain of Tm dependent loads).
  const float* values = a.values + (size_t)mrow * Tm * E;
  if (vec4 && E <= 1024) {
    const int NC4 = E >> 2;
    const int TS = NT / NC4 > 0 ? NT / NC4 : 1;            // E <= 1024 -> NC4 <= 256
    const int cg = tid % NC4, ts = tid / NC4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int lim = min(te, len);                          // this CTA's valid positions end here
    if (ts < TS) {
      int t = tb + ts;
#pragma unroll 1
      for (; t + 3 * TS < lim; t += 4 * TS) {
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(values + (size_t)(t) * E) + cg);
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(values + (size_t)(t + TS) * E) + cg);
        const float4 x2 = __ldg(reinterpret_cast<const float4*>(values + (size_t)(t + 2 * TS) * E) + cg);
        const float4 x3 = __ldg(reinterpret_cast<const float4*>(values + (size_t)(t + 3 * TS) * E) + cg);
        const float e0 = e[t], e1 = e[t + TS], e2 = e[t + 2 * TS], e3 = e[t + 3 * TS];
        acc.x += (e0 * x0.x + e1 * x1.x) + (e2 * x2.x + e3 * x3.x);
        acc.y += (e0 * x0.y + e1 * x1.y) + (e2 * x2.y + e3 * x3.y);
        acc.z += (e0 * x0.z + e1 * x1.z) + (e2 * x2.z + e3 * x3.z);
        acc.w += (e0 * x0.w + e1 * x1.w) + (e2 * x2.w + e3 * x3.w);
      }
      for (; t < lim; t += TS) {
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(values + (size_t)t * E) + cg);
        const float e0 = e[t];
        acc.x = fmaf(e0, x0.x, acc.x); acc.y = fmaf(e0, x0.y, acc.y); acc.z = fmaf(e0, x0.z, acc.z); acc.w = fmaf(e0, x0.w, acc.w);
      }
      *reinterpret_cast<float4*>(cpart + (size_t)tid * 4) = acc;
    }
    __syncthreads();
    for (int i = tid; i < E; i += NT) {
      float s = 0.f;
      for (int k = 0; k < TS; ++k) s += cpart[(size_t)(k * NC4 + (i >> 2)) * 4 + (i & 3)];
      ctx[i] = s;
    }
  } else {
    for (int i = tid; i < E; i += NT) {
      float s = 0.f;
      for (int t = tb; t < min(te, len); ++t) s = fmaf(e[t], values[(size_t)t * E + i], s);
      ctx[i] = s;
    }
  }
  if (CS > 1) {
    // partial contexts into slot cr of every other CTA; summed in CTA order so that every CTA holds the same bits
    for (int i = tid; i < E; i += NT) {
      const float s = ctx[i];
      for (int pc = 0; pc < CS; ++pc)
        if (pc != cr) cluster.map_shared_rank(xctx, pc)[(size_t)cr * E + i] = s;
    }
    cluster.sync();                          // no remote access after this
    if (cr != 0) return;                     // the output projection and the row's outputs are CTA 0's
    for (int i = tid; i < E; i += NT) {
      float s = 0.f;
      for (int pc = 0; pc < CS; ++pc) s += pc == cr ? ctx[i] : xctx[(size_t)pc * E + i];
      ctx[i] = s;
    }
  }
  for (int i = tid; i < E; i += NT) {        // (each thread reads back what it wrote)
    const float s = ctx[i];
    a.ctx_new[(size_t)r * E + i] = s;
    a.ctxT_new[(size_t)i * a.R + r] = s;
  }
  __syncthreads();
  if (a.outin_save) {
    float* o = a.outin_save + r * a.outin_row_stride;
    for (int i = tid; i < H; i += NT) o[i] = query[i];
    for (int i = tid; i < E; i += NT) o[H + i] = ctx[i];
  }
  // phase 5: logits = [query, ctx] . Wo + bo ; thread (vcol = lane, kchunk = warp)
  for (int v0 = 0; v0 < V; v0 += 32) {
    const int vc = v0 + lane;
    float s = 0.f;
    if (vc < V) {
      float s1 = 0.f;
      int k = warp;
#pragma unroll 4
      for (; k + NW < H; k += 2 * NW) {
        s = fmaf(query[k], __ldg(a.Wo + (size_t)k * V + vc), s);
        s1 = fmaf(query[k + NW], __ldg(a.Wo + (size_t)(k + NW) * V + vc), s1);
      }
      for (; k < H; k += NW) s = fmaf(query[k], __ldg(a.Wo + (size_t)k * V + vc), s);
      k = warp;
#pragma unroll 4
      for (; k + NW < E; k += 2 * NW) {
        s = fmaf(ctx[k], __ldg(a.Wo + (size_t)(H + k) * V + vc), s);
        s1 = fmaf(ctx[k + NW], __ldg(a.Wo + (size_t)(H + k + NW) * V + vc), s1);
      }
      for (; k < E; k += NW) s = fmaf(ctx[k], __ldg(a.Wo + (size_t)(H + k) * V + vc), s);
      s += s1;
    }
    part[warp * 32 + lane] = s;
    __syncthreads();
    if (warp == 0 && vc < V) {
      float t = a.bo[vc];
      for (int w = 0; w < NW; ++w) t += part[w * 32 + lane];
      a.logits[r * a.logits_row_stride + vc] = t / a.temperature;
    }
    __syncthreads();
  }
}

inline size_t attn_step_smem(int Tm, int E, int H, int A, int F, int ksz, int cs = 1) {
  // the arrays read with 128-bit accesses (q, wd, vs, cpart) must start on 16-byte boundaries: round every piece up to 4
  auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  return (r4(H) + r4(A) + r4(Tm + ksz + 4) + r4(Tm) + r4(E) + 32 + 512 + r4((size_t)Tm * F) + r4((size_t)F * A) +
          r4((size_t)ksz * F) + r4(A) + 2048 + 16 + (cs > 1 ? (size_t)cs * E : 0)) * sizeof(float);
}
// CTAs per decoder row of the attention step: clusters only for training-size batches (one memory per row) that cannot
// fill the SMs on their own (NABU_ATTN_FWD_CLUSTER=1|2|4 forces)
inline int attn_step_cluster(int rows, int rows_per_mem) {
  static int forced = -1;
  if (forced < 0) forced = getenv("NABU_ATTN_FWD_CLUSTER") ? atoi(getenv("NABU_ATTN_FWD_CLUSTER")) : 0;
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  if (rows_per_mem != 1) return 1;
  return rows * 4 <= 160 ? 4 : rows * 2 <= 160 ? 2 : 1;
}
inline cudaError_t attn_step_launch(const AttnStepArgs& a, int rows, size_t* smem_out, cudaStream_t stream) {
  const int cs = attn_step_cluster(rows, a.rows_per_mem);
  const size_t smem = attn_step_smem(a.Tm, a.E, a.H, a.A, a.F, a.ksz, cs);
  if (smem_out) *smem_out = smem;
  void (*fn)(const AttnStepArgs) = cs == 4 ? dec_attn_step_kernel<4> : cs == 2 ? dec_attn_step_kernel<2> : dec_attn_step_kernel<1>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  return chain_launch(fn, dim3(rows * cs), dim3(512), smem, stream, a);
}

// ------------------------------------------------------------------------------------------------
// Backward of one LSTM cell unit (row r, unit j) given the gradient `dha` wrt the cell OUTPUT; fused into the kernel
// that produces dha (round 2: the attention backward for the top layer, the transposed matmul of the layer above for the
// others) instead of a kernel of its own.  gates [R][4H] holds the activated i,g,f,o and receives dz; dzT [4H][R].
// ------------------------------------------------------------------------------------------------
struct LstmBwdPw {
  float* gates; const float* c_new; const float* c_prev; const float* dh_carry; float* dc_carry; float* dzT;
  const int* tlen; int u; float keep; unsigned seed; int layer;
};
__device__ __forceinline__ void lstm_bwd_unit(const LstmBwdPw& p, int R, int H, int r, int j, float dha) {
  const size_t i = (size_t)r * H + j;
  float* gp = p.gates + (size_t)r * 4 * H + j;
  const bool active = p.u < p.tlen[r];
  float dz[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    const float ig = gp[0], gg = gp[H], fg = gp[2 * H], og = gp[3 * H];
    // dha is the gradient wrt the cell OUTPUT (dropped), dh_carry wrt the state h
    if (p.keep > 0.f && p.keep < 1.f)
      dha = dec_uniform(p.seed, (uint32_t)(p.layer * 65536 + p.u), (uint32_t)r, (uint32_t)j) < p.keep ? dha / p.keep : 0.f;
    const float dh = dha + p.dh_carry[i];
    const float tc = tanhf(p.c_new[i]);
    const float dc = p.dc_carry[i] + dh * og * (1.f - tc * tc);
    dz[0] = dc * gg * ig * (1.f - ig);
    dz[1] = dc * ig * (1.f - gg * gg);
    dz[2] = dc * p.c_prev[i] * fg * (1.f - fg);
    dz[3] = dh * tc * og * (1.f - og);
    p.dc_carry[i] = dc * fg;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    gp[g * H] = dz[g];
    p.dzT[(size_t)(g * H + j) * R + r] = dz[g];
  }
}

// ------------------------------------------------------------------------------------------------
// Skinny transposed-weight matmul: out[r][n] = sum_k xT[k][r] * W[row0 + n][k]  (d(input) = dz.K^T).
// grid = (N/8, ceil(R/ROWS)).  Output columns [0,N0) go to out0, [N0,N) to out1.
// ------------------------------------------------------------------------------------------------
struct MatmulTArgs {
  const float* xT; int K; int R;          // [K][R]
  const float* W; int ldw; int row0;      // W[(row0+n)][k]
  const float* Wr;                        // optional [N/8][K][8]: the CTAs' weight slices laid out contiguously (relayout_bwd)
  int N, N0;
  float* out0; int ld0; float* out1; int ld1;
  // pw.gates != nullptr: columns [0, N0) are d(output) of the LSTM layer below (N0 = its num_units); instead of being
  // stored they go straight through that layer's cell backward (lstm_bwd_unit)
  LstmBwdPw pw;
};

__global__ void __launch_bounds__(MT_THREADS) dec_matmul_t_kernel(const MatmulTArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                          // [K][8]
  float* red = sm + (size_t)a.K * 8;       // [MT_KSPLIT][ROWS][8]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * 8, r0 = blockIdx.y * ROWS;
  __shared__ uint64_t bars[3];
  if (tid == 0) { dk_bar_init(&bars[0]); dk_bar_init(&bars[1]); dk_bar_init(&bars[2]); }
  if (a.Wr) {
    if (tid == 0) dk_bulk_load(Ws, a.Wr + (size_t)blockIdx.x * a.K * 8, (unsigned)(a.K * 8 * sizeof(float)), &bars[2]);
  } else {
    for (int i = tid; i < a.K * 8; i += MT_THREADS) {
      const int c = i / a.K, k = i % a.K;    // k fastest: coalesced rows of W
      const int n = n0 + c;
      Ws[k * 8 + c] = (n < a.N) ? a.W[(size_t)(a.row0 + n) * a.ldw + k] : 0.f;
    }
  }
  __syncthreads();                         // (the barriers are initialised for every thread)
  chain_wait();                            // the weight slice was staged next to the previous kernel
  if (a.Wr) dk_bar_wait(&bars[2], 0);
  __syncthreads();
  const int rl = tid % ROWS, ks = tid / ROWS, r = r0 + rl;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (dk_bulk_ok(a.xT, a.xT, a.R)) {
    float* xs = red + MT_KSPLIT * ROWS * 8;            // [2][XC][R]
    unsigned issued = 0, consumed = 0;
    dk_stream_fma<MT_THREADS>(a.xT, a.K, a.R, r0, Ws, 0, xs, bars, issued, consumed, rl, ks, r < a.R, acc);
  } else if (r < a.R) {
#pragma unroll 16
    for (int k = ks; k < a.K; k += MT_KSPLIT) {
      const float x = __ldcg(a.xT + (size_t)k * a.R + r);
      const float4 w0 = *reinterpret_cast<const float4*>(Ws + k * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(Ws + k * 8 + 4);
      acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
      acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
      acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
      acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) red[(ks * ROWS + rl) * 8 + c] = acc[c];
  __syncthreads();
  for (int i = tid; i < ROWS * 8; i += MT_THREADS) {
    const int rl2 = i / 8, c = i % 8, r2 = r0 + rl2, n = n0 + c;
    if (r2 < a.R && n < a.N) {
      float s = 0.f;
#pragma unroll
      for (int k2 = 0; k2 < MT_KSPLIT; ++k2) s += red[(k2 * ROWS + rl2) * 8 + c];
      if (n >= a.N0) a.out1[(size_t)r2 * a.ld1 + (n - a.N0)] = s;
      else if (a.pw.gates) lstm_bwd_unit(a.pw, a.R, a.N0, r2, n, s);
      else a.out0[(size_t)r2 * a.ld0 + n] = s;
    }
  }
}

}  // namespace dec
}  // namespace nabu
