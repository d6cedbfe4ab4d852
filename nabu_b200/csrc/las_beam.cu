// LAS beam search (replaces neuralnetworks/decoders/
// beam_search_decoder.py:30-112 driving components/beam_search_decoder.py:136-485 under
// tf.contrib.seq2seq.dynamic_decode(maximum_iterations=max_steps)).
//
// Per step, on R = B*W decoder rows (row b*W+w, the tile_batch order):
//   speller step (dec_lstm_step x layers + dec_attn_step) on the un-tiled memory (rows_per_mem = W
//   instead of tiling the encoder output W times),
//   las_prune_kernel   one CTA per utterance: log-softmax, finished rows -> -FLT_MAX, add to the beam
//                      log-probs, append the W "stay" hypotheses, length-penalised score, top-W with
//                      tf.nn.top_k's tie order (lowest index first), parent / id / length / finished,
//   las_gather_kernel  new state of slot w = new cell state of its parent (expansion) or the OLD state
//                      of its own slot (stay) -- the reference tiles every state V times and gathers;
//                      here only W rows move -- and the alignment-history write,
//   las_done_kernel    dynamic_decode's sticky `finished |= step_finished` per slot; once every slot of
//                      every utterance has held EOS (or max_steps) a device flag freezes all later
//                      launches, so the host loop needs no synchronisation.
// las_finalize_kernel back-traces parents exactly like BeamSearchDecoder.finalize.
#include "common.cuh"
#include "speller_api.h"
#include "nabu_b200.h"
#include <math_constants.h>
#include <float.h>

namespace nabu {
namespace {

struct BeamBuf {
  // three rotating state sets
  float* hT[3][4]; float* h[3][4]; float* c[3][4];
  float* ctx[3]; float* ctxT[3]; float* align[3];
  float* values; float* keys; float* logits;        // [B][Tm][E], [B][Tm][A], [R][V]
  int* ids; float* logprobs; int* lengths; int* finished; int* loop_finished;   // [R]
  int* parent; int* is_stay;                        // [R]
  int* pred_hist; int* parent_hist;                 // [max_steps][R]
  float* align_hist;                                // [max_steps][R][Tm]
  int* done; int* n_steps;                          // device scalars
  float* wf[4];                                     // the cells' weight slices as the step kernel stages them
  size_t total;
};

BeamBuf carve_beam(void* base, const nabu_speller_desc_t& d, int W, int max_steps) {
  BeamBuf b;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t n4) { void* q = (void*)(p + off); off += align_up(n4 * 4, 256); return q; };
  const size_t R = (size_t)d.B * W, H = d.H, E = d.E, Tm = d.Tm;
  for (int s = 0; s < 3; ++s) {
    for (int l = 0; l < d.num_layers; ++l) {
      b.hT[s][l] = (float*)take(H * R); b.h[s][l] = (float*)take(R * H); b.c[s][l] = (float*)take(R * H);
    }
    b.ctx[s] = (float*)take(R * E); b.ctxT[s] = (float*)take(E * R); b.align[s] = (float*)take(R * Tm);
  }
  b.values = (float*)take((size_t)d.B * Tm * E);
  b.keys = (float*)take((size_t)d.B * Tm * d.A);
  b.logits = (float*)take(R * d.V);
  b.ids = (int*)take(R); b.logprobs = (float*)take(R); b.lengths = (int*)take(R);
  b.finished = (int*)take(R); b.loop_finished = (int*)take(R); b.parent = (int*)take(R); b.is_stay = (int*)take(R);
  b.pred_hist = (int*)take((size_t)max_steps * R); b.parent_hist = (int*)take((size_t)max_steps * R);
  b.align_hist = (float*)take((size_t)max_steps * R * Tm);
  b.done = (int*)take(2); b.n_steps = b.done + 1;
  for (int l = 0; l < d.num_layers; ++l) b.wf[l] = (float*)take(dec::relayout_fwd_floats(d, l));
  b.total = off;
  return b;
}

__global__ void las_init_kernel(int* ids, float* logprobs, int* lengths, int* finished, int* loop_finished, int R,
                                int W, int V, int* done, int* n_steps, int max_steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { *done = (max_steps <= 0) ? 1 : 0; *n_steps = 0; }
  if (i >= R) return;
  ids[i] = V - 1;                                         // start tokens = EOS/SOS label
  logprobs[i] = (i % W == 0) ? 0.f : -CUDART_INF_F;       // beam_search_decoder.py:158-160
  lengths[i] = 0;
  finished[i] = 0;
  loop_finished[i] = 0;
}

__device__ __forceinline__ float las_score(float lp, int len, float w) {
  if (w == 0.f) return lp;
  // ((5 + len)^w) / (6^w), then logprob / penalty  (beam_search_decoder.py:482-485)
  const float pen = (w == 1.f) ? (5.f + (float)len) / 6.f : powf(5.f + (float)len, w) / powf(6.f, w);
  return lp / pen;
}

// one CTA per utterance; dynamic smem: lp[W*V + W] scores, cand ids/lengths implicit
__global__ void __launch_bounds__(256) las_prune_kernel(const float* logits, int W, int V, float lpw, int* ids,
                                                        float* logprobs, int* lengths, int* finished, int* parent,
                                                        int* is_stay, int* pred_hist_t, int* parent_hist_t,
                                                        const int* done) {
  extern __shared__ float sm[];
  if (*done) return;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int eos = V - 1;
  const int NC = W * V + W;
  float* clp = sm;                   // [NC] candidate log-probs
  float* csc = clp + NC;             // [NC] candidate scores
  int* taken = (int*)(csc + NC);     // [NC]
  float* lse = (float*)(taken + NC); // [2][W] row max, log-sum
  __shared__ float rbest[8];
  __shared__ int ridx[8];
  __shared__ int sel[64];

  // log-softmax normaliser per beam row (warp per row)
  for (int w = warp; w < W; w += 8) {
    const float* x = logits + ((size_t)b * W + w) * V;
    float mx = -CUDART_INF_F;
    for (int k = lane; k < V; k += 32) mx = fmaxf(mx, x[k]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int k = lane; k < V; k += 32) s += expf(x[k] - mx);
    s = warp_sum(s);
    if (lane == 0) { lse[w] = mx; lse[W + w] = logf(s); }
  }
  __syncthreads();
  for (int i = tid; i < NC; i += 256) {
    float lp; int len;
    if (i < W * V) {
      const int w = i / V, k = i % V;
      const int r = b * W + w;
      const float nl = finished[r] ? -FLT_MAX : ((logits[(size_t)r * V + k] - lse[w]) - lse[W + w]);
      lp = logprobs[r] + nl;
      len = lengths[r] + (k == eos ? 0 : 1);
    } else {
      const int r = b * W + (i - W * V);
      lp = finished[r] ? logprobs[r] : -FLT_MAX;
      len = lengths[r];
    }
    clp[i] = lp;
    csc[i] = las_score(lp, len, lpw);
    taken[i] = 0;
  }
  __syncthreads();
  // top-W: W rounds of argmax with (score desc, index asc)
  for (int round = 0; round < W; ++round) {
    float best = -CUDART_INF_F; int bi = 0x7fffffff;
    for (int i = tid; i < NC; i += 256) {
      if (taken[i]) continue;
      const float s = csc[i];
      if (bi == 0x7fffffff || s > best || (s == best && i < bi)) { best = s; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
    }
    if (lane == 0) { rbest[warp] = best; ridx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float fb = rbest[0]; int fi = ridx[0];
      for (int w = 1; w < 8; ++w) {
        const float ob = rbest[w]; const int oi = ridx[w];
        if (oi != 0x7fffffff && (fi == 0x7fffffff || ob > fb || (ob == fb && oi < fi))) { fb = ob; fi = oi; }
      }
      sel[round] = fi;
      taken[fi] = 1;
    }
    __syncthreads();
  }
  // read everything the winners need, then (after a barrier) overwrite the beam in place
  int par = 0, id = 0, len = 0, stay = 0;
  float lp = 0.f;
  if (tid < W) {
    const int i = sel[tid];
    if (i < W * V) { par = i / V; id = i % V; stay = 0; len = lengths[b * W + par] + (id == eos ? 0 : 1); }
    else { par = i - W * V; id = eos; stay = 1; len = lengths[b * W + par]; }
    lp = clp[i];
  }
  __syncthreads();
  if (tid < W) {
    const int r = b * W + tid;
    parent[r] = par; is_stay[r] = stay;
    pred_hist_t[r] = id; parent_hist_t[r] = par;
    ids[r] = id; logprobs[r] = lp; lengths[r] = len; finished[r] = (id == eos);
  }
}

// state gather: one CTA per new row
struct GatherArgs {
  int R, W, H, E, Tm, NL;
  const int* parent; const int* is_stay;
  const float* h_old[4]; const float* c_old[4]; const float* h_new[4]; const float* c_new[4];
  const float* ctx_old; const float* ctx_new; const float* al_old; const float* al_new;
  float* h_out[4]; float* hT_out[4]; float* c_out[4]; float* ctx_out; float* ctxT_out; float* al_out;
  float* align_hist_t;
  const int* done;
};
__global__ void las_gather_kernel(const GatherArgs a) {
  if (*a.done) return;
  const int r = blockIdx.x, tid = threadIdx.x;
  const int b = r / a.W;
  const int src = b * a.W + a.parent[r];
  const bool stay = a.is_stay[r] != 0;
  for (int l = 0; l < a.NL; ++l) {
    const float* hs = stay ? a.h_old[l] : a.h_new[l];
    const float* cs = stay ? a.c_old[l] : a.c_new[l];
    for (int j = tid; j < a.H; j += blockDim.x) {
      const float hv = hs[(size_t)src * a.H + j];
      a.h_out[l][(size_t)r * a.H + j] = hv;
      a.hT_out[l][(size_t)j * a.R + r] = hv;
      a.c_out[l][(size_t)r * a.H + j] = cs[(size_t)src * a.H + j];
    }
  }
  const float* cx = stay ? a.ctx_old : a.ctx_new;
  for (int i = tid; i < a.E; i += blockDim.x) {
    const float v = cx[(size_t)src * a.E + i];
    a.ctx_out[(size_t)r * a.E + i] = v;
    a.ctxT_out[(size_t)i * a.R + r] = v;
  }
  const float* al = stay ? a.al_old : a.al_new;
  for (int t = tid; t < a.Tm; t += blockDim.x) {
    const float v = al[(size_t)src * a.Tm + t];
    a.al_out[(size_t)r * a.Tm + t] = v;
    a.align_hist_t[(size_t)r * a.Tm + t] = v;
  }
}

__global__ void las_done_kernel(const int* finished, int* loop_finished, int R, int t, int max_steps, int* done,
                                int* n_steps) {
  __shared__ int all;
  if (*done) return;
  if (threadIdx.x == 0) all = 1;
  __syncthreads();
  int mine = 1;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const int f = loop_finished[i] | finished[i] | (t + 1 >= max_steps ? 1 : 0);
    loop_finished[i] = f;
    mine &= f;
  }
  if (!mine) all = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    *n_steps = t + 1;
    if (all) *done = 1;
  }
}

// one CTA per (b, w): back-trace
__global__ void las_finalize_kernel(const int* pred_hist, const int* parent_hist, const float* align_hist,
                                    const float* logprobs, const int* lengths, const int* n_steps_p, int R, int W,
                                    int Tm, int max_steps, float lpw, int* sequences, int* out_lengths,
                                    float* scores, float* alignments) {
  const int r = blockIdx.x, b = r / W;
  const int n = *n_steps_p;
  int beam = r % W;
  for (int tt = n - 1; tt >= 0; --tt) {
    const int src = b * W + beam;
    if (threadIdx.x == 0) sequences[(size_t)r * max_steps + tt] = pred_hist[(size_t)tt * R + src];
    for (int t = threadIdx.x; t < Tm; t += blockDim.x)
      alignments[((size_t)r * max_steps + tt) * Tm + t] = align_hist[((size_t)tt * R + src) * Tm + t];
    beam = parent_hist[(size_t)tt * R + src];
  }
  if (threadIdx.x == 0) {
    out_lengths[r] = lengths[r];
    scores[r] = las_score(logprobs[r], lengths[r], lpw);
  }
}

}  // namespace
}  // namespace nabu

using namespace nabu;

extern "C" size_t nabu_las_beam_workspace_bytes(const nabu_speller_desc_t* d, int W, int max_steps) {
  if (dec::check_desc(*d) || W < 1 || max_steps < 0) return 0;
  return carve_beam(nullptr, *d, W, max_steps > 0 ? max_steps : 1).total;
}

extern "C" int nabu_las_beam_search(const nabu_speller_desc_t* dp, const nabu_speller_params_t* p,
                                    const float* memory, const int* mem_len, int W, int max_steps,
                                    float length_penalty, float temperature, int* sequences, int* lengths,
                                    float* scores, float* alignments, int* n_steps, void* workspace,
                                    size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const nabu_speller_desc_t& d = *dp;
  if (int e = dec::check_desc(d)) return e;
  NABU_REQUIRE(W >= 1 && W <= 64, "las_beam_search: beam_width=%d not in 1..64", W);
  NABU_REQUIRE(W <= d.V, "las_beam_search: beam_width=%d > output classes %d (the reference's parent trick needs W <= V)", W, d.V);
  NABU_REQUIRE(max_steps >= 1, "las_beam_search: max_steps=%d", max_steps);
  BeamBuf bb = carve_beam(workspace, d, W, max_steps);
  NABU_REQUIRE(ws_bytes >= bb.total, "las_beam_search: workspace %zu < %zu bytes", ws_bytes, bb.total);
  const int R = d.B * W, H = d.H, E = d.E, Tm = d.Tm, V = d.V, NL = d.num_layers;
  if (int e = dec::prepare_memory(d, *p, memory, mem_len, bb.values, bb.keys, stream)) return e;
  // zero initial state in set 0
  for (int l = 0; l < NL; ++l) {
    NABU_CHECK_CUDA(cudaMemsetAsync(bb.hT[0][l], 0, (size_t)H * R * 4, stream));
    NABU_CHECK_CUDA(cudaMemsetAsync(bb.h[0][l], 0, (size_t)R * H * 4, stream));
    NABU_CHECK_CUDA(cudaMemsetAsync(bb.c[0][l], 0, (size_t)R * H * 4, stream));
  }
  NABU_CHECK_CUDA(cudaMemsetAsync(bb.ctx[0], 0, (size_t)R * E * 4, stream));
  NABU_CHECK_CUDA(cudaMemsetAsync(bb.ctxT[0], 0, (size_t)E * R * 4, stream));
  NABU_CHECK_CUDA(cudaMemsetAsync(bb.align[0], 0, (size_t)R * Tm * 4, stream));
  if (d.attention == 2)
    if (int e = dec::init_window_alignments(bb.align[0], R, Tm, stream)) return e;
  {
    KernelScope ks("las_init", stream);
    las_init_kernel<<<ceil_div(R, 256), 256, 0, stream>>>(bb.ids, bb.logprobs, bb.lengths, bb.finished, bb.loop_finished,
                                                          R, W, V, bb.done, bb.n_steps, max_steps);
    NABU_CHECK_LAUNCH();
  }
  const size_t prune_smem = ((size_t)3 * (W * V + W) + 2 * W) * sizeof(float);
  NABU_REQUIRE(prune_smem <= 200 * 1024, "las_beam_search: W*V too large");
  if (prune_smem > 48 * 1024)
    NABU_CHECK_CUDA(cudaFuncSetAttribute(las_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prune_smem));
  if (int e = dec::relayout_fwd(d, *p, bb.wf, stream)) return e;
  int P = 0;   // index of the "previous" state set
  for (int t = 0; t < max_steps; ++t) {
    const int N = (P + 1) % 3, G = (P + 2) % 3;
    if (int e = dec::launch_step(d, *p, R, W, bb.ids, bb.keys, bb.values, mem_len, bb.hT[P], bb.h[P], bb.c[P], bb.ctx[P],
                                 bb.ctxT[P], bb.align[P], bb.hT[N], bb.h[N], bb.c[N], bb.ctx[N], bb.ctxT[N], bb.align[N],
                                 nullptr, bb.logits, V, temperature, nullptr, nullptr, nullptr, 0, nullptr, t,
                                 bb.done, stream, nullptr, nullptr, nullptr, 1.f, 0, bb.wf))
      return e;
    {
      KernelScope ks("las_prune", stream);
      las_prune_kernel<<<d.B, 256, prune_smem, stream>>>(bb.logits, W, V, length_penalty, bb.ids, bb.logprobs, bb.lengths,
                                                        bb.finished, bb.parent, bb.is_stay,
                                                        bb.pred_hist + (size_t)t * R, bb.parent_hist + (size_t)t * R,
                                                        bb.done);
      NABU_CHECK_LAUNCH();
    }
    GatherArgs g = {};
    g.R = R; g.W = W; g.H = H; g.E = E; g.Tm = Tm; g.NL = NL; g.parent = bb.parent; g.is_stay = bb.is_stay;
    for (int l = 0; l < NL; ++l) {
      g.h_old[l] = bb.h[P][l]; g.c_old[l] = bb.c[P][l]; g.h_new[l] = bb.h[N][l]; g.c_new[l] = bb.c[N][l];
      g.h_out[l] = bb.h[G][l]; g.hT_out[l] = bb.hT[G][l]; g.c_out[l] = bb.c[G][l];
    }
    g.ctx_old = bb.ctx[P]; g.ctx_new = bb.ctx[N]; g.al_old = bb.align[P]; g.al_new = bb.align[N];
    g.ctx_out = bb.ctx[G]; g.ctxT_out = bb.ctxT[G]; g.al_out = bb.align[G];
    g.align_hist_t = bb.align_hist + (size_t)t * R * Tm; g.done = bb.done;
    {
      KernelScope ks("las_gather", stream);
      las_gather_kernel<<<R, 128, 0, stream>>>(g);
      NABU_CHECK_LAUNCH();
    }
    {
      KernelScope ks("las_done", stream);
      las_done_kernel<<<1, 256, 0, stream>>>(bb.finished, bb.loop_finished, R, t, max_steps, bb.done, bb.n_steps);
      NABU_CHECK_LAUNCH();
    }
    P = G;
  }
  {
    KernelScope ks("las_finalize", stream);
    las_finalize_kernel<<<R, 64, 0, stream>>>(bb.pred_hist, bb.parent_hist, bb.align_hist, bb.logprobs, bb.lengths,
                                              bb.n_steps, R, W, Tm, max_steps, length_penalty, sequences, lengths, scores,
                                              alignments);
    NABU_CHECK_LAUNCH();
  }
  NABU_CHECK_CUDA(cudaMemcpyAsync(n_steps, bb.n_steps, sizeof(int), cudaMemcpyDeviceToHost, stream));
  NABU_CHECK_CUDA(cudaStreamSynchronize(stream));
  return 0;
}
