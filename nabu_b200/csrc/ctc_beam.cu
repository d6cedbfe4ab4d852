// CTC prefix beam search (replaces neuralnetworks/decoders/ctc_decoder.py:57
// -> tf.nn.ctc_beam_search_decoder(beam_width=100, top_paths=1, merge_repeated=True), which TF-1.8 only
// runs on the CPU: tensorflow/core/util/ctc/ctc_beam_search.h, restated in SURVEY appendix B8).
//
// One CTA per utterance walks its frames.  Per frame the existing leaves are updated in parallel, but
// the "grow new leaves" phase is replayed in TF's exact order (branches by descending oldp, labels
// ascending) by one warp: TF's Step() has an order-dependent side effect -- a rejected child gets
// oldp.Reset(), and when that child is itself a leaf that was evicted earlier in the same frame and has
// not had its turn as a branch yet, it silently loses its chance to grow -- so "the beam_width best of
// all candidates" is NOT what the reference computes (measured: 1 in 4 random utterances differs).
// The warp keeps the bounded top-N as a flat array + its current bottom (warp arg-min), scores the
// labels of one branch across lanes and settles them with ballots.  The prefix tree lives in global memory as
// (parent, label) arrays plus an open-addressing hash (parent, label) -> node, so a prefix that drops
// out and later re-enters is the SAME node, as in TF (its children keep their parent link).
#include "common.cuh"
#include "nabu_b200.h"
#include <math_constants.h>

namespace nabu {
namespace {

constexpr int CB_THREADS = 128;
constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ float lse_tf(float a, float b) {     // ctc_loss_util.h LogSumExp
  if (a == -CUDART_INF_F) return b;
  if (b == -CUDART_INF_F) return a;
  return (a > b) ? log1pf(expf(b - a)) + a : log1pf(expf(a - b)) + b;
}

// order-preserving map float -> uint (larger float = larger uint)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct UttBuf {
  int* node_parent; int* node_label; int* node_slot;
  unsigned long long* hkeys; int* hvals;
  int hcap; int nmax;
};

__device__ int hash_find_or_insert(const UttBuf& u, int parent, int label, int V, int* node_count) {
  const unsigned long long key = (unsigned long long)parent * (unsigned long long)V + (unsigned long long)label;
  unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40) & (unsigned)(u.hcap - 1);
  while (true) {
    const unsigned long long prev = atomicCAS(&u.hkeys[h], EMPTY_KEY, key);
    if (prev == EMPTY_KEY) {                       // we own the slot: allocate the node
      const int n = atomicAdd(node_count, 1);
      u.node_parent[n] = parent;
      u.node_label[n] = label;
      u.node_slot[n] = -1;
      __threadfence_block();
      atomicExch(&u.hvals[h], n);
      return n;
    }
    if (prev == key) {                             // exists (maybe being created by nobody else: keys are
      int n;                                       // unique per step, so the value is already published)
      while ((n = atomicAdd(&u.hvals[h], 0)) < 0) { }
      return n;
    }
    h = (h + 1) & (unsigned)(u.hcap - 1);
  }
}

// dynamic smem layout (W = beam width, V classes):
//   float old_t/old_b/old_l/new_t/new_b/new_l/tmp_t/tmp_b/tmp_l [W]; float ent_val[W];
//   int slot_node[W], tmp_node[W], ent_ref[W]; int alive[W], dead_old[W]; float x[V];
//   unsigned short child_slot[W*V]
constexpr unsigned short NO_CHILD = 0xFFFFu;

__global__ void __launch_bounds__(CB_THREADS) ctc_beam_kernel(
    const float* logits, const int* logit_len, int T, int V, int W, int merge_repeated,
    int* g_node_parent, int* g_node_label, int* g_node_slot, unsigned long long* g_hkeys, int* g_hvals,
    int* g_node_count, int nmax, int hcap, int* out_ids, int* out_len, float* out_neg_logprob) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blank = V - 1;
  const float NINF = -CUDART_INF_F;
  float* old_t = reinterpret_cast<float*>(smraw);
  float* old_b = old_t + W; float* old_l = old_b + W;
  float* new_t = old_l + W; float* new_b = new_t + W; float* new_l = new_b + W;
  float* tmp_t = new_l + W; float* tmp_b = tmp_t + W; float* tmp_l = tmp_b + W;
  float* ent_val = tmp_l + W;
  int* slot_node = reinterpret_cast<int*>(ent_val + W);
  int* tmp_node = slot_node + W;
  int* ent_ref = tmp_node + W;
  int* alive = ent_ref + W;
  int* dead_old = alive + W;
  float* x = reinterpret_cast<float*>(dead_old + W);
  unsigned short* child_slot = reinterpret_cast<unsigned short*>(x + V);
  __shared__ int s_nleaf, s_n;
  __shared__ float s_max;

  UttBuf u;
  u.node_parent = g_node_parent + (size_t)b * nmax;
  u.node_label = g_node_label + (size_t)b * nmax;
  u.node_slot = g_node_slot + (size_t)b * nmax;
  u.hkeys = g_hkeys + (size_t)b * hcap;
  u.hvals = g_hvals + (size_t)b * hcap;
  u.hcap = hcap; u.nmax = nmax;
  int* node_count = g_node_count + b;

  if (tid == 0) {
    // root: node 0, total = blank = log 1, label = log 0
    u.node_parent[0] = -1; u.node_label[0] = -1; u.node_slot[0] = 0;
    *node_count = 1;
    slot_node[0] = 0;
    new_t[0] = 0.f; new_b[0] = 0.f; new_l[0] = NINF;
    s_nleaf = 1;
  }
  __syncthreads();
  const int Tb = min(logit_len[b], T);
  const float* lg = logits + (size_t)b * T * V;

  for (int t = 0; t < Tb; ++t) {
    const int nleaf = s_nleaf;          // leaves are stored sorted by newp.total, descending (Extract())
    // ---- A: x = logits[t] - max   (Step(): "remove the max for stability") ------------------------
    if (warp == 0) {
      float m = NINF;
      for (int k = lane; k < V; k += 32) m = fmaxf(m, lg[(size_t)t * V + k]);
      m = warp_max(m);
      if (lane == 0) s_max = m;
    }
    for (int i = tid; i < W * V; i += CB_THREADS) child_slot[i] = NO_CHILD;
    __syncthreads();
    for (int k = tid; k < V; k += CB_THREADS) x[k] = lg[(size_t)t * V + k] - s_max;
    // ---- B: oldp = newp -------------------------------------------------------------------------------
    for (int i = tid; i < nleaf; i += CB_THREADS) { old_t[i] = new_t[i]; old_b[i] = new_b[i]; old_l[i] = new_l[i]; }
    __syncthreads();
    // ---- C: existing leaves extend by blank / by a repeat of their own label --------------------------
    for (int i = tid; i < nleaf; i += CB_THREADS) {
      const int node = slot_node[i];
      float nl = old_l[i];
      if (node != 0) {
        const int label = u.node_label[node];
        const int par = u.node_parent[node];
        const int ps = u.node_slot[par];
        if (ps >= 0) {                                      // parent->Active()
          const int plabel = u.node_label[par];
          const float prev = (label == plabel) ? old_b[ps] : old_t[ps];
          nl = lse_tf(nl, prev);
          child_slot[ps * V + label] = (unsigned short)i;   // this (branch, label) child is a leaf already
        }
        nl = nl + x[label];
      }
      const float nb = old_t[i] + x[blank];
      const float nt = lse_tf(nb, nl);
      new_l[i] = nl; new_b[i] = nb; new_t[i] = nt;
      ent_val[i] = nt; ent_ref[i] = i; alive[i] = 1; dead_old[i] = 0;
    }
    __syncthreads();
    // ---- D: "grow new leaves", in TF's order: branches by descending oldp, labels ascending.  One warp
    //      replays the bounded top-N insertions; lanes hold the labels of the current branch. ----------
    if (warp == 0) {
      int n = nleaf;
      float bot_val = NINF; int bot_idx = -1;
      auto find_bottom = [&]() {
        float bv = CUDART_INF_F; int bi = 0x7fffffff;
        for (int k = lane; k < n; k += 32) {
          const float v = ent_val[k];
          if (v < bv || (v == bv && k < bi)) { bv = v; bi = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        bot_val = bv; bot_idx = bi;
      };
      if (n == W) find_bottom();
      for (int i = 0; i < nleaf; ++i) {
        const float ot = old_t[i], ob = old_b[i];
        // is_candidate(b->oldp); a branch whose oldp was Reset() by the quirk below has total = log 0
        if (dead_old[i] || !(ot > NINF)) continue;
        if (n == W && !(ot > bot_val)) continue;
        const int bnode = slot_node[i];
        const int blabel = (bnode == 0) ? -1 : u.node_label[bnode];
        for (int c0 = 0; c0 < blank; c0 += 32) {
          const int c = c0 + lane;
          const bool has = c < blank;
          const int cs = has ? (int)child_slot[i * V + c] : (int)NO_CHILD;
          const float val = has ? x[c] + ((c == blabel) ? ob : ot) : NINF;
          unsigned remaining = __ballot_sync(0xffffffffu, has);
          while (remaining) {
            const bool mine = (remaining >> lane) & 1u;
            const bool active_now = mine && cs != (int)NO_CHILD && alive[cs];
            const bool pass = mine && !active_now && val > NINF && (n < W || val > bot_val);
            const unsigned pmask = __ballot_sync(0xffffffffu, pass);
            const int L = pmask ? (__ffs(pmask) - 1) : 32;
            // labels before L are settled now: an inactive child that is not a candidate gets
            // oldp.Reset()/newp.Reset() -- if that child is itself a pending branch it will not grow
            if (mine && lane < L && !active_now && cs != (int)NO_CHILD) dead_old[cs] = 1;
            if (L == 32) break;
            if (lane == L) {
              int pos;
              if (n < W) {
                pos = n;
              } else {
                pos = bot_idx;
                const int ref = ent_ref[pos];
                if (ref < W) alive[ref] = 0;              // bottom->newp.Reset()
              }
              ent_val[pos] = val;
              ent_ref[pos] = W + i * V + c;
            }
            __syncwarp();
            if (n < W) ++n;
            if (n == W) find_bottom();
            remaining &= ~((2u << L) - 1u);               // lanes <= L are done
          }
          __syncwarp();
        }
      }
      if (lane == 0) s_n = n;
    }
    __syncthreads();
    // ---- E: materialise the survivors -------------------------------------------------------------------
    const int n = s_n;
    for (int i = tid; i < nleaf; i += CB_THREADS) u.node_slot[slot_node[i]] = -1;
    __syncthreads();
    for (int k = tid; k < n; k += CB_THREADS) {
      const int ref = ent_ref[k];
      if (ref < W) {
        tmp_node[k] = slot_node[ref]; tmp_t[k] = new_t[ref]; tmp_b[k] = new_b[ref]; tmp_l[k] = new_l[ref];
      } else {
        const int j = ref - W, li = j / V, c = j % V;
        tmp_node[k] = hash_find_or_insert(u, slot_node[li], c, V, node_count);
        tmp_t[k] = ent_val[k]; tmp_b[k] = NINF; tmp_l[k] = ent_val[k];
      }
    }
    __syncthreads();
    // ---- F: next frame's branch order = descending newp.total (rank sort, stable) ---------------------
    for (int k = tid; k < n; k += CB_THREADS) {
      const float v = tmp_t[k];
      int r = 0;
      for (int j = 0; j < n; ++j) {
        const float o = tmp_t[j];
        r += (o > v || (o == v && j < k)) ? 1 : 0;
      }
      slot_node[r] = tmp_node[k]; new_t[r] = v; new_b[r] = tmp_b[k]; new_l[r] = tmp_l[k];
      u.node_slot[tmp_node[k]] = r;
    }
    if (tid == 0) s_nleaf = n;
    __syncthreads();
  }

  if (tid == 0) {
    // TopPaths(1): the best leaf is slot 0 (leaves are kept sorted)
    const int best = 0;
    // LabelSeq(merge_repeated): walk to the root, dropping a label equal to the one emitted after it
    int* out = out_ids + (size_t)b * T;
    int n = 0, prev_label = -1;
    for (int node = slot_node[best]; node != 0; node = u.node_parent[node]) {
      const int label = u.node_label[node];
      if (!merge_repeated || label != prev_label) out[n++] = label;
      prev_label = label;
    }
    for (int i = 0; i < n / 2; ++i) { const int tmp = out[i]; out[i] = out[n - 1 - i]; out[n - 1 - i] = tmp; }
    for (int i = n; i < T; ++i) out[i] = 0;
    out_len[b] = n;
    out_neg_logprob[b] = -new_t[best];
  }
}

struct CbWs {
  int* node_parent; int* node_label; int* node_slot; unsigned long long* hkeys; int* hvals; int* node_count;
  int nmax, hcap; size_t hkeys_bytes, hvals_bytes, total;
};
CbWs cb_carve(void* base, int B, int T, int W) {
  CbWs w;
  w.nmax = 1 + W * T + W;
  int cap = 1;
  while (cap < 2 * w.nmax) cap <<= 1;
  w.hcap = cap;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* q = (void*)(p + off); off += align_up(bytes, 256); return q; };
  w.node_parent = (int*)take((size_t)B * w.nmax * 4);
  w.node_label = (int*)take((size_t)B * w.nmax * 4);
  w.node_slot = (int*)take((size_t)B * w.nmax * 4);
  w.hkeys_bytes = (size_t)B * cap * 8;
  w.hkeys = (unsigned long long*)take(w.hkeys_bytes);
  w.hvals_bytes = (size_t)B * cap * 4;
  w.hvals = (int*)take(w.hvals_bytes);
  w.node_count = (int*)take((size_t)B * 4);
  w.total = off;
  return w;
}

}  // namespace
}  // namespace nabu

using namespace nabu;

extern "C" size_t nabu_ctc_beam_workspace_bytes(int B, int T, int V, int beam_width) {
  (void)V;
  if (B <= 0 || T <= 0 || beam_width <= 0) return 0;
  return cb_carve(nullptr, B, T, beam_width).total;
}

extern "C" int nabu_ctc_beam_search(const float* logits, const int* logit_len, int B, int T, int V, int beam_width,
                                    int merge_repeated, int* out_ids, int* out_len, float* out_neg_logprob,
                                    void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NABU_REQUIRE(B > 0 && T > 0 && V > 1 && beam_width > 0, "ctc_beam_search: bad shape");
  const int W = beam_width;
  CbWs w = cb_carve(workspace, B, T, W);
  NABU_REQUIRE(ws_bytes >= w.total, "ctc_beam_search: workspace %zu < %zu bytes", ws_bytes, w.total);
  NABU_REQUIRE(W < 65535, "ctc_beam_search: beam_width=%d too large", W);
  const size_t smem = (size_t)15 * W * 4 + (size_t)V * 4 + (size_t)W * V * 2 + 16;
  NABU_REQUIRE(smem <= (size_t)max_smem_optin(), "ctc_beam_search: beam_width*classes too large (%d x %d)", W, V);
  NABU_CHECK_CUDA(cudaMemsetAsync(w.hkeys, 0xFF, w.hkeys_bytes, stream));
  NABU_CHECK_CUDA(cudaMemsetAsync(w.hvals, 0xFF, w.hvals_bytes, stream));
  if (smem > 48 * 1024)
    NABU_CHECK_CUDA(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KernelScope ks("ctc_beam", stream);
  ctc_beam_kernel<<<B, CB_THREADS, smem, stream>>>(logits, logit_len, T, V, W, merge_repeated, w.node_parent,
                                                   w.node_label, w.node_slot, w.hkeys, w.hvals, w.node_count, w.nmax,
                                                   w.hcap, out_ids, out_len, out_neg_logprob);
  NABU_CHECK_LAUNCH();
  return 0;
}

// The name SURVEY.md section 8b lists for this entry point.
extern "C" int nabu_ctc_prefix_beam(const float* logits, const int* logit_len, int B, int T, int V, int beam_width,
                                    int merge_repeated, int* out_ids, int* out_len, float* out_neg_logprob,
                                    void* workspace, size_t ws_bytes, void* stream) {
  return nabu_ctc_beam_search(logits, logit_len, B, T, V, beam_width, merge_repeated, out_ids, out_len, out_neg_logprob,
                              workspace, ws_bytes, stream);
}
