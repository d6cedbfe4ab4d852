// CTC prefix beam search (replaces neuralnetworks/decoders/ctc_decoder.py:57
// -> tf.nn.ctc_beam_search_decoder(beam_width=100, top_paths=1, merge_repeated=True), which TF-1.8 only
// runs on the CPU: tensorflow/core/util/ctc/ctc_beam_search.h, restated in SURVEY appendix B8).
//
// One CTA per utterance walks its frames.  TF's Step() inserts candidates one by one into a bounded
// top-N heap; because a child's score never exceeds its parent's, and the heap's bottom only rises,
// the leaf set after a frame is exactly the `beam_width` best of
//     { updated existing leaves }  U  { (leaf b, label c) children that are not currently leaves },
// so the frame is evaluated in parallel: every thread scores candidates, a bitonic sort on
// (score desc, index asc) picks the survivors.  The prefix tree lives in global memory as
// (parent, label) arrays plus an open-addressing hash (parent, label) -> node, so a prefix that drops
// out and later re-enters is the SAME node, as in TF (its children keep their parent link).
#include "common.cuh"
#include "nabu_b200.h"
#include <math_constants.h>

namespace nabu {
namespace {

constexpr int CB_THREADS = 256;
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

// dynamic smem layout (W = beam width, NS = sort size, V classes)
//   float old_t/old_b/old_l/new_t/new_b/new_l [W]; int slot_node[W]; int sel_node[W]; float sel_val[3][W]
//   unsigned long long sortbuf[NS]; float x[V]; unsigned char active_child[W*V]
__global__ void __launch_bounds__(CB_THREADS) ctc_beam_kernel(
    const float* logits, const int* logit_len, int T, int V, int W, int NS, int merge_repeated,
    int* g_node_parent, int* g_node_label, int* g_node_slot, unsigned long long* g_hkeys, int* g_hvals,
    int* g_node_count, int nmax, int hcap, int* out_ids, int* out_len, float* out_neg_logprob) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int blank = V - 1;
  const float NINF = -CUDART_INF_F;
  unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(smraw);
  float* old_t = reinterpret_cast<float*>(sortbuf + NS);
  float* old_b = old_t + W; float* old_l = old_b + W;
  float* new_t = old_l + W; float* new_b = new_t + W; float* new_l = new_b + W;
  float* sv_t = new_l + W; float* sv_b = sv_t + W; float* sv_l = sv_b + W;
  int* slot_node = reinterpret_cast<int*>(sv_l + W);
  int* sel_node = slot_node + W;
  float* x = reinterpret_cast<float*>(sel_node + W);
  unsigned char* active_child = reinterpret_cast<unsigned char*>(x + V);
  __shared__ int s_nleaf;
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
    const int nleaf = s_nleaf;
    // x = logits[t] - max   (Step(): "remove the max for stability")
    if (tid < 32) {
      float m = NINF;
      for (int k = tid; k < V; k += 32) m = fmaxf(m, lg[(size_t)t * V + k]);
      m = warp_max(m);
      if (tid == 0) s_max = m;
    }
    for (int i = tid; i < W * V; i += CB_THREADS) active_child[i] = 0;
    __syncthreads();
    for (int k = tid; k < V; k += CB_THREADS) x[k] = lg[(size_t)t * V + k] - s_max;
    // oldp = newp
    for (int i = tid; i < nleaf; i += CB_THREADS) { old_t[i] = new_t[i]; old_b[i] = new_b[i]; old_l[i] = new_l[i]; }
    __syncthreads();
    // existing leaves: extend by blank / repeat of their own label; mark the (parent slot, label) pairs that
    // are leaves already so they are not spawned again
    for (int i = tid; i < nleaf; i += CB_THREADS) {
      const int node = slot_node[i];
      float nl = old_l[i];
      if (node != 0) {
        const int label = u.node_label[node];
        const int par = u.node_parent[node];
        const int ps = u.node_slot[par];
        if (ps >= 0) {
          const int plabel = u.node_label[par];
          const float prev = (label == plabel) ? old_b[ps] : old_t[ps];
          nl = lse_tf(nl, prev);
          active_child[ps * V + label] = 1;
        }
        nl = nl + x[label];
      }
      const float nb = old_t[i] + x[blank];
      new_l[i] = nl; new_b[i] = nb; new_t[i] = lse_tf(nb, nl);
    }
    __syncthreads();
    // candidates -> sort keys.  index space: [0, W) existing leaves, W + i*V + c child c of leaf i.
    for (int i = tid; i < NS; i += CB_THREADS) {
      float val = NINF;
      if (i < W) {
        if (i < nleaf) val = new_t[i];
      } else {
        const int j = i - W, li = j / V, c = j % V;
        if (li < nleaf && c != blank && !active_child[li * V + c]) {
          const int node = slot_node[li];
          const int label = (node == 0) ? -1 : u.node_label[node];
          const float prev = (c == label) ? old_b[li] : old_t[li];
          val = x[c] + prev;
        }
      }
      sortbuf[i] = ((unsigned long long)f2ord(val) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    }
    __syncthreads();
    // bitonic sort, descending
    for (int k = 2; k <= NS; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < NS; i += CB_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = sortbuf[i], c2 = sortbuf[ixj];
            const bool desc = ((i & k) == 0);
            if (desc ? (a < c2) : (a > c2)) { sortbuf[i] = c2; sortbuf[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    // survivors: the first W entries with a finite score
    for (int i = tid; i < nleaf; i += CB_THREADS) u.node_slot[slot_node[i]] = -1;
    __syncthreads();
    const unsigned ninf_ord = f2ord(NINF);
    for (int k = tid; k < W; k += CB_THREADS) {
      const unsigned long long key = sortbuf[k];
      const unsigned ord = (unsigned)(key >> 32);
      int node = -1;
      float vt = NINF, vb = NINF, vl = NINF;
      if (ord != ninf_ord) {
        const int idx = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        if (idx < W) {
          node = slot_node[idx]; vt = new_t[idx]; vb = new_b[idx]; vl = new_l[idx];
        } else {
          const int j = idx - W, li = j / V, c = j % V;
          const int pnode = slot_node[li];
          const int plabel = (pnode == 0) ? -1 : u.node_label[pnode];
          const float prev = (c == plabel) ? old_b[li] : old_t[li];
          vl = x[c] + prev; vt = vl; vb = NINF;
          node = hash_find_or_insert(u, pnode, c, V, node_count);
        }
      }
      sel_node[k] = node; sv_t[k] = vt; sv_b[k] = vb; sv_l[k] = vl;
    }
    __syncthreads();
    if (tid == 0) {
      int n = 0;
      while (n < W && sel_node[n] >= 0) ++n;
      s_nleaf = n;
    }
    for (int k = tid; k < W; k += CB_THREADS) {
      const int node = sel_node[k];
      if (node >= 0) {
        slot_node[k] = node; new_t[k] = sv_t[k]; new_b[k] = sv_b[k]; new_l[k] = sv_l[k];
        u.node_slot[node] = k;
      }
    }
    __threadfence_block();
    __syncthreads();
  }

  if (tid == 0) {
    const int nleaf = s_nleaf;
    int best = 0;
    for (int i = 1; i < nleaf; ++i)
      if (new_t[i] > new_t[best]) best = i;
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
  int NS = 1;
  while (NS < W + W * V) NS <<= 1;
  const size_t smem = (size_t)NS * 8 + (size_t)9 * W * 4 + (size_t)2 * W * 4 + (size_t)V * 4 + (size_t)W * V + 16;
  NABU_REQUIRE(smem <= (size_t)max_smem_optin(), "ctc_beam_search: beam_width*classes too large (%d x %d)", W, V);
  NABU_CHECK_CUDA(cudaMemsetAsync(w.hkeys, 0xFF, w.hkeys_bytes, stream));
  NABU_CHECK_CUDA(cudaMemsetAsync(w.hvals, 0xFF, w.hvals_bytes, stream));
  if (smem > 48 * 1024)
    NABU_CHECK_CUDA(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KernelScope ks("ctc_beam", stream);
  ctc_beam_kernel<<<B, CB_THREADS, smem, stream>>>(logits, logit_len, T, V, W, NS, merge_repeated, w.node_parent,
                                                   w.node_label, w.node_slot, w.hkeys, w.hvals, w.node_count, w.nmax,
                                                   w.hcap, out_ids, out_len, out_neg_logprob);
  NABU_CHECK_LAUNCH();
  return 0;
}
