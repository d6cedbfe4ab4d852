// tcgen05 version of the persistent BLSTM forward recurrence (row a1; see blstm.cu for the semantics).
//
// The FFMA kernel in blstm.cu spends 28 us per time step at cfg-3 on a [128 x 512] x [512 x 32] product
// per CTA (issue/latency bound, 42 % issue-slot utilisation).  Here the same product runs on the tensor
// core with the 3xTF32 split of gemm_tc.cu, so it keeps fp32-grade accuracy:
//   * the CTA's slice of the recurrent matrix (HS hidden units = 4*HS gate columns) is split ONCE into
//     hi/lo TF32 and kept resident in shared memory as K-major SWIZZLE_128B UMMA tiles (128 KB at H=512);
//   * per step, h_{t-1} [128 x H] is streamed from the L2-resident exchange buffer by TMA in 32-wide
//     K chunks through a 3-stage ring; four converter warps split each landed chunk into hi/lo in place;
//   * one thread issues 3 x tcgen05.mma.kind::tf32 (M=128, N=4*HS, K=8) per k-step into a [128 x 4*HS] fp32
//     accumulator in TMEM; the same four warps then read their batch row back with tcgen05.ld and do the
//     LSTM pointwise update, the length masking and the stores with 128-bit accesses;
//   * CTAs of one direction hand h_t to each other through the exchange buffer + release/acquire counter
//     exactly like the FFMA kernel (generic-proxy stores -> fence.proxy.async -> TMA loads on the reader).
#include "common.cuh"
#include "tc_common.cuh"
#include "blstm_tc.h"

namespace nabu {
namespace {

using namespace tc;

constexpr int RT_THREADS = 192;
constexpr int A_CHUNK_BYTES = 128 * 32 * 4;        // [128 batch rows][32 k] fp32 = 16 KB

struct RecTcParams {
  const float* kernel[2];
  float* gates[2];
  float* cells[2];
  float* y;
  float* hrow;              // [2 dir][2 parity][128][H]
  unsigned* counters;
  const int* len;
  int B, T, yT, D, H, nsl, NS;
};

template <int HS>
__global__ void __launch_bounds__(RT_THREADS, 1)
blstm_rec_fwd_tc_kernel(const __grid_constant__ CUtensorMap hmap, const RecTcParams p) {
  constexpr int N = 4 * HS;                        // gate columns of this CTA
  constexpr int WT_BYTES = N * 128;                // one 32-k weight tile [N rows][32 k]
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int H = p.H, H4 = 4 * p.H, NC = p.H / 32, NS = p.NS;
  const uint32_t w_hi = base, w_lo = base + NC * WT_BYTES;
  const uint32_t ring = base + 2 * NC * WT_BYTES;
  const uint32_t bars = ring + NS * 2 * A_CHUNK_BYTES;
  uint8_t* ring_ptr = base_ptr + 2 * NC * WT_BYTES;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_conv = [&](int s) { return bars + 8u * (NS + s); };
  auto bar_empty = [&](int s) { return bars + 8u * (2 * NS + s); };
  const uint32_t bar_tfull = bars + 8u * (3 * NS), bar_tempty = bars + 8u * (3 * NS + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + 2 * NC * WT_BYTES + NS * 2 * A_CHUNK_BYTES + 8 * (3 * NS + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.x / p.nsl, slice = blockIdx.x % p.nsl;
  const int j0 = slice * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  unsigned* counter = p.counters + dir;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_conv(s), 128); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 32);
  // resident weights: row n = gate*HS + jl  <->  Kh[k][gate*H + j0 + jl], split into hi / lo TF32
  for (int i = threadIdx.x; i < H * N; i += RT_THREADS) {
    const int n = i % N, k = i / N;
    const float w = Kh[(size_t)k * H4 + (n / HS) * H + j0 + (n % HS)];
    const float hi = rn_tf32(w);
    const uint32_t off = (uint32_t)(k >> 5) * WT_BYTES + kmajor_sw128_offset(n, k & 31);
    *reinterpret_cast<float*>(base_ptr + off) = hi;
    *reinterpret_cast<float*>(base_ptr + NC * WT_BYTES + off) = w - hi;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int g = 0;
      for (int s = 1; s < p.T; ++s) {
        const unsigned target = (unsigned)p.nsl * (unsigned)s;
        while (ld_acquire_gpu(counter) < target) { }
        fence_proxy_async_all();                      // other CTAs' generic-proxy stores -> our async-proxy loads
        const int slab = dir * 2 + ((s + 1) & 1);     // h_{s-1} lives in parity (s-1)&1
        for (int c = 0; c < NC; ++c, ++g) {
          const int st = g % NS;
          mbar_wait(bar_empty(st), ((g / NS) & 1) ^ 1);
          mbar_expect_tx(bar_full(st), A_CHUNK_BYTES);
          tma_load_3d(ring + st * 2 * A_CHUNK_BYTES, &hmap, bar_full(st), 32 * c, 0, slab);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
      int g = 0;
      for (int s = 1; s < p.T; ++s) {
        mbar_wait(bar_tempty, (s - 1) & 1);           // epilogue of step s-1 has drained the accumulator
        tc_fence_after();
        for (int c = 0; c < NC; ++c, ++g) {
          const int st = g % NS;
          mbar_wait(bar_conv(st), (g / NS) & 1);
          tc_fence_after();
          const uint32_t a_hi = ring + st * 2 * A_CHUNK_BYTES, a_lo = a_hi + A_CHUNK_BYTES;
          const uint32_t b_hi = w_hi + c * WT_BYTES, b_lo = w_lo + c * WT_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dah = make_desc(a_hi + 32 * k, 16, 1024, 2), dal = make_desc(a_lo + 32 * k, 16, 1024, 2);
            const uint64_t dbh = make_desc(b_hi + 32 * k, 16, 1024, 2), dbl = make_desc(b_lo + 32 * k, 16, 1024, 2);
            umma_tf32(tmem_d, dal, dbh, idesc, (c > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
          }
          umma_commit(bar_empty(st));
        }
        umma_commit(bar_tfull);
      }
    }
  } else {
    // ===== converters + pointwise epilogue: thread <-> batch row =====
    const int ct = threadIdx.x - 64;
    const int q = warp & 3;
    const int b = 32 * q + lane;
    const bool row_ok = b < p.B;
    const int L = row_ok ? p.len[b] : 0;
    float* gates = p.gates[dir];
    float* cells = p.cells[dir];
    int g = 0;
    for (int s = 0; s < p.T; ++s) {
      // operands of the pointwise update, issued before the matmul so their latency is hidden
      const bool valid = row_ok && s < L;
      const int t = valid ? (dir ? L - 1 - s : s) : s;
      float gx[4][HS], cprev[HS];
#pragma unroll
      for (int j = 0; j < HS; ++j) { cprev[j] = 0.f; gx[0][j] = gx[1][j] = gx[2][j] = gx[3][j] = 0.f; }
      if (valid) {
        const float* gp = gates + ((size_t)b * p.T + t) * H4 + j0;
#pragma unroll
        for (int gt = 0; gt < 4; ++gt)
#pragma unroll
          for (int v = 0; v < HS / 4; ++v) {
            const float4 x4 = __ldcg(reinterpret_cast<const float4*>(gp + gt * H) + v);
            gx[gt][4 * v] = x4.x; gx[gt][4 * v + 1] = x4.y; gx[gt][4 * v + 2] = x4.z; gx[gt][4 * v + 3] = x4.w;
          }
        if (s > 0) {
          const float* cp = cells + ((size_t)b * p.T + (dir ? t + 1 : t - 1)) * H + j0;
#pragma unroll
          for (int v = 0; v < HS / 4; ++v) {
            const float4 x4 = __ldcg(reinterpret_cast<const float4*>(cp) + v);
            cprev[4 * v] = x4.x; cprev[4 * v + 1] = x4.y; cprev[4 * v + 2] = x4.z; cprev[4 * v + 3] = x4.w;
          }
        }
      }
      uint32_t r[N];
      if (s > 0) {
        for (int c = 0; c < NC; ++c, ++g) {
          const int st = g % NS;
          mbar_wait(bar_full(st), (g / NS) & 1);
          split_tile(ring_ptr + st * 2 * A_CHUNK_BYTES, A_CHUNK_BYTES, ct, 128);
          fence_proxy_async_smem();
          mbar_arrive(bar_conv(st));
        }
        mbar_wait(bar_tfull, (s - 1) & 1);
        tc_fence_after();
        if constexpr (N == 32) tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16), r);
        else tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = 0u;
      }
      tc_fence_before();
      mbar_arrive(bar_tempty);
      // LSTM cell update for the HS hidden units of this row
      float hn[HS];
      if (row_ok) {
        float ig[HS], gg[HS], fg[HS], og[HS], cn[HS];
#pragma unroll
        for (int j = 0; j < HS; ++j) {
          ig[j] = sigmoid_acc(__uint_as_float(r[j]) + gx[0][j]);
          gg[j] = tanhf(__uint_as_float(r[HS + j]) + gx[1][j]);
          fg[j] = sigmoid_acc(__uint_as_float(r[2 * HS + j]) + gx[2][j] + 1.0f);
          og[j] = sigmoid_acc(__uint_as_float(r[3 * HS + j]) + gx[3][j]);
          cn[j] = cprev[j] * fg[j] + ig[j] * gg[j];
          hn[j] = valid ? tanhf(cn[j]) * og[j] : 0.f;
        }
        if (valid) {
          float* gp = gates + ((size_t)b * p.T + t) * H4 + j0;
          float* cp = cells + ((size_t)b * p.T + t) * H + j0;
#pragma unroll
          for (int v = 0; v < HS / 4; ++v) {
            __stcg(reinterpret_cast<float4*>(gp) + v, make_float4(ig[4 * v], ig[4 * v + 1], ig[4 * v + 2], ig[4 * v + 3]));
            __stcg(reinterpret_cast<float4*>(gp + H) + v, make_float4(gg[4 * v], gg[4 * v + 1], gg[4 * v + 2], gg[4 * v + 3]));
            __stcg(reinterpret_cast<float4*>(gp + 2 * H) + v, make_float4(fg[4 * v], fg[4 * v + 1], fg[4 * v + 2], fg[4 * v + 3]));
            __stcg(reinterpret_cast<float4*>(gp + 3 * H) + v, make_float4(og[4 * v], og[4 * v + 1], og[4 * v + 2], og[4 * v + 3]));
            __stcg(reinterpret_cast<float4*>(cp) + v, make_float4(cn[4 * v], cn[4 * v + 1], cn[4 * v + 2], cn[4 * v + 3]));
          }
        }
        float* yp = p.y + ((size_t)b * p.yT + t) * 2 * H + dir * H + j0;
        float* hp = p.hrow + ((size_t)(dir * 2 + (s & 1)) * 128 + b) * H + j0;
#pragma unroll
        for (int v = 0; v < HS / 4; ++v) {
          const float4 h4 = make_float4(hn[4 * v], hn[4 * v + 1], hn[4 * v + 2], hn[4 * v + 3]);
          __stcg(reinterpret_cast<float4*>(yp) + v, h4);
          __stcg(reinterpret_cast<float4*>(hp) + v, h4);
        }
      }
      // publish h_s: our generic-proxy stores must be visible to the other CTAs' TMA (async proxy) loads
      fence_proxy_async_all();
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (ct == 0) red_release_gpu_add(counter, 1u);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, 32);
}

}  // namespace

bool blstm_tc_plan(int B, int H, BlstmTcPlan* pl) {
  if (B > 128 || H % 32 != 0 || H < 32) return false;
  const int sms = num_sms();
  const size_t cap = (size_t)max_smem_optin();
  for (int hs = 4; hs <= 8; hs += 4) {
    if (H % hs) continue;
    const int nsl = H / hs;
    if (2 * nsl > sms) continue;
    const size_t weights = (size_t)2 * (H / 32) * (4 * hs) * 128;
    for (int ns = 4; ns >= 2; --ns) {
      const size_t need = weights + (size_t)ns * 2 * A_CHUNK_BYTES + 1024 + 256;
      if (need <= cap) {
        pl->hs = hs; pl->nsl = nsl; pl->ns = ns; pl->smem = need;
        return true;
      }
    }
  }
  return false;
}

int blstm_rec_fwd_tc(const BlstmTcPlan& pl, const float* const kernel[2], float* const gates[2], float* const cells[2],
                     float* y, float* hrow, unsigned* counters, const int* len, int B, int T, int yT, int D, int H,
                     cudaStream_t stream) {
  CUtensorMap hmap;
  const int r = encode_map_3d(&hmap, hrow, (uint64_t)H, 128, 4, (uint64_t)H, (uint64_t)128 * H, 32, 128, false);
  NABU_REQUIRE(r == 0, "blstm_tc: cuTensorMapEncodeTiled failed (%d)", r);
  RecTcParams p = {};
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.y = y; p.hrow = hrow; p.counters = counters; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H; p.nsl = pl.nsl; p.NS = pl.ns;
  const void* fn = pl.hs == 8 ? (const void*)blstm_rec_fwd_tc_kernel<8> : (const void*)blstm_rec_fwd_tc_kernel<4>;
  NABU_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  void* args[] = {(void*)&hmap, (void*)&p};
  KernelScope ks("blstm_rec_fwd_tc", stream);
  NABU_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(2 * pl.nsl), dim3(RT_THREADS), args, pl.smem, stream));
  return 0;
}

}  // namespace nabu
