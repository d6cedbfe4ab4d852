// tcgen05 version of the cluster K-split forward recurrence (row a1; semantics in blstm.cu, partition in blstm_cl.cu).
//
// The phase stamps of the FFMA cluster kernel (NABU_REC_TRACE, cfg-3) show where a 23 us time step goes:
// 2.4 us waiting for the producers, 1.1 us until the first h block lands, 12 us in the FFMA loop (70 % of the FFMA
// issue rate), 4 us scatter + cluster barrier, 2.6 us pointwise, 1.2 us publish.  This kernel moves the partial
// product  P[128 b x 16*HS n] = h_{t-1}[128 b x 16*HS k] . Wl  onto the tensor cores at fp32-grade accuracy:
//   * every fp32 value x is carried as two fp16 numbers, hi = fp16(x) and lo = fp16((x - hi) * 2^11).  fp16 has the
//     same 11-bit significand as TF32, so hi + lo*2^-11 keeps 22 bits -- the precision of the 3xTF32 GEMMs -- at
//     4 bytes per value, and kind::f16 runs at twice the TF32 rate;
//   * D1 += Ah.Bh and D2 += Ah.Bl + Al.Bh are two TMEM accumulators (fp32), z = D1 + D2 * 2^-11.  Only Al.Bl
//     (2^-22 relative) is dropped.  h is in (-1, 1) and weights are O(1), so fp16's range is not an issue on this
//     path (values below 6e-5 lose relative but not absolute precision; |w| >= 65504 would overflow -- a documented
//     limit; NABU_REC_FWD=ffma selects the fp32 FFMA kernel);
//   * the producers write h_t to the L2 exchange buffer already split and already in the UMMA canonical layout
//     (K-major, SWIZZLE_128B, 64-column K blocks), so the consumer needs ONE 32 KB bulk copy per K block, no tensor
//     map and no converter warps; the weight block is split once at kernel start and stays in shared memory;
//   * the epilogue reads TMEM (one batch row x 4*HS columns per thread = exactly what one peer CTA owns) and
//     stores it straight into that peer's receive buffer over DSMEM, chunk-swizzled so the 128-byte-stride stores
//     do not bank-conflict.  The receive buffer is single (shared memory: 64 KB weights + 64 KB h + 64 KB receive),
//     so a second, split-phase cluster barrier ("receive buffer free") brackets the pointwise stage.
#include "cl_common.cuh"
#include "tc_common.cuh"
#include "blstm_cl.h"
#include <cuda_fp16.h>
#include <string.h>

namespace nabu {
namespace {

using namespace tc;

constexpr int TC_CLS = 4;
constexpr int A_TILE = 128 * 128;        // bytes of one [128 rows x 64 fp16] K-major tile

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// kind::f16, A and B fp16 K-major, fp32 accumulate
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of fp16 element (row, k < 64) inside a K-major SWIZZLE_128B tile (rows of 128 bytes)
__host__ __device__ inline uint32_t sw128_h(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)k & 7u) << 1);
}
__device__ __forceinline__ void split_h(float x, __half* hi, __half* lo) {
  const __half h = __float2half_rn(x);
  *hi = h;
  *lo = __float2half_rn((x - __half2float(h)) * 2048.f);
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float sigmoid_tc(float x) { return 1.f / (1.f + expf(-x)); }

// receive buffer: [4 src][128 rows][4*HS floats], 16-byte chunks XOR-swizzled per row
template <int HS>
__device__ __forceinline__ int rb_chunk(int row, int c4) {
  return HS == 8 ? (c4 ^ (row & 7)) : (c4 ^ ((row >> 1) & 3));
}

template <int HS>
__global__ void __launch_bounds__(CL_THREADS, 1)
blstm_rec_fwd_cluster_tc_kernel(const ClParams p) {
  constexpr int CLS = TC_CLS;
  constexpr int BT = 128;
  constexpr int NC = CLS * HS;             // hidden units per cluster
  constexpr int GC = 4 * NC;               // gate columns per cluster = MMA N
  constexpr int KS = 64 * HS / CLS;        // h rows per K-slice
  constexpr int KB = KS / 64;              // 64-wide K blocks per slice
  constexpr int CPS = 64 / CLS / CLS;      // producer clusters per K-slice
  constexpr int B_TILE = GC * 128;         // bytes of one [GC rows x 64 fp16] tile
  constexpr int RW = 4 * HS;               // floats per receive row
  constexpr int TCOLS = 2 * GC;            // TMEM columns: D1 | D2
  constexpr int PAIRS = BT * HS;
  constexpr int PP = PAIRS / CL_THREADS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Bs = sm;                                    // [KB][hi|lo][B_TILE]
  uint8_t* As = Bs + KB * 2 * B_TILE;                  // [KB][hi|lo][A_TILE]
  float* rbuf = reinterpret_cast<float*>(As + KB * 2 * A_TILE);   // [CLS][BT][RW]
  __shared__ __align__(8) uint64_t a_bar[2];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;

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
  uint8_t* hx = reinterpret_cast<uint8_t*>(p.xchg) + (size_t)dir * 2 * H * BT * 4;   // [2 parity][slice][KB][hi|lo][A_TILE]

  // resident weights, split: B[n][k] = Kh[r*KS + k][g*H + NC*q + d*HS + u],  n = d*4*HS + g*HS + u
  for (int i = tid; i < KS * GC; i += CL_THREADS) {
    const int u = i % HS, d = (i / HS) % CLS, g = (i / NC) % 4, k = i / GC;
    const float w = Kh[(size_t)(r * KS + k) * H4 + g * H + NC * q + d * HS + u];
    __half hi, lo;
    split_h(w, &hi, &lo);
    const int n = d * 4 * HS + g * HS + u;
    uint8_t* t = Bs + (size_t)(k / 64) * 2 * B_TILE + sw128_h(n, k % 64);
    *reinterpret_cast<__half*>(t) = hi;
    *reinterpret_cast<__half*>(t + B_TILE) = lo;
  }
  if (tid == 0) {
    mbar_init(smem_u32(&a_bar[0]), 1);
    mbar_init(smem_u32(&a_bar[1]), 1);
    mbar_init(smem_u32(&mma_bar), 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();                            // generic writes of Bs -> tensor-core (async proxy) reads
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  cluster_arrive();
  cluster_wait();                                      // peers' smem exists before anyone stores into it

  const uint32_t idesc = make_idesc_f16(128, GC);
  const uint32_t As_u = smem_u32(As), Bs_u = smem_u32(Bs);
  const int lg = warp & 3, ch = warp >> 2;             // TMEM lane group, column half
  const int row = lg * 32 + lane;

  for (int s = 0; s < p.T; ++s) {
    const uint8_t* hprev = hx + (size_t)((s + 1) & 1) * H * BT * 4;
    uint8_t* hnext = hx + (size_t)(s & 1) * H * BT * 4;
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
      if (b < p.B) {
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
      const unsigned par = (unsigned)(s - 1) & 1u;
      if (tid == 0) {
        const unsigned target = (unsigned)CLS * (unsigned)s;
        for (int c = 0; c < CPS; ++c)
          while (ld_acquire_gpu(cnt + r * CPS + c) < target) { }
        CL_STAMP(s, 1);
        __threadfence();
        fence_proxy_async_all();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          mbar_expect_tx(smem_u32(&a_bar[kb]), 2 * A_TILE);
          cb_bulk(As + (size_t)kb * 2 * A_TILE, hprev + (size_t)(r * KB + kb) * 2 * A_TILE, 2 * A_TILE, &a_bar[kb]);
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&a_bar[kb]), par);
          if (kb == 0) CL_STAMP(s, 2);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = make_desc(As_u + (kb * 2 + 0) * A_TILE + ks * 32, 16, 1024, 2);
            const uint64_t al = make_desc(As_u + (kb * 2 + 1) * A_TILE + ks * 32, 16, 1024, 2);
            const uint64_t bh = make_desc(Bs_u + (kb * 2 + 0) * B_TILE + ks * 32, 16, 1024, 2);
            const uint64_t bl = make_desc(Bs_u + (kb * 2 + 1) * B_TILE + ks * 32, 16, 1024, 2);
            const uint32_t acc = (kb | ks) != 0;
            umma_f16(tm, ah, bh, idesc, acc);
            umma_f16(tm + GC, ah, bl, idesc, acc);
            umma_f16(tm + GC, al, bh, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&mma_bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&mma_bar), par);
      tc_fence_after();
      CL_STAMP(s, 3);
      cluster_wait();                                  // every peer has finished reading its receive buffer (step s-1)
      // ---- TMEM -> peers' receive buffers (slot r) ----------------------------------------------------------
#pragma unroll
      for (int dd = 0; dd < 2; ++dd) {
        const int d = 2 * ch + dd;
        const uint32_t taddr = tm + ((uint32_t)(lg * 32) << 16) + (uint32_t)(d * RW);
        uint32_t v1[RW], v2[RW];
        if constexpr (HS == 8) {
          tmem_ld32(taddr, reinterpret_cast<uint32_t(&)[32]>(v1));
          tmem_ld32(taddr + GC, reinterpret_cast<uint32_t(&)[32]>(v2));
        } else {
          tmem_ld16(taddr, reinterpret_cast<uint32_t(&)[16]>(v1));
          tmem_ld16(taddr + GC, reinterpret_cast<uint32_t(&)[16]>(v2));
        }
        tmem_ld_wait();
        const uint32_t dst = map_to_rank(smem_u32(rbuf), (uint32_t)d) + (uint32_t)((r * BT + row) * RW) * 4u;
#pragma unroll
        for (int c4 = 0; c4 < RW / 4; ++c4) {
          float z[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            z[e] = fmaf(__uint_as_float(v2[c4 * 4 + e]), 1.f / 2048.f, __uint_as_float(v1[c4 * 4 + e]));
          st_cluster_v4(dst + (uint32_t)rb_chunk<HS>(row, c4) * 16u, z[0], z[1], z[2], z[3]);
        }
      }
      tc_fence_before();
      CL_STAMP(s, 4);
      cluster_arrive();
      cluster_wait();
      CL_STAMP(s, 5);
    }

    // ---- pointwise cell update for my HS units ----------------------------------------------------------
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const int pr = tid + k * CL_THREADS;
      const int jl = pr % HS, b = pr / HS;
      float hn = 0.f;
      if (b < p.B) {
        float z[4] = {gx[k][0], gx[k][1], gx[k][2], gx[k][3]};
        if (s > 0) {
#pragma unroll
          for (int src = 0; src < CLS; ++src)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int c = g * HS + jl;
              z[g] += rbuf[((size_t)src * BT + b) * RW + rb_chunk<HS>(b, c >> 2) * 4 + (c & 3)];
            }
        }
        const float ig = sigmoid_tc(z[0]);
        const float gg = tanhf(z[1]);
        const float fg = sigmoid_tc(z[2] + 1.0f);
        const float og = sigmoid_tc(z[3]);
        const float cn = cprev[k] * fg + ig * gg;
        hn = valid[k] ? tanhf(cn) * og : 0.f;
        const int t = tb[k];
        if (valid[k]) {
          float* gp = gates + ((size_t)b * p.T + t) * H4 + j0 + jl;
          __stcg(gp, ig); __stcg(gp + H, gg); __stcg(gp + 2 * H, fg); __stcg(gp + 3 * H, og);
          __stcg(cells + ((size_t)b * p.T + t) * H + j0 + jl, cn);
        }
        __stcg(p.y + ((size_t)b * p.yT + t) * 2 * H + dir * H + j0 + jl, hn);
      }
      // h_t, split, in the consumer's UMMA layout (rows b >= B stay zero from the host memset)
      if (b < p.B) {
        const int j = j0 + jl;
        __half hi, lo;
        split_h(hn, &hi, &lo);
        uint8_t* t = hnext + (size_t)((j / KS) * KB + (j % KS) / 64) * 2 * A_TILE + sw128_h(b, j % 64);
        __stcg(reinterpret_cast<unsigned short*>(t), __half_as_ushort(hi));
        __stcg(reinterpret_cast<unsigned short*>(t + A_TILE), __half_as_ushort(lo));
      }
    }
    CL_STAMP(s, 6);
    fence_proxy_async_all();
    __threadfence();
    CL_STAMP(s, 7);
    if (s + 1 < p.T) cluster_arrive();                 // my receive buffer is free for step s+1
    __syncthreads();
    CL_STAMP(s, 8);
    if (tid == 0) red_release_gpu_add(cnt + q, 1u);
    CL_STAMP(s, 9);
  }
  tc_fence_before();
  cluster_arrive();
  cluster_wait();                                      // nobody exits while a peer may still store into it
  if (warp == 0) tmem_dealloc(tm, TCOLS);
}

template <int HS>
int launch_fwd_tc(const ClParams& p, cudaStream_t stream, bool* launched) {
  constexpr int CLS = TC_CLS, GC = 16 * HS, KB = (64 * HS / CLS) / 64;
  const size_t smem = 1024 + (size_t)KB * 2 * GC * 128 + (size_t)KB * 2 * A_TILE + (size_t)CLS * 128 * 4 * HS * sizeof(float);
  auto* fn = blstm_rec_fwd_cluster_tc_kernel<HS>;
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
    fprintf(stderr, "[nabu] fwd tcgen05 cluster kernel HS=%d: smem %zu B, max active clusters %d (%s), need %d\n", HS, smem,
            nclusters, cudaGetErrorString(oe), (int)cfg.gridDim.x / CLS);
  if (oe != cudaSuccess || nclusters * CLS < (int)cfg.gridDim.x) {
    cudaGetLastError();
    return 0;
  }
  KernelScope ks("blstm_rec_fwd_cluster_tc", stream);
  ClParams pt = p;
  pt.trace = trace_buffer();
  NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt));
  trace_dump("fwd_tc", pt.trace, stream);
  *launched = true;
  return 0;
}

}  // namespace

bool blstm_fwd_cluster_tc_eligible(int B, int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NABU_REC_FWD");
    enabled = (e && (strcmp(e, "flat") == 0 || strcmp(e, "ffma") == 0)) ? 0 : 1;
  }
  if (!enabled) return false;
  return B <= 128 && B > 0 && (H == 256 || H == 512);
}

int blstm_rec_fwd_cluster_tc(const float* const kernel[2], float* const gates[2], float* const cells[2], float* y,
                             float* xchg, unsigned* counters, const int* len, int B, int T, int yT, int D, int H,
                             cudaStream_t stream, bool* launched) {
  ClParams p = {};
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.y = y; p.xchg = xchg; p.counters = counters; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  return H == 512 ? launch_fwd_tc<8>(p, stream, launched) : launch_fwd_tc<4>(p, stream, launched);
}

}  // namespace nabu
