// tcgen05 / TMEM / TMA dense contraction for sm_100a with fp32-grade accuracy ("3xTF32").
//
// Every big GEMM of the hot path (input projection X.Kx, dKx = X^T.dZ, dKh = Hprev^T.dZ,
// dX = dZ.Kx^T) must hold the 1e-4 fp32 parity bar, which a single TF32 or BF16 tensor-core pass
// does not.  So each fp32 operand is split on the fly into hi = rn_tf32(a) and
// lo = a - hi (exact in fp32) and the tensor core accumulates  hi.hi + hi.lo + lo.hi  in fp32 TMEM:
// the dropped lo.lo term is ~2^-22 relative.  Cost: 3 TF32 MMAs per k-step, still ~6x the fp32 FFMA
// roofline of the SIMT kernel in sgemm.cu.
//
// Kernel anatomy (one CTA = one 128 x 256 output tile, 192 threads, 2-stage ring, BK = 32):
//   warp 0     TMA producer: cp.async.bulk.tensor (SWIZZLE_128B boxes, OOB rows/cols zero-filled, so
//              M/N/K tails and the per-utterance row segmentation of dKh need no special cases)
//   warps 2-5  converters: rewrite the landed fp32 tile in place as `hi` and write `lo` next to it
//              (generic proxy -> fence.proxy.async -> mbarrier), later the epilogue warps
//   warp 1     MMA issuer: one thread issues tcgen05.mma.kind::tf32 (M=128, N=256, K=8) x 4 k-steps x 3
//              products per stage, tcgen05.commit frees the stage / signals the epilogue
//   epilogue   tcgen05.ld 32x32b -> warp-private smem transpose -> alpha, beta, bias -> coalesced 128-bit stores
// Operands may be K-major (row = M/N index, K contiguous) or MN-major (row = K index): both are legal
// UMMA layouts for TF32, selected in the instruction descriptor, so NN / NT / TN need no transposes.
#include "common.cuh"
#include "gemm.h"
#include "tc_common.cuh"

namespace nabu {
namespace {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 32, TC_STAGES = 2;
constexpr int TC_THREADS = 192;
constexpr int A_TILE_BYTES = TC_BM * TC_BK * 4;          // 16 KB
constexpr int B_TILE_BYTES = TC_BN * TC_BK * 4;          // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // hi+lo of both = 96 KB
constexpr int TC_TPAD = 36;                              // padded row of the epilogue transpose buffer (floats)
constexpr int TC_SMEM_BYTES = TC_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * 32 * TC_TPAD * 4;

struct TcArgs {
  int M, N;
  int kblocks;            // total k-blocks of 32
  int kps;                // k-blocks per row segment (MN-major operands); == kblocks when unsegmented
  int kb_per_split;       // k-blocks per blockIdx.z
  int a_mn_major, b_mn_major;
  float alpha, beta;
  const float* bias;
  float* C; int ldc;
  float* part;            // split-K partials [splits][M][N] or nullptr
};

using namespace tc;

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B needs 1024-B alignment
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + TC_STAGES * STAGE_BYTES;
  // barrier slots (8 B each): full[s], conv[s], empty[s], tmem_full ; then the TMEM base address
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_conv = [&](int s) { return bars + 8u * (TC_STAGES + s); };
  auto bar_empty = [&](int s) { return bars + 8u * (2 * TC_STAGES + s); };
  const uint32_t bar_tmem = bars + 8u * (3 * TC_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + TC_STAGES * STAGE_BYTES + 8 * (3 * TC_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int kb_begin = blockIdx.z * g.kb_per_split;
  const int kb_end = min(g.kblocks, kb_begin + g.kb_per_split);
  const int nkb = max(0, kb_end - kb_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_conv(s), 128);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_tmem, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256u);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb_begin + i;
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(bar_empty(s), ph ^ 1);
        mbar_expect_tx(bar_full(s), A_TILE_BYTES + B_TILE_BYTES);
        const uint32_t sa = base + s * STAGE_BYTES;
        const uint32_t sb = sa + 2 * A_TILE_BYTES;
        const int seg = kb / g.kps, kk = (kb % g.kps) * TC_BK;
        if (g.a_mn_major) {
          for (int j = 0; j < TC_BM / 32; ++j) tma_load_3d(sa + j * 4096, &mapA, bar_full(s), m0 + 32 * j, kk, seg);
        } else {
          tma_load_3d(sa, &mapA, bar_full(s), kb * TC_BK, m0, 0);
        }
        if (g.b_mn_major) {
          for (int j = 0; j < TC_BN / 32; ++j) tma_load_3d(sb + j * 4096, &mapB, bar_full(s), n0 + 32 * j, kk, seg);
        } else {
          tma_load_3d(sb, &mapB, bar_full(s), kb * TC_BK, n0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BM, TC_BN, g.a_mn_major, g.b_mn_major);
      // K-major (SWIZZLE_128B): 8-row groups 1024 B apart (SBO), k-step = +32 B inside the swizzled row.
      // MN-major: TF32 only supports SWIZZLE_128B_BASE32B there (Swizzle<2,5,2>: 32-B chunks XOR row%4):
      // 32-element column blocks 4096 B apart (LBO), 4-row k-groups 512 B apart (SBO), k-step (8 rows) = 1024 B.
      const uint32_t a_lbo = g.a_mn_major ? 4096u : 16u, b_lbo = g.b_mn_major ? 4096u : 16u;
      const uint32_t a_sbo = g.a_mn_major ? 512u : 1024u, b_sbo = g.b_mn_major ? 512u : 1024u;
      const uint32_t a_lt = g.a_mn_major ? 1u : 2u, b_lt = g.b_mn_major ? 1u : 2u;
      const uint32_t a_kstep = g.a_mn_major ? 1024u : 32u, b_kstep = g.b_mn_major ? 1024u : 32u;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(bar_conv(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * STAGE_BYTES;
        const uint32_t sb = sa + 2 * A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t a_hi = make_desc(sa + k * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t a_lo = make_desc(sa + A_TILE_BYTES + k * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t b_hi = make_desc(sb + k * b_kstep, b_lbo, b_sbo, b_lt);
          const uint64_t b_lo = make_desc(sb + B_TILE_BYTES + k * b_kstep, b_lbo, b_sbo, b_lt);
          umma_tf32(tmem_d, a_lo, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);   // small terms first
          umma_tf32(tmem_d, a_hi, b_lo, idesc, 1u);
          umma_tf32(tmem_d, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(bar_empty(s));                 // frees the stage once these MMAs have read it
      }
      umma_commit(bar_tmem);                       // accumulator complete
    }
  } else {
    // ===== converters (then epilogue) =====
    const int ct = threadIdx.x - 64;               // 0..127
    for (int i = 0; i < nkb; ++i) {
      const int s = i % TC_STAGES;
      const uint32_t ph = (i / TC_STAGES) & 1;
      mbar_wait(bar_full(s), ph);
      uint8_t* sa = base_ptr + s * STAGE_BYTES;
      uint8_t* sb = sa + 2 * A_TILE_BYTES;
split_tile(sa, A_TILE_BYTES, ct, 128);
      split_tile(sb, B_TILE_BYTES, ct, 128);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
      mbar_arrive(bar_conv(s));
    }
    // ===== epilogue: TMEM -> registers -> (warp-private smem transpose) -> coalesced global stores =====
    // (see gemm_h2.cu: storing the row each thread gets from tcgen05.ld directly touches 32 rows per instruction)
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const bool split = g.part != nullptr;
    float* Cout = split ? g.part + (size_t)blockIdx.z * g.M * g.N : g.C;
    const int ldc = split ? g.N : g.ldc;
    const bool vec = ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0) && (ldc % 4 == 0);
    float* tbuf = reinterpret_cast<float*>(base_ptr + TC_STAGES * STAGE_BYTES + 256) + (warp - 2) * (32 * TC_TPAD);
    const int rl = lane >> 3, c4 = (lane & 7) * 4;
    if (nkb > 0) {
      mbar_wait(bar_tmem, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    for (int c0 = 0; c0 < TC_BN; c0 += 32) {
      if (n0 + c0 >= g.N) break;
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4)
        *reinterpret_cast<float4*>(tbuf + lane * TC_TPAD + j4) =
            make_float4(__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3]));
      __syncwarp();
      const int n = n0 + c0 + c4;
      float cs = split ? 1.f : g.alpha, bs[4] = {0.f, 0.f, 0.f, 0.f};
      if (!split && g.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < g.N) bs[j] = __ldg(g.bias + n + j);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + rl;
        const int mm = m0 + 32 * q + rr;
        const float4 t4 = *reinterpret_cast<const float4*>(tbuf + rr * TC_TPAD + c4);
        float v[4] = {fmaf(t4.x, cs, bs[0]), fmaf(t4.y, cs, bs[1]), fmaf(t4.z, cs, bs[2]), fmaf(t4.w, cs, bs[3])};
        if (mm < g.M && n < g.N) {
          float* cp = Cout + (size_t)mm * ldc + n;
          const int nv = min(4, g.N - n);
          if (nv == 4 && vec) {
            if (!split && g.beta != 0.f) {
              const float4 o = *reinterpret_cast<const float4*>(cp);
              v[0] += g.beta * o.x; v[1] += g.beta * o.y; v[2] += g.beta * o.z; v[3] += g.beta * o.w;
            }
            *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
            for (int j = 0; j < nv; ++j) cp[j] = (!split && g.beta != 0.f) ? v[j] + g.beta * cp[j] : v[j];
          }
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------
int make_map(CUtensorMap* map, const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
             uint64_t stride2, uint32_t box0, uint32_t box1, bool mn_major) {
  const int r = encode_map_3d(map, ptr, d0, d1, d2, stride1, stride2, box0, box1, mn_major);
  NABU_REQUIRE(r == 0, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", r);
  return 0;
}

}  // namespace

bool gemm_tc_eligible(GemmMode mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb) {
  (void)mode;
  if (M < 1 || N < 1 || K < 1) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if (lda % 4 || ldb % 4) return false;
  // tiny problems are not worth a 128x256 tile
  if ((long)M * N < 128L * 128L) return false;
  return true;
}

int gemm_tc(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
            float beta, float* C, int ldc, const float* bias, const GemmSeg* segp, float* workspace,
            size_t ws_bytes, cudaStream_t stream) {
  NABU_REQUIRE(!(segp && mode != GEMM_TN), "gemm_tc: row segmentation only in TN mode");
  CUtensorMap mapA, mapB;
  TcArgs g = {};
  g.M = M; g.N = N; g.alpha = alpha; g.beta = beta; g.bias = bias; g.C = C; g.ldc = ldc;
  if (mode == GEMM_TN) {
    const int seg = segp ? segp->seg : K;
    const int nseg = segp ? K / segp->seg : 1;
    NABU_REQUIRE(!segp || K % segp->seg == 0, "gemm_tc: K must be a multiple of the segment length");
    const uint64_t sA = segp ? (uint64_t)segp->segA : (uint64_t)K, sB = segp ? (uint64_t)segp->segB : (uint64_t)K;
    const float* Ap = A + (segp ? (size_t)segp->offA * lda : 0);
    const float* Bp = B + (segp ? (size_t)segp->offB * ldb : 0);
    if (int e = make_map(&mapA, Ap, M, seg, nseg, lda, sA * lda, 32, 32, true)) return e;
    if (int e = make_map(&mapB, Bp, N, seg, nseg, ldb, sB * ldb, 32, 32, true)) return e;
    g.a_mn_major = 1; g.b_mn_major = 1;
    g.kps = ceil_div(seg, TC_BK);
    g.kblocks = g.kps * nseg;
  } else {
    if (int e = make_map(&mapA, A, K, M, 1, lda, (uint64_t)M * lda, 32, TC_BM, false)) return e;
    g.a_mn_major = 0;
    if (mode == GEMM_NN) {
      if (int e = make_map(&mapB, B, N, K, 1, ldb, (uint64_t)K * ldb, 32, 32, true)) return e;
      g.b_mn_major = 1;
    } else {
      if (int e = make_map(&mapB, B, K, N, 1, ldb, (uint64_t)N * ldb, 32, TC_BN, false)) return e;
      g.b_mn_major = 0;
    }
    g.kblocks = ceil_div(K, TC_BK);
    g.kps = g.kblocks;
  }
  const int tiles = ceil_div(M, TC_BM) * ceil_div(N, TC_BN);
  int splits = 1;
  if (workspace != nullptr && tiles < num_sms() && g.kblocks >= 64) {
    splits = min(num_sms() / tiles, g.kblocks / 16);
    const size_t per = (size_t)M * N * sizeof(float);
    if ((size_t)splits * per > ws_bytes) splits = (int)(ws_bytes / per);
    if (splits < 1) splits = 1;
  }
  g.kb_per_split = ceil_div(g.kblocks, splits);
  splits = ceil_div(g.kblocks, g.kb_per_split);
  g.part = splits > 1 ? workspace : nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    NABU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  {
    KernelScope ks(mode == GEMM_NN ? "gemm_tc_nn" : mode == GEMM_NT ? "gemm_tc_nt" : "gemm_tc_tn", stream);
    gemm_tc_kernel<<<dim3(ceil_div(N, TC_BN), ceil_div(M, TC_BM), splits), TC_THREADS, TC_SMEM_BYTES, stream>>>(mapA, mapB, g);
    NABU_CHECK_LAUNCH();
  }
  if (splits > 1) return splitk_reduce(workspace, splits, C, M, N, ldc, alpha, beta, bias, stream);
  return 0;
}

}  // namespace nabu
