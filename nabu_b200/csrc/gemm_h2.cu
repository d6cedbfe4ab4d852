// tcgen05 dense contraction on pre-split fp16 operands ("2xFP16 block floating point"), fp32-grade accuracy.
//
// gemm_tc.cu (3xTF32) spends its time in the converter warps: every landed fp32 tile is rewritten as hi/lo in
// shared memory before the tensor core may read it, and kind::tf32 runs at half the kind::f16 rate.  Here the
// split is done ONCE per operand by a bandwidth-bound pass (split_rows / split_global below): x * S is stored as
// hi = fp16(x S) and lo = fp16((x S - hi) * 2^11) -- two 11-bit significands = the 22 bits of the 3xTF32 scheme, in
// the same 4 bytes per value -- with S a power of two that puts the largest magnitude of the scaled group in
// [32, 64):
//   * operands whose rows run along K (A of NN/NT, B of NT) get one S per row (= per output row / column), so every
//     dot product is scaled by its own operand rows: 20 binades below the row maximum keep all 22 bits, smaller
//     values keep an absolute error of 2^-36 of the row maximum, which no fp32 dot product resolves either;
//   * operands whose rows run along M/N (both operands of TN -- the contraction is over 1e5 time-batch rows --
//     and the weight operand of NN) get one global S; the quantisation floor 2^-36 * max is far below the
//     2^-24 * sqrt(K) rounding noise of the fp32 accumulation itself.
// The kernel then needs no converter: TMA -> tcgen05.mma.kind::f16 x 3 per k-step (D1 += Ah.Bh, D2 += Ah.Bl + Al.Bh),
// epilogue out = (D1 + D2 * 2^-11) / (S_A S_B).  One CTA = one 128 x 256 tile, TMEM 512 columns (two fp32 accumulators).
// K-major (SWIZZLE_64B) and MN-major (SWIZZLE_128B) operand layouts, so NN / NT / TN and the per-utterance row
// segmentation of dKh need no transposes.
// Round 2, last revision (tools/gemm_bench.py, 192 000 x 2048 outputs: time per tile = 1.19 us per 64 of K + 7.5 us):
//   * the 1.19 us per 64 of K are 12 MMAs of 128 x 256 x 16 = 1536 tensor-pipe cycles at the ~1.3 GHz the SMs hold
//     under tensor load (MEASURED_PEAKS.json: 1327 MHz): the main loop is MMA-bound -- BK = 32 in four stages instead of
//     BK = 64 in two left it unchanged -- and the 7.5 us per tile were launch, pipeline fill and a serial epilogue;
//   * so the kernel is PERSISTENT (one CTA per SM, items = (split, row tile, column tile), column tile fastest): the TMA
//     warp runs ahead into the next tile (BK = 32, three stages of 48 KB), EIGHT epilogue warps (two per TMEM lane
//     quadrant) first pull their 32 x 128 share of D1 + D2 * 2^-11 into registers, hand tensor memory back to the MMA
//     warp ("drained" barrier) and only then transpose and store, under the next tile's main loop.
#include "common.cuh"
#include "gemm.h"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <algorithm>

namespace nabu {
namespace {

using namespace tc;

constexpr int H2_BM = 128, H2_BN = 256, H2_BK = 32, H2_STAGES = 3;
constexpr int H2_EPI_WARPS = 8;
constexpr int H2_THREADS = 64 + 32 * H2_EPI_WARPS;
constexpr int H2_A_TILE = H2_BM * H2_BK * 2;             // 8 KB
constexpr int H2_B_TILE = H2_BN * H2_BK * 2;             // 16 KB
constexpr int H2_STAGE = 2 * H2_A_TILE + 2 * H2_B_TILE;  // 48 KB
constexpr int H2_MN_BLK = 64 * H2_BK * 2;                // one 64-wide M/N block of an MN-major tile: BK rows of 128 B
constexpr int H2_TPAD = 36;                              // padded row of the epilogue transpose buffer (floats)
// stages | barriers (256 B) | column scale and bias of the tile | one transposition buffer per epilogue warp
constexpr int H2_SMEM = H2_STAGES * H2_STAGE + 1024 + 256 + 2 * H2_BN * 4 + H2_EPI_WARPS * 32 * H2_TPAD * 4;
static_assert(H2_SMEM <= 227 * 1024, "gemm_h2: shared memory");

struct H2Args {
  int M, N;
  int kblocks, kps, kb_per_split;
  int a_mn_major, b_mn_major;
  float alpha, beta;
  const float* bias;
  const float* a_row_inv; const float* a_glob_inv;       // 1 / S of A: per output row, or one value, or neither
  const float* b_row_inv; const float* b_glob_inv;       // 1 / S of B: per output column, or one value, or neither
  float* C; int ldc;
  float* part;
  int tiles_m, tiles_n, splits;                          // work items of the persistent grid
  unsigned* absmax_out;                                  // optional: max |C| over the outputs this launch writes (bits of a float, atomicMax)
};

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(H2_THREADS, 1)
gemm_h2_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const H2Args g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + H2_STAGES * H2_STAGE;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (H2_STAGES + s); };
  const uint32_t bar_tmem = bars + 8u * (2 * H2_STAGES);            // accumulators of a tile complete   (MMA -> epilogue)
  const uint32_t bar_drained = bars + 8u * (2 * H2_STAGES + 1);     // accumulators read into registers  (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + H2_STAGES * H2_STAGE + 8 * (2 * H2_STAGES + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Persistent: CTA c works on items c, c + gridDim.x, ...; item = (split z, row tile, column tile), column tile fastest
  // (the CTAs that run at the same time share their A rows through L2).
  const int total = g.tiles_n * g.tiles_m * g.splits;
  struct Item { int m0, n0, kb_begin, nkb, z; };
  auto decode = [&](int item) {
    Item t;
    const int tn = item % g.tiles_n, tm = (item / g.tiles_n) % g.tiles_m;
    t.z = item / (g.tiles_n * g.tiles_m);
    t.m0 = tm * H2_BM; t.n0 = tn * H2_BN;
    t.kb_begin = t.z * g.kb_per_split;
    t.nkb = max(0, min(g.kblocks, t.kb_begin + g.kb_per_split) - t.kb_begin);
    return t;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < H2_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_tmem, 1);
    mbar_init(bar_drained, H2_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: runs ahead into the next tile while the epilogue of this one drains =====
    if (lane == 0) {
      int it = 0;                                        // k blocks issued by this CTA so far (stage / phase counter)
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const Item ti = decode(item);
        const int m0 = ti.m0, n0 = ti.n0, kb_begin = ti.kb_begin, nkb = ti.nkb;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int kb = kb_begin + i;
          const int s = it % H2_STAGES;
          const uint32_t ph = (it / H2_STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          mbar_expect_tx(bar_full(s), H2_STAGE);
          const uint32_t sa = base + s * H2_STAGE;
          const uint32_t sb = sa + 2 * H2_A_TILE;
          const int seg = kb / g.kps, kk = (kb % g.kps) * H2_BK;
          if (g.a_mn_major) {
            for (int j = 0; j < H2_BM / 64; ++j) {
              tma_load_3d(sa + j * H2_MN_BLK, &mapAh, bar_full(s), m0 + 64 * j, kk, seg);
              tma_load_3d(sa + H2_A_TILE + j * H2_MN_BLK, &mapAl, bar_full(s), m0 + 64 * j, kk, seg);
            }
          } else {
            tma_load_3d(sa, &mapAh, bar_full(s), kb * H2_BK, m0, 0);
            tma_load_3d(sa + H2_A_TILE, &mapAl, bar_full(s), kb * H2_BK, m0, 0);
          }
          if (g.b_mn_major) {
            for (int j = 0; j < H2_BN / 64; ++j) {
              tma_load_3d(sb + j * H2_MN_BLK, &mapBh, bar_full(s), n0 + 64 * j, kk, seg);
              tma_load_3d(sb + H2_B_TILE + j * H2_MN_BLK, &mapBl, bar_full(s), n0 + 64 * j, kk, seg);
            }
          } else {
            tma_load_3d(sb, &mapBh, bar_full(s), kb * H2_BK, n0, 0);
            tma_load_3d(sb + H2_B_TILE, &mapBl, bar_full(s), kb * H2_BK, n0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(H2_BM, H2_BN, g.a_mn_major, g.b_mn_major);
      // K-major SWIZZLE_64B: rows of 64 B (32 fp16 of K), 8-row groups 512 B apart (SBO), k-step (16) = +32 B.
      // MN-major SWIZZLE_128B: rows of 128 B (64 fp16 of M/N) per k, 8-k groups 1024 B apart (SBO), 64-wide M/N
      // blocks H2_MN_BLK = 4096 B apart (LBO), k-step (16 rows) = +2048 B.
      const uint32_t a_lbo = g.a_mn_major ? (uint32_t)H2_MN_BLK : 16u, b_lbo = g.b_mn_major ? (uint32_t)H2_MN_BLK : 16u;
      const uint32_t a_kstep = g.a_mn_major ? 2048u : 32u, b_kstep = g.b_mn_major ? 2048u : 32u;
      const uint32_t a_sbo = g.a_mn_major ? 1024u : 512u, b_sbo = g.b_mn_major ? 1024u : 512u;
      const uint32_t a_lay = g.a_mn_major ? 2u : 4u, b_lay = g.b_mn_major ? 2u : 4u;      // SWIZZLE_128B : SWIZZLE_64B
      int it = 0, j = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x, ++j) {
        const int nkb = decode(item).nkb;
        // the epilogue warps hold the previous tile's accumulators in registers: tensor memory may be overwritten
        mbar_wait(bar_drained, (uint32_t)(j & 1) ^ 1u);
        tc_fence_after();
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % H2_STAGES;
          const uint32_t ph = (it / H2_STAGES) & 1;
          mbar_wait(bar_full(s), ph);
          tc_fence_after();
          const uint32_t sa = base + s * H2_STAGE;
          const uint32_t sb = sa + 2 * H2_A_TILE;
#pragma unroll
          for (int k = 0; k < H2_BK / 16; ++k) {
            const uint64_t a_hi = make_desc(sa + k * a_kstep, a_lbo, a_sbo, a_lay);
            const uint64_t a_lo = make_desc(sa + H2_A_TILE + k * a_kstep, a_lbo, a_sbo, a_lay);
            const uint64_t b_hi = make_desc(sb + k * b_kstep, b_lbo, b_sbo, b_lay);
            const uint64_t b_lo = make_desc(sb + H2_B_TILE + k * b_kstep, b_lbo, b_sbo, b_lay);
            const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
            umma_f16(tmem_d, a_hi, b_hi, idesc, acc);
            umma_f16(tmem_d + H2_BN, a_hi, b_lo, idesc, acc);
            umma_f16(tmem_d + H2_BN, a_lo, b_hi, idesc, 1u);
          }
          umma_commit(bar_empty(s));
        }
        umma_commit(bar_tmem);
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers (then the MMA warp may start the next tile) -> warp-private smem transpose ->
    // coalesced global stores, overlapped with the next tile's main loop =====
    // tcgen05.ld hands every thread one ROW of the tile; storing that directly makes each warp instruction touch 32
    // rows (measured: 29 us per tile, 0.7 TB/s).  Each warp therefore transposes its 32 x 32 chunks through a padded
    // shared buffer and writes full 128-byte row segments (8 lanes x float4 per row, 4 rows per instruction).
    const int q = warp & 3;
    const bool split = g.part != nullptr;
    const int ldc = split ? g.N : g.ldc;
    float* tbuf = reinterpret_cast<float*>(base_ptr + H2_STAGES * H2_STAGE + 256 + 2 * H2_BN * 4) + (warp - 2) * (32 * H2_TPAD);
    const int chalf = (warp - 2) >> 2;                   // which four 32-column chunks of the tile this warp drains
    constexpr int NCH = H2_BN / 64;                      // chunks per warp
    float* cs_s = reinterpret_cast<float*>(base_ptr + H2_STAGES * H2_STAGE + 256);
    float* bs_s = cs_s + H2_BN;
    const int rl = lane >> 3, c4 = (lane & 7) * 4;       // read phase: row within a group of 4, first of 4 columns
    const float ag = (g.a_glob_inv ? __ldg(g.a_glob_inv) : 1.f) * (g.b_glob_inv ? __ldg(g.b_glob_inv) : 1.f);
    float vmax = 0.f;                                     // max |value stored| by this thread (absmax_out)
    int j = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++j) {
      const Item ti = decode(item);
      const int m0 = ti.m0, n0 = ti.n0, nkb = ti.nkb, z = ti.z;
      const int m = m0 + 32 * q + lane;
      float* Cout = split ? g.part + (size_t)z * g.M * g.N : g.C;
      const bool vec = ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0) && (ldc % 4 == 0);
      float sa_inv = ag;
      if (g.a_row_inv && m < g.M) sa_inv *= __ldg(g.a_row_inv + m);
      // column scale and bias of this tile -> shared memory (every epilogue warp is done with the previous tile's)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * H2_EPI_WARPS) : "memory");
      for (int c = threadIdx.x - 64; c < H2_BN; c += 32 * H2_EPI_WARPS) {
        const int n = n0 + c;
        float cs = 1.f, bs = 0.f;
        if (n < g.N) {
          if (g.b_row_inv) cs = __ldg(g.b_row_inv + n);
          if (!split) {
            cs *= g.alpha;
            if (g.bias) bs = __ldg(g.bias + n);
          }
        }
        cs_s[c] = cs;
        bs_s[c] = bs;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * H2_EPI_WARPS) : "memory");
      const int ch0 = chalf * NCH;
      const int nchunks = max(ch0, min(ch0 + NCH, (g.N - n0 + 31) / 32));     // this warp: chunks [ch0, nchunks)
      mbar_wait(bar_tmem, (uint32_t)(j & 1));
      tc_fence_after();
      // phase 1: my 32 rows x 128 columns of D1 + D2 * 2^-11, scaled by the row's 1 / S, into registers
      float acc[NCH][32];
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t ra[16], rb[16];
          if (nkb > 0 && ch0 + cc < nchunks) {
            tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)((ch0 + cc) * 32 + h * 16), ra);
            tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(H2_BN + (ch0 + cc) * 32 + h * 16), rb);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) ra[i] = rb[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i)
            acc[cc][h * 16 + i] = fmaf(__uint_as_float(rb[i]), 1.f / 2048.f, __uint_as_float(ra[i])) * sa_inv;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_drained);
      // phase 2: transpose and store
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        const int ch = ch0 + cc;
        if (ch >= nchunks) break;
        const int c0 = ch * 32;
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4)
          *reinterpret_cast<float4*>(tbuf + lane * H2_TPAD + j4) = make_float4(acc[cc][j4], acc[cc][j4 + 1], acc[cc][j4 + 2], acc[cc][j4 + 3]);
        __syncwarp();
        const int n = n0 + c0 + c4;
        const float4 cs4 = *reinterpret_cast<const float4*>(cs_s + c0 + c4);
        const float4 bs4 = *reinterpret_cast<const float4*>(bs_s + c0 + c4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rl;
          const int mm = m0 + 32 * q + rr;
          const float4 t4 = *reinterpret_cast<const float4*>(tbuf + rr * H2_TPAD + c4);
          float v[4] = {fmaf(t4.x, cs4.x, bs4.x), fmaf(t4.y, cs4.y, bs4.y), fmaf(t4.z, cs4.z, bs4.z), fmaf(t4.w, cs4.w, bs4.w)};
          if (mm < g.M && n < g.N) {
            float* cp = Cout + (size_t)mm * ldc + n;
            const int nv = min(4, g.N - n);
            if (nv == 4 && vec) {
              if (!split && g.beta != 0.f) {
                const float4 o = *reinterpret_cast<const float4*>(cp);
                v[0] += g.beta * o.x; v[1] += g.beta * o.y; v[2] += g.beta * o.z; v[3] += g.beta * o.w;
              }
              *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
              vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));
            } else {
              for (int jj = 0; jj < nv; ++jj) {
                const float o = (!split && g.beta != 0.f) ? v[jj] + g.beta * cp[jj] : v[jj];
                cp[jj] = o;
                vmax = fmaxf(vmax, fabsf(o));
              }
            }
          }
        }
        __syncwarp();
      }
    }
    if (g.absmax_out) {
      vmax = warp_max(vmax);
      if (lane == 0 && vmax > 0.f) atomicMax(g.absmax_out, __float_as_uint(vmax));     // positive floats order like their bits
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, 512u);
}

// ---- split passes -----------------------------------------------------------------------------------------
__device__ __forceinline__ float pow2_scale(float maxabs) {     // S = 2^k with maxabs * S in [32, 64); 1 for 0 / inf / nan
  if (!(maxabs > 0.f) || !(maxabs < 3.0e38f)) return 1.f;
  int e;
  frexpf(maxabs, &e);
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return ldexpf(1.f, 6 - e);
}
__device__ __forceinline__ __half sat_h(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return __ushort_as_half(h);
}
__device__ __forceinline__ void split2(float x, __half* hi, __half* lo) {
  const __half h = sat_h(x);
  *hi = h;
  *lo = sat_h((x - __half2float(h)) * 2048.f);
}

// 8 consecutive values -> 8 hi + 8 lo halves (one 16-byte store each)
__device__ __forceinline__ void split8(const float* v, float S, __half* hi, __half* lo) {
  union { __half h[8]; uint4 u; } a, b;
#pragma unroll
  for (int j = 0; j < 8; ++j) split2(v[j] * S, &a.h[j], &b.h[j]);
  *reinterpret_cast<uint4*>(hi) = a.u;
  *reinterpret_cast<uint4*>(lo) = b.u;
}
__device__ __forceinline__ void load8(const float* x, bool vec, int c, int C, float* v) {
  if (vec && c + 8 <= C) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(x + c));
    const float4 q = __ldg(reinterpret_cast<const float4*>(x + c + 4));
    v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = c + j < C ? __ldg(x + c + j) : 0.f;
  }
}

// TPR threads per row (256 / TPR rows per block): max |x| over the row, then hi/lo of x * S(row).  Columns [C, ldo)
// are zero-filled; ldo % 8 == 0.
template <int TPR>
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ src, int ld, int R, int C, __half* __restrict__ hi, __half* __restrict__ lo,
                  int ldo, float* __restrict__ row_inv) {
  constexpr int RPB = 256 / TPR;
  const int t = threadIdx.x % TPR;
  const size_t row = (size_t)blockIdx.x * RPB + threadIdx.x / TPR;
  const bool live = row < (size_t)R;
  const float* x = src + (live ? row : 0) * ld;
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  __shared__ float wmax[8];
  float m = 0.f;
  if (live)
    for (int c = t * 8; c < C; c += TPR * 8) {
      float v[8];
      load8(x, vec, c, C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(v[j]));
    }
  m = warp_max(m);
  if (TPR > 32) {
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    m = wmax[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, wmax[i]);
  }
  if (!live) return;
  const float S = pow2_scale(m);
  if (t == 0) row_inv[row] = 1.f / S;
  for (int c = t * 8; c < ldo; c += TPR * 8) {
    float v[8];
    load8(x, vec, c, C, v);
    split8(v, S, hi + row * ldo + c, lo + row * ldo + c);
  }
}

// rows are distributed over blocks, 8-column groups over threads
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ src, int ld, size_t R, size_t C, unsigned* out) {
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  float m = 0.f;
  for (size_t r = blockIdx.x; r < R; r += gridDim.x) {
    const float* x = src + r * ld;
    for (size_t c = (size_t)threadIdx.x * 8; c < C; c += 256 * 8) {
      float v[8];
      if (vec && c + 8 <= C) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(x + c));
        const float4 q = __ldg(reinterpret_cast<const float4*>(x + c + 4));
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = c + j < C ? __ldg(x + c + j) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(v[j]));
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

__global__ void __launch_bounds__(256)
split_global_kernel(const float* __restrict__ src, int ld, size_t R, size_t C, const unsigned* __restrict__ maxbits,
                    __half* __restrict__ hi, __half* __restrict__ lo, size_t ldo, float* __restrict__ glob_inv) {
  const float S = pow2_scale(__uint_as_float(*maxbits));
  if (blockIdx.x == 0 && threadIdx.x == 0) *glob_inv = 1.f / S;
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (size_t r = blockIdx.x; r < R; r += gridDim.x) {
    const float* x = src + r * ld;
    for (size_t c = (size_t)threadIdx.x * 8; c < ldo; c += 256 * 8) {
      float v[8];
      if (vec && c + 8 <= C) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(x + c));
        const float4 q = __ldg(reinterpret_cast<const float4*>(x + c + 4));
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = c + j < C ? __ldg(x + c + j) : 0.f;
      }
      split8(v, S, hi + r * ldo + c, lo + r * ldo + c);
    }
  }
}

// Two [R, C] matrices side by side -> one [R, 2C] operand with ONE power-of-two scale per row (the two directions' dZ
// for the single dX GEMM over K = 8H).  One block per row, C % 8 == 0, 16-byte aligned sources.
__global__ void __launch_bounds__(256)
split_rows_pair_kernel(const float* __restrict__ s0, const float* __restrict__ s1, int ld, int C, __half* __restrict__ hi,
                       __half* __restrict__ lo, float* __restrict__ row_inv) {
  const size_t row = blockIdx.x;
  const float* x0 = s0 + row * ld;
  const float* x1 = s1 + row * ld;
  __shared__ float wmax[8];
  float m = 0.f;
  for (int c = threadIdx.x * 8; c < 2 * C; c += 256 * 8) {
    float v[8];
    load8(c < C ? x0 : x1, true, c < C ? c : c - C, C, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(v[j]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  m = wmax[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, wmax[i]);
  const float S = pow2_scale(m);
  if (threadIdx.x == 0) row_inv[row] = 1.f / S;
  for (int c = threadIdx.x * 8; c < 2 * C; c += 256 * 8) {
    float v[8];
    load8(c < C ? x0 : x1, true, c < C ? c : c - C, C, v);
    split8(v, S, hi + row * 2 * C + c, lo + row * 2 * C + c);
  }
}

// C columns (C % 8 == 0) of an [R, C] matrix into columns [col0, col0 + C) of planes with row pitch ldo; the scale is
// read from maxbits (accumulated by absmax_kernel over every matrix that shares it).
__global__ void __launch_bounds__(256)
split_global_sub_kernel(const float* __restrict__ src, int ld, int R, int C, const unsigned* __restrict__ maxbits,
                        __half* __restrict__ hi, __half* __restrict__ lo, size_t ldo, int col0, float* __restrict__ glob_inv) {
  const float S = pow2_scale(__uint_as_float(*maxbits));
  if (blockIdx.x == 0 && threadIdx.x == 0) *glob_inv = 1.f / S;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const float* x = src + (size_t)r * ld;
    for (int c = threadIdx.x * 8; c < C; c += 256 * 8) {
      float v[8];
      load8(x, true, c, C, v);
      split8(v, S, hi + (size_t)r * ldo + col0 + c, lo + (size_t)r * ldo + col0 + c);
    }
  }
}

int make_map_h(CUtensorMap* map, const __half* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
               uint64_t stride2, uint32_t box0, uint32_t box1) {
  // the box's inner extent IS the swizzle span: 64 fp16 = SWIZZLE_128B (MN-major operands), 32 fp16 = SWIZZLE_64B (K-major)
  const CUtensorMapSwizzle swz = box0 * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  EncodeTiledFn enc = get_encode();
  NABU_REQUIRE(enc != nullptr, "gemm_h2: cuTensorMapEncodeTiled entry point missing");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const int r = (int)enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NABU_REQUIRE(r == 0, "gemm_h2: cuTensorMapEncodeTiled failed (%d)", r);
  return 0;
}

}  // namespace

int split_rows(const float* src, int ld, int R, int C, void* hi, void* lo, int ldo, float* row_inv, cudaStream_t stream) {
  NABU_REQUIRE(ldo % 8 == 0 && ldo >= C, "split_rows: ldo must be a multiple of 8 and >= C");
  KernelScope ks("split_rows", stream);
  if (C <= 256) split_rows_kernel<32><<<ceil_div(R, 8), 256, 0, stream>>>(src, ld, R, C, (__half*)hi, (__half*)lo, ldo, row_inv);
  else split_rows_kernel<256><<<R, 256, 0, stream>>>(src, ld, R, C, (__half*)hi, (__half*)lo, ldo, row_inv);
  NABU_CHECK_LAUNCH();
  return 0;
}

int split_rows_pair(const float* s0, const float* s1, int ld, int R, int C, void* hi, void* lo, float* row_inv,
                    cudaStream_t stream) {
  NABU_REQUIRE(C % 8 == 0 && ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(s0) | reinterpret_cast<uintptr_t>(s1)) & 15) == 0,
               "split_rows_pair: C %% 8, ld %% 4 and 16-byte alignment required");
  KernelScope ks("split_rows", stream);
  split_rows_pair_kernel<<<R, 256, 0, stream>>>(s0, s1, ld, C, (__half*)hi, (__half*)lo, row_inv);
  NABU_CHECK_LAUNCH();
  return 0;
}

int split_global_pair(const float* s0, const float* s1, int ld, int R, int C, void* hi, void* lo, unsigned* maxbits,
                      float* glob_inv, cudaStream_t stream) {
  NABU_REQUIRE(C % 8 == 0 && ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(s0) | reinterpret_cast<uintptr_t>(s1)) & 15) == 0,
               "split_global_pair: C %% 8, ld %% 4 and 16-byte alignment required");
  NABU_CHECK_CUDA(cudaMemsetAsync(maxbits, 0, sizeof(unsigned), stream));
  const int blocks = (int)min((size_t)num_sms() * 4, (size_t)R);
  const float* src[2] = {s0, s1};
  for (int d = 0; d < 2; ++d) {
    KernelScope ks("absmax", stream);
    absmax_kernel<<<blocks, 256, 0, stream>>>(src[d], ld, (size_t)R, (size_t)C, maxbits);
    NABU_CHECK_LAUNCH();
  }
  for (int d = 0; d < 2; ++d) {
    KernelScope ks("split_global", stream);
    split_global_sub_kernel<<<blocks, 256, 0, stream>>>(src[d], ld, R, C, maxbits, (__half*)hi, (__half*)lo, (size_t)2 * C, d * C,
                                                        glob_inv);
    NABU_CHECK_LAUNCH();
  }
  return 0;
}

// A matrix that is contiguous (ld == C == ldo) is processed as one long row, so narrow matrices keep all lanes busy.
int split_global(const float* src, int ld, int R, int C, void* hi, void* lo, int ldo, unsigned* maxbits, float* glob_inv,
                 cudaStream_t stream) {
  NABU_REQUIRE(ldo % 8 == 0 && ldo >= C, "split_global: ldo must be a multiple of 8 and >= C");
  NABU_CHECK_CUDA(cudaMemsetAsync(maxbits, 0, sizeof(unsigned), stream));
  size_t R2 = (size_t)R, C2 = (size_t)C, ldo2 = (size_t)ldo;
  int ld2 = ld;
  if (ld == C && ldo == C) {                           // flatten into rows of 2048 values
    const size_t n = (size_t)R * C;
    if (n % 2048 == 0) { R2 = n / 2048; C2 = 2048; ldo2 = 2048; ld2 = 2048; }
  }
  const int blocks = (int)min((size_t)num_sms() * 16, R2);
  {
    KernelScope ks("absmax", stream);
    absmax_kernel<<<blocks, 256, 0, stream>>>(src, ld2, R2, C2, maxbits);
    NABU_CHECK_LAUNCH();
  }
  KernelScope ks("split_global", stream);
  split_global_kernel<<<blocks, 256, 0, stream>>>(src, ld2, R2, C2, maxbits, (__half*)hi, (__half*)lo, ldo2, glob_inv);
  NABU_CHECK_LAUNCH();
  return 0;
}

int absmax_accumulate(const float* src, int ld, size_t R, size_t C, unsigned* out, cudaStream_t stream) {
  const int blocks = (int)min((size_t)num_sms() * 16, R);
  KernelScope ks("absmax", stream);
  absmax_kernel<<<blocks, 256, 0, stream>>>(src, ld, R, C, out);
  NABU_CHECK_LAUNCH();
  return 0;
}

int gemm_h2(GemmMode mode, int M, int N, int K, float alpha, const H2Operand& A, const H2Operand& B, float beta, float* C,
            int ldc, const float* bias, const GemmSeg* segp, float* workspace, size_t ws_bytes, cudaStream_t stream,
            unsigned* absmax_out) {
  NABU_REQUIRE(!(segp && mode != GEMM_TN), "gemm_h2: row segmentation only in TN mode");
  NABU_REQUIRE(A.ld % 8 == 0 && B.ld % 8 == 0, "gemm_h2: operand leading dimensions must be multiples of 8");
  CUtensorMap mAh, mAl, mBh, mBl;
  H2Args g = {};
  g.M = M; g.N = N; g.alpha = alpha; g.beta = beta; g.bias = bias; g.C = C; g.ldc = ldc;
  g.a_row_inv = A.row_inv; g.a_glob_inv = A.glob_inv; g.b_row_inv = B.row_inv; g.b_glob_inv = B.glob_inv;
  const __half* Ah = (const __half*)A.hi; const __half* Al = (const __half*)A.lo;
  const __half* Bh = (const __half*)B.hi; const __half* Bl = (const __half*)B.lo;
  if (mode == GEMM_TN) {
    const int seg = segp ? segp->seg : K;
    const int nseg = segp ? K / segp->seg : 1;
    NABU_REQUIRE(!segp || K % segp->seg == 0, "gemm_h2: K must be a multiple of the segment length");
    NABU_REQUIRE(!A.row_inv && !B.row_inv, "gemm_h2: TN operands take a global scale");
    const uint64_t sA = segp ? (uint64_t)segp->segA : (uint64_t)K, sB = segp ? (uint64_t)segp->segB : (uint64_t)K;
    const size_t oa = segp ? (size_t)segp->offA * A.ld : 0, ob = segp ? (size_t)segp->offB * B.ld : 0;
    if (int e = make_map_h(&mAh, Ah + oa, M, seg, nseg, A.ld, sA * A.ld, 64, H2_BK)) return e;
    if (int e = make_map_h(&mAl, Al + oa, M, seg, nseg, A.ld, sA * A.ld, 64, H2_BK)) return e;
    if (int e = make_map_h(&mBh, Bh + ob, N, seg, nseg, B.ld, sB * B.ld, 64, H2_BK)) return e;
    if (int e = make_map_h(&mBl, Bl + ob, N, seg, nseg, B.ld, sB * B.ld, 64, H2_BK)) return e;
    g.a_mn_major = 1; g.b_mn_major = 1;
    g.kps = ceil_div(seg, H2_BK);
    g.kblocks = g.kps * nseg;
  } else {
    if (int e = make_map_h(&mAh, Ah, K, M, 1, A.ld, (uint64_t)M * A.ld, H2_BK, H2_BM)) return e;
    if (int e = make_map_h(&mAl, Al, K, M, 1, A.ld, (uint64_t)M * A.ld, H2_BK, H2_BM)) return e;
    g.a_mn_major = 0;
    if (mode == GEMM_NN) {
      NABU_REQUIRE(!B.row_inv, "gemm_h2: the NN weight operand takes a global scale");
      if (int e = make_map_h(&mBh, Bh, N, K, 1, B.ld, (uint64_t)K * B.ld, 64, H2_BK)) return e;
      if (int e = make_map_h(&mBl, Bl, N, K, 1, B.ld, (uint64_t)K * B.ld, 64, H2_BK)) return e;
      g.b_mn_major = 1;
    } else {
      if (int e = make_map_h(&mBh, Bh, K, N, 1, B.ld, (uint64_t)N * B.ld, H2_BK, H2_BN)) return e;
      if (int e = make_map_h(&mBl, Bl, K, N, 1, B.ld, (uint64_t)N * B.ld, H2_BK, H2_BN)) return e;
      g.b_mn_major = 0;
    }
    g.kblocks = ceil_div(K, H2_BK);
    g.kps = g.kblocks;
  }
  const int tiles = ceil_div(M, H2_BM) * ceil_div(N, H2_BN);
  int splits = 1;
  if (workspace != nullptr && tiles < num_sms() && g.kblocks >= 64) {          // k blocks of 32: at least 512 of K per split
    splits = min(num_sms() / tiles, g.kblocks / 16);
    const size_t per = (size_t)M * N * sizeof(float);
    if ((size_t)splits * per > ws_bytes) splits = (int)(ws_bytes / per);
    if (splits < 1) splits = 1;
  }
  g.kb_per_split = ceil_div(g.kblocks, splits);
  splits = ceil_div(g.kblocks, g.kb_per_split);
  g.part = splits > 1 ? workspace : nullptr;
  g.absmax_out = splits > 1 ? nullptr : absmax_out;       // split-K: the outputs only exist after the reduction (below)
  static bool attr_set = false;
  if (!attr_set) {
    NABU_CHECK_CUDA(cudaFuncSetAttribute(gemm_h2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H2_SMEM));
    attr_set = true;
  }
  {
    KernelScope ks(mode == GEMM_NN ? "gemm_h2_nn" : mode == GEMM_NT ? "gemm_h2_nt" : "gemm_h2_tn", stream);
    g.tiles_n = ceil_div(N, H2_BN); g.tiles_m = ceil_div(M, H2_BM); g.splits = splits;
    const long total = (long)g.tiles_n * g.tiles_m * splits;
    NABU_REQUIRE(total < (1L << 31), "gemm_h2: too many tiles");
    // TN = the weight gradients, which run on a side stream beside a recurrence that holds 64 SMs: one item per CTA, so
    // that the hardware scheduler balances them over whatever SMs are free (their few tiles with K ~ 1e5 have nothing to
    // gain from persistence); NN / NT run alone and persist
    const int grid = mode == GEMM_TN ? (int)total : (int)std::min<long>(total, num_sms());
    gemm_h2_kernel<<<grid, H2_THREADS, H2_SMEM, stream>>>(mAh, mAl, mBh, mBl, g);
    NABU_CHECK_LAUNCH();
  }
  if (splits > 1) {
    if (int e = splitk_reduce(workspace, splits, C, M, N, ldc, alpha, beta, bias, stream)) return e;
    if (absmax_out) return absmax_accumulate(C, ldc, (size_t)M, (size_t)N, absmax_out, stream);
  }
  return 0;
}


// ---- fp32 in, fp32 out: split both operands into `workspace`, then contract ------------------------------------
namespace {
struct AutoWs {
  float* splitk; size_t splitk_bytes;
  void *ah, *al, *bh, *bl;
  float *a_row, *b_row, *a_glob, *b_glob;
  unsigned *a_max, *b_max;
  size_t total;
};
AutoWs carve_auto(void* base, GemmMode mode, int M, int N, int K) {
  AutoWs w;
  char* b = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = b + off; off += align_up(bytes, 256); return p; };
  w.splitk_bytes = sgemm_workspace_bytes();
  w.splitk = (float*)take(w.splitk_bytes);
  const size_t ra = mode == GEMM_TN ? K : M, ca = mode == GEMM_TN ? M : K;     // A as stored: [ra][ca]
  const size_t rb = mode == GEMM_NT ? N : K, cb = mode == GEMM_NT ? K : N;     // B as stored: [rb][cb]
  const size_t ea = ra * align_up(ca, 8), eb = rb * align_up(cb, 8);
  w.ah = take(ea * 2); w.al = take(ea * 2);
  w.bh = take(eb * 2); w.bl = take(eb * 2);
  w.a_row = (float*)take((size_t)M * 4); w.b_row = (float*)take((size_t)N * 4);
  w.a_glob = (float*)take(256); w.b_glob = w.a_glob + 1;
  w.a_max = (unsigned*)(w.a_glob + 2); w.b_max = w.a_max + 1;
  w.total = off;
  return w;
}
}  // namespace

size_t gemm_h2_auto_workspace_bytes(GemmMode mode, int M, int N, int K) { return carve_auto(nullptr, mode, M, N, K).total; }

bool gemm_h2_eligible(GemmMode mode, int M, int N, int K) {
  (void)mode;
  return M >= 1 && N >= 1 && K >= 1 && (long)M * N >= 128L * 128L;
}

int gemm_h2_auto(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
                 float beta, float* C, int ldc, const float* bias, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  AutoWs w = carve_auto(workspace, mode, M, N, K);
  NABU_REQUIRE(workspace && ws_bytes >= w.total, "gemm_h2: workspace %zu < %zu bytes", ws_bytes, w.total);
  H2Operand a = {}, b = {};
  a.hi = w.ah; a.lo = w.al; b.hi = w.bh; b.lo = w.bl;
  if (mode == GEMM_TN) {
    a.ld = (int)align_up(M, 8); b.ld = (int)align_up(N, 8);
    if (int e = split_global(A, lda, K, M, w.ah, w.al, a.ld, w.a_max, w.a_glob, stream)) return e;
    if (int e = split_global(B, ldb, K, N, w.bh, w.bl, b.ld, w.b_max, w.b_glob, stream)) return e;
    a.glob_inv = w.a_glob; b.glob_inv = w.b_glob;
  } else {
    a.ld = (int)align_up(K, 8);
    if (int e = split_rows(A, lda, M, K, w.ah, w.al, a.ld, w.a_row, stream)) return e;
    a.row_inv = w.a_row;
    if (mode == GEMM_NN) {
      b.ld = (int)align_up(N, 8);
      if (int e = split_global(B, ldb, K, N, w.bh, w.bl, b.ld, w.b_max, w.b_glob, stream)) return e;
      b.glob_inv = w.b_glob;
    } else {
      b.ld = (int)align_up(K, 8);
      if (int e = split_rows(B, ldb, N, K, w.bh, w.bl, b.ld, w.b_row, stream)) return e;
      b.row_inv = w.b_row;
    }
  }
  return gemm_h2(mode, M, N, K, alpha, a, b, beta, C, ldc, bias, nullptr, w.splitk, w.splitk_bytes, stream);
}

}  // namespace nabu
