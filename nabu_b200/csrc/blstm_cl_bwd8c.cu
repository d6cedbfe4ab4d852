// Backward recurrence of a BLSTM layer, "chains" version of blstm_cl_bwd8.cu (same partition: clusters of 8, a cluster
// owns 128 units, CTA r multiplies K-slice r of dz against its block of Kh resident in TMEM, transposed product
// D^T[128 units x batch] = W . dz^T, bulk-DSMEM reduce-scatter, flag-in-data exchange through L2).
//
// What changes.  The phase trace of blstm_cl_bwd8 (profiles/r1d_trace_bwd8.txt) shows a time step as a chain of stages
// that each keep ONE resource busy while the others idle: L2 hand-off and staging of the dz slice (LSU), 48 MMAs (tensor
// pipe, 1.6 us at the measured 64 cycles per 128x128x16), TMEM drain, 56 KB of DSMEM copies (SM-to-SM network,
// 1.8 us), the pointwise stage (LSU / MUFU).  Batch rows are independent, so the 128-row tile is cut into NCH = 2
// CHAINS of NB = 64 rows, each run by its own warpgroup with its own barriers, shared-memory buffers and TMEM
// accumulators and NO common synchronisation: while one chain waits for its operands or its copies, the other one
// computes.  tools/probe_mma_rate.cu: a TS-form MMA of N = 64 costs 47 cycles from one issuing thread and 33 from two,
// against 64 for N = 128 -- two issuers at N = 64 keep the tensor pipe as busy as one at N = 128.
// The same kernel with ONE chain of NB = 16 / 32 / 64 rows serves small batches (the strong-scaling split of a
// minibatch over 8 GPUs leaves 16 rows per GPU): exchange volume, DSMEM volume and pointwise work shrink with the batch
// instead of being paid for 128 rows.
//
// Per-chain hand-back of the receive / staging buffers: an mbarrier ("free") in every CTA that collects one arrival per
// warp of the chain from all 8 CTAs of the cluster (remote mbarrier.arrive) replaces the CTA-wide cluster barrier.
// TMEM columns (512): accumulators of chain c at [c*2*NB, (c+1)*2*NB) (D1 | D2) | W hi [256, 256+K/2) | W lo [.., 256+K).
#include "cl_tc_common.cuh"
#include "blstm_cl.h"
#include <algorithm>

namespace nabu {
namespace {

__device__ __forceinline__ void umma_f16_ts_c(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st8_c(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait_c() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void split_h_sat_flag_c(float x, unsigned short fb, unsigned short* hi, unsigned short* lo) {
  const unsigned short h = (unsigned short)((__half_as_ushort(sat_half(x)) & 0xFFFEu) | fb);
  *hi = h;
  const float res = (x - __half2float(__ushort_as_half(h))) * 2048.f;
  *lo = (unsigned short)((__half_as_ushort(sat_half(res)) & 0xFFFEu) | fb);
}
__device__ __forceinline__ void ld4c(const float* p, float (&v)[4]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
// 128 threads of one chain (named barrier 1 + chain; barrier 0 is __syncthreads)
__device__ __forceinline__ void bar_chain(int ch) { asm volatile("bar.sync %0, 128;" ::"r"(ch + 1) : "memory"); }
// Relaxed on purpose: the arrival only says "my loads of the receive buffer have returned" -- `dep` is a value computed
// from every one of them, so the instruction cannot issue before they have -- and publishes no writes.  The .release form
// compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive: 0.5 us of every time step (ncu, profiles/r2_ncu_full.md).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster, float dep) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster), "f"(dep) : "memory");
}
// Keeps the arithmetic on a prefetched value BELOW this point of the instruction stream: without it the compiler starts
// tanh(c) right behind the prefetch loads and the chain stalls on their L2 / HBM latency in front of the MMAs.
__device__ __forceinline__ void pin(float& x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}

template <int NQ, int NB, int NCH>                    // clusters per direction, batch rows per chain, chains
__global__ void __launch_bounds__(128 * NCH, 1)
blstm_rec_bwd_chain_kernel(const ClParams p, const unsigned* __restrict__ rowmax) {
  constexpr int CLS = 8, HS = 16, NC = 128, BT = NB * NCH;
  constexpr int KBN = NQ;                             // 64-column K blocks of my slice, one per producer cluster
  constexpr int TILE = NB * 128;                      // bytes of a chain's [NB rows x 64 fp16] K-major tile
  constexpr int XTILE = BT * 128;                     // bytes of the exchange's tile (the rows of all chains)
  constexpr int CSLICE = KBN * 2 * TILE;              // shared-memory image of a chain's dz slice (hi | lo per K block)
  constexpr int XSLICE = KBN * 2 * XTILE;             // a slice in the exchange buffer
  constexpr int BLK = NB * 16 * 4;                    // one (source CTA, destination CTA) block: [NB batch][16 units] fp32
  constexpr int BST = BLK + 64;                       // block stride (+16 banks)
  constexpr int CSTRIDE = (CSLICE + CLS * BST + 1023) / 1024 * 1024;   // shared memory of one chain
  constexpr int ACOLS = KBN * 32;                     // TMEM columns of one half of the weights
  // num_units = 1024 (NQ = 8): the hi halves of the CTA's [128 units x 512] weight block fill the 256 TMEM columns the
  // accumulators leave, the lo halves sit in shared memory as K-major SWIZZLE_128B tiles (the A operand of an SS-form MMA)
  constexpr bool LOS = NQ > 4;
  constexpr int WLO = LOS ? KBN * A_TILE : 0;         // bytes of the lo halves in shared memory
  constexpr uint32_t TM_AH = 256, TM_AL = 256 + ACOLS;
  constexpr int RPT = NB >= 32 ? NB / 32 : 1;         // batch rows per pointwise thread
  constexpr int CPB = NB / 8;                         // 16-byte chunks per thread and K block (hi tile, then lo tile)
  constexpr int CPH = CPB / 2;                        // ... per tile
  static_assert(NB == 16 || NB == 32 || NB == 64, "NB");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t rx_bar[NCH];
  __shared__ __align__(8) uint64_t mma_bar[NCH];
  __shared__ __align__(8) uint64_t free_bar[NCH];
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned gmax_bits;

  const int H = p.H, H4 = 4 * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ch = __shfl_sync(0xffffffffu, tid >> 7, 0);          // chain of this warpgroup (warp-uniform)
  const int t = tid & 127, wq = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int per_dir = NQ * CLS;
  const int dir = p.dir0 + blockIdx.x / per_dir;
  const int q = (blockIdx.x % per_dir) / CLS;
  const int r = blockIdx.x % CLS;
  const int j0 = (q * CLS + r) * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  const float* cells = p.cells[dir];
  uint8_t* dzx = reinterpret_cast<uint8_t*>(p.xchg) + (size_t)dir * 2 * CLS * XSLICE;   // [2 parity][8 slices][XSLICE]
  uint8_t* Wlo = sm;                                  // (LOS) [KBN][128 units x 64 k] fp16
  uint8_t* Bs = sm + WLO + (size_t)ch * CSTRIDE;      // my chain's dz slice; after the MMAs: staging of 7 blocks
  float* rbuf = reinterpret_cast<float*>(Bs + CSLICE);   // [CLS src][NB batch][16 units], blocks BST bytes apart

  if (tid == 0) gmax_bits = 0u;
  __syncthreads();
  for (int i = tid; i < p.B; i += 128 * NCH) {
    const float G = __uint_as_float(rowmax[i]);
    if (G > 0.f && G < 3.0e38f) atomicMax(&gmax_bits, __float_as_uint(G));      // positive floats order like their bits
  }
  if (tid == 0) {
    for (int c = 0; c < NCH; ++c) {
      mbar_init(smem_u32(&rx_bar[c]), 1);
      mbar_init(smem_u32(&mma_bar[c]), 1);
      mbar_init(smem_u32(&free_bar[c]), 4 * CLS);     // one arrival per warp of the chain from every CTA of the cluster
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;

  // resident weights -> TMEM (see blstm_cl_bwd8.cu): A[m][kl] = Kh[NC*q + m][g*H + (q'*8 + r)*16 + u],
  // kl = q'*64 + (u/4)*16 + g*4 + u%4.  Warps with the same lane quadrant share the (q', quad) groups.
  {
    const int m = (warp & 3) * 32 + lane;
    const float* wrow = Kh + (size_t)(NC * q + m) * H4;
    const uint32_t tbase = tm + ((uint32_t)((warp & 3) * 32) << 16);
    for (int grp = warp >> 2; grp < KBN * 4; grp += NCH) {
      const int qq = grp >> 2, uq = grp & 3;
      uint32_t vh[8], vl[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + g * H + (qq * CLS + r) * HS + uq * 4);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          __half h0, l0, h1, l1;
          split_h(wv[2 * c], &h0, &l0);
          split_h(wv[2 * c + 1], &h1, &l1);
          vh[g * 2 + c] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          vl[g * 2 + c] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
      }
      tmem_st8_c(tbase + TM_AH + (uint32_t)grp * 8, vh);
      if constexpr (LOS) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint32_t*>(Wlo + (size_t)qq * A_TILE + sw128_h(m, uq * 16 + 2 * c)) = vl[c];
      } else {
        tmem_st8_c(tbase + TM_AL + (uint32_t)grp * 8, vl);
      }
    }
    tmem_st_wait_c();
  }
  if constexpr (LOS) fence_proxy_async_smem();         // generic writes of Wlo -> tensor-core (async proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_arrive();
  cluster_wait();                                      // peers' shared memory and barriers exist before anyone copies into them

  const uint32_t idesc = make_idesc_f16(128, NB);
  const uint32_t Bs_u = smem_u32(Bs), rbuf_u = smem_u32(rbuf);
  const uint32_t rx_u = smem_u32(&rx_bar[ch]), mma_u = smem_u32(&mma_bar[ch]), free_u = smem_u32(&free_bar[ch]);
  const uint32_t tm_d1 = tm + (uint32_t)(ch * 2 * NB), tm_d2 = tm_d1 + NB;
  // pointwise mapping inside the chain: 4 consecutive units (ug) of rows rw + 32*rr (see blstm_cl_bwd8.cu)
  const int ug = (t & 3) * 4, rw = t >> 2;
  float zS = 1.f;                                      // ONE power-of-two scale for the batch (blstm_cl_bwd8.cu)
  {
    const float G = __uint_as_float(gmax_bits);
    if (G > 0.f) {
      int e;
      frexpf(G, &e);
      e = e < -100 ? -100 : (e > 100 ? 100 : e);
      zS = ldexpf(1.f, 6 - e);
    }
    if (p.zinv && blockIdx.x == 0 && tid == 0) *p.zinv = 1.f / zS;
  }
  const float zSi = 1.f / zS;
  int plen[RPT];
  bool rowact[RPT];
#pragma unroll
  for (int rr = 0; rr < RPT; ++rr) {
    const int rl = rw + 32 * rr;
    rowact[rr] = rl < NB;
    const int b = ch * NB + rl;
    plen[rr] = (rowact[rr] && b < p.B) ? p.len[b] : 0;
  }
  float dbacc[4][4], dcc[RPT][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) dcc[rr][u] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) dbacc[g][u] = 0.f;
  }

  int iter = 0;
  for (int s = p.T - 1; s >= 0; --s, ++iter) {
    const uint8_t* dzprev = dzx + (size_t)((iter + 1) & 1) * CLS * XSLICE;
    uint8_t* dznext = dzx + (size_t)(iter & 1) * CLS * XSLICE;
    if (ch == 0) CL_STAMP(iter, 0);
    const unsigned par = (unsigned)(iter - 1) & 1u;
    const uint32_t fl = ll_flag(iter - 1) ? 0x00010001u : 0u;
    // my chain's rows of slice r: chunk (kb, hl, i) of thread t = 16 bytes at tile (kb*2 + hl), sub-tile ch, chunk i*128 + t
    const uint8_t* srcb = dzprev + (size_t)r * XSLICE + (size_t)ch * TILE + (size_t)t * 16;
    auto chunk = [&](int kb, int j) -> const uint4* {      // j in [0, CPB): hi chunks first
      return reinterpret_cast<const uint4*>(srcb + (size_t)(kb * 2 + j / CPH) * XTILE + (size_t)(j % CPH) * 2048);
    };
    uint4 v[2][CPB];
    if (iter > 0) {
      if (t == 0) mbar_expect_tx(rx_u, (CLS - 1) * BLK);
      do { v[0][0] = ld_relaxed_v4(chunk(0, 0)); } while (!ll_ok(v[0][0], fl));
      if (ch == 0) CL_STAMP(iter, 1);
#pragma unroll
      for (int i = 1; i < CPB; ++i) v[0][i] = ld_relaxed_v4(chunk(0, i));
    }
    // ---- pointwise operands: L2 prefetch now (no registers held), the loads themselves behind the MMA issue -----------
    bool valid[RPT];
    int tt[RPT];
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
      const int b = ch * NB + rw + 32 * rr;
      valid[rr] = s < plen[rr];
      tt[rr] = valid[rr] ? (dir ? plen[rr] - 1 - s : s) : s;
      if (valid[rr]) {
        const float* gp = gates + ((size_t)b * p.T + tt[rr]) * H4 + j0 + ug;
#pragma unroll
        for (int g = 0; g < 4; ++g) prefetch_l2(gp + g * H);
        prefetch_l2(cells + ((size_t)b * p.T + tt[rr]) * H + j0 + ug);
        prefetch_l2(p.dy + ((size_t)b * p.yT + tt[rr]) * 2 * H + dir * H + j0 + ug);
      }
    }

    if (iter > 0) {
#pragma unroll
      for (int kb = 0; kb < KBN; ++kb) {
        if (kb + 1 < KBN) {
#pragma unroll
          for (int i = 0; i < CPB; ++i) v[(kb + 1) & 1][i] = ld_relaxed_v4(chunk(kb + 1, i));
        }
#pragma unroll
        for (int i = 0; i < CPB; ++i)
          while (!ll_ok(v[kb & 1][i], fl)) v[kb & 1][i] = ld_relaxed_v4(chunk(kb, i));
        // every CTA's chain has read its receive buffer of the previous step, hence received my blocks: my slice buffer
        // (the copies' source) and the peers' receive buffers are free
        if (kb == 0) mbar_wait_cluster(free_u, par);
#pragma unroll
        for (int i = 0; i < CPB; ++i)
          *reinterpret_cast<uint4*>(Bs + (size_t)(kb * 2 + i / CPH) * TILE + (size_t)(i % CPH) * 2048 + (size_t)t * 16) = v[kb & 1][i];
        fence_proxy_async_smem();
        bar_chain(ch);
        if (wq == 0) {                                 // converged warp; one elected lane issues
          if (kb == 0 && ch == 0) CL_STAMP(iter, 2);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bh = make_desc(Bs_u + (kb * 2 + 0) * TILE + ks * 32, 16, 1024, 2);
            const uint64_t bl = make_desc(Bs_u + (kb * 2 + 1) * TILE + ks * 32, 16, 1024, 2);
            const uint32_t ah = tm + TM_AH + (uint32_t)(kb * 4 + ks) * 8;
            const uint32_t al = tm + TM_AL + (uint32_t)(kb * 4 + ks) * 8;
            const uint32_t acc = (kb | ks) != 0;
            if (elect_one()) {
              umma_f16_ts_c(tm_d1, ah, bh, idesc, acc);
              umma_f16_ts_c(tm_d2, ah, bl, idesc, acc);
              if constexpr (LOS) umma_f16_ss_c(tm_d2, make_desc(smem_u32(Wlo) + kb * A_TILE + ks * 32, 16, 1024, 2), bh, idesc, 1u);
              else umma_f16_ts_c(tm_d2, al, bh, idesc, 1u);
            }
          }
          if (kb == KBN - 1 && elect_one()) umma_commit(mma_u);
        }
        __syncwarp();
      }
    }
    // (the staging registers v[][] are dead here: the operands of the pointwise stage take their place while the tensor
    // pipe, the TMEM drain and the reduce-scatter run)
    // ---- prefetch pointwise operands -------------------------------------------------------------------------
    float gt[RPT][4][4], ct[RPT][4], cprev[RPT][4], dyv[RPT][4];
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
      const int b = ch * NB + rw + 32 * rr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ct[rr][u] = cprev[rr][u] = dyv[rr][u] = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) gt[rr][g][u] = 0.f;
      }
      if (valid[rr]) {
        const float* gp = gates + ((size_t)b * p.T + tt[rr]) * H4 + j0 + ug;
#pragma unroll
        for (int g = 0; g < 4; ++g) ld4c(gp + g * H, gt[rr][g]);
        ld4c(cells + ((size_t)b * p.T + tt[rr]) * H + j0 + ug, ct[rr]);
        if (s > 0) ld4c(cells + ((size_t)b * p.T + (dir ? tt[rr] + 1 : tt[rr] - 1)) * H + j0 + ug, cprev[rr]);
        ld4c(p.dy + ((size_t)b * p.yT + tt[rr]) * 2 * H + dir * H + j0 + ug, dyv[rr]);
      }
    }

    if (iter > 0) {
      mbar_wait(mma_u, par);
      tc_fence_after();
      if (ch == 0) CL_STAMP(iter, 3);
      // ---- TMEM -> owners.  Lane = unit m of the cluster (owner CTA m / 16), columns = my chain's batch rows ------------
      {
        const int m = wq * 32 + lane, d = m >> 4, u = m & 15;
        float* dstcol = (d == r) ? rbuf + (size_t)r * (BST / 4) + u
                                 : reinterpret_cast<float*>(Bs) + (size_t)(d < r ? d : d - 1) * (BST / 4) + u;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        if constexpr (NB >= 32 && NCH == 4) {
          // 512 threads leave 128 registers each: 16 columns at a time (64 live registers of a 32-column drain next to
          // the prefetched pointwise operands spilled to local memory, and a reload behind the polling loads in the LSU
          // queue costs an L2 round trip)
#pragma unroll
          for (int k = 0; k < NB / 16; ++k) {
            uint32_t v1[16], v2[16];
            tmem_ld16(tm_d1 + lane_base + (uint32_t)(k * 16), v1);
            tmem_ld16(tm_d2 + lane_base + (uint32_t)(k * 16), v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              dstcol[(k * 16 + i) * 16] = fmaf(__uint_as_float(v2[i]), 1.f / 2048.f, __uint_as_float(v1[i]));
          }
        } else if constexpr (NB >= 32) {
#pragma unroll
          for (int k = 0; k < NB / 32; ++k) {
            uint32_t v1[32], v2[32];
            tmem_ld32(tm_d1 + lane_base + (uint32_t)(k * 32), v1);
            tmem_ld32(tm_d2 + lane_base + (uint32_t)(k * 32), v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              dstcol[(k * 32 + i) * 16] = fmaf(__uint_as_float(v2[i]), 1.f / 2048.f, __uint_as_float(v1[i]));
          }
        } else {
          uint32_t v1[16], v2[16];
          tmem_ld16(tm_d1 + lane_base, v1);
          tmem_ld16(tm_d2 + lane_base, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) dstcol[i * 16] = fmaf(__uint_as_float(v2[i]), 1.f / 2048.f, __uint_as_float(v1[i]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      bar_chain(ch);
      if (ch == 0) CL_STAMP(iter, 4);
      if (t < CLS && t != r)
        bulk_s2s(map_to_rank(rbuf_u + (uint32_t)r * BST, (uint32_t)t), Bs_u + (uint32_t)(t < r ? t : t - 1) * BST, BLK,
                 map_to_rank(rx_u, (uint32_t)t));
      mbar_wait(rx_u, par);                            // the 7 remote blocks have landed in my buffer
      if (ch == 0) CL_STAMP(iter, 5);
    }
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pin(ct[rr][u]); pin(cprev[rr][u]); pin(dyv[rr][u]);
#pragma unroll
        for (int g = 0; g < 4; ++g) pin(gt[rr][g][u]);
      }

    // ---- pointwise gate gradients for my 16 units --------------------------------------------------------------
    float dh[RPT][4];
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
      for (int u = 0; u < 4; ++u) dh[rr][u] = 0.f;
    if (iter > 0) {
#pragma unroll
      for (int src = 0; src < CLS; ++src)
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr)
          if (rowact[rr]) {
            const float4 x = *reinterpret_cast<const float4*>(rbuf + (size_t)src * (BST / 4) + (rw + 32 * rr) * 16 + ug);
            dh[rr][0] += x.x; dh[rr][1] += x.y; dh[rr][2] += x.z; dh[rr][3] += x.w;
          }
    }
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
      for (int u = 0; u < 4; ++u) dh[rr][u] = fmaf(dh[rr][u], zSi, dyv[rr][u]);
    // my receive buffer is free for the next step once these loads have returned: one arrival per warp on the chain's
    // "free" barrier of every CTA of the cluster (lane d -> CTA d)
    if (s > 0) {
      __syncwarp();
      float dep = 0.f;
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr) dep += (dh[rr][0] + dh[rr][1]) + (dh[rr][2] + dh[rr][3]);
      if (lane < CLS) mbar_arrive_remote(map_to_rank(free_u, (uint32_t)lane), dep);
    }
    float dzv[RPT][4][4];
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dzv[rr][0][u] = dzv[rr][1][u] = dzv[rr][2][u] = dzv[rr][3][u] = 0.f;
        float dcn = 0.f;
        if (valid[rr]) {
          const float ig = gt[rr][0][u], gg = gt[rr][1][u], fg = gt[rr][2][u], og = gt[rr][3][u];
          const float tc_ = tanh_tc(ct[rr][u]);
          const float d_o = dh[rr][u] * tc_;
          const float dc = dcc[rr][u] + dh[rr][u] * og * (1.f - tc_ * tc_);
          dzv[rr][0][u] = dc * gg * ig * (1.f - ig);
          dzv[rr][1][u] = dc * ig * (1.f - gg * gg);
          dzv[rr][2][u] = dc * cprev[rr][u] * fg * (1.f - fg);
          dzv[rr][3][u] = d_o * og * (1.f - og);
          dcn = dc * fg;
        }
        dcc[rr][u] = dcn;
      }
    {
      // dz_t, scaled, split and flagged, into K block q of slice r and (the same words) into the dZ operand planes
      const unsigned short fb = (unsigned short)ll_flag(iter);
      uint8_t* blk = dznext + (size_t)r * XSLICE + (size_t)q * 2 * XTILE;
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr) {
        if (!rowact[rr]) continue;
        uint32_t wh[8], wl[8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          unsigned short hh[4], hl[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            split_h_sat_flag_c(dzv[rr][g][u] * zS, fb, &hh[u], &hl[u]);
            dbacc[g][u] += dzv[rr][g][u];
          }
          wh[2 * g] = (uint32_t)hh[0] | ((uint32_t)hh[1] << 16); wh[2 * g + 1] = (uint32_t)hh[2] | ((uint32_t)hh[3] << 16);
          wl[2 * g] = (uint32_t)hl[0] | ((uint32_t)hl[1] << 16); wl[2 * g + 1] = (uint32_t)hl[2] | ((uint32_t)hl[3] << 16);
        }
        const int row = ch * NB + rw + 32 * rr;               // row of the exchange tile = batch row
        uint8_t* t0 = blk + sw128_h(row, ug * 4);             // k = (ug / 4) * 16: gates 0, 1
        uint8_t* t1 = blk + sw128_h(row, ug * 4 + 8);         // gates 2, 3
        __stcg(reinterpret_cast<uint4*>(t0), make_uint4(wh[0], wh[1], wh[2], wh[3]));
        __stcg(reinterpret_cast<uint4*>(t1), make_uint4(wh[4], wh[5], wh[6], wh[7]));
        __stcg(reinterpret_cast<uint4*>(t0 + XTILE), make_uint4(wl[0], wl[1], wl[2], wl[3]));
        __stcg(reinterpret_cast<uint4*>(t1 + XTILE), make_uint4(wl[4], wl[5], wl[6], wl[7]));
        if (row < p.B) {
          if (p.zh) {
            const size_t zo = ((size_t)row * p.T + tt[rr]) * (2 * H4) + (size_t)dir * H4 + (size_t)(j0 + ug) * 4;
            uint4* zh = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.zh) + zo);
            uint4* zl = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.zl) + zo);
            const uint32_t km = valid[rr] ? 0xFFFFFFFFu : 0u;   // frames past the length: exact zeros
            __stcg(zh, make_uint4(wh[0] & km, wh[1] & km, wh[2] & km, wh[3] & km));
            __stcg(zh + 1, make_uint4(wh[4] & km, wh[5] & km, wh[6] & km, wh[7] & km));
            __stcg(zl, make_uint4(wl[0] & km, wl[1] & km, wl[2] & km, wl[3] & km));
            __stcg(zl + 1, make_uint4(wl[4] & km, wl[5] & km, wl[6] & km, wl[7] & km));
          } else {
            float* gp = gates + ((size_t)row * p.T + tt[rr]) * H4 + j0 + ug;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              __stcg(reinterpret_cast<float4*>(gp + g * H), make_float4(dzv[rr][g][0], dzv[rr][g][1], dzv[rr][g][2], dzv[rr][g][3]));
          }
        }
      }
    }
    if (ch == 0) { CL_STAMP(iter, 6); CL_STAMP(iter, 7); CL_STAMP(iter, 8); CL_STAMP(iter, 9); }
  }

  // bias gradient: fixed-order sum over my chain's rows of every (gate, unit) -> partial slot `ch`
  {
    float* red = reinterpret_cast<float*>(Bs);         // [128 threads][4 gates][4 units]
    bar_chain(ch);
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int u = 0; u < 4; ++u) red[(t * 4 + g) * 4 + u] = dbacc[g][u];
    bar_chain(ch);
    if (t < 4 * HS) {
      const int g = t / HS, j = t % HS;
      float sum = 0.f;
      for (int i = 0; i < 32; ++i) sum += red[((i * 4 + (j >> 2)) * 4 + g) * 4 + (j & 3)];
      p.dbpart[((size_t)dir * 8 + ch) * H4 + g * H + j0 + j] = sum;
    }
  }
  tc_fence_before();
  cluster_arrive();
  cluster_wait();                                      // nobody exits while a peer's copy may still target it
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int NQ, int NB, int NCH>
int launch_chain(const ClParams& p, unsigned* rowmax, cudaStream_t stream, bool* launched, bool rowmax_ready) {
  constexpr int CLS = 8;
  constexpr int CSLICE = NQ * 2 * NB * 128, BST = NB * 16 * 4 + 64;
  constexpr int CSTRIDE = (CSLICE + CLS * BST + 1023) / 1024 * 1024;
  // Every CTA allocates all 512 TMEM columns, so two CTAs of this kernel on one SM deadlock (the second one waits in
  // tcgen05.alloc for the first, which spins on data the second one would produce): the small-batch variants ask for
  // more than half of an SM's shared memory to keep the scheduler from co-locating them.
  constexpr int WLO = NQ > 4 ? NQ * A_TILE : 0;
  const size_t smem = std::max<size_t>(1024 + WLO + (size_t)NCH * CSTRIDE, (size_t)max_smem_optin() / 2 + 2048);
  auto* fn = blstm_rec_bwd_chain_kernel<NQ, NB, NCH>;
  *launched = false;
  if (smem > (size_t)max_smem_optin()) return 0;
  NABU_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * NQ * CLS);
  cfg.blockDim = dim3(128 * NCH);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = coop_attr() ? 2 : 1;
  int nclusters = 0;
  const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg);
  if (getenv("NABU_DEBUG"))
    fprintf(stderr, "[nabu] bwd chain kernel NQ=%d NB=%d NCH=%d: smem %zu B, max active clusters %d (%s), need %d\n", NQ, NB, NCH,
            smem, nclusters, cudaGetErrorString(oe), 2 * NQ);
  // both directions in one launch when their clusters are co-resident (2 NQ clusters of 8), else one after the other
  // (num_units = 1024: 16 clusters of 8 do not fit the 148 SMs)
  int ndir = 2;
  if (oe == cudaSuccess && nclusters < 2 * NQ && nclusters >= NQ) {
    ndir = 1;
    cfg.gridDim = dim3(NQ * CLS);
  } else if (oe != cudaSuccess || nclusters < 2 * NQ) {
    cudaGetLastError();
    return 0;
  }
  if (!rowmax_ready) {
    KernelScope ks("row_absmax", stream);
    row_absmax_kernel<<<dim3(32, p.B), 256, 0, stream>>>(p.dy, p.len, p.yT, 2 * p.H, rowmax);
    NABU_CHECK_LAUNCH();
  }
  ClParams pt = p;
  pt.trace = trace_buffer();
  const unsigned* rm = rowmax;
  for (int d0 = 0; d0 < 2; d0 += ndir) {
    KernelScope ks(NCH == 4 ? "blstm_rec_bwd_chain4" : NCH == 2 ? "blstm_rec_bwd_chain2" : "blstm_rec_bwd_chain1", stream);
    pt.dir0 = d0;
    NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt, rm));
  }
  trace_dump("bwd8c", pt.trace, stream);
  *launched = true;
  return 0;
}

template <int NQ>
int dispatch_chain(const ClParams& p, unsigned* rowmax, cudaStream_t stream, bool* launched, bool rmr) {
  static int force = -2;          // NABU_BWD_CHAINS = "NB x NCH" override, e.g. 32x2 (profiling)
  static int fnb = 0, fnch = 0;
  if (force == -2) {
    force = 0;
    if (const char* e = getenv("NABU_BWD_CHAINS"))
      if (sscanf(e, "%dx%d", &fnb, &fnch) == 2) force = 1;
  }
  // measured on a B200, one cfg-3 layer, T = 600 (tools/smallb_probe.py, profiles/r2_smallb_probe.txt): chains of 16 / 32
  // rows beat one chain of the whole batch below 128 rows (B = 64: 6.1 ms as 2 x 32 against 7.9 ms as 1 x 64)
  // and four chains of 32 beat two of 64 at B = 128 (cfg-3 step 195 ms against 203 ms; 10.5 us per time step alone)
  int nb = p.B <= 32 ? 16 : 32;
  int nch = p.B <= 16 ? 1 : p.B <= 64 ? 2 : 4;
  if constexpr (NQ > 4) {                             // num_units = 1024: at most 32 rows per launch (shared memory)
    NABU_REQUIRE(p.B <= 32, "blstm bwd chain kernel: num_units = 1024 takes at most 32 rows per launch (got %d)", p.B);
    nb = 16; nch = p.B <= 16 ? 1 : 2;
    if (force && fnb * fnch >= p.B && fnb * fnch <= 32) { nb = fnb; nch = fnch; }
    if (nb == 32) return launch_chain<NQ, 32, 1>(p, rowmax, stream, launched, rmr);
    if (nch == 1) return launch_chain<NQ, 16, 1>(p, rowmax, stream, launched, rmr);
    return launch_chain<NQ, 16, 2>(p, rowmax, stream, launched, rmr);
  } else {
  if (force && fnb * fnch >= p.B) { nb = fnb; nch = fnch; }
  if (nb == 32 && nch == 4) return launch_chain<NQ, 32, 4>(p, rowmax, stream, launched, rmr);
  if (nb == 16 && nch == 1) return launch_chain<NQ, 16, 1>(p, rowmax, stream, launched, rmr);
  if (nb == 32 && nch == 1) return launch_chain<NQ, 32, 1>(p, rowmax, stream, launched, rmr);
  if (nb == 64 && nch == 1) return launch_chain<NQ, 64, 1>(p, rowmax, stream, launched, rmr);
  if (nb == 16 && nch == 2) return launch_chain<NQ, 16, 2>(p, rowmax, stream, launched, rmr);
  if (nb == 32 && nch == 2) return launch_chain<NQ, 32, 2>(p, rowmax, stream, launched, rmr);
  return launch_chain<NQ, 64, 2>(p, rowmax, stream, launched, rmr);
  }
}

}  // namespace

bool blstm_bwd_chain_eligible(int B, int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NABU_REC_BWD");
    enabled = (e && (strcmp(e, "flat") == 0 || strcmp(e, "ffma") == 0 || strcmp(e, "cl4") == 0 || strcmp(e, "cl8") == 0)) ? 0 : 1;
  }
  if (!enabled) return false;
  return B > 0 && ((B <= 128 && (H == 256 || H == 512)) || (B <= 32 && H == 1024));
}

int blstm_rec_bwd_chain(const float* const kernel[2], float* const gates[2], const float* const cells[2], const float* dy,
                        float* dbpart, float* xchg, unsigned* rowmax, const int* len, int B, int T, int yT, int D, int H,
                        cudaStream_t stream, bool* launched, int* nslots, void* zh, void* zl, float* zinv, bool rowmax_ready) {
  ClParams p = {};
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.dy = dy; p.dbpart = dbpart; p.xchg = xchg; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  p.zh = zh; p.zl = zl; p.zinv = zinv;
  *nslots = B <= 16 ? 1 : B <= 64 ? 2 : 4;
  if (H == 1024) *nslots = B <= 16 ? 1 : 2;
  if (const char* e = getenv("NABU_BWD_CHAINS")) {
    int a = 0, b = 0;
    if (sscanf(e, "%dx%d", &a, &b) == 2 && a * b >= B) *nslots = b;
  }
  return H == 1024 ? dispatch_chain<8>(p, rowmax, stream, launched, rowmax_ready)
                   : H == 512 ? dispatch_chain<4>(p, rowmax, stream, launched, rowmax_ready) : dispatch_chain<2>(p, rowmax, stream, launched, rowmax_ready);
}

}  // namespace nabu
