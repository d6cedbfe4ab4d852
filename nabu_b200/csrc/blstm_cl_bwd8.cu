// Backward recurrence of a BLSTM layer (row a1; semantics in blstm.cu) on clusters of 8 with the recurrent weights
// RESIDENT IN TENSOR MEMORY:  dh_{t-1}[b, j] = sum_c dz_t[b, c] * Kh[j, c],  c over the 4H gate columns.
//
// Why a second partition.  The cluster-of-4 kernel (blstm_cl_tc.cu) gives every CTA 8 units and a quarter of the dz
// columns: 96 tcgen05.mma of N = 32 per time step and 256 KB of dz per CTA per step.  Its phase trace shows the K loop
// at 4.4 us of an 11 us step and a tcgen05.mma costs ~42 ns whatever its N (24 MMAs of N = 128 take 1.0 us in the
// forward kernel, 96 of N = 32 take 4.3 us here): the instruction count is what has to come down.  Here
//   * a cluster of 8 CTAs owns 128 units (16 per CTA); H/128 clusters per direction, 64 CTAs in all at H = 512;
//   * CTA r multiplies K-slice r (4H/8 dz columns, produced by the rank-r CTAs of the direction's clusters) against
//     its block of Kh for ALL 128 units of the cluster, TRANSPOSED:  D^T[128 units x 128 batch] = W[128 x K] . dz^T.
//     The weights are the A operand and live in TMEM (fp16 hi and lo halves, packed two per 32-bit column, lane =
//     unit: tcgen05.mma's "TS" form), so shared memory holds only the dz slice; 48 MMAs of N = 128 per step at
//     H = 512 instead of 96, and 128 KB of dz per CTA per step over 64 CTAs instead of 256 KB over 128 (L2 traffic / 4);
//   * everything else is the forward kernel's machinery: flag-in-data exchange through L2 (no counters, no release
//     fences), partial sums staged in the dz buffer once the MMAs have consumed it and pushed to their owners by one
//     bulk DSMEM copy per peer that completes on the peer's mbarrier, one relaxed cluster barrier per step for the
//     buffer hand-back, per-row power-of-two scaling of the exchanged gradients, gate math from MUFU.
// TMEM columns (512): D1 [0,128) | D2 [128,256) | W hi [256, 256+K/2) | W lo [256+K/2, 256+K).
#include "cl_tc_common.cuh"
#include "blstm_cl.h"

namespace nabu {
namespace {

constexpr int B8_BLK = 128 * 16 * 4;         // bytes of one (source CTA, destination CTA) block: [128 batch][16 units] fp32
constexpr int B8_BST = B8_BLK + 64;          // block stride in shared memory (+16 banks: the two owners a warp feeds do not collide)

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// saturating split with the exchange flag in the lowest bit of both halves (cl_tc_common.cuh, "LL")
__device__ __forceinline__ void split_h_sat_flag(float x, unsigned short fb, unsigned short* hi, unsigned short* lo) {
  const unsigned short h = (unsigned short)((__half_as_ushort(sat_half(x)) & 0xFFFEu) | fb);
  *hi = h;
  const float res = (x - __half2float(__ushort_as_half(h))) * 2048.f;
  *lo = (unsigned short)((__half_as_ushort(sat_half(res)) & 0xFFFEu) | fb);
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

template <int NQ>                                     // clusters per direction = H / 128 = K blocks per slice
__global__ void __launch_bounds__(CL_THREADS, 1)
blstm_rec_bwd_cluster8_kernel(const ClParams p, const unsigned* __restrict__ rowmax) {
  constexpr int CLS = 8, HS = 16, NC = 128, BT = 128;
  constexpr int KBN = NQ;                             // 64-column K blocks of my slice, one per producer cluster
  constexpr int SLICE = KBN * 2 * A_TILE;             // bytes of the UMMA image of a dz slice (hi | lo per K block)
  constexpr int ACOLS = KBN * 32;                     // TMEM columns of one half of the weights
  constexpr uint32_t TM_D1 = 0, TM_D2 = 128, TM_AH = 256, TM_AL = 256 + ACOLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Bs = sm;                                   // dz slice [KBN][hi|lo][A_TILE]; after the MMAs: staging of 7 blocks
  float* rbuf = reinterpret_cast<float*>(sm + SLICE); // [CLS src][128 batch][16 units], blocks B8_BST bytes apart
  __shared__ __align__(8) uint64_t rx_bar;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned gmax_bits;                      // max over the batch of the rows' max |dy| (the planes' scale)

  const int H = p.H, H4 = 4 * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int per_dir = NQ * CLS;
  const int dir = blockIdx.x / per_dir;
  const int q = (blockIdx.x % per_dir) / CLS;
  const int r = blockIdx.x % CLS;
  const int j0 = (q * CLS + r) * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  const float* cells = p.cells[dir];
  uint8_t* dzx = reinterpret_cast<uint8_t*>(p.xchg) + (size_t)dir * 2 * H4 * BT * 4;   // [2 parity][8 slices][SLICE]

  if (tid == 0) gmax_bits = 0u;
  __syncthreads();
  if (tid < BT) {
    const float G = tid < p.B ? __uint_as_float(rowmax[tid]) : 0.f;
    if (G > 0.f && G < 3.0e38f) atomicMax(&gmax_bits, __float_as_uint(G));      // positive floats order like their bits
  }
  if (tid == 0) {
    mbar_init(smem_u32(&rx_bar), 1);
    mbar_init(smem_u32(&mma_bar), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;

  // resident weights -> TMEM: A[m][kl] = Kh[NC*q + m][g*H + (q'*8 + r)*16 + u],  kl = q'*64 + (u/4)*16 + g*4 + u%4:
  // inside a producer CTA's 64 columns the order is [unit quad][gate][4 units], so that the 16 values a pointwise thread
  // owns (4 gates x 4 units of one row) are 32 contiguous bytes of the exchanged operand -- two 16-byte stores instead
  // of four 8-byte ones (the LSU transaction count bounds the pointwise stage).  One (q', quad) group = 16 consecutive
  // k = 8 packed columns; warps w and w+4 share a lane quadrant.
  {
    const int m = (warp & 3) * 32 + lane;
    const float* wrow = Kh + (size_t)(NC * q + m) * H4;
    const uint32_t tbase = tm + ((uint32_t)((warp & 3) * 32) << 16);
    for (int grp = warp >> 2; grp < KBN * 4; grp += 2) {
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
      tmem_st8(tbase + TM_AH + (uint32_t)grp * 8, vh);
      tmem_st8(tbase + TM_AL + (uint32_t)grp * 8, vl);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_arrive();
  cluster_wait();                                      // peers' shared memory and barriers exist before anyone copies into them

  const uint32_t idesc = make_idesc_f16(128, 128);
  const uint32_t Bs_u = smem_u32(Bs), rbuf_u = smem_u32(rbuf);
  // A thread owns 4 consecutive units (ug) of batch rows rw and rw + 64 at every step: 4 lanes cover the CTA's 64
  // contiguous bytes of a (row, gate) in every global array, so a warp access is 8 rows x 2 full sectors (the LSU
  // transaction count, not bytes, bounded the first version of this kernel: one 16-byte access per lane and line).
  const int ug = (tid & 3) * 4, rw = tid >> 2;
  // ONE power-of-two scale for the exchanged gate gradients and their operand planes: the largest |dy| of the batch lands
  // in [32, 64) (every CTA computes the same value).  Gate derivatives are <= 1, so dz starts below that bound and would
  // have to grow 1000x through the recurrence to reach fp16's maximum (conversions saturate instead of producing inf);
  // values down to 1e-6 of the batch maximum keep the full 22 bits, smaller ones an absolute error of 2^-41 of it.
  float zS = 1.f;
  {
    const float G = __uint_as_float(gmax_bits);
    if (G > 0.f) {
      int e;
      frexpf(G, &e);                                   // G in [2^(e-1), 2^e)
      e = e < -100 ? -100 : (e > 100 ? 100 : e);
      zS = ldexpf(1.f, 6 - e);                         // G * zS in [32, 64)
    }
    if (p.zinv && blockIdx.x == 0 && tid == 0) *p.zinv = 1.f / zS;
  }
  const float zSi = 1.f / zS;
  int plen[2];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int b = rw + 64 * rr;
    plen[rr] = b < p.B ? p.len[b] : 0;
  }
  float dbacc[4][4], dcc[2][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    dcc[0][u] = dcc[1][u] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) dbacc[g][u] = 0.f;
  }

  int iter = 0;
  for (int s = p.T - 1; s >= 0; --s, ++iter) {
    const uint8_t* dzprev = dzx + (size_t)((iter + 1) & 1) * H4 * BT * 4;
    uint8_t* dznext = dzx + (size_t)(iter & 1) * H4 * BT * 4;
    CL_STAMP(iter, 0);
    // ---- first the critical load: my dz slice (self-validating data, see "LL" in cl_tc_common.cuh).  The poll and the
    // first K block's loads are issued BEFORE the pointwise operands' 14 loads per thread, which would otherwise sit in
    // front of them in the LSU queue (~0.9 us on the slowest CTA's path). -------------------------------------------------
    const unsigned par = (unsigned)(iter - 1) & 1u;
    const uint32_t fl = ll_flag(iter - 1) ? 0x00010001u : 0u;
    uint4 v[2][8];
    const uint4* src = reinterpret_cast<const uint4*>(dzprev + (size_t)r * SLICE) + tid;
    if (iter > 0) {
      if (tid == 0) mbar_expect_tx(smem_u32(&rx_bar), (CLS - 1) * B8_BLK);
      do { v[0][0] = ld_relaxed_v4(src); } while (!ll_ok(v[0][0], fl));
      CL_STAMP(iter, 1);
#pragma unroll
      for (int i = 1; i < 8; ++i) v[0][i] = ld_relaxed_v4(src + i * CL_THREADS);
    }
    // ---- prefetch pointwise operands -------------------------------------------------------------------------
    float gt[2][4][4], ct[2][4], cprev[2][4], dyv[2][4];
    bool valid[2];
    int tt[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int b = rw + 64 * rr;
      valid[rr] = s < plen[rr];
      tt[rr] = valid[rr] ? (dir ? plen[rr] - 1 - s : s) : s;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ct[rr][u] = cprev[rr][u] = dyv[rr][u] = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) gt[rr][g][u] = 0.f;
      }
      if (valid[rr]) {
        const float* gp = gates + ((size_t)b * p.T + tt[rr]) * H4 + j0 + ug;
#pragma unroll
        for (int g = 0; g < 4; ++g) ld4(gp + g * H, gt[rr][g]);
        ld4(cells + ((size_t)b * p.T + tt[rr]) * H + j0 + ug, ct[rr]);
        if (s > 0) ld4(cells + ((size_t)b * p.T + (dir ? tt[rr] + 1 : tt[rr] - 1)) * H + j0 + ug, cprev[rr]);
        ld4(p.dy + ((size_t)b * p.yT + tt[rr]) * 2 * H + dir * H + j0 + ug, dyv[rr]);
      }
    }

    if (iter > 0) {
      // One K block (hi and lo tile, 8 chunks of 16 bytes per thread) at a time, the next block's loads in flight while
      // this one is validated, stored and multiplied.
#pragma unroll
      for (int kb = 0; kb < KBN; ++kb) {
        const uint4* cur = src + (size_t)kb * (2 * A_TILE / 16);
        if (kb + 1 < KBN) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[(kb + 1) & 1][i] = ld_relaxed_v4(cur + (2 * A_TILE / 16) + i * CL_THREADS);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          while (!ll_ok(v[kb & 1][i], fl)) v[kb & 1][i] = ld_relaxed_v4(cur + i * CL_THREADS);
        // peers have read their receive buffers of the previous step, hence received my blocks: Bs and theirs are free
        if (kb == 0) cluster_wait();
        uint4* dst = reinterpret_cast<uint4*>(Bs + (size_t)kb * 2 * A_TILE) + tid;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i * CL_THREADS] = v[kb & 1][i];
        fence_proxy_async_smem();
        __syncthreads();
        if (warp_u == 0) {                             // converged warp; one elected lane issues
          if (kb == 0) CL_STAMP(iter, 2);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bh = make_desc(Bs_u + (kb * 2 + 0) * A_TILE + ks * 32, 16, 1024, 2);
            const uint64_t bl = make_desc(Bs_u + (kb * 2 + 1) * A_TILE + ks * 32, 16, 1024, 2);
            const uint32_t ah = tm + TM_AH + (uint32_t)(kb * 4 + ks) * 8;
            const uint32_t al = tm + TM_AL + (uint32_t)(kb * 4 + ks) * 8;
            const uint32_t acc = (kb | ks) != 0;
            if (elect_one()) {
              umma_f16_ts(tm + TM_D1, ah, bh, idesc, acc);
              umma_f16_ts(tm + TM_D2, ah, bl, idesc, acc);
              umma_f16_ts(tm + TM_D2, al, bh, idesc, 1u);
            }
          }
          if (kb == KBN - 1 && elect_one()) umma_commit(smem_u32(&mma_bar));
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&mma_bar), par);
      tc_fence_after();
      CL_STAMP(iter, 3);
      // ---- TMEM -> owners.  Lane = unit m of the cluster (owner CTA m / 16), columns = batch rows.  My own block goes
      // straight into my receive buffer, the 7 others are staged in Bs (the MMAs have consumed it) and pushed by one bulk
      // DSMEM copy each that completes on the owner's mbarrier. ----------------------------------------------------------
      {
        const int lq = warp & 3, chh = warp >> 2;
        const int m = lq * 32 + lane, d = m >> 4, u = m & 15;
        float* dstcol = (d == r) ? rbuf + (size_t)r * (B8_BST / 4) + u
                                 : reinterpret_cast<float*>(Bs) + (size_t)(d < r ? d : d - 1) * (B8_BST / 4) + u;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c0 = chh * 64 + k * 32;
          uint32_t v1[32], v2[32];
          const uint32_t taddr = tm + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
          tmem_ld32(taddr + TM_D1, v1);
          tmem_ld32(taddr + TM_D2, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            dstcol[(c0 + i) * 16] = fmaf(__uint_as_float(v2[i]), 1.f / 2048.f, __uint_as_float(v1[i]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      CL_STAMP(iter, 4);
      if (tid < CLS && tid != r)
        bulk_s2s(map_to_rank(rbuf_u + (uint32_t)r * B8_BST, (uint32_t)tid), Bs_u + (uint32_t)(tid < r ? tid : tid - 1) * B8_BST,
                 B8_BLK, map_to_rank(smem_u32(&rx_bar), (uint32_t)tid));
      mbar_wait(smem_u32(&rx_bar), par);               // the 7 remote blocks have landed in my buffer
      CL_STAMP(iter, 5);
    }

    // ---- pointwise gate gradients for my 16 units --------------------------------------------------------------
    float dh[2][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int u = 0; u < 4; ++u) dh[rr][u] = 0.f;
    if (iter > 0) {
#pragma unroll
      for (int src = 0; src < CLS; ++src)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(rbuf + (size_t)src * (B8_BST / 4) + (rw + 64 * rr) * 16 + ug);
          dh[rr][0] += v.x; dh[rr][1] += v.y; dh[rr][2] += v.z; dh[rr][3] += v.w;
        }
    }
    // the exchanged gradients carry the row's power-of-two scale: divide it out (exact) and add this layer's dy
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int u = 0; u < 4; ++u) dh[rr][u] = fmaf(dh[rr][u], zSi, dyv[rr][u]);
    // my receive buffer is free for the next step once these loads have returned (nothing to publish: relaxed)
    if (s > 0) cluster_arrive_relaxed();
    float dzv[2][4][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
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
      // dz_t, scaled, split and flagged: the thread's 16 values of a row (k = quad*16 + g*4 + u) are two 16-byte chunks of
      // the hi tile and two of the lo tile of K block q of slice r (all 128 rows: the consumers wait for every half).
      // The SAME words are the row's piece of the dZ operand planes of the weight-gradient / dX GEMMs (column order
      // [unit quad][gate][4 units] inside a direction, undone by the host: blstm.cu), so no pass ever re-reads dZ; the
      // exchange flag stays in the lowest bit of both halves there (hi + lo / 2048 still carries dz to 2^-20).
      const unsigned short fb = (unsigned short)ll_flag(iter);
      uint8_t* blk = dznext + (size_t)r * SLICE + (size_t)q * 2 * A_TILE;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        uint32_t wh[8], wl[8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          unsigned short hh[4], hl[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            split_h_sat_flag(dzv[rr][g][u] * zS, fb, &hh[u], &hl[u]);
            dbacc[g][u] += dzv[rr][g][u];
          }
          wh[2 * g] = (uint32_t)hh[0] | ((uint32_t)hh[1] << 16); wh[2 * g + 1] = (uint32_t)hh[2] | ((uint32_t)hh[3] << 16);
          wl[2 * g] = (uint32_t)hl[0] | ((uint32_t)hl[1] << 16); wl[2 * g + 1] = (uint32_t)hl[2] | ((uint32_t)hl[3] << 16);
        }
        const int row = rw + 64 * rr;
        uint8_t* t0 = blk + sw128_h(row, ug * 4);             // k = (ug / 4) * 16: gates 0, 1
        uint8_t* t1 = blk + sw128_h(row, ug * 4 + 8);         // gates 2, 3
        __stcg(reinterpret_cast<uint4*>(t0), make_uint4(wh[0], wh[1], wh[2], wh[3]));
        __stcg(reinterpret_cast<uint4*>(t1), make_uint4(wh[4], wh[5], wh[6], wh[7]));
        __stcg(reinterpret_cast<uint4*>(t0 + A_TILE), make_uint4(wl[0], wl[1], wl[2], wl[3]));
        __stcg(reinterpret_cast<uint4*>(t1 + A_TILE), make_uint4(wl[4], wl[5], wl[6], wl[7]));
        if (p.zh && row < p.B) {
          const size_t zo = ((size_t)row * p.T + tt[rr]) * (2 * H4) + (size_t)dir * H4 + (size_t)(j0 + ug) * 4;
          uint4* zh = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.zh) + zo);
          uint4* zl = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.zl) + zo);
          // frames past the utterance's length: exact zeros (a flagged zero is -2^-34, and dX must be 0 there exactly)
          const uint32_t km = valid[rr] ? 0xFFFFFFFFu : 0u;
          __stcg(zh, make_uint4(wh[0] & km, wh[1] & km, wh[2] & km, wh[3] & km));
          __stcg(zh + 1, make_uint4(wh[4] & km, wh[5] & km, wh[6] & km, wh[7] & km));
          __stcg(zl, make_uint4(wl[0] & km, wl[1] & km, wl[2] & km, wl[3] & km));
          __stcg(zl + 1, make_uint4(wl[4] & km, wl[5] & km, wl[6] & km, wl[7] & km));
        }
      }
    }
    CL_STAMP(iter, 6); CL_STAMP(iter, 7); CL_STAMP(iter, 8); CL_STAMP(iter, 9);
    // ---- without operand planes: the fp32 dz into gates[] for the split passes / fp32 GEMMs of the host -----------------
    if (!p.zh) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int b = rw + 64 * rr;
        if (b < p.B) {
          float* gp = gates + ((size_t)b * p.T + tt[rr]) * H4 + j0 + ug;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            __stcg(reinterpret_cast<float4*>(gp + g * H), make_float4(dzv[rr][g][0], dzv[rr][g][1], dzv[rr][g][2], dzv[rr][g][3]));
        }
      }
    }
  }

  // bias gradient: fixed-order sum over the 128 rows of every (gate, unit)
  {
    float* red = reinterpret_cast<float*>(Bs);         // [256 threads][4 gates][4 units]
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int u = 0; u < 4; ++u) red[(tid * 4 + g) * 4 + u] = dbacc[g][u];
    __syncthreads();
    if (tid < 4 * HS) {
      const int g = tid / HS, j = tid % HS;
      float sum = 0.f;
      for (int i = 0; i < 64; ++i) sum += red[((i * 4 + (j >> 2)) * 4 + g) * 4 + (j & 3)];
      p.dbpart[((size_t)dir * 8) * H4 + g * H + j0 + j] = sum;
    }
  }
  tc_fence_before();
  cluster_arrive();
  cluster_wait();                                      // nobody exits while a peer's copy may still target it
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int NQ>
int launch_bwd8(const ClParams& p, unsigned* rowmax, cudaStream_t stream, bool* launched) {
  constexpr int CLS = 8;
  const size_t smem = 1024 + (size_t)NQ * 2 * A_TILE + (size_t)CLS * B8_BST;
  auto* fn = blstm_rec_bwd_cluster8_kernel<NQ>;
  *launched = false;
  if (smem > (size_t)max_smem_optin()) return 0;
  NABU_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * NQ * CLS);
  cfg.blockDim = dim3(CL_THREADS);
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
    fprintf(stderr, "[nabu] bwd cluster-of-8 TMEM-resident kernel NQ=%d: smem %zu B, max active clusters %d (%s), need %d\n", NQ,
            smem, nclusters, cudaGetErrorString(oe), 2 * NQ);
  if (oe != cudaSuccess || nclusters < 2 * NQ) {
    cudaGetLastError();
    return 0;
  }
  {
    KernelScope ks("row_absmax", stream);
    row_absmax_kernel<<<dim3(32, p.B), 256, 0, stream>>>(p.dy, p.len, p.yT, 2 * p.H, rowmax);
    NABU_CHECK_LAUNCH();
  }
  KernelScope ks("blstm_rec_bwd_cluster8", stream);
  ClParams pt = p;
  pt.trace = trace_buffer();
  const unsigned* rm = rowmax;
  NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt, rm));
  trace_dump("bwd8", pt.trace, stream);
  *launched = true;
  return 0;
}

}  // namespace

bool blstm_bwd_cluster8_eligible(int B, int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NABU_REC_BWD");
    enabled = (e && (strcmp(e, "flat") == 0 || strcmp(e, "ffma") == 0 || strcmp(e, "cl4") == 0)) ? 0 : 1;
  }
  if (!enabled) return false;
  return B <= 128 && B > 0 && (H == 256 || H == 512);
}

int blstm_rec_bwd_cluster8(const float* const kernel[2], float* const gates[2], const float* const cells[2],
                           const float* dy, float* dbpart, float* xchg, unsigned* rowmax, const int* len, int B, int T, int yT,
                           int D, int H, cudaStream_t stream, bool* launched, void* zh, void* zl, float* zinv) {
  ClParams p = {};
  p.zh = zh; p.zl = zl; p.zinv = zinv;
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.dy = dy; p.dbpart = dbpart; p.xchg = xchg; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  return H == 512 ? launch_bwd8<4>(p, rowmax, stream, launched) : launch_bwd8<2>(p, rowmax, stream, launched);
}

}  // namespace nabu
