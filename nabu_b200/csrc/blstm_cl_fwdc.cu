// Forward recurrence of a BLSTM layer for SMALL batches (<= 64 rows per GPU: the strong-scaling split of a minibatch, the
// decode batches), "chains" version with the product transposed -- the forward twin of blstm_cl_bwd8c.cu.
//
// blstm_cl_tc.cu multiplies h[128 batch x K] (the A operand, M = 128) against the resident weight block: whatever the
// batch, a time step moves and multiplies a 128-row tile.  Here the roles are swapped,
//     z^T[128 gate columns x NB batch] = W^T[128 x K] . h^T[K x NB],
// the weights are the A operand and live in TENSOR MEMORY (fp16 hi | lo halves packed two per 32-bit column, lane = gate
// column: tcgen05.mma's TS form), the batch is the MMA's N = NB in {16, 32}: exchange volume, DSMEM volume, TMEM drain
// and pointwise work all shrink with the batch.  Same partition as blstm_cl_tc.cu otherwise: clusters of 4, a cluster
// owns 32 units = 128 gate columns, CTA r multiplies K-slice r of h (H/4 units) and owns 8 units in the pointwise
// stage; flag-in-data exchange of h through L2, partial sums reduce-scattered by bulk DSMEM copies.  One or two
// independent chains of NB rows per CTA, each run by its own warpgroup (see blstm_cl_bwd8c.cu).
// TMEM columns (512): accumulators of chain c at [c*2*NB, (c+1)*2*NB) (D1 | D2) | W hi [256, 256+KS/2) | W lo [.., 256+KS).
#include "cl_tc_common.cuh"
#include "blstm_cl.h"
#include <algorithm>

namespace nabu {
namespace {

__device__ __forceinline__ void umma_f16_ts_f(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st8_f(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait_f() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bar_chain_f(int ch) { asm volatile("bar.sync %0, 128;" ::"r"(ch + 1) : "memory"); }
// relaxed, ordered behind the loads of the receive buffer by the data dependency on `dep` (see blstm_cl_bwd8c.cu)
__device__ __forceinline__ void mbar_arrive_remote_f(uint32_t bar_cluster, float dep) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster), "f"(dep) : "memory");
}
__device__ __forceinline__ void pin_f(float& x) { asm volatile("" : "+f"(x)); }
// N = 1, 2 or 4 consecutive floats / packed fp16 values (the pointwise stage gives every thread NB / 16 units of a row)
template <int N> __device__ __forceinline__ void ldcg_n(const float* p, float (&v)[N]) {
  if constexpr (N == 4) { const float4 a = __ldcg(reinterpret_cast<const float4*>(p)); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
  else if constexpr (N == 2) { const float2 a = __ldcg(reinterpret_cast<const float2*>(p)); v[0] = a.x; v[1] = a.y; }
  else v[0] = __ldcg(p);
}
template <int N> __device__ __forceinline__ void lds_add_n(const float* p, float (&v)[N]) {
  if constexpr (N == 4) { const float4 a = *reinterpret_cast<const float4*>(p); v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; }
  else if constexpr (N == 2) { const float2 a = *reinterpret_cast<const float2*>(p); v[0] += a.x; v[1] += a.y; }
  else v[0] += *p;
}
template <int N> __device__ __forceinline__ void stcg_n(float* p, const float (&v)[N]) {
  if constexpr (N == 4) __stcg(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  else if constexpr (N == 2) __stcg(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  else __stcg(p, v[0]);
}
template <int N> __device__ __forceinline__ void stcg_h(void* p, const unsigned short (&h)[N]) {
  if constexpr (N == 4) __stcg(reinterpret_cast<uint2*>(p), make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16)));
  else if constexpr (N == 2) __stcg(reinterpret_cast<unsigned*>(p), (uint32_t)h[0] | ((uint32_t)h[1] << 16));
  else __stcg(reinterpret_cast<unsigned short*>(p), h[0]);
}
__device__ __forceinline__ void mbar_wait_cluster_f(uint32_t bar, uint32_t parity) {
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

template <int KB, int NB, int NCH>                    // 64-unit K blocks per slice (H / 256), batch rows per chain, chains
__global__ void __launch_bounds__(128 * NCH, 1)
blstm_rec_fwd_chain_kernel(const ClParams p) {
  constexpr int CLS = 4, HS = 8, NC = 32, BT = NB * NCH;
  constexpr int KS = 64 * KB;                         // h units per K-slice
  constexpr int TILE = NB * 128;                      // bytes of a chain's [NB rows x 64 fp16] K-major tile
  constexpr int XTILE = BT * 128;
  constexpr int CSLICE = KB * 2 * TILE;
  constexpr int XSLICE = KB * 2 * XTILE;
  constexpr int BLK = NB * 32 * 4;                    // one (source, destination) block: [NB batch][32 gate columns] fp32
  constexpr int BST = BLK + 64;
  constexpr int SBUF = CSLICE > (CLS - 1) * BST ? CSLICE : (CLS - 1) * BST;      // slice image, then staging of 3 blocks
  constexpr int CSTRIDE = (SBUF + CLS * BST + 1023) / 1024 * 1024;
  constexpr int ACOLS = KS / 2;                       // TMEM columns of one half of the weights
  constexpr uint32_t TM_AH = 256, TM_AL = 256 + ACOLS;
  constexpr int CPB = NB / 8, CPH = CPB / 2;          // 16-byte chunks per thread and K block / per tile
  static_assert(NB == 16 || NB == 32 || NB == 64, "NB");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t rx_bar[NCH];
  __shared__ __align__(8) uint64_t mma_bar[NCH];
  __shared__ __align__(8) uint64_t free_bar[NCH];
  __shared__ uint32_t tmem_slot;

  const int H = p.H, H4 = 4 * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ch = __shfl_sync(0xffffffffu, tid >> 7, 0);
  const int t = tid & 127, wq = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int per_dir = H / HS;                          // CTAs per direction
  const int dir = p.dir0 + blockIdx.x / per_dir;
  const int q = (blockIdx.x % per_dir) / CLS;
  const int r = blockIdx.x % CLS;
  const int j0 = (q * CLS + r) * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  float* cells = const_cast<float*>(p.cells[dir]);
  uint8_t* hx = reinterpret_cast<uint8_t*>(p.xchg) + (size_t)dir * 2 * CLS * XSLICE;   // [2 parity][4 slices][XSLICE]
  uint8_t* Bs = sm + (size_t)ch * CSTRIDE;
  float* rbuf = reinterpret_cast<float*>(Bs + SBUF);   // [CLS src][NB batch][32 gate columns], blocks BST bytes apart

  if (tid == 0) {
    for (int c = 0; c < NCH; ++c) {
      mbar_init(smem_u32(&rx_bar[c]), 1);
      mbar_init(smem_u32(&mma_bar[c]), 1);
      mbar_init(smem_u32(&free_bar[c]), 4 * CLS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;

  // resident weights -> TMEM: lane m = gate column n = d*32 + g*8 + u of the cluster (owner d, gate g, unit u), i.e.
  // column g*H + 32*q + d*8 + u of Kh; k = h unit r*KS + k.  8 packed columns (16 consecutive k) per tcgen05.st.
  {
    const int m = (warp & 3) * 32 + lane;
    const int d = m >> 5, g = (m >> 3) & 3, u = m & 7;
    const float* wcol = Kh + (size_t)(r * KS) * H4 + g * H + NC * q + d * HS + u;
    const uint32_t tbase = tm + ((uint32_t)((warp & 3) * 32) << 16);
    for (int grp = warp >> 2; grp < KS / 16; grp += NCH) {
      uint32_t vh[8], vl[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        __half h0, l0, h1, l1;
        split_h(wcol[(size_t)(grp * 16 + 2 * c) * H4], &h0, &l0);
        split_h(wcol[(size_t)(grp * 16 + 2 * c + 1) * H4], &h1, &l1);
        vh[c] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        vl[c] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      tmem_st8_f(tbase + TM_AH + (uint32_t)grp * 8, vh);
      tmem_st8_f(tbase + TM_AL + (uint32_t)grp * 8, vl);
    }
    tmem_st_wait_f();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_arrive();
  cluster_wait();

  const uint32_t idesc = make_idesc_f16(128, NB);
  const uint32_t Bs_u = smem_u32(Bs), rbuf_u = smem_u32(rbuf);
  const uint32_t rx_u = smem_u32(&rx_bar[ch]), mma_u = smem_u32(&mma_bar[ch]), free_u = smem_u32(&free_bar[ch]);
  const uint32_t tm_d1 = tm + (uint32_t)(ch * 2 * NB), tm_d2 = tm_d1 + NB;
  // pointwise: all 128 threads of the chain, thread = (row, UPT of the CTA's 8 units).  (With 4 units per thread whatever
  // NB, half of the threads idled at NB = 32 and the stage is latency-bound on the per-thread instruction chain.)
  constexpr int UPT = NB >= 64 ? 4 : NB / 16, TPR = HS / UPT;
  const int prl = t / TPR, pu = (t % TPR) * UPT;
  const bool pact = prl < NB;
  const int pb = ch * NB + prl;                        // batch row
  const bool prow = pact && pb < p.B;
  const int plen = prow ? p.len[pb] : 0;
  float ccarry[UPT];
#pragma unroll
  for (int u = 0; u < UPT; ++u) ccarry[u] = 0.f;

  for (int s = 0; s < p.T; ++s) {
    const uint8_t* hprev = hx + (size_t)((s + 1) & 1) * CLS * XSLICE;
    uint8_t* hnext = hx + (size_t)(s & 1) * CLS * XSLICE;
    if (ch == 0) CL_STAMP(s, 0);
    float gx[4][UPT];
    const bool valid = s < plen;
    const int tt = valid ? (dir ? plen - 1 - s : s) : s;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int u = 0; u < UPT; ++u) gx[g][u] = 0.f;
    auto prefetch_gx = [&]() {
      if (valid) {
        const float* gp = gates + ((size_t)pb * p.T + tt) * H4 + j0 + pu;
#pragma unroll
        for (int g = 0; g < 4; ++g) ldcg_n<UPT>(gp + g * H, gx[g]);
      }
    };
    if (s == 0) prefetch_gx();

    if (s > 0) {
      const unsigned par = (unsigned)(s - 1) & 1u;
      const uint32_t fl = ll_flag(s - 1) ? 0x00010001u : 0u;
      if (t == 0) mbar_expect_tx(rx_u, (CLS - 1) * BLK);
      const uint8_t* srcb = hprev + (size_t)r * XSLICE + (size_t)ch * TILE + (size_t)t * 16;
      auto chunk = [&](int kb, int j) -> const uint4* {
        return reinterpret_cast<const uint4*>(srcb + (size_t)(kb * 2 + j / CPH) * XTILE + (size_t)(j % CPH) * 2048);
      };
      uint4 v[KB][CPB];
      do { v[0][0] = ld_relaxed_v4(chunk(0, 0)); } while (!ll_ok(v[0][0], fl));
      if (ch == 0) CL_STAMP(s, 1);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int i = 0; i < CPB; ++i)
          if (kb | i) v[kb][i] = ld_relaxed_v4(chunk(kb, i));
      prefetch_gx();
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int i = 0; i < CPB; ++i)
          while (!ll_ok(v[kb][i], fl)) v[kb][i] = ld_relaxed_v4(chunk(kb, i));
        if (kb == 0) mbar_wait_cluster_f(free_u, par);          // slice / staging buffer and the peers' receive buffers are free
#pragma unroll
        for (int i = 0; i < CPB; ++i)
          *reinterpret_cast<uint4*>(Bs + (size_t)(kb * 2 + i / CPH) * TILE + (size_t)(i % CPH) * 2048 + (size_t)t * 16) = v[kb][i];
        fence_proxy_async_smem();
        bar_chain_f(ch);
        if (wq == 0) {
          if (kb == 0 && ch == 0) CL_STAMP(s, 2);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bh = make_desc(Bs_u + (kb * 2 + 0) * TILE + ks * 32, 16, 1024, 2);
            const uint64_t bl = make_desc(Bs_u + (kb * 2 + 1) * TILE + ks * 32, 16, 1024, 2);
            const uint32_t ah = tm + TM_AH + (uint32_t)(kb * 4 + ks) * 8;
            const uint32_t al = tm + TM_AL + (uint32_t)(kb * 4 + ks) * 8;
            const uint32_t acc = (kb | ks) != 0;
            if (elect_one()) {
              umma_f16_ts_f(tm_d1, ah, bh, idesc, acc);
              umma_f16_ts_f(tm_d2, ah, bl, idesc, acc);
              umma_f16_ts_f(tm_d2, al, bh, idesc, 1u);
            }
          }
          if (kb == KB - 1 && elect_one()) umma_commit(mma_u);
        }
        __syncwarp();
      }
      mbar_wait(mma_u, par);
      tc_fence_after();
      if (ch == 0) CL_STAMP(s, 3);
      // ---- TMEM -> owners: warp wq holds the 32 gate columns of owner CTA wq; columns of TMEM = my chain's batch rows ----
      {
        const int d = wq;
        float* dstcol = (d == r) ? rbuf + (size_t)r * (BST / 4) + lane
                                 : reinterpret_cast<float*>(Bs) + (size_t)(d < r ? d : d - 1) * (BST / 4) + lane;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        if constexpr (NB >= 32) {
#pragma unroll
          for (int k = 0; k < NB / 32; ++k) {
            uint32_t v1[32], v2[32];
            tmem_ld32(tm_d1 + lane_base + (uint32_t)(k * 32), v1);
            tmem_ld32(tm_d2 + lane_base + (uint32_t)(k * 32), v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              dstcol[(k * 32 + i) * 32] = fmaf(__uint_as_float(v2[i]), 1.f / 2048.f, __uint_as_float(v1[i]));
          }
        } else {
          uint32_t v1[16], v2[16];
          tmem_ld16(tm_d1 + lane_base, v1);
          tmem_ld16(tm_d2 + lane_base, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) dstcol[i * 32] = fmaf(__uint_as_float(v2[i]), 1.f / 2048.f, __uint_as_float(v1[i]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      bar_chain_f(ch);
      if (ch == 0) CL_STAMP(s, 4);
      if (t < CLS && t != r)
        bulk_s2s(map_to_rank(rbuf_u + (uint32_t)r * BST, (uint32_t)t), Bs_u + (uint32_t)(t < r ? t : t - 1) * BST, BLK,
                 map_to_rank(rx_u, (uint32_t)t));
      mbar_wait(rx_u, par);
      if (ch == 0) CL_STAMP(s, 5);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int u = 0; u < UPT; ++u) pin_f(gx[g][u]);

    // ---- pointwise cell update for my 8 units -------------------------------------------------------------------
    float av[5][UPT], hn[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) hn[u] = 0.f;
    if (s > 0 && pact) {
#pragma unroll
      for (int src = 0; src < CLS; ++src)
#pragma unroll
        for (int g = 0; g < 4; ++g) lds_add_n<UPT>(rbuf + (size_t)src * (BST / 4) + prl * 32 + g * 8 + pu, gx[g]);
    }
    if (s + 1 < p.T) {                                 // my receive buffer is free once these loads have returned
      __syncwarp();
      const float dep = (gx[0][0] + gx[1][0]) + (gx[2][0] + gx[3][0]);
      if (lane < CLS) mbar_arrive_remote_f(map_to_rank(free_u, (uint32_t)lane), dep);
    }
    if (prow) {
#pragma unroll
      for (int u = 0; u < UPT; ++u) {
        const float ig = sigmoid_tc(gx[0][u]);
        const float gg = tanh_tc(gx[1][u]);
        const float fg = sigmoid_tc(gx[2][u] + 1.0f);
        const float og = sigmoid_tc(gx[3][u]);
        const float cn = ccarry[u] * fg + ig * gg;
        av[0][u] = ig; av[1][u] = gg; av[2][u] = fg; av[3][u] = og; av[4][u] = cn;
        if (valid) {
          hn[u] = tanh_tc(cn) * og;
          ccarry[u] = cn;
        }
      }
    }
    if (pact) {
      // h_t, split, flagged, in the consumers' UMMA layout (every row of the tile: rows b >= B as zeros)
      const unsigned short fb = (unsigned short)ll_flag(s);
      unsigned short hh[UPT], hl[UPT];
#pragma unroll
      for (int u = 0; u < UPT; ++u) split_h_flag(hn[u], fb, &hh[u], &hl[u]);
      const int j = j0 + pu;
      uint8_t* tp = hnext + (size_t)(j / KS) * XSLICE + (size_t)((j % KS) / 64) * 2 * XTILE + sw128_h(pb, j % 64);
      stcg_h<UPT>(tp, hh);
      stcg_h<UPT>(tp + XTILE, hl);
    }
    if (ch == 0) { CL_STAMP(s, 6); CL_STAMP(s, 7); CL_STAMP(s, 8); CL_STAMP(s, 9); }
    // ---- off the critical path: what the backward pass and the next layer need --------------------------------------
    if (prow) {
      if (valid) {
        float* gp = gates + ((size_t)pb * p.T + tt) * H4 + j0 + pu;
        float* cp = cells + ((size_t)pb * p.T + tt) * H + j0 + pu;
#pragma unroll
        for (int g = 0; g < 4; ++g) stcg_n<UPT>(gp + g * H, av[g]);
        stcg_n<UPT>(cp, av[4]);
      }
      const size_t yo = ((size_t)pb * p.yT + tt) * 2 * H + dir * H + j0 + pu;
      stcg_n<UPT>(p.y + yo, hn);
      if (p.yh) {
        unsigned short ph[UPT], pl[UPT];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
          __half hi, lo;
          split_h(hn[u] * Y_PLANE_SCALE, &hi, &lo);
          ph[u] = __half_as_ushort(hi); pl[u] = __half_as_ushort(lo);
        }
        stcg_h<UPT>(reinterpret_cast<__half*>(p.yh) + yo, ph);
        stcg_h<UPT>(reinterpret_cast<__half*>(p.yl) + yo, pl);
      }
    }
  }
  tc_fence_before();
  cluster_arrive();
  cluster_wait();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int KB, int NB, int NCH>
int launch_fwd_chain(const ClParams& p, cudaStream_t stream, bool* launched) {
  constexpr int CLS = 4;
  constexpr int CSLICE = KB * 2 * NB * 128, BST = NB * 32 * 4 + 64;
  constexpr int SBUF = CSLICE > (CLS - 1) * BST ? CSLICE : (CLS - 1) * BST;
  constexpr int CSTRIDE = (SBUF + CLS * BST + 1023) / 1024 * 1024;
  // one CTA per SM (every CTA allocates all of TMEM, see blstm_cl_bwd8c.cu)
  const size_t smem = std::max<size_t>(1024 + (size_t)NCH * CSTRIDE, (size_t)max_smem_optin() / 2 + 2048);
  auto* fn = blstm_rec_fwd_chain_kernel<KB, NB, NCH>;
  *launched = false;
  if (smem > (size_t)max_smem_optin()) return 0;
  NABU_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // num_units = 1024: the 2 x 128 CTAs of both directions do not fit the GPU, the directions run one after the other
  const int ndir = 2 * (p.H / 8) <= num_sms() ? 2 : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ndir * (p.H / 8));
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
    fprintf(stderr, "[nabu] fwd chain kernel KB=%d NB=%d NCH=%d: smem %zu B, max active clusters %d (%s), need %d\n", KB, NB, NCH,
            smem, nclusters, cudaGetErrorString(oe), (int)cfg.gridDim.x / CLS);
  if (oe != cudaSuccess || nclusters * CLS < (int)cfg.gridDim.x) {
    cudaGetLastError();
    return 0;
  }
  ClParams pt = p;
  pt.trace = trace_buffer();
  for (int d0 = 0; d0 < 2; d0 += ndir) {
    KernelScope ks(NCH == 4 ? "blstm_rec_fwd_chain4" : NCH == 2 ? "blstm_rec_fwd_chain2" : "blstm_rec_fwd_chain1", stream);
    pt.dir0 = d0;
    NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt));
  }
  trace_dump("fwdc", pt.trace, stream);
  *launched = true;
  return 0;
}

template <int KB>
int dispatch_fwd_chain(const ClParams& p, cudaStream_t stream, bool* launched) {
  static int force = -2, fnb = 0, fnch = 0;           // NABU_FWD_CHAINS = "NB x NCH" override (profiling)
  if (force == -2) {
    force = 0;
    if (const char* e = getenv("NABU_FWD_CHAINS"))
      if (sscanf(e, "%dx%d", &fnb, &fnch) == 2) force = 1;
  }
  // measured (tools/smallb_probe.py, one cfg-3 layer, T = 600): B = 16: 2.9 ms as 16x1 (4.7 ms on the 128-row kernel),
  // B = 32: 3.5 ms as 32x1 (3.7 as 16x2, 4.9), B = 64: 4.5 ms as 32x2 (5.6); at B = 128 two chains of 64 lose (7.2 vs 6.9)
  // and four chains of 32 beat the 128-row kernel at B = 128 as well (7.9 against 8.5 us per time step in the cfg-3 step)
  int nb = p.B <= 16 ? 16 : 32;
  int nch = p.B <= 32 ? 1 : p.B <= 64 ? 2 : 4;
  if (force && fnb * fnch >= p.B) { nb = fnb; nch = fnch; }
  if (nb == 32 && nch == 4) return launch_fwd_chain<KB, 32, 4>(p, stream, launched);
  if (nb == 16 && nch == 1) return launch_fwd_chain<KB, 16, 1>(p, stream, launched);
  if (nb == 16 && nch == 2) return launch_fwd_chain<KB, 16, 2>(p, stream, launched);
  if (nb == 32 && nch == 1) return launch_fwd_chain<KB, 32, 1>(p, stream, launched);
  if (nb == 64 && nch == 1) return launch_fwd_chain<KB, 64, 1>(p, stream, launched);
  if (nb == 64 && nch == 2) return launch_fwd_chain<KB, 64, 2>(p, stream, launched);
  return launch_fwd_chain<KB, 32, 2>(p, stream, launched);
}

}  // namespace

bool blstm_fwd_chain_eligible(int B, int H) {
  static int enabled = -1, maxb = 128;
  if (enabled < 0) {
    const char* e = getenv("NABU_REC_FWD");
    enabled = (e && (strcmp(e, "flat") == 0 || strcmp(e, "ffma") == 0 || strcmp(e, "cl4") == 0)) ? 0 : 1;
    if (const char* m = getenv("NABU_FWD_CHAIN_MAXB")) maxb = atoi(m);
  }
  if (!enabled) return false;
  return B <= maxb && B <= 128 && B > 0 && (H == 256 || H == 512 || H == 1024);
}

int blstm_rec_fwd_chain(const float* const kernel[2], float* const gates[2], float* const cells[2], float* y, float* xchg,
                        const int* len, int B, int T, int yT, int D, int H, cudaStream_t stream, bool* launched, void* yh,
                        void* yl) {
  ClParams p = {};
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.y = y; p.xchg = xchg; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  p.yh = yh; p.yl = yl;
  return H == 1024 ? dispatch_fwd_chain<4>(p, stream, launched)
                   : H == 512 ? dispatch_fwd_chain<2>(p, stream, launched) : dispatch_fwd_chain<1>(p, stream, launched);
}

}  // namespace nabu
