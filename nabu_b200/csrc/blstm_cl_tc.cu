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
#include "cl_tc_common.cuh"
#include "blstm_cl.h"

namespace nabu {
namespace {

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
  constexpr int B_TILE = GC * 128;         // bytes of one [GC rows x 64 fp16] tile
  constexpr int RW = 4 * HS;               // floats per receive row
  constexpr int TCOLS = 2 * GC;            // TMEM columns: D1 | D2
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Bs = sm;                                    // [KB][hi|lo][B_TILE]
  uint8_t* As = Bs + KB * 2 * B_TILE;                  // [KB][hi|lo][A_TILE]
  float* rbuf = reinterpret_cast<float*>(As + KB * 2 * A_TILE);   // [CLS][BT][RW]
  __shared__ __align__(8) uint64_t rx_bar;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;

  const int H = p.H, H4 = 4 * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform
  const int per_dir = H / HS;
  const int dir = blockIdx.x / per_dir;
  const int q = (blockIdx.x % per_dir) / CLS;
  const int r = blockIdx.x % CLS;
  const int j0 = (q * CLS + r) * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  float* cells = const_cast<float*>(p.cells[dir]);
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
    mbar_init(smem_u32(&rx_bar), 1);
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
  // A thread owns batch row pb and UPT consecutive units at every step (one 16- or 8-byte access per array): cell
  // state and length stay in registers.
  constexpr int UPT = HS / 2;
  constexpr int NCH = KB * 2 * A_TILE / 16 / CL_THREADS;   // 16-byte chunks of my K-slice per thread
  const int pb = tid >> 1, pu = (tid & 1) * UPT;
  const bool prow = pb < p.B;
  const int plen = prow ? p.len[pb] : 0;
  float ccarry[UPT];
#pragma unroll
  for (int u = 0; u < UPT; ++u) ccarry[u] = 0.f;
  const uint32_t rbuf_u = smem_u32(rbuf);

  for (int s = 0; s < p.T; ++s) {
    const uint8_t* hprev = hx + (size_t)((s + 1) & 1) * H * BT * 4;
    uint8_t* hnext = hx + (size_t)(s & 1) * H * BT * 4;
    CL_STAMP(s, 0);
    // ---- prefetch pointwise operands -------------------------------------------------------------
    float gx[4][UPT];
    const bool valid = s < plen;
    const int tt = valid ? (dir ? plen - 1 - s : s) : s;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int u = 0; u < UPT; ++u) gx[g][u] = 0.f;
    // issued AFTER the critical loads of the h slice below (it would sit in front of them in the LSU queue)
    auto prefetch_gx = [&]() {
      if (valid) {
        const float* gp = gates + ((size_t)pb * p.T + tt) * H4 + j0 + pu;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if constexpr (UPT == 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(gp + g * H));
            gx[g][0] = v.x; gx[g][1] = v.y; gx[g][2] = v.z; gx[g][3] = v.w;
          } else {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(gp + g * H));
            gx[g][0] = v.x; gx[g][1] = v.y;
          }
        }
      }
    };
    if (s == 0) prefetch_gx();

    if (s > 0) {
      const unsigned par = (unsigned)(s - 1) & 1u;
      if (tid == 0) mbar_expect_tx(smem_u32(&rx_bar), (CLS - 1) * BT * RW * 4);
      // ---- fetch my K-slice of h_{s-1}.  No counters and no release/acquire round trips: every 16-bit half carries
      // the flag of the step that wrote it in its lowest bit (see the producer below), so the data validate themselves.
      // A thread spins on its first 16-byte chunk only (4 KB of polling traffic per CTA and round trip), then
      // fetches the rest and re-reads whatever is still stale; the slice is already the UMMA image of the A operand.
      {
        const uint4* src = reinterpret_cast<const uint4*>(hprev + (size_t)r * KB * 2 * A_TILE) + tid;
        const uint32_t fl = ll_flag(s - 1) ? 0x00010001u : 0u;
        uint4 v[NCH];
        do { v[0] = ld_relaxed_v4(src); } while (!ll_ok(v[0], fl));
        CL_STAMP(s, 1);
#pragma unroll
        for (int i = 1; i < NCH; ++i) v[i] = ld_relaxed_v4(src + i * CL_THREADS);
        prefetch_gx();
        // K block by K block: the MMAs of block kb run while block kb+1 is validated and stored (its loads are in flight)
        constexpr int CPB = NCH / KB;                  // chunks per thread and K block (hi tile then lo tile)
        uint4* dstA = reinterpret_cast<uint4*>(As) + tid;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int i = kb * CPB; i < (kb + 1) * CPB; ++i)
            while (!ll_ok(v[i], fl)) v[i] = ld_relaxed_v4(src + i * CL_THREADS);
          // peers have read their receive buffers of step s-1, hence received my quarters: As and their buffers are free
          if (kb == 0) cluster_wait();
#pragma unroll
          for (int i = kb * CPB; i < (kb + 1) * CPB; ++i) dstA[i * CL_THREADS] = v[i];
          fence_proxy_async_smem();
          __syncthreads();
          if (warp_u == 0) {                           // converged warp; one elected lane issues
            if (kb == 0) CL_STAMP(s, 2);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ah = make_desc(As_u + (kb * 2 + 0) * A_TILE + ks * 32, 16, 1024, 2);
              const uint64_t al = make_desc(As_u + (kb * 2 + 1) * A_TILE + ks * 32, 16, 1024, 2);
              const uint64_t bh = make_desc(Bs_u + (kb * 2 + 0) * B_TILE + ks * 32, 16, 1024, 2);
              const uint64_t bl = make_desc(Bs_u + (kb * 2 + 1) * B_TILE + ks * 32, 16, 1024, 2);
              const uint32_t acc = (kb | ks) != 0;
              if (elect_one()) {
                umma_f16(tm, ah, bh, idesc, acc);
                umma_f16(tm + GC, ah, bl, idesc, acc);
                umma_f16(tm + GC, al, bh, idesc, 1u);
              }
            }
            if (kb == KB - 1 && elect_one()) umma_commit(smem_u32(&mma_bar));
          }
          __syncwarp();
        }
      }
      mbar_wait(smem_u32(&mma_bar), par);
      tc_fence_after();
      CL_STAMP(s, 3);
      // ---- TMEM -> receive buffers (slot r).  My own quarter goes straight into my buffer; the three remote quarters
      // are staged in As (free: the MMAs have consumed it) and pushed by ONE bulk DSMEM copy each, which completes on
      // the destination's mbarrier: no remote store instructions and no closing cluster barrier (those two cost
      // 4.5 us of an 10.5 us time step; st.shared::cluster moved ~15 bytes/ns).  As is free again for the next step
      // because the cluster_wait before the next step's As stores implies the peers' rx_bar phases completed. ----------
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
        float* dstl = (d == r) ? rbuf + (size_t)(r * BT + row) * RW
                               : reinterpret_cast<float*>(As) + (size_t)(d * BT + row) * RW;
#pragma unroll
        for (int c4 = 0; c4 < RW / 4; ++c4) {
          float4 z;
          z.x = fmaf(__uint_as_float(v2[c4 * 4 + 0]), 1.f / 2048.f, __uint_as_float(v1[c4 * 4 + 0]));
          z.y = fmaf(__uint_as_float(v2[c4 * 4 + 1]), 1.f / 2048.f, __uint_as_float(v1[c4 * 4 + 1]));
          z.z = fmaf(__uint_as_float(v2[c4 * 4 + 2]), 1.f / 2048.f, __uint_as_float(v1[c4 * 4 + 2]));
          z.w = fmaf(__uint_as_float(v2[c4 * 4 + 3]), 1.f / 2048.f, __uint_as_float(v1[c4 * 4 + 3]));
          *reinterpret_cast<float4*>(dstl + rb_chunk<HS>(row, c4) * 4) = z;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      CL_STAMP(s, 4);
      if (tid < CLS && tid != r) {
        constexpr uint32_t QB = BT * RW * 4;           // bytes of one quarter
        bulk_s2s(map_to_rank(rbuf_u + (uint32_t)r * QB, (uint32_t)tid), As_u + (uint32_t)tid * QB, QB,
                 map_to_rank(smem_u32(&rx_bar), (uint32_t)tid));
      }
      mbar_wait(smem_u32(&rx_bar), par);               // the three remote quarters have landed in my buffer
      CL_STAMP(s, 5);
    }

    // ---- pointwise cell update for my units.  Only h_t (the exchange) is on the critical path of the next time
    // step: it is stored first and published; gates, cell and output go to memory after the release. ---------------
    float av[5][UPT], hn[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) hn[u] = 0.f;
    if (s > 0) {
#pragma unroll
      for (int src = 0; src < CLS; ++src)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float* rp = rbuf + ((size_t)src * BT + pb) * RW;
          if constexpr (UPT == 4) {
            const float4 v = *reinterpret_cast<const float4*>(rp + rb_chunk<HS>(pb, g * 2 + (tid & 1)) * 4);
            gx[g][0] += v.x; gx[g][1] += v.y; gx[g][2] += v.z; gx[g][3] += v.w;
          } else {
            const float2 v = *reinterpret_cast<const float2*>(rp + rb_chunk<HS>(pb, g) * 4 + (tid & 1) * 2);
            gx[g][0] += v.x; gx[g][1] += v.y;
          }
        }
    }
    // My receive buffer is free for step s+1 as soon as these loads have returned: a relaxed arrive (nothing to
    // publish; the release form would wait for the h stores below and sat on the critical path for 0.6 us).
    if (s + 1 < p.T) cluster_arrive_relaxed();
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
          ccarry[u] = cn;                              // valid steps are s = 0 .. len-1, so the carry is c_{s-1}
        }
      }
    }
    {
      // h_t, split, flagged, in the consumer's UMMA layout.  All 128 rows are written (rows b >= B as zeros): the
      // consumers wait for every half of their slice to carry this step's flag.
      const unsigned short fb = (unsigned short)ll_flag(s);
      unsigned short hh[UPT], hl[UPT];
#pragma unroll
      for (int u = 0; u < UPT; ++u) split_h_flag(hn[u], fb, &hh[u], &hl[u]);
      const int j = j0 + pu;
      uint8_t* t = hnext + (size_t)((j / KS) * KB + (j % KS) / 64) * 2 * A_TILE + sw128_h(pb, j % 64);
      if constexpr (UPT == 4) {
        __stcg(reinterpret_cast<uint2*>(t), make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16)));
        __stcg(reinterpret_cast<uint2*>(t + A_TILE), make_uint2((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16)));
      } else {
        __stcg(reinterpret_cast<unsigned*>(t), (uint32_t)hh[0] | ((uint32_t)hh[1] << 16));
        __stcg(reinterpret_cast<unsigned*>(t + A_TILE), (uint32_t)hl[0] | ((uint32_t)hl[1] << 16));
      }
    }
    CL_STAMP(s, 6); CL_STAMP(s, 7); CL_STAMP(s, 8); CL_STAMP(s, 9);
    // ---- off the critical path: what the backward pass and the next layer need --------------------------
    if (prow) {
      if (valid) {
        float* gp = gates + ((size_t)pb * p.T + tt) * H4 + j0 + pu;
        float* cp = cells + ((size_t)pb * p.T + tt) * H + j0 + pu;
        if constexpr (UPT == 4) {
#pragma unroll
          for (int g = 0; g < 4; ++g) __stcg(reinterpret_cast<float4*>(gp + g * H), make_float4(av[g][0], av[g][1], av[g][2], av[g][3]));
          __stcg(reinterpret_cast<float4*>(cp), make_float4(av[4][0], av[4][1], av[4][2], av[4][3]));
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) __stcg(reinterpret_cast<float2*>(gp + g * H), make_float2(av[g][0], av[g][1]));
          __stcg(reinterpret_cast<float2*>(cp), make_float2(av[4][0], av[4][1]));
        }
      }
      const size_t yo = ((size_t)pb * p.yT + tt) * 2 * H + dir * H + j0 + pu;
      float* yp = p.y + yo;
      if constexpr (UPT == 4) __stcg(reinterpret_cast<float4*>(yp), make_float4(hn[0], hn[1], hn[2], hn[3]));
      else __stcg(reinterpret_cast<float2*>(yp), make_float2(hn[0], hn[1]));
      if (p.yh) {
        // the operand planes of the GEMMs that read y (next layer's input projection, dKh, next layer's dKx): clean split
        // (no exchange flag) of y * 32, so no separate split / absmax pass ever re-reads y
        unsigned short ph[UPT], pl[UPT];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
          __half hi, lo;
          split_h(hn[u] * Y_PLANE_SCALE, &hi, &lo);
          ph[u] = __half_as_ushort(hi); pl[u] = __half_as_ushort(lo);
        }
        __half* yh = reinterpret_cast<__half*>(p.yh) + yo;
        __half* yl = reinterpret_cast<__half*>(p.yl) + yo;
        if constexpr (UPT == 4) {
          __stcg(reinterpret_cast<uint2*>(yh), make_uint2((uint32_t)ph[0] | ((uint32_t)ph[1] << 16), (uint32_t)ph[2] | ((uint32_t)ph[3] << 16)));
          __stcg(reinterpret_cast<uint2*>(yl), make_uint2((uint32_t)pl[0] | ((uint32_t)pl[1] << 16), (uint32_t)pl[2] | ((uint32_t)pl[3] << 16)));
        } else {
          __stcg(reinterpret_cast<unsigned*>(yh), (uint32_t)ph[0] | ((uint32_t)ph[1] << 16));
          __stcg(reinterpret_cast<unsigned*>(yl), (uint32_t)pl[0] | ((uint32_t)pl[1] << 16));
        }
      }
    }
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
  cfg.numAttrs = coop_attr() ? 2 : 1;
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
  pt.fences = getenv("NABU_REC_FENCES") ? 1 : 0;
  NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt));
  trace_dump("fwd_tc", pt.trace, stream);
  *launched = true;
  return 0;
}


// ---------------------------------------------------------------------------------------------------------
// backward: dh_{t-1}[B, H] = dz_t[B, 4H] . Kh^T with the partition of blstm_rec_bwd_cluster_kernel<CLS = 4>: CTA r of
// a cluster multiplies the 1024*HS/8... K-slice r (4H/4 dz columns, produced by 4 clusters) against its resident
// [4H/4 x 4*HS] block of Kh^T: M = 128 batch rows, N = 4*HS, K = 64*HS, as HS K-blocks of 64 streamed through a
// 3-stage bulk-copy ring (32 KB per stage, hi|lo), stages recycled by tcgen05.commit.
//
// dz is a gradient: its magnitude is arbitrary, fp16's range is not.  Every batch row b is therefore exchanged
// multiplied by a power of two S_b chosen so that max_t,j |dy[b,t,j]| lands in [32, 64) (row_absmax_kernel runs
// before the recurrence).  Gate derivatives are <= 1, so dz starts below that bound and would have to grow 1000x
// through the recurrence to reach fp16's maximum (conversions saturate instead of producing inf); 20 binades below
// the row maximum keep the full 22 bits, smaller values keep an absolute error of 2^-36 of the row maximum.  The
// scale is exact (power of two) and is divided out in the epilogue, per row.
// ---------------------------------------------------------------------------------------------------------

template <int HS>
__global__ void __launch_bounds__(CL_THREADS, 1)
blstm_rec_bwd_cluster_tc_kernel(const ClParams p, const unsigned* __restrict__ rowmax) {
  constexpr int CLS = TC_CLS;
  constexpr int BT = 128;
  constexpr int NC = CLS * HS;             // hidden units (= output columns) per cluster = MMA N
  constexpr int CPS = 64 / CLS / CLS;      // producer clusters per K-slice
  constexpr int SLAB = 4 * NC;             // dz columns per producer cluster
  constexpr int KBN = CPS * SLAB / 64;     // K blocks per slice (= HS)
  constexpr int KBC = KBN / CPS;           // K blocks per producer cluster
  constexpr int NST = HS == 8 ? 4 : 3;     // ring stages (all the shared memory that is left)
  constexpr int B_TILE = NC * 128;         // bytes of one [NC rows x 64 fp16] tile
  constexpr int TCOLS = 4 * NC < 32 ? 32 : 4 * NC;   // D1 | D2a | D2b | (unused): independent accumulate chains
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Bs = sm;                                    // [KBN][hi|lo][B_TILE]
  uint8_t* ring = Bs + KBN * 2 * B_TILE;               // [NST][hi|lo][A_TILE]
  float* rbuf = reinterpret_cast<float*>(ring + NST * 2 * A_TILE);   // [2 parity][CLS src][BT][HS]
  float* red = reinterpret_cast<float*>(ring);          // [CL_THREADS][4] bias-gradient scratch (after the loop)
  __shared__ __align__(8) uint64_t full_bar[NST];
  __shared__ __align__(8) uint64_t empty_bar[NST];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float scale[BT];

  const int H = p.H, H4 = 4 * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform
  const int per_dir = H / HS;
  const int dir = blockIdx.x / per_dir;
  const int q = (blockIdx.x % per_dir) / CLS;
  const int r = blockIdx.x % CLS;
  const int j0 = (q * CLS + r) * HS;
  const float* Kh = p.kernel[dir] + (size_t)p.D * H4;
  float* gates = p.gates[dir];
  const float* cells = p.cells[dir];
  unsigned* cnt = p.counters + dir * 16;
  uint8_t* dzx = reinterpret_cast<uint8_t*>(p.xchg) + (size_t)dir * 2 * H4 * BT * 4;   // [2 parity][slice][KBN][hi|lo][A_TILE]

  // resident weights, split: B[n][kl] = Kh[NC*q + n][g*H + NC*(r*CPS + cl) + u],  kl = (cl*4 + g)*NC + u
  for (int i = tid; i < CPS * SLAB * NC; i += CL_THREADS) {
    const int kl = i % (CPS * SLAB), n = i / (CPS * SLAB);
    const int cl = kl / SLAB, g = (kl % SLAB) / NC, u = kl % NC;
    const float w = Kh[(size_t)(NC * q + n) * H4 + g * H + NC * (r * CPS + cl) + u];
    __half hi, lo;
    split_h(w, &hi, &lo);
    uint8_t* t = Bs + (size_t)(kl / 64) * 2 * B_TILE + sw128_h(n, kl % 64);
    *reinterpret_cast<__half*>(t) = hi;
    *reinterpret_cast<__half*>(t + B_TILE) = lo;
  }
  if (tid < BT) {
    const float G = tid < p.B ? __uint_as_float(rowmax[tid]) : 0.f;
    float S = 1.f;
    if (G > 0.f && G < 3.0e38f) {
      int e;
      frexpf(G, &e);                                   // G in [2^(e-1), 2^e)
      e = e < -100 ? -100 : (e > 100 ? 100 : e);
      S = ldexpf(1.f, 6 - e);                          // G * S in [32, 64)
    }
    scale[tid] = S;
  }
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(&mma_bar), 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  cluster_arrive();
  cluster_wait();

  const uint32_t idesc = make_idesc_f16(128, NC);
  const uint32_t ring_u = smem_u32(ring), Bs_u = smem_u32(Bs);
  constexpr int UPT = HS / 2;                          // a thread owns batch row pb and UPT consecutive units
  const int pb = tid >> 1, pu = (tid & 1) * UPT;
  const bool prow = pb < p.B;
  const int plen = prow ? p.len[pb] : 0;
  float dbacc[4][UPT], dcc[UPT];
#pragma unroll
  for (int u = 0; u < UPT; ++u) {
    dcc[u] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) dbacc[g][u] = 0.f;
  }
  unsigned gq = 0;                                     // K blocks consumed so far (ring position)
  const float pS = scale[pb];

  int iter = 0;
  for (int s = p.T - 1; s >= 0; --s, ++iter) {
    const uint8_t* dzprev = dzx + (size_t)((iter + 1) & 1) * H4 * BT * 4;
    uint8_t* dznext = dzx + (size_t)(iter & 1) * H4 * BT * 4;
    float* rb = rbuf + (size_t)(iter & 1) * CLS * BT * HS;
    CL_STAMP(iter, 0);
    // ---- prefetch pointwise operands -----------------------------------------------------------
    float gt[4][UPT], ct[UPT], cprev[UPT], dyv[UPT];
    const bool valid = s < plen;
    const int tt = valid ? (dir ? plen - 1 - s : s) : s;
#pragma unroll
    for (int u = 0; u < UPT; ++u) {
      ct[u] = cprev[u] = dyv[u] = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) gt[g][u] = 0.f;
    }
    if (valid) {
      const float* gp = gates + ((size_t)pb * p.T + tt) * H4 + j0 + pu;
      const float* cp = cells + ((size_t)pb * p.T + tt) * H + j0 + pu;
      const float* cq = cells + ((size_t)pb * p.T + (dir ? tt + 1 : tt - 1)) * H + j0 + pu;
      const float* yp = p.dy + ((size_t)pb * p.yT + tt) * 2 * H + dir * H + j0 + pu;
      if constexpr (UPT == 4) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(gp + g * H));
          gt[g][0] = v.x; gt[g][1] = v.y; gt[g][2] = v.z; gt[g][3] = v.w;
        }
        float4 v = __ldcg(reinterpret_cast<const float4*>(cp));
        ct[0] = v.x; ct[1] = v.y; ct[2] = v.z; ct[3] = v.w;
        if (s > 0) {
          v = __ldcg(reinterpret_cast<const float4*>(cq));
          cprev[0] = v.x; cprev[1] = v.y; cprev[2] = v.z; cprev[3] = v.w;
        }
        v = __ldcg(reinterpret_cast<const float4*>(yp));
        dyv[0] = v.x; dyv[1] = v.y; dyv[2] = v.z; dyv[3] = v.w;
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float2 v = __ldcg(reinterpret_cast<const float2*>(gp + g * H));
          gt[g][0] = v.x; gt[g][1] = v.y;
        }
        float2 v = __ldcg(reinterpret_cast<const float2*>(cp));
        ct[0] = v.x; ct[1] = v.y;
        if (s > 0) {
          v = __ldcg(reinterpret_cast<const float2*>(cq));
          cprev[0] = v.x; cprev[1] = v.y;
        }
        v = __ldcg(reinterpret_cast<const float2*>(yp));
        dyv[0] = v.x; dyv[1] = v.y;
      }
    }

    if (iter > 0) {
      const unsigned g0 = gq;
      if (tid == 32) {
        // ---- loader (warp 1): K block kb of this step -> ring position g0 + kb, as soon as its producer cluster has
        // published and the previous tenant of the stage has been consumed by the tensor core ----------------------
        const unsigned target = (unsigned)CLS * (unsigned)iter;
        const uint8_t* slab = dzprev + (size_t)r * KBN * 2 * A_TILE;
#pragma unroll 1
        for (int kb = 0; kb < KBN; ++kb) {
          if (kb % KBC == 0) {
            while (ld_acquire_gpu(cnt + r * CPS + kb / KBC) < target) { }
            fence_proxy_async_all();
          }
          const unsigned pos = g0 + kb, st = pos % NST;
          if (pos >= NST) mbar_wait(smem_u32(&empty_bar[st]), (pos / NST - 1) & 1);
          mbar_expect_tx(smem_u32(&full_bar[st]), 2 * A_TILE);
          cb_bulk(ring + (size_t)st * 2 * A_TILE, slab + (size_t)kb * 2 * A_TILE, 2 * A_TILE, &full_bar[st]);
        }
      } else if (warp_u == 0) {
        // ---- MMA issuer (warp 0, converged; one elected lane issues) --------------------------------------------
#pragma unroll 1
        for (int kb = 0; kb < KBN; ++kb) {
          const unsigned pos = g0 + kb, st = pos % NST;
          mbar_wait(smem_u32(&full_bar[st]), (pos / NST) & 1);
          if (kb == 0) { CL_STAMP(iter, 1); CL_STAMP(iter, 2); }
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = make_desc(ring_u + (st * 2 + 0) * A_TILE + ks * 32, 16, 1024, 2);
            const uint64_t al = make_desc(ring_u + (st * 2 + 1) * A_TILE + ks * 32, 16, 1024, 2);
            const uint64_t bh = make_desc(Bs_u + (kb * 2 + 0) * B_TILE + ks * 32, 16, 1024, 2);
            const uint64_t bl = make_desc(Bs_u + (kb * 2 + 1) * B_TILE + ks * 32, 16, 1024, 2);
            const uint32_t acc = (kb | ks) != 0;
            if (elect_one()) {
              umma_f16(tm, ah, bh, idesc, acc);
              umma_f16(tm + NC, ah, bl, idesc, acc);
              umma_f16(tm + 2 * NC, al, bh, idesc, acc);
            }
          }
          if (elect_one()) umma_commit(smem_u32(&empty_bar[st]));
        }
        if (elect_one()) umma_commit(smem_u32(&mma_bar));
      }
      gq = g0 + KBN;
      __syncwarp();
      mbar_wait(smem_u32(&mma_bar), (unsigned)(iter - 1) & 1u);
      tc_fence_after();
      CL_STAMP(iter, 3);
      // ---- TMEM -> receive buffers: thread = one batch row, columns [d*HS, (d+1)*HS) go to CTA d, slot r --------
      if (warp < 4) {
        const int row = warp * 32 + lane;
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
        uint32_t v1[NC], v2[NC], v3[NC];
        if constexpr (NC == 32) {
          tmem_ld32(taddr, reinterpret_cast<uint32_t(&)[32]>(v1));
          tmem_ld32(taddr + NC, reinterpret_cast<uint32_t(&)[32]>(v2));
          tmem_ld32(taddr + 2 * NC, reinterpret_cast<uint32_t(&)[32]>(v3));
        } else {
          tmem_ld16(taddr, reinterpret_cast<uint32_t(&)[16]>(v1));
          tmem_ld16(taddr + NC, reinterpret_cast<uint32_t(&)[16]>(v2));
          tmem_ld16(taddr + 2 * NC, reinterpret_cast<uint32_t(&)[16]>(v3));
        }
        tmem_ld_wait();
        const float is = 1.f / scale[row];                // exact: S is a power of two
#pragma unroll
        for (int d = 0; d < CLS; ++d) {
          const uint32_t dst = map_to_rank(smem_u32(rbuf), (uint32_t)d) +
                               (uint32_t)((((iter & 1) * CLS + r) * BT + row) * HS) * 4u;
#pragma unroll
          for (int c4 = 0; c4 < HS / 4; ++c4) {
            float z[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = d * HS + c4 * 4 + e;
              z[e] = fmaf(__uint_as_float(v2[c]) + __uint_as_float(v3[c]), 1.f / 2048.f, __uint_as_float(v1[c])) * is;
            }
            const int pos = HS == 8 ? (c4 ^ ((row >> 2) & 1)) : c4;
            st_cluster_v4(dst + (uint32_t)pos * 16u, z[0], z[1], z[2], z[3]);
          }
        }
      }
      tc_fence_before();
      CL_STAMP(iter, 4);
      cluster_arrive();
      cluster_wait();
      CL_STAMP(iter, 5);
    }

    // ---- pointwise gate gradients for my units.  Only the exchanged dz is on the critical path of the next time
    // step; the fp32 dz the GEMMs read (gates[]) goes to memory after the release. ---------------------------------
    float dzv[4][UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) dzv[0][u] = dzv[1][u] = dzv[2][u] = dzv[3][u] = 0.f;
    if (prow) {
      float dh[UPT];
#pragma unroll
      for (int u = 0; u < UPT; ++u) dh[u] = dyv[u];
      if (iter > 0) {
#pragma unroll
        for (int src = 0; src < CLS; ++src) {
          const float* rp = rb + ((size_t)src * BT + pb) * HS;
          if constexpr (UPT == 4) {
            const float4 v = *reinterpret_cast<const float4*>(rp + ((tid & 1) ^ ((pb >> 2) & 1)) * 4);
            dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
          } else {
            const float2 v = *reinterpret_cast<const float2*>(rp + (tid & 1) * 2);
            dh[0] += v.x; dh[1] += v.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UPT; ++u) {
        float dcn = 0.f;
        if (valid) {
          const float ig = gt[0][u], gg = gt[1][u], fg = gt[2][u], og = gt[3][u];
          const float tc_ = tanh_tc(ct[u]);
          const float d_o = dh[u] * tc_;
          const float dc = dcc[u] + dh[u] * og * (1.f - tc_ * tc_);
          dzv[0][u] = dc * gg * ig * (1.f - ig);
          dzv[1][u] = dc * ig * (1.f - gg * gg);
          dzv[2][u] = dc * cprev[u] * fg * (1.f - fg);
          dzv[3][u] = d_o * og * (1.f - og);
          dcn = dc * fg;
        }
        dcc[u] = dcn;
      }
      // my cluster is producer (q % CPS) of K-slice q / CPS; column kl = ((q % CPS)*4 + g)*NC + r*HS + pu + u
      uint8_t* xs = dznext + (size_t)(q / CPS) * KBN * 2 * A_TILE;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int kl = ((q % CPS) * 4 + g) * NC + r * HS + pu;
        unsigned short hh[UPT], hl[UPT];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
          __half hi, lo;
          split_h_sat(dzv[g][u] * pS, &hi, &lo);
          hh[u] = __half_as_ushort(hi); hl[u] = __half_as_ushort(lo);
          dbacc[g][u] += dzv[g][u];
        }
        uint8_t* tptr = xs + (size_t)(kl / 64) * 2 * A_TILE + sw128_h(pb, kl % 64);
        if constexpr (UPT == 4) {
          __stcg(reinterpret_cast<uint2*>(tptr), make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16)));
          __stcg(reinterpret_cast<uint2*>(tptr + A_TILE), make_uint2((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16)));
        } else {
          __stcg(reinterpret_cast<unsigned*>(tptr), (uint32_t)hh[0] | ((uint32_t)hh[1] << 16));
          __stcg(reinterpret_cast<unsigned*>(tptr + A_TILE), (uint32_t)hl[0] | ((uint32_t)hl[1] << 16));
        }
      }
    }
    CL_STAMP(iter, 6);
    if (p.fences) {
      fence_proxy_async_all();
      __threadfence();
    }
    CL_STAMP(iter, 7);
    __syncthreads();
    CL_STAMP(iter, 8);
    if (tid == 0) red_release_gpu_add(cnt + q, 1u);
    CL_STAMP(iter, 9);
    if (prow) {
      float* gp = gates + ((size_t)pb * p.T + tt) * H4 + j0 + pu;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if constexpr (UPT == 4) __stcg(reinterpret_cast<float4*>(gp + g * H), make_float4(dzv[g][0], dzv[g][1], dzv[g][2], dzv[g][3]));
        else __stcg(reinterpret_cast<float2*>(gp + g * H), make_float2(dzv[g][0], dzv[g][1]));
      }
    }
  }

  // bias gradient: thread (row, half) holds the sums of its row for units half*UPT + u; fixed-order sum over rows
  {
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int u = 0; u < UPT; ++u) red[(tid * 4 + g) * UPT + u] = dbacc[g][u];
    __syncthreads();
    if (tid < 4 * HS) {
      const int g = tid / HS, j = tid % HS;
      float sum = 0.f;
      for (int i = j / UPT; i < CL_THREADS; i += 2) sum += red[(i * 4 + g) * UPT + (j % UPT)];
      p.dbpart[((size_t)dir * 8) * H4 + g * H + j0 + j] = sum;
    }
  }
  tc_fence_before();
  cluster_arrive();
  cluster_wait();
  if (warp == 0) tmem_dealloc(tm, TCOLS);
}

template <int HS>
int launch_bwd_tc(const ClParams& p, unsigned* rowmax, cudaStream_t stream, bool* launched) {
  constexpr int CLS = TC_CLS, NC = CLS * HS, KBN = HS, NST = HS == 8 ? 4 : 3;
  const size_t smem = 1024 + (size_t)KBN * 2 * NC * 128 + (size_t)NST * 2 * A_TILE + (size_t)2 * CLS * 128 * HS * sizeof(float);
  auto* fn = blstm_rec_bwd_cluster_tc_kernel<HS>;
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
  cfg.numAttrs = coop_attr() ? 2 : 1;
  int nclusters = 0;
  const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg);
  if (getenv("NABU_DEBUG"))
    fprintf(stderr, "[nabu] bwd tcgen05 cluster kernel HS=%d: smem %zu B, max active clusters %d (%s), need %d\n", HS, smem,
            nclusters, cudaGetErrorString(oe), (int)cfg.gridDim.x / CLS);
  if (oe != cudaSuccess || nclusters * CLS < (int)cfg.gridDim.x) {
    cudaGetLastError();
    return 0;
  }
  {
    KernelScope ks("row_absmax", stream);
    row_absmax_kernel<<<dim3(32, p.B), 256, 0, stream>>>(p.dy, p.len, p.yT, 2 * p.H, rowmax);
    NABU_CHECK_LAUNCH();
  }
  KernelScope ks("blstm_rec_bwd_cluster_tc", stream);
  ClParams pt = p;
  pt.trace = trace_buffer();
  pt.fences = getenv("NABU_REC_FENCES") ? 1 : 0;
  const unsigned* rm = rowmax;
  NABU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, pt, rm));
  trace_dump("bwd_tc", pt.trace, stream);
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
                             cudaStream_t stream, bool* launched, void* yh, void* yl) {
  ClParams p = {};
  p.yh = yh; p.yl = yl;
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.y = y; p.xchg = xchg; p.counters = counters; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  return H == 512 ? launch_fwd_tc<8>(p, stream, launched) : launch_fwd_tc<4>(p, stream, launched);
}

bool blstm_bwd_cluster_tc_eligible(int B, int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NABU_REC_BWD");
    enabled = (e && (strcmp(e, "flat") == 0 || strcmp(e, "ffma") == 0)) ? 0 : 1;
  }
  if (!enabled) return false;
  return B <= 128 && B > 0 && (H == 256 || H == 512);
}

int blstm_rec_bwd_cluster_tc(const float* const kernel[2], float* const gates[2], const float* const cells[2],
                             const float* dy, float* dbpart, float* xchg, float* dcbuf, unsigned* counters,
                             unsigned* rowmax, const int* len, int B, int T, int yT, int D, int H, cudaStream_t stream,
                             bool* launched) {
  ClParams p = {};
  p.kernel[0] = kernel[0]; p.kernel[1] = kernel[1];
  p.gates[0] = gates[0]; p.gates[1] = gates[1];
  p.cells[0] = cells[0]; p.cells[1] = cells[1];
  p.dy = dy; p.dbpart = dbpart; p.xchg = xchg; p.dcbuf = dcbuf; p.counters = counters; p.len = len;
  p.B = B; p.T = T; p.yT = yT; p.D = D; p.H = H;
  return H == 512 ? launch_bwd_tc<8>(p, rowmax, stream, launched) : launch_bwd_tc<4>(p, rowmax, stream, launched);
}

}  // namespace nabu
