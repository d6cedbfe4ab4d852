// FP32 SIMT GEMM used by every dense contraction of the hot path that must hold the
// 1e-4 fp32 parity bar without a split-precision tensor-core pass:
//   NN  C[M,N] = A[M,K]   . B[K,N]      (input projection X.Kx, output layer)
//   NT  C[M,N] = A[M,K]   . B[N,K]^T    (dX = dZ.Kx^T)
//   TN  C[M,N] = A[R,M]^T . B[R,N]      (dK = X^T.dZ, reduction over R = B*T rows;
//                                        optional per-utterance row segmentation so the
//                                        recurrent-weight gradient Hprev^T.dZ needs no
//                                        shifted copy of the hidden sequence)
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered shared
// memory with register prefetch; split-K (deterministic two-pass) when the output is
// too small to fill 148 SMs.
#include <algorithm>
#include "common.cuh"
#include "gemm.h"
#include <stdlib.h>
#include <string.h>

namespace nabu {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

struct GemmArgs {
  const float* A; const float* B; float* C; const float* bias;
  int M, N, K;            // output MxN, reduction K
  int lda, ldb, ldc;
  float alpha, beta;
  // TN row segmentation: reduction row r -> A row (r/seg)*segA + r%seg + offA
  int seg, segA, segB, offA, offB;
  int ksplit_len;         // reduction length per blockIdx.z
  float* part;            // split-K partials [splits, M, N] or nullptr
};

__device__ __forceinline__ float4 ldg4(const float* p, int nvalid, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid >= 4 && vec) {
    v = *reinterpret_cast<const float4*>(p);
  } else {
    if (nvalid > 0) v.x = p[0];
    if (nvalid > 1) v.y = p[1];
    if (nvalid > 2) v.z = p[2];
    if (nvalid > 3) v.w = p[3];
  }
  return v;
}

// KC_A: A operand is contiguous along the reduction dim (needs transposing store).
// KC_B: same for B.
template <bool KC_A, bool KC_B, bool SEG>
__global__ void __launch_bounds__(NT, 2)
sgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.ksplit_len;
  const int kend = min(g.K, kbeg + g.ksplit_len);

  const bool vecA = ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && (g.lda % 4 == 0);
  const bool vecB = ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && (g.ldb % 4 == 0);

  float4 ra[2], rb[2];

  auto rowA = [&](int r) -> long {
    if (SEG) return (long)(r / g.seg) * g.segA + (r % g.seg) + g.offA;
    return r;
  };
  auto rowB = [&](int r) -> long {
    if (SEG) return (long)(r / g.seg) * g.segB + (r % g.seg) + g.offB;
    return r;
  };

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i;
      if (KC_A) {           // A[m][k], 4 consecutive k
        const int r = idx / 4, kq = (idx % 4) * 4;
        const int m = m0 + r, k = k0 + kq;
        const int nv = (m < g.M) ? max(0, min(4, kend - k)) : 0;
        ra[i] = ldg4(g.A + (long)m * g.lda + k, nv, vecA);
      } else {              // A[k][m], 4 consecutive m
        const int kr = idx / 32, c = (idx % 32) * 4;
        const int k = k0 + kr, m = m0 + c;
        const int nv = (k < kend) ? max(0, min(4, g.M - m)) : 0;
        ra[i] = ldg4(g.A + rowA(k) * g.lda + m, nv, vecA);
      }
      if (KC_B) {           // B[n][k]
        const int r = idx / 4, kq = (idx % 4) * 4;
        const int n = n0 + r, k = k0 + kq;
        const int nv = (n < g.N) ? max(0, min(4, kend - k)) : 0;
        rb[i] = ldg4(g.B + (long)n * g.ldb + k, nv, vecB);
      } else {              // B[k][n]
        const int kr = idx / 32, c = (idx % 32) * 4;
        const int k = k0 + kr, n = n0 + c;
        const int nv = (k < kend) ? max(0, min(4, g.N - n)) : 0;
        rb[i] = ldg4(g.B + rowB(k) * g.ldb + n, nv, vecB);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + NT * i;
      if (KC_A) {
        const int r = idx / 4, kq = (idx % 4) * 4;
        As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
        As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
      } else {
        const int kr = idx / 32, c = (idx % 32) * 4;
        *reinterpret_cast<float4*>(&As[buf][kr][c]) = ra[i];
      }
      if (KC_B) {
        const int r = idx / 4, kq = (idx % 4) * 4;
        Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y;
        Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
      } else {
        const int kr = idx / 32, c = (idx % 32) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kr][c]) = rb[i];
      }
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int ntiles = (kend > kbeg) ? ceil_div(kend - kbeg, BK) : 0;
  if (ntiles > 0) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) gload(kbeg + (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const bool split = (g.part != nullptr);
  float* Cout = split ? g.part + (size_t)blockIdx.z * g.M * g.N : g.C;
  const int ldc = split ? g.N : g.ldc;
  const bool vecC = ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0) && (ldc % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n >= g.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      float* cp = Cout + (size_t)m * ldc + n;
      const int nv = min(4, g.N - n);
      if (!split) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nv) {
            float r = g.alpha * v[j];
            if (g.bias) r += g.bias[n + j];
            if (g.beta != 0.f) r += g.beta * cp[j];
            v[j] = r;
          }
        }
      }
      if (nv == 4 && vecC) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int j = 0; j < nv; ++j) cp[j] = v[j];
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* part, int splits, float* C, int M, int N, int ldc,
                                     float alpha, float beta, const float* bias) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)M * N) return;
  const int m = i / N, n = i % N;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(size_t)z * M * N + i];
  float r = alpha * s;
  if (bias) r += bias[n];
  float* cp = C + (size_t)m * ldc + n;
  if (beta != 0.f) r += beta * *cp;
  *cp = r;
}

// column sums: out[n] = sum_m X[m, n]  (bias gradients).  Rows are sliced over gridDim.y blocks (a 29-column matrix of
// 192 000 rows used to be summed by ONE block: 1.9 ms); the slices' partial sums are added in a fixed order.
__global__ void colsum_kernel(const float* X, int M, int N, int ldx, float* part) {
  __shared__ float sm[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int rows = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * rows, m1 = min(M, m0 + rows);
  float s = 0.f;
  if (n < N)
    for (int m = m0 + threadIdx.y; m < m1; m += 8) s += X[(size_t)m * ldx + n];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
    for (int j = 0; j < 8; ++j) t += sm[j][threadIdx.x];
    part[(size_t)blockIdx.y * N + n] = t;
  }
}
__global__ void colsum_final_kernel(const float* part, int S, int N, float* out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int s = 0; s < S; ++s) t += part[(size_t)s * N + n];
  out[n] = t;
}

}  // namespace

int sgemm(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
          int ldb, float beta, float* C, int ldc, const float* bias, const GemmSeg* segp,
          float* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.bias = bias;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.alpha = alpha; g.beta = beta;
  g.seg = segp ? segp->seg : 0; g.segA = segp ? segp->segA : 0; g.segB = segp ? segp->segB : 0;
  g.offA = segp ? segp->offA : 0; g.offB = segp ? segp->offB : 0;
  NABU_REQUIRE(!(segp && mode != GEMM_TN), "sgemm: row segmentation only in TN mode");

  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  int splits = 1;
  if (mode == GEMM_TN && workspace != nullptr) {
    const int sms = num_sms();
    if (tiles < sms && K >= 8 * BK * 4) {
      splits = min(ceil_div(2 * sms, tiles), K / (8 * BK));
      const size_t need = (size_t)splits * M * N * sizeof(float);
      if (need > ws_bytes) splits = (int)(ws_bytes / ((size_t)M * N * sizeof(float)));
      if (splits < 1) splits = 1;
    }
  }
  int klen = ceil_div(K, splits);
  klen = ceil_div(klen, BK) * BK;
  splits = max(1, ceil_div(K, klen));
  g.ksplit_len = klen;
  g.part = (splits > 1) ? workspace : nullptr;

  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits), block(NT);
  {
    KernelScope ks(mode == GEMM_NN ? "sgemm_nn" : mode == GEMM_NT ? "sgemm_nt" : "sgemm_tn", stream);
    switch (mode) {
      case GEMM_NN: sgemm_kernel<true, false, false><<<grid, block, 0, stream>>>(g); break;
      case GEMM_NT: sgemm_kernel<true, true, false><<<grid, block, 0, stream>>>(g); break;
      case GEMM_TN:
        if (segp) sgemm_kernel<false, false, true><<<grid, block, 0, stream>>>(g);
        else sgemm_kernel<false, false, false><<<grid, block, 0, stream>>>(g);
        break;
    }
  }
  NABU_CHECK_LAUNCH();
  if (splits > 1) {
    const long tot = (long)M * N;
    KernelScope ks("splitk_reduce", stream);
    splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(workspace, splits, C, M, N, ldc,
                                                                          alpha, beta, bias);
    NABU_CHECK_LAUNCH();
  }
  return 0;
}

int splitk_reduce(const float* part, int splits, float* C, int M, int N, int ldc, float alpha, float beta,
                  const float* bias, cudaStream_t stream) {
  const long tot = (long)M * N;
  KernelScope ks("splitk_reduce", stream);
  splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(part, splits, C, M, N, ldc, alpha, beta, bias);
  NABU_CHECK_LAUNCH();
  return 0;
}

int gemm(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
         float beta, float* C, int ldc, const float* bias, const GemmSeg* seg, float* workspace,
         size_t ws_bytes, cudaStream_t stream) {
  static int use_tc = -1;
  if (use_tc < 0) {
    const char* e = getenv("NABU_GEMM");
    use_tc = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  if (use_tc && gemm_tc_eligible(mode, M, N, K, A, lda, B, ldb))
    return gemm_tc(mode, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, seg, workspace, ws_bytes, stream);
  return sgemm(mode, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, seg, workspace, ws_bytes, stream);
}

int colsum(const float* X, int M, int N, int ldx, float* out, cudaStream_t stream) {
  // library-owned scratch for the slices' partial sums (grow-only; a handful of KB to a few MB)
  static float* part = nullptr;
  static size_t part_floats = 0;
  const int S = std::max(1, std::min(M / 256, 4 * num_sms() / ceil_div(N, 32)));
  if (S == 1) {
    KernelScope ks("colsum", stream);
    colsum_kernel<<<dim3(ceil_div(N, 32), 1), dim3(32, 8), 0, stream>>>(X, M, N, ldx, out);
    NABU_CHECK_LAUNCH();
    return 0;
  }
  if ((size_t)S * N > part_floats) {
    if (part) { NABU_CHECK_CUDA(cudaDeviceSynchronize()); NABU_CHECK_CUDA(cudaFree(part)); part = nullptr; part_floats = 0; }
    NABU_CHECK_CUDA(cudaMalloc(&part, (size_t)S * N * sizeof(float)));
    part_floats = (size_t)S * N;
  }
  KernelScope ks("colsum", stream);
  colsum_kernel<<<dim3(ceil_div(N, 32), S), dim3(32, 8), 0, stream>>>(X, M, N, ldx, part);
  NABU_CHECK_LAUNCH();
  colsum_final_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(part, S, N, out);
  NABU_CHECK_LAUNCH();
  return 0;
}

}  // namespace nabu
