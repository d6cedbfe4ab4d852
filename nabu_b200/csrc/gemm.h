// Internal dense-contraction interface shared by the layer kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace nabu {

enum GemmMode { GEMM_NN = 0, GEMM_NT = 1, GEMM_TN = 2 };

// TN only: reduction row r maps to A row (r/seg)*segA + r%seg + offA (same for B).
struct GemmSeg { int seg, segA, segB, offA, offB; };

// Scratch needed for deterministic split-K (TN with small outputs).
inline size_t sgemm_workspace_bytes() { return (size_t)32 << 20; }

int sgemm(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
          int ldb, float beta, float* C, int ldc, const float* bias, const GemmSeg* seg,
          float* workspace, size_t ws_bytes, cudaStream_t stream);

// Deterministic second pass of split-K: C = alpha * sum_z part[z] + beta*C + bias.
int splitk_reduce(const float* part, int splits, float* C, int M, int N, int ldc, float alpha, float beta,
                  const float* bias, cudaStream_t stream);

// tcgen05 / TMEM / TMA path with 3xTF32 split precision (gemm_tc.cu).  Same contract as sgemm().
bool gemm_tc_eligible(GemmMode mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb);
int gemm_tc(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
            float beta, float* C, int ldc, const float* bias, const GemmSeg* seg, float* workspace,
            size_t ws_bytes, cudaStream_t stream);

// Dispatcher used by the layer code: tensor-core path when eligible (and not disabled with
// NABU_GEMM=simt), FFMA path otherwise.
int gemm(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
         float beta, float* C, int ldc, const float* bias, const GemmSeg* seg, float* workspace,
         size_t ws_bytes, cudaStream_t stream);

// out[n] = sum_m X[m,n]
int colsum(const float* X, int M, int N, int ldx, float* out, cudaStream_t stream);

}  // namespace nabu
