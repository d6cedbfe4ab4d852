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

// Pre-split fp16 path (gemm_h2.cu): x * S stored as hi = fp16(x S), lo = fp16((x S - hi) * 2^11), S a power of two
// per row (split_rows: operands whose rows run along K) or per matrix (split_global: operands whose rows run along
// M/N, i.e. both TN operands and the NN weight).  hi/lo are [R][ldo] fp16 planes, ldo % 8 == 0, columns >= C zero.
struct H2Operand {
  const void* hi; const void* lo; int ld;
  const float* row_inv;    // 1/S per row (device) or nullptr
  const float* glob_inv;   // 1/S of the matrix (device scalar) or nullptr
};
int split_rows(const float* src, int ld, int R, int C, void* hi, void* lo, int ldo, float* row_inv, cudaStream_t stream);
int split_global(const float* src, int ld, int R, int C, void* hi, void* lo, int ldo, unsigned* maxbits /*device scratch*/,
                 float* glob_inv, cudaStream_t stream);
// [s0 | s1] side by side (both [R, C], row pitch ld) as one [R, 2C] operand: one scale per row / one global scale.
int split_rows_pair(const float* s0, const float* s1, int ld, int R, int C, void* hi, void* lo, float* row_inv,
                    cudaStream_t stream);
int split_global_pair(const float* s0, const float* s1, int ld, int R, int C, void* hi, void* lo, unsigned* maxbits,
                      float* glob_inv, cudaStream_t stream);
// Same contract as gemm_tc() on pre-split operands.  NN: A rows / B global; NT: A rows, B rows or global; TN: global.
// absmax_out (optional, device): atomicMax of the bits of max |C| over the outputs written (the caller zeroes it).
int gemm_h2(GemmMode mode, int M, int N, int K, float alpha, const H2Operand& A, const H2Operand& B, float beta, float* C,
            int ldc, const float* bias, const GemmSeg* seg, float* workspace, size_t ws_bytes, cudaStream_t stream,
            unsigned* absmax_out = nullptr);
// atomicMax of the bits of max |src| over an [R, C] matrix with row pitch ld into *out (not zeroed here)
int absmax_accumulate(const float* src, int ld, size_t R, size_t C, unsigned* out, cudaStream_t stream);

// fp32 operands: split both into `workspace` (gemm_h2_auto_workspace_bytes), then gemm_h2.
size_t gemm_h2_auto_workspace_bytes(GemmMode mode, int M, int N, int K);
bool gemm_h2_eligible(GemmMode mode, int M, int N, int K);
int gemm_h2_auto(GemmMode mode, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
                 float beta, float* C, int ldc, const float* bias, void* workspace, size_t ws_bytes, cudaStream_t stream);

// Output layers with V <= 32 units (linear_skinny.cu): HBM-bound FFMA kernels, y = x.W + b and dW = x^T.dy
bool linear_skinny_eligible(const float* x, int N, int D, int V);
int linear_skinny_fwd(const float* x, int N, int D, int V, const float* W, const float* b, float* y, cudaStream_t stream);
int linear_skinny_dw(const float* x, int N, int D, int V, const float* dy, float* dW, float* workspace, size_t ws_bytes,
                     cudaStream_t stream);

// out[n] = sum_m X[m,n]
int colsum(const float* X, int M, int N, int ldx, float* out, cudaStream_t stream);

}  // namespace nabu
