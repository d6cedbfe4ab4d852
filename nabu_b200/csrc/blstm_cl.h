// Cluster K-split recurrences (blstm_cl.cu), used by nabu_blstm_fwd / nabu_blstm_bwd when eligible.
#pragma once
#include <cuda_runtime.h>

namespace nabu {

// fixed power-of-two scale of the fp16 hi/lo operand planes of a BLSTM output: |h| < 1, so h * 32 < 32 sits where the
// per-row / per-matrix scales of gemm_h2.cu put an operand's maximum ([32, 64))
constexpr float Y_PLANE_SCALE = 32.f;

// B <= 128 and num_units in {128, 256, 512} (64 CTAs per direction = 8 clusters of 8); NABU_REC_BWD=flat disables.
bool blstm_bwd_cluster_eligible(int B, int H);

// *launched = false (and status 0) when the clusters are not co-resident on this device: use the flat kernel.
int blstm_rec_bwd_cluster(const float* const kernel[2], float* const gates[2], const float* const cells[2],
                          const float* dy, float* dbpart, float* xchg, float* dcbuf, unsigned* counters,
                          const int* len, int B, int T, int yT, int D, int H, cudaStream_t stream, bool* launched);

// forward twin (clusters of 4): same eligibility; NABU_REC_FWD=flat disables.
bool blstm_fwd_cluster_eligible(int B, int H);
int blstm_rec_fwd_cluster(const float* const kernel[2], float* const gates[2], float* const cells[2], float* y,
                          float* xchg, unsigned* counters, const int* len, int B, int T, int yT, int D, int H,
                          cudaStream_t stream, bool* launched);

// tcgen05 forward (blstm_cl_tc.cu): B <= 128, num_units in {256, 512}; NABU_REC_FWD=ffma|flat disables.
bool blstm_fwd_cluster_tc_eligible(int B, int H);
// yh / yl (optional): fp16 hi / lo planes [B, yT, 2H] of y * 32 written by the kernel (cl_common.cuh).
int blstm_rec_fwd_cluster_tc(const float* const kernel[2], float* const gates[2], float* const cells[2], float* y,
                             float* xchg, unsigned* counters, const int* len, int B, int T, int yT, int D, int H,
                             cudaStream_t stream, bool* launched, void* yh = nullptr, void* yl = nullptr);

// tcgen05 backward: same eligibility; rowmax = 128 zeroed words of scratch (per-row max |dy|, filled by a pre-pass);
// the exchange buffer must be zeroed by the caller (rows b >= B are read); NABU_REC_BWD=ffma|flat disables.
bool blstm_bwd_cluster_tc_eligible(int B, int H);
int blstm_rec_bwd_cluster_tc(const float* const kernel[2], float* const gates[2], const float* const cells[2],
                             const float* dy, float* dbpart, float* xchg, float* dcbuf, unsigned* counters,
                             unsigned* rowmax, const int* len, int B, int T, int yT, int D, int H, cudaStream_t stream,
                             bool* launched);

// tcgen05 backward on clusters of 8 with the recurrent weights resident in TMEM (blstm_cl_bwd8.cu): B <= 128,
// num_units in {256, 512}; preferred over the cluster-of-4 kernel; NABU_REC_BWD=cl4|ffma|flat disables.  rowmax and the
// (zeroed) exchange buffer as for blstm_rec_bwd_cluster_tc.
bool blstm_bwd_cluster8_eligible(int B, int H);
// zh / zl / zinv (optional): the kernel writes the fp16 hi / lo planes [B*T, 8H] of dZ (fw | bw) scaled by one power of two
// and 1/scale to *zinv INSTEAD of the fp32 dZ in gates[] (which then keeps the forward's activations).
int blstm_rec_bwd_cluster8(const float* const kernel[2], float* const gates[2], const float* const cells[2],
                           const float* dy, float* dbpart, float* xchg, unsigned* rowmax, const int* len, int B, int T, int yT,
                           int D, int H, cudaStream_t stream, bool* launched, void* zh = nullptr, void* zl = nullptr,
                           float* zinv = nullptr);

// "Chains" version of the cluster-of-8 backward (blstm_cl_bwd8c.cu): the batch tile is cut into independent chains of 16 /
// 32 / 64 rows, each run by its own warpgroup (two chains of 64 for B > 64, one chain sized to the batch below that).
// Same arguments as blstm_rec_bwd_cluster8; *nslots = the number of bias-gradient partial slots written (dbpart[dir][slot]).
// NABU_REC_BWD=cl8 keeps the single-chain kernel.
bool blstm_bwd_chain_eligible(int B, int H);
int blstm_rec_bwd_chain(const float* const kernel[2], float* const gates[2], const float* const cells[2], const float* dy,
                        float* dbpart, float* xchg, unsigned* rowmax, const int* len, int B, int T, int yT, int D, int H,
                        cudaStream_t stream, bool* launched, int* nslots, void* zh = nullptr, void* zl = nullptr,
                        float* zinv = nullptr, bool rowmax_ready = false);
// rowmax_ready: `rowmax` already holds the rows' max |dy| (a batch cut into tiles of 128 rows passes ONE word, the maximum
// of the whole batch, so that every tile scales its gradients alike).

// Small-batch forward (blstm_cl_fwdc.cu): transposed product with the weights in TMEM and the batch as the MMA's N, one or
// one, two or four chains of 16 / 32 rows: B <= 128 (NABU_FWD_CHAIN_MAXB lowers that), num_units in {256, 512};
// NABU_REC_FWD=cl4 keeps the 128-row kernel (blstm_cl_tc.cu).
bool blstm_fwd_chain_eligible(int B, int H);
int blstm_rec_fwd_chain(const float* const kernel[2], float* const gates[2], float* const cells[2], float* y, float* xchg,
                        const int* len, int B, int T, int yT, int D, int H, cudaStream_t stream, bool* launched,
                        void* yh = nullptr, void* yl = nullptr);

}  // namespace nabu
