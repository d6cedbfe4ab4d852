// tcgen05 forward recurrence (blstm_tc.cu), used by nabu_blstm_fwd when the shape is eligible.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace nabu {

struct BlstmTcPlan { int hs, nsl, ns; size_t smem; };

// B <= 128 rows (one UMMA M tile), H % 32 == 0, both directions co-resident, weights fit in smem.
bool blstm_tc_plan(int B, int H, BlstmTcPlan* pl);

// hrow: zero-initialised exchange buffer of 2*2*128*H floats; counters: 2 zeroed uints.
int blstm_rec_fwd_tc(const BlstmTcPlan& pl, const float* const kernel[2], float* const gates[2], float* const cells[2],
                     float* y, float* hrow, unsigned* counters, const int* len, int B, int T, int yT, int D, int H,
                     cudaStream_t stream);

}  // namespace nabu
