// Decoder-step entry points shared between speller.cu (training) and las_beam.cu (decoding).
#pragma once
#include <cuda_runtime.h>
#include "nabu_b200.h"

namespace nabu {
namespace dec {

int check_desc(const nabu_speller_desc_t& d);

int launch_step(const nabu_speller_desc_t& d, const nabu_speller_params_t& p, int R, int rows_per_mem,
                const int* ids, const float* keys, const float* values, const int* mem_len,
                float* const* hT_prev, float* const* h_prev, float* const* c_prev, const float* ctx_prev,
                const float* ctxT_prev, const float* align_prev,
                float* const* hT_new, float* const* h_new, float* const* c_new, float* ctx_new, float* ctxT_new,
                float* align_new, float* const* gates_out, float* logits, long logits_row_stride,
                float temperature, float* q_save, float* cf_save, float* outin_save, long outin_row_stride,
                const int* tlen, int u, const int* done, cudaStream_t stream, float* asum_save = nullptr,
                float* const* out_new = nullptr, float* const* outT_new = nullptr, float keep = 1.f, unsigned seed = 0,
                const float* const* cell_relayout = nullptr);

// the LSTM cells' weight slices in the layout the step kernel stages (launch_step's cell_relayout): floats per layer, and
// the kernels that fill out[l]
size_t relayout_fwd_floats(const nabu_speller_desc_t& d, int l);
int relayout_fwd(const nabu_speller_desc_t& d, const nabu_speller_params_t& p, float* const* out, cudaStream_t stream);

// WindowedAttention's initial alignments: align [R][Tm] (already zeroed) gets 1 at frame 0 of every row
int init_window_alignments(float* align, int R, int Tm, cudaStream_t stream);

int prepare_memory(const nabu_speller_desc_t& d, const nabu_speller_params_t& p, const float* memory,
                   const int* mem_len, float* values, float* keys, cudaStream_t stream);

}  // namespace dec
}  // namespace nabu
