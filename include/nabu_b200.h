/* nabu_b200 -- C-ABI of the B200-native engine for nabu's per-utterance hot path.
 *
 * Conventions (all entry points):
 *   - plain C: device pointers + sizes, no torch / C++ types;
 *   - every array is fp32 (lengths / labels / ids int32), row-major, batch-major, DEVICE memory
 *     unless the parameter is documented "host";
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *     nothing synchronises, allocates or frees: the caller owns every buffer, including the
 *     scratch sized by the matching *_workspace_bytes();
 *   - return 0 on success, non-zero on error with a message in nabu_last_error();
 *   - there is no CPU fallback: without a sm_100a device every call fails.
 *
 * "Replaces" names the reference interface (vrenkens/nabu @ 39deb62, paths under
 * nabu/neuralnetworks/) whose arithmetic the entry point performs.
 */
#ifndef NABU_B200_H_
#define NABU_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* nabu_last_error(void);
int nabu_version(void);

/* Instrumentation used by bench.py: kernels launched by this library since it was loaded, and
 * optional CUDA-event bracketing of every launch (on the launch stream).  nabu_profile_collect
 * synchronises the device and writes {"kernel": [launches, total_ms], ...} as JSON. */
unsigned long long nabu_kernel_launches(void);
int nabu_profile_enable(int on);
int nabu_profile_collect(char* json_out, size_t cap);

/* Deferred weight gradients (off by default).  With nabu_set_overlap(1), nabu_blstm_bwd returns once the recurrence
 * and dx are enqueued on `stream`; the weight gradients (dkernel_*) of that call are computed on a library-owned side
 * stream, in library-owned scratch, concurrently with whatever the caller enqueues next (the next layer's backward
 * recurrence leaves more than half of the SMs idle).  The caller must then (1) keep x, y and gates of that call alive
 * and unmodified and (2) not read dkernel_* until nabu_side_join(stream) has made `stream` wait for the deferred work.
 * nabu_blstm_fwd and nabu_clip_adam_step join implicitly.  (The reference has no counterpart: TF's executor schedules
 * independent gradient ops of `tf.gradients` concurrently on its own, trainers/trainer.py:556-558.) */
int nabu_set_overlap(int on);
int nabu_side_join(void* stream);

/* ---- dense contraction (building block; exported for tests) ------------------------------------
 * mode 0: C[M,N] = alpha*A[M,K].B[K,N]   + beta*C + bias[N]
 * mode 1: C[M,N] = alpha*A[M,K].B[N,K]^T + beta*C + bias[N]
 * mode 2: C[M,N] = alpha*A[K,M]^T.B[K,N] + beta*C + bias[N]
 * precision 0: fp32 FFMA (bit-reproducible); 1: tcgen05 3xTF32 split (fp32-grade, tensor cores);
 * 2: tcgen05 on operands pre-split into two scaled fp16 terms (fp32-grade, twice the TF32 rate; workspace from
 * nabu_gemm_h2_workspace_bytes).
 * Replaces: the tf.matmul inside every TF cell / layer the hot path touches. */
size_t nabu_gemm_workspace_bytes(void);
size_t nabu_gemm_h2_workspace_bytes(int mode, int M, int N, int K);
int nabu_gemm(int mode, int precision, int M, int N, int K, float alpha, const float* A, int lda,
              const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
              void* workspace, size_t ws_bytes, void* stream);

/* ---- a1: BLSTM layer ----------------------------------------------------------------------------
 * Replaces components/layer.py:8-51 `blstm` (LayerNormBasicLSTMCell(layer_norm=False) under
 * bidirectional_dynamic_rnn, sequence_length=len).
 *   x [B,T,D]; len [B]; kernel_{fw,bw} [(D+H),4H] (rows: input then hidden; columns i,j,f,o);
 *   bias_{fw,bw} [4H];  y [B,yT,2H] (fw | bw), rows t>=len[b] and t in [T,yT) are zero -- yT>T
 *   lets the caller get the zero padding ops.pyramid_stack (components/ops.py:30-38) adds.
 *   gates [2,B,T,4H] and cells [2,B,T,H] are saved activations for nabu_blstm_bwd. */
size_t nabu_blstm_workspace_bytes(int B, int T, int D, int H);
int nabu_blstm_fwd(const float* x, const int* len, int B, int T, int D, int H,
                   const float* kernel_fw, const float* bias_fw,
                   const float* kernel_bw, const float* bias_bw,
                   float* y, int yT, float* gates, float* cells,
                   void* workspace, size_t ws_bytes, void* stream);
/* dy [B,yT,2H].  gates is overwritten (it becomes dZ).  dx may be NULL (first layer). */
int nabu_blstm_bwd(const float* x, const int* len, int B, int T, int D, int H,
                   const float* kernel_fw, const float* kernel_bw,
                   const float* y, int yT, float* gates, const float* cells, const float* dy,
                   float* dx, float* dkernel_fw, float* dbias_fw, float* dkernel_bw, float* dbias_bw,
                   void* workspace, size_t ws_bytes, void* stream);

/* The same layer with the tensor-core GEMMs' operand planes travelling with the activations.  The dense contractions of a
 * layer (input projection, dKx, dKh, dX) run on tcgen05 with every fp32 operand carried as two fp16 planes ("hi", "lo");
 * producing those planes by separate passes over y / dZ cost 18 % of the cfg-3 step in round 1.  With these entry points
 * the forward recurrence writes the planes of y (y * 32 split in two fp16 numbers) next to y, the next layer reads them
 * as x_planes, and the backward recurrence hands dZ to its three GEMMs as planes in library-owned scratch.
 *   y_planes : nabu_blstm_planes_bytes(B, yT, H) bytes, caller-owned, written by _fwd_planes, to be passed again to
 *              _bwd_planes of the same layer (dKh) and as x_planes to the layer that consumes y unchanged -- also after
 *              pyramid_stack, which is a reshape of y and of its planes alike.  NULL: no planes are written / used.
 *   x_planes : planes of x (the y_planes of the producer of x; needs D % 8 == 0) or NULL (the library splits x itself).
 * Results are identical in accuracy class to nabu_blstm_fwd / _bwd (22-bit operands, fp32 accumulation); gates[] is NOT
 * overwritten with dZ by _bwd_planes when the tcgen05 recurrence runs (it is still clobbered on the fallback kernels). */
size_t nabu_blstm_planes_bytes(int B, int yT, int H);
int nabu_blstm_fwd_planes(const float* x, const void* x_planes, const int* len, int B, int T, int D, int H,
                          const float* kernel_fw, const float* bias_fw,
                          const float* kernel_bw, const float* bias_bw,
                          float* y, void* y_planes, int yT, float* gates, float* cells,
                          void* workspace, size_t ws_bytes, void* stream);
int nabu_blstm_bwd_planes(const float* x, const void* x_planes, const int* len, int B, int T, int D, int H,
                          const float* kernel_fw, const float* kernel_bw,
                          const float* y, const void* y_planes, int yT, float* gates, const float* cells, const float* dy,
                          float* dx, float* dkernel_fw, float* dbias_fw, float* dkernel_bw, float* dbias_bw,
                          void* workspace, size_t ws_bytes, void* stream);
/* Optional, one-shot (consumed by the NEXT nabu_blstm_bwd / nabu_blstm_bwd_planes call of the calling thread, whatever
 * its outcome): hand-over of max |dx| between the backward calls of stacked layers (components/layer.py:8-51 stacked by
 * models/ed_encoders/{dblstm,listener}.py).  The backward recurrence exchanges its gate gradients under one power-of-two
 * scale taken from max |dy| over the batch; without a hint every call reads dy once to find it.
 *   dx_absmax_out  device buffer of 128 words: the call also leaves max |dx| there (word 0 = the bits of the float, the
 *                  other words 0), written by the dX contraction's epilogue on `stream`;
 *   dy_absmax_in   such a buffer, filled by the call whose dx IS this call's dy (same memory, unmodified -- the caller's
 *                  responsibility): the pass over dy is skipped.
 * Either may be NULL.  Results are bit-identical with and without hints. */
int nabu_blstm_bwd_hints(unsigned* dx_absmax_out, const unsigned* dy_absmax_in);

/* ---- a2: pyramid_stack lengths -------------------------------------------------------------------
 * Replaces components/ops.py:55-58: out[b] = ceil(len[b] / numsteps).  (The data movement of
 * pyramid_stack is a free reshape of the yT-padded BLSTM output.) */
int nabu_pyramid_lengths(const int* len, int B, int numsteps, int* out, void* stream);

/* ---- a5: output layer ----------------------------------------------------------------------------
 * Replaces models/ed_decoders/dnn_decoder.py:53-57 (tf.contrib.layers.linear): y[N,V]=x[N,D].W+b. */
int nabu_linear_fwd(const float* x, int N, int D, int V, const float* W, const float* b, float* y,
                    void* workspace, size_t ws_bytes, void* stream);
int nabu_linear_bwd(const float* x, int N, int D, int V, const float* W, const float* dy,
                    float* dx, float* dW, float* db, void* workspace, size_t ws_bytes, void* stream);

/* ---- a9: CTC loss --------------------------------------------------------------------------------
 * Replaces trainers/loss_functions.py:203-210 (tf.nn.ctc_loss, time_major=False, defaults): softmax
 * over V inside, blank = V-1.  loss[b] = -log p(labels_b | logits_b); grad [B,T,V] = grad_scale *
 * d loss[b] / d logits (zero for t >= logit_len[b]).  Infeasible utterances get loss = +inf. */
size_t nabu_ctc_workspace_bytes(int B, int T, int V, int Lmax);
int nabu_ctc_loss_fwd_bwd(const float* logits, const int* logit_len, const int* labels, int Lmax,
                          const int* label_len, int B, int T, int V, float grad_scale,
                          float* loss, float* grad, void* workspace, size_t ws_bytes, void* stream);

/* ---- a10: masked cross-entropy ---------------------------------------------------------------------
 * Replaces trainers/loss_functions.py:78-109,155-165 (average_cross_entropy): per utterance
 * sum_{u<logit_len} -log softmax(logits[b,u])[targets[b,u]] / target_len[b].  loss [B];
 * grad [B,U,V] = grad_scale * d loss[b] / d logits. */
int nabu_masked_ce_fwd_bwd(const float* logits, const int* targets, int ldt, const int* logit_len,
                           const int* target_len, int B, int U, int V, float grad_scale,
                           float* loss, float* grad, void* stream);

/* ---- a11: update ---------------------------------------------------------------------------------
 * Replaces trainers/trainer.py:556-569: g = clip_by_value(g,-clip,clip); tf.train.AdamOptimizer step
 * `t` (>=1): m,v update; theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps).  Flat buffers [n].
 * grad_scale multiplies g before the clip (1/world for an averaged all-reduce; 1 otherwise). */
int nabu_clip_adam_step(float* theta, const float* grad, float* m, float* v, size_t n, float lr,
                        int t, float beta1, float beta2, float eps, float clip, float grad_scale,
                        void* stream);

/* ---- a6-a8: Speller (attention decoder), whole teacher-forced sequence ---------------------------
 * Replaces models/ed_decoders/rnn_decoder.py:40-82 + speller.py:29-69 + components/rnn_cell.py:
 * 145-155 + components/attention.py:142-240 with sample_prob=0 (teacher forcing).
 * Filled in by nabu_speller_desc_t below. */
typedef struct {
  int B, Tm, E;            /* memory [B,Tm,E], mem_len [B] */
  int V, H, num_layers;    /* output classes (incl. EOS/SOS = V-1), LSTM units, layers (<=4) */
  int A;                   /* attention units (= H in the reference) */
  int attention;           /* 0 vanilla Bahdanau, 1 location_aware, 2 windowed (attention.py:294-396) */
  int numfilt, filtersize; /* location_aware: conv filters / taps; windowed: left_window_width / right_window_width */
  int U;                   /* decoder steps = max target length */
  int probability_fn;      /* alignments from the masked scores (components/attention.py:9-13, 41-55):
                            * 0 softmax, 1 normalized_sigmoid (sigmoid / its sum over the memory), 2 sigmoid */
  /* the stochastic parts of training (speller.py:37-41, rnn_decoder.py:59-64); all zero = off (beam search ignores them) */
  float dropout_keep;      /* DropoutWrapper(output_keep_prob) on every LSTM layer; 0 or >= 1: no dropout */
  float sample_prob;       /* ScheduledEmbeddingTrainingHelper sampling probability */
  unsigned seed;           /* counter-based generator (csrc/speller_kernels.cuh: dec_rng_u32); change it every step */
} nabu_speller_desc_t;

/* Parameter pack (device pointers).  Same field order for the gradient pack. */
typedef struct {
  float* cell_kernel[4];   /* layer 0: [(V+E+H),4H]; layer l>0: [(H+H),4H] */
  float* cell_bias[4];     /* [4H] */
  float* memory_kernel;    /* [E,A]  (memory_layer, no bias) */
  float* query_kernel;     /* [H,A]  (query_layer, no bias) */
  float* attention_v;      /* [A] */
  float* conv_kernel;      /* [filtersize,1,numfilt] or NULL */
  float* conv_dense_kernel;/* [numfilt,A] or NULL (process_conv_features) */
  float* out_kernel;       /* [(H+E),V] */
  float* out_bias;         /* [V] */
} nabu_speller_params_t;

size_t nabu_speller_workspace_bytes(const nabu_speller_desc_t* d);
size_t nabu_speller_saved_bytes(const nabu_speller_desc_t* d);
/* targets [B,ldt] int32 (no SOS; the entry point prepends V-1), target_len [B].
 * logits [B,U,V] (zero rows for u >= target_len[b]); saved = activations for the backward. */
int nabu_speller_fwd(const nabu_speller_desc_t* d, const nabu_speller_params_t* p,
                     const float* memory, const int* mem_len, const int* targets, int ldt,
                     const int* target_len, float* logits, void* saved,
                     void* workspace, size_t ws_bytes, void* stream);
int nabu_speller_bwd(const nabu_speller_desc_t* d, const nabu_speller_params_t* p,
                     const float* memory, const int* mem_len, const int* targets, int ldt,
                     const int* target_len, const float* dlogits, void* saved,
                     float* dmemory, const nabu_speller_params_t* grads,
                     void* workspace, size_t ws_bytes, void* stream);

/* ---- a8 on its own: the attention mechanism step by step -------------------------------------------
 * Replaces components/attention.py:142-184 (LocationAwareAttention.__call__), :186-240 (_bahdanau_location_score),
 * :24-30 (vanilla Bahdanau), :294-396 (windowed) and the memory_layer / context computation of TF's BahdanauAttention
 * / AttentionWrapper -- what models/ed_decoders/speller.py:57-61 builds as a separate object and
 * components/beam_search_decoder.py:176 steps.  nabu_speller_fwd/bwd and nabu_las_beam_search run the same kernels
 * fused with the LSTM cell and the projection; these entry points expose the mechanism alone.
 *   nabu_attn_keys:      values [B,Tm,E] = memory * sequence_mask(mem_len); keys [B,Tm,A] = values . memory_kernel
 *   nabu_attn_step_fwd:  query [R,H] (R = B * rows_per_mem decoder rows, rows_per_mem consecutive rows share a memory
 *                        row = tile_batch) and align_prev [R,Tm]  ->  align_new [R,Tm], context [R,E] = align_new . values.
 *                        q_save [R,A], cf_save [R,Tm,numfilt] (location_aware), asum_save [R] (normalized_sigmoid) are the
 *                        activations nabu_attn_step_bwd needs; each may be NULL when no backward follows.
 *   nabu_attn_step_bwd:  rows_per_mem = 1.  dalign_new [R,Tm] (NULL = 0) and dcontext [R,E] in; dquery [R,H] and
 *                        dalign_prev [R,Tm] out; dkeys [B,Tm,A] and dvalues [B,Tm,E] are ACCUMULATED (+=: they collect
 *                        over the decoder steps); grads->{query_kernel, attention_v, conv_kernel, conv_dense_kernel}
 *                        receive THIS step's parameter gradients (overwritten). */
size_t nabu_attn_workspace_bytes(const nabu_speller_desc_t* d, int R);
int nabu_attn_keys(const nabu_speller_desc_t* d, const nabu_speller_params_t* p, const float* memory,
                   const int* mem_len, float* values, float* keys, void* stream);
int nabu_attn_step_fwd(const nabu_speller_desc_t* d, const nabu_speller_params_t* p, const float* query, int R,
                       int rows_per_mem, const float* keys, const float* values, const int* mem_len,
                       const float* align_prev, float* align_new, float* context, float* q_save, float* cf_save,
                       float* asum_save, void* workspace, size_t ws_bytes, void* stream);
int nabu_attn_step_bwd(const nabu_speller_desc_t* d, const nabu_speller_params_t* p, const float* query, int R,
                       const float* keys, const float* values, const int* mem_len, const float* align_prev,
                       const float* align_new, const float* q_save, const float* cf_save, const float* asum_save,
                       const float* dalign_new, const float* dcontext, float* dquery, float* dalign_prev,
                       float* dkeys, float* dvalues, const nabu_speller_params_t* grads, void* workspace,
                       size_t ws_bytes, void* stream);

/* ---- a13: LAS beam search -------------------------------------------------------------------------
 * Replaces decoders/beam_search_decoder.py:30-112 + components/beam_search_decoder.py:136-485
 * (initialize / step / finalize under dynamic_decode(maximum_iterations=max_steps)).
 * memory [B,Tm,E] un-tiled.  Outputs: sequences [B,W,max_steps] int32 (only the first *n_steps
 * columns are meaningful), lengths [B,W] int32, scores [B,W], alignments [B,W,max_steps,Tm].
 * n_steps: host int, written after a stream synchronise inside the call (the loop length is
 * data dependent: the reference stops when every beam slot has held EOS once). */
size_t nabu_las_beam_workspace_bytes(const nabu_speller_desc_t* d, int W, int max_steps);
int nabu_las_beam_search(const nabu_speller_desc_t* d, const nabu_speller_params_t* p,
                         const float* memory, const int* mem_len, int W, int max_steps,
                         float length_penalty, float temperature,
                         int* sequences, int* lengths, float* scores, float* alignments,
                         int* n_steps, void* workspace, size_t ws_bytes, void* stream);

/* ---- a12: CTC prefix beam search ------------------------------------------------------------------
 * Replaces decoders/ctc_decoder.py:57-59 (tf.nn.ctc_beam_search_decoder: beam_width=100,
 * top_paths=1, merge_repeated=True).  logits [B,T,V] batch-major raw scores, blank = V-1.
 * out_ids [B,T] int32 (best path per utterance), out_len [B]. */
size_t nabu_ctc_beam_workspace_bytes(int B, int T, int V, int beam_width);
int nabu_ctc_beam_search(const float* logits, const int* logit_len, int B, int T, int V,
                         int beam_width, int merge_repeated, int* out_ids, int* out_len,
                         float* out_neg_logprob, void* workspace, size_t ws_bytes, void* stream);

/* The name SURVEY.md section 8b lists for the same entry point. */
int nabu_ctc_prefix_beam(const float* logits, const int* logit_len, int B, int T, int V,
                         int beam_width, int merge_repeated, int* out_ids, int* out_len,
                         float* out_neg_logprob, void* workspace, size_t ws_bytes, void* stream);

/* ---- e: the data-parallel step's one collective ------------------------------------------------------
 * Replaces the gradient exchange of trainers/trainer.py:479-510, 556-569 (asynchronous parameter servers in the
 * reference; here ONE synchronous ncclAllReduce(SUM, fp32) of the flat gradient buffer over NVLink, enqueued on the
 * caller's stream like every other entry point).  One process per GPU: rank 0 calls nabu_comm_unique_id (HOST buffer
 * of 128 bytes), hands the bytes to the other ranks by any means (torch.distributed, MPI, a file) and every rank calls
 * nabu_comm_init with them.  Without a communicator nabu_allreduce_grads is the identity (a one-GPU job).
 * NCCL is resolved with dlopen("libnccl.so.2") on the first of these calls. */
int nabu_comm_unique_id(void* id128_host);
int nabu_comm_init(const void* id128_host, int rank, int world);
int nabu_comm_world(void);
int nabu_comm_destroy(void);
int nabu_allreduce_grads(float* grads, size_t n, void* stream);

/* ---- host helper (rows f1 / f3: TFRecord frames, TF checkpoint bundles) ---------------------------
 * CRC-32C (Castagnoli, reflected 0x82F63B78) of `nbytes` HOST bytes, continuing from `crc` (0 to start).
 * Replaces tensorflow/core/lib/hash/crc32c (external) behind tf.python_io.TFRecordWriter
 * (processing/tfwriters/tfwriter.py:34-55) and tf.train.Saver (components/hooks.py:6-52).  Needs no device. */
unsigned int nabu_crc32c(const void* data, size_t nbytes, unsigned int crc);

#ifdef __cplusplus
}
#endif
#endif /* NABU_B200_H_ */
