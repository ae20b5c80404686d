/* libmultinn_sm100.so -- C ABI of the B200-native MultINN hot path.
 *
 * The reference (ilya16/MultINN) has no FFI: its boundary is the Python module interface
 * (Encoder / Generator / RnnEstimator / NADE / RBM classes) over TF 1.13.1 library ops. Each entry
 * point below replaces the TF ops behind one reference call site (cited as file:line relative to
 * /root/reference/multinn/). The host-side mirror of the Python interface lives in multinn_b200/.
 *
 * Conventions: every function returns 0 on success, a negative MNN_ERR_* code for argument errors,
 * or a positive cudaError_t; mnn_last_error_string() describes the last failure of the calling
 * thread. All pointers are DEVICE pointers owned by the caller (workspace included); fp32,
 * row-major; `ld*` are row strides in elements; the last argument is the CUDA stream. No function
 * allocates, synchronises or throws. Sequence tensors are time-major: row n' = t * B + b.
 */
#ifndef MULTINN_B200_H_
#define MULTINN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mnn_stream_t;

#define MNN_OK 0
#define MNN_ERR_ARG (-1)
#define MNN_ERR_UNSUPPORTED (-2)
#define MNN_ERR_WORKSPACE (-3)

int mnn_version(void);
const char* mnn_last_error_string(void);
/* Number of CUDA kernels this library has launched in the calling process (bench.py gpu_launches). */
unsigned long long mnn_launch_count(void);

/* K0 -- input staging. core/multi_encoder_nn.py:66-87 (_build_inputs zero-pad + unstack, _build_targets),
 * multinn_composer.py:73-87 (stack axis=3, reshape, [:, :-1] / [:, 1:] shift), multinn_jamming.py:60-68.
 * x[B,T,D,M] -> xin[(T+1),B,D*M] (slot 0 zero; feature d*M+m), xtr[M,(T+1),B,D] (optional),
 * bits[M,T*B,4] target bit masks. Any of xin/xtr/bits may be NULL. D <= 128. */
int mnn_pack_pianoroll(const float* x, float* xin, float* xtr, uint32_t* bits, int B, int T, int D, int M,
                       mnn_stream_t stream);
/* Same, x given as bytes (bool / uint8 piano-rolls as stored by prepare_data.py:56). */
int mnn_pack_pianoroll_u8(const uint8_t* x, float* xin, float* xtr, uint32_t* bits, int B, int T, int D, int M,
                          mnn_stream_t stream);
/* Bit masks of an already flattened binary matrix v[N,D] (element (n,d) at v[n*ld + d*dim_stride]). */
int mnn_pack_rows(const float* v, long long ld, int dim_stride, uint32_t* bits, int N, int D, mnn_stream_t stream);

/* fp32 GEMM: C = alpha*op(A)*op(B) + beta*C (+ bias[n]). tf.matmul / tf.layers.Dense at common/rnn.py:124
 * (LSTMBlockCell xh.W), generators/rnn_nade.py:54-57,212 (Dense output_layer), generators/rnn_rbm.py:252-253
 * (Wuh, Wuv), common/rbm.py:351,370, common/dnn.py:56-60. transA=0: A is [M,K]; 1: A is stored [K,M]. */
int mnn_gemm_f32(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                 long long ldc, const float* bias, float alpha, float beta, int M, int N, int K, mnn_stream_t stream);

/* Same contract on the tcgen05 tensor cores with fp32 accuracy (3xTF32 error-compensated splitting, fp32 accumulation
 * in TMEM); a_exact != 0 promises that A is exactly representable in tf32 (binary piano-roll rows), which drops one
 * of the three products. Needs 16-byte aligned A/B and lda/ldb % 4 == 0 (TMA); mnn_gemm_tc_supported() tells. */
int mnn_gemm_tc_supported(const float* A, long long lda, const float* B, long long ldb);
int mnn_gemm_tc(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                long long ldc, const float* bias, float alpha, float beta, int M, int N, int K, int a_exact,
                mnn_stream_t stream);

/* Caps the grids of this thread's following mnn_gemm_tc launches at `sms` SMs (0 = whole device): a projection GEMM of
 * one time chunk can then run beside the persistent recurrence kernels of other streams (layer wavefront at small
 * per-GPU batches) instead of queueing CTAs behind them. Host-side, thread-local; no reference counterpart. */
int mnn_set_sm_budget(int sms);

/* Operand split of this thread's following mnn_gemm_tc launches on the CTA-pair kernel (host-side, thread-local; no
 * reference counterpart -- the reference's tf.matmul is plain fp32):
 *   0 (default)  "2.5 products": tf32(A).tf32(B) + bf16(A).bf16(B - tf32(B)) + bf16(A - tf32(A)).bf16(B), about 2^-20
 *                per product; what evaluate(), generation and every parity test of a single GEMM use
 *   1            bf16 PAIRS x = x1 + x2: A1.B1 + A1.B2 + A2.B1, all kind::f16 MMAs, about 2^-17 per product, 9 % faster
 *                (half the operand bytes of the main product). The training step (modes/core.py step()) selects it:
 *                loss and gradients stay inside the 1e-4 parity bar (tests/test_gpu_model.py trajectories and gradients
 *                run through it). The environment variable MNN_GEMM_BF16X (0 / 1 / 2) overrides both. */
int mnn_set_gemm_split(int mode);

/* Weights as bf16 pairs, split once per training step instead of once per output tile: mnn_split_bf16_pair writes
 * W[rows, cols] (row stride ld floats) as two bf16 planes [rows][ld_elems] (hi = bf16(w), then lo = bf16(w - hi); ld_elems
 * a multiple of 8, pad columns zero) into dst (2 * rows * ld_elems * 2 bytes, 16-byte aligned). mnn_gemm_tc_bpair is
 * mnn_gemm_tc with B given as such planes (Bpair, row stride ldb_elems; the planes' logical shape is [K, N] for
 * transB == 0 and [N, K] otherwise): TMA drops the planes straight into the operand tiles of the bf16-pair split, so the
 * B half of the in-kernel conversion and its shared-memory traffic disappear. CTA-pair kernel only (M >= 256, N > 128,
 * K >= 64: MNN_ERR_UNSUPPORTED otherwise); used for the input-projection, Dense and data-gradient GEMMs of the
 * training step (common/rnn.py, generators/rnn_nade.py here; tf.matmul in common/rnn.py:124, rnn_nade.py:54-57,212). */
int mnn_split_bf16_pair(const float* src, long long ld, int rows, int cols, void* dst, long long ld_elems,
                        mnn_stream_t stream);
int mnn_gemm_tc_bpair(const float* A, long long lda, int transA, const void* Bpair, long long ldb_elems, int transB,
                      float* C, long long ldc, const float* bias, float alpha, float beta, int M, int N, int K, int a_exact,
                      mnn_stream_t stream);

/* Binary A operands as an exact bf16 plane (a piano-roll bit is exact in bf16): mnn_pack_stacked_bf16 writes the stacked,
 * zero-padded input rows of core/multi_encoder_nn.py:66-76 + multinn_composer.py:73-80 -- xin16[(t + 1) * B + b][i] =
 * x[b][t][i], i = d * M + m, slot t = 0 and the pad columns zero, row stride ld16 elements (multiple of 8) -- from the
 * [B, T, I] piano-roll batch (float32, or uint8 when x_is_u8). mnn_gemm_tc_abf16 is mnn_gemm_tc with A given as such a
 * plane (row stride lda_elems; [M, K] for transA == 0, [K, M] otherwise) and B either fp32 (Bpair == NULL) or pre-split
 * (Bpair, ldb_elems as in mnn_gemm_tc_bpair): no raw A stage and no A conversion in the kernel, half the A bytes from
 * HBM. CTA-pair kernel only. Used for the layer-0 input projection and the x-rows weight gradient of the training step. */
int mnn_pack_stacked_bf16(const void* x, int x_is_u8, void* xin16, long long ld16, int B, int T, int I, mnn_stream_t stream);
int mnn_gemm_tc_abf16(const void* A16, long long lda_elems, int transA, const float* B, long long ldb, const void* Bpair,
                      long long ldb_elems, int transB, float* C, long long ldc, const float* bias, float alpha, float beta,
                      int M, int N, int K, mnn_stream_t stream);

/* Data-parallel noise keying (no reference counterpart: the reference is single-device; SURVEY 8(e) asks that results
 * do not depend on the GPU count). Every entry point that can draw Philox noise (dropout in mnn_lstm_*_fwd*, the
 * Bernoulli draws of mnn_bias_sigmoid_sample, mnn_rbm_gibbs, mnn_nade_sample, mnn_sample_steps) keys the counter by the
 * GLOBAL row of its local row r:  (r / rows_local) * rows_global + row_base + r % rows_local  -- for a time-major
 * [T, B_local, .] tensor of a rank that owns sequences [row_base, row_base + B_local) of a global batch B_global:
 * rows_local = B_local, rows_global = B_global. Thread-local, applies to the calling thread's following launches;
 * (0, 0, 0) restores the identity. Supplied uniforms are unaffected. */
int mnn_set_row_map(long long rows_local, long long rows_global, long long row_base);
/* Index of the first time step of the calling thread's following mnn_lstm_* launches inside the whole sequence: the
 * dropout counter is (global batch row, unit, GLOBAL time step), so a sequence run in time chunks (the small-batch
 * pipelines) draws the same masks as one launch over all T steps. 0 by default. */
int mnn_set_time_base(long long t_base);

/* K2 -- LSTM temporal unit. common/rnn.py:104-145 (CudnnCompatibleLSTMCell, gate blocks i,j,f,o, forget_bias 0;
 * DropoutWrapper output_keep_prob; MultiRNNCell), driven like dynamic_decode at generators/rnn_nade.py:204-218.
 * One cell step: gates[B,4R] in = pre-activations, out = activations; out = h/keep*floor(keep+u). */
int mnn_lstm_cell_fwd(float* gates, const float* c_prev, float* c, float* h, float* out, float* dscale,
                      const float* u, float keep, unsigned long long seed, unsigned long long offset, int B, int R,
                      mnn_stream_t stream);
/* Whole sequence, one layer. gates[T,B,4R] in = x.Wx + b for every step, out = gate activations (saved for
 * backward); wh[R,4R] = kernel rows I..I+R-1; hbuf/cbuf[(T+1),B,R] slot 0 = initial state, slot t+1 = state
 * after step t; out[T,B,R] = dropped-out h (NULL to skip); dscale[T,B,R] required when keep < 1;
 * u[T,B,R] uniforms or NULL (Philox(seed)). */
int mnn_lstm_seq_fwd(float* gates, const float* wh, float* hbuf, float* cbuf, float* out, float* dscale,
                     const float* u, float keep, unsigned long long seed, int T, int B, int R, mnn_stream_t stream);
/* BPTT of the same: gates in = activations, out = d(pre-activations) [T,B,4R]; dout[T,B,R] = grad wrt the
 * dropped-out outputs; dh_work/dc_work[B,R] scratch. Weight/input grads are then GEMMs over all rows. */
int mnn_lstm_seq_bwd(float* gates, const float* wh, const float* cbuf, const float* dout, const float* dscale,
                     float* dh_work, float* dc_work, int T, int B, int R, mnn_stream_t stream);
/* The same two sequence passes on the tensor cores: per step one tcgen05 3xTF32 GEMM (h_{t-1}.Wh, resp. dG_{t+1}.Wh^T)
 * with the LSTM cell (resp. its backward) fused into the epilogue; persistent != 0 runs all T steps in ONE cooperative
 * launch (CTAs sharing a 128-row batch slab hand h_t / dG_t over through per-slab counters in `ws`). Needs R % 8 == 0.
 * ws >= mnn_lstm_workspace_bytes(B, R); dc_work[B,R] scratch. Philox dropout differs in stream from the elementwise path. */
size_t mnn_lstm_workspace_bytes(int B, int R);
int mnn_lstm_tc_supported(int B, int R);
/* SMs (one CTA each) the persistent forward / BPTT kernel of a layer occupies under the calling thread's SM budget
 * (mnn_set_sm_budget): the host sizes the budget of the batched work it runs beside the recurrences from these. */
int mnn_lstm_seq_fwd_ctas(int T, int B, int R);
int mnn_lstm_seq_bwd_ctas(int T, int B, int R);
int mnn_lstm_seq_fwd_tc(float* gates, const float* wh, float* hbuf, float* cbuf, float* out, float* dscale,
                        const float* u, float keep, unsigned long long seed, int T, int B, int R, void* ws,
                        int persistent, mnn_stream_t stream);
int mnn_lstm_seq_bwd_tc(float* gates, const float* wh, const float* cbuf, const float* dout, const float* dscale,
                        float* dc_work, int T, int B, int R, void* ws, int persistent, mnn_stream_t stream);
/* BPTT of one time chunk [t0, t0+T) of a longer sequence (pointers at the chunk's first slot). has_next != 0: the chunk
 * after it has already been back-propagated, i.e. gates slot T holds dG of step t0+T and dc_work the carried cell
 * gradient, so the chunk's last step takes its recurrent term like every other step. */
int mnn_lstm_seq_bwd_tc_chunk(float* gates, const float* wh, const float* cbuf, const float* dout, const float* dscale,
                              float* dc_work, int T, int B, int R, void* ws, int persistent, int has_next,
                              mnn_stream_t stream);
/* out[c] (+)= sum_r A[r,c] (bias gradients); deterministic two-stage sum, ws >= mnn_colsum_workspace_bytes(cols). */
size_t mnn_colsum_workspace_bytes(int cols);
int mnn_colsum(const float* A, long long ld, int rows, int cols, float* out, int accumulate, void* ws,
               mnn_stream_t stream);

/* K4 -- NADE / MultiNADE teacher-forced log-likelihood. common/nade.py:155-229 (log_prob), :310-329
 * (_cond_prob), utils/auxiliary.py:9-11 (safe_log), generators/rnn_multinade.py:231-290 (bias split, per-track
 * loop). fc[N,ld] is the Dense output read in place: b_enc of track m at column enc_col0 + m*H, b_dec at
 * dec_col0 + m*D. w_enc/w_dec[M,D,H]. Outputs nll[M,N] (positive), cond_p[M,N,D] (optional). When dfc != NULL
 * (training) the d b_dec columns of dfc[N,ld] receive gscale * dNLL/dl. track_stride (0 = N): rows between
 * consecutive tracks in bits / nll / cond_p, so that a chunk of N rows of longer [M,TS,..] buffers can be processed
 * (all row pointers pre-offset by the caller). The grid honours mnn_set_sm_budget. */
int mnn_nade_logprob_fwd(const uint32_t* bits, const float* fc, long long ld, int enc_col0, int dec_col0,
                         const float* w_enc, const float* w_dec, float* nll, float* cond_p, float* dfc, float gscale,
                         int N, int M, int D, int H, long long track_stride, mnn_stream_t stream);
/* Which forward kernel mnn_nade_logprob_fwd runs: 0 = the tcgen05 segment-row kernel (nade_tc.cu: H rows synthesised
 * from the bit masks as bf16 hi/lo tiles, W_dec split and resident in shared memory, logits in TMEM) where the shape
 * fits (H in {128, 256}, D <= 96), 1 (default: the faster one today) = the SIMT segment kernel always. Process-wide;
 * the environment variable MNN_NADE_MODE=0 makes the tensor-core kernel the default. */
int mnn_set_nade_mode(int mode);
/* K5 -- its backward (what tf.gradients builds through nade.py:199-226). Needs the d b_dec columns written by
 * the forward; writes the d b_enc columns of dfc and ACCUMULATES into dw_enc/dw_dec[M,D,H]. */
int mnn_nade_logprob_bwd(const uint32_t* bits, const float* fc, long long ld, int enc_col0, int dec_col0,
                         const float* w_enc, const float* w_dec, float* dfc, float* dw_enc, float* dw_dec, int N,
                         int M, int D, int H, long long track_stride, mnn_stream_t stream);
/* K6 -- NADE ancestral sampling, common/nade.py:231-308 (+ tfp Bernoulli(logits).sample(), :283-287;
 * rnn_multinade.py:292-317 stacks tracks with axis=2). u[M,N,D] uniforms; u == NULL and use_philox == 0 means
 * temperature=None (p >= .5). out[row*out_ld + i*out_dim_stride + m*out_track_stride] in {0,1}. */
int mnn_nade_sample(const float* fc, long long ld, int enc_col0, int dec_col0, const float* w_enc,
                    const float* w_dec, const float* u, int use_philox, unsigned long long seed,
                    unsigned long long offset, float* out, long long out_ld, int out_dim_stride, int out_track_stride,
                    float* nll, int N, int M, int D, int H, mnn_stream_t stream);

/* K6 fused -- the whole generation scan of generators/rnn_estimator.py:271-323 (`generate`: per step sample_single =
 * M x NADE.sample, then single_step = MultiRNNCell + Dense output layer, rnn_nade.py:253-277) for num_steps steps in ONE
 * cooperative launch; B <= 128 (sample.py's batches: num_songs x intros). The grid's CTAs split into sampler / LSTM
 * layer / Dense groups that keep their weights in shared memory for the whole call and hand the state over through
 * counters in `ws` (gen_fused.cu). State after the intro scan comes in, state after the last step goes out: c_l, h_l
 * [B, r_l] per layer (1 or 2 layers; layer 1 pointers NULL for one), fc[B, ldfc] = the Dense output (NADE biases,
 * columns [M*H b_enc | M*D b_dec]). Layer-0 input = the sampled frame (num_inputs == M*D, feature d*M + m). u[S, M, B, D]
 * uniforms or NULL (+ use_philox: counter ((global row * M + m) * D + i, offset0 + step), key seed; neither: p >= .5).
 * out[b*out_ld + s*out_step + d*M + m] in {0,1}. ws >= mnn_generate_fused_workspace_bytes(...) (0: shape not taken).
 * Arithmetic: bf16-pair operands (16 mantissa bits) on tcgen05, fp32 accumulation; the multi-launch path
 * (mnn_nade_sample + mnn_gemm_tc + mnn_lstm_cell_fwd per step) stays the fp32-accurate one. */
size_t mnn_generate_fused_workspace_bytes(int num_layers, int num_inputs, int r0, int r1, int B, int M, int D, int H);
int mnn_generate_fused(int num_layers, int num_inputs, const float* kern0, const float* bias0, float* c0, float* h0, int r0,
                       const float* kern1, const float* bias1, float* c1, float* h1, int r1, const float* dense_kernel,
                       const float* dense_bias, const float* w_enc, const float* w_dec, float* fc, long long ldfc,
                       const float* u, int use_philox, unsigned long long seed, unsigned long long offset0, float* out,
                       long long out_ld, long long out_step, int B, int S, int M, int D, int H, void* ws,
                       mnn_stream_t stream);

/* K7 -- RBM half-steps: p = sigmoid(pre + bias), s = float(u < p). common/rbm.py:337-387, used by forward
 * (:148-167), reconstruct (:169-190), the Gibbs chain (:192-231), DBN (common/dbn.py:136-180) and the sigmoid
 * Dense feedback module (common/dnn.py:97-116). ld_bias == 0 broadcasts one bias row. */
int mnn_bias_sigmoid_sample(const float* pre, long long ld_pre, const float* bias, long long ld_bias, const float* u,
                            long long ld_u, int use_philox, unsigned long long seed, unsigned long long offset,
                            float* p, long long ld_p, float* s, long long ld_s, int N, int C, mnn_stream_t stream);
/* K7 fused -- the whole k-step Gibbs chain of common/rbm.py:192-231 (tf.while_loop of 2k {matmul, sigmoid, Bernoulli}
 * pairs) in ONE launch: v0[N,D] -> p_v = p(v | h_k) of the last step, v_k, and optionally h_k. W[D,H] (contiguous) and its
 * transpose are staged in shared memory; bh/bv are per-row [N,.] (ld > 0) or one broadcast row (ld == 0) or NULL;
 * uh[k,N,H] / uv[k,N,D] contiguous uniforms (parity runs) or both NULL with use_philox (counter = (offset + row,
 * half-step, column group), key = seed: a function of the global row only). mnn_rbm_gibbs_smem_bytes() is 0 for shapes
 * the kernel does not take (D, H multiples of 4, <= 256, 2*D*H floats + buffers within 227 KB): callers then run the
 * chain as GEMM + mnn_bias_sigmoid_sample half-steps. All float pointers but v0 16-byte aligned, strides % 4 == 0. */
size_t mnn_rbm_gibbs_smem_bytes(int D, int H);
int mnn_rbm_gibbs(const float* v0, long long ld_v, const float* W, const float* bh, long long ld_bh, const float* bv,
                  long long ld_bv, const float* uh, const float* uv, int use_philox, unsigned long long seed,
                  unsigned long long offset, float* p_v, long long ld_p, float* v_k, long long ld_vk, float* h_k,
                  long long ld_hk, int N, int D, int H, int k, mnn_stream_t stream);
/* d(pre) = dy * y * (1 - y): backward of a sigmoid layer (Dense feedback module, common/dnn.py:56-60). */
int mnn_sigmoid_bwd(const float* y, long long ld_y, const float* dy, long long ld_dy, float* dpre, long long ld_d, int N,
                    int C, mnn_stream_t stream);
/* F(v)[n] = -sum_j softplus(pre[n,j] + bh[j]) - v[n].bv, common/rbm.py:256-258 (pre = v.W). */
int mnn_rbm_free_energy(const float* pre, long long ld_pre, const float* bh, long long ld_bh, const float* v,
                        long long ld_v, const float* bv, long long ld_bv, float* F, int N, int H, int D,
                        mnn_stream_t stream);

/* K9 -- reductions and the optimiser. utils/training.py:151-177 (clip_by_global_norm(5.)), train.py:61-64
 * (AdamOptimizer(lr, epsilon=1e-4) / GradientDescentOptimizer), metrics/statistical.py:34 (reduce_mean). */
size_t mnn_reduce_workspace_bytes(void);
int mnn_sum(const float* x, size_t n, void* ws, float* out, float scale, int accumulate, mnn_stream_t stream);
int mnn_sqnorm(const float* x, size_t n, void* ws, float* out, mnn_stream_t stream);
int mnn_clip_adam(float* p, const float* g, float* m, float* v, size_t n, const float* sqnorm, float grad_scale,
                  float clip_norm, float lr, float beta1, float beta2, float eps, int step, mnn_stream_t stream);
/* x[r*ld + c] *= w[r % period], c < ncols. Variable sequence lengths: utils/sequences.py:6-37 drops the rows t >= lengths[b]
 * from every per-row result (flatten_maybe_padded_sequences); here such rows get weight 0 in the NLL and in dNLL/dl. */
int mnn_scale_rows(float* x, long long ld, int ncols, const float* w, long long rows, int period, mnn_stream_t stream);
/* y += alpha * x (CD-k assign_add, common/rbm.py:322-330). */
int mnn_axpy(float* y, const float* x, float alpha, size_t n, mnn_stream_t stream);
int mnn_clip_sgd(float* p, const float* g, size_t n, const float* sqnorm, float grad_scale, float clip_norm,
                 float lr, mnn_stream_t stream);

/* Roofline probe (no reference counterpart): `iters` rounds of 8 independent MUFU chains per thread -- kind 0:
 * ex2.approx, 1: rcp.approx, 2: the kernels' sigmoid (ex2 + rcp + 2 FMA-pipe ops). MUFU instructions executed =
 * blocks * threads * iters * 8 (x2 for kind 2); bench.py times it with CUDA events to put an XU-pipe peak beside the HBM and
 * tensor peaks of MEASURED_PEAKS.json. out: >= blocks floats. */
int mnn_probe_mufu(float* out, int blocks, int threads, int iters, int kind, mnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTINN_B200_H_ */
