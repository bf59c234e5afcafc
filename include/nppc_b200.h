/*
 * nppc_b200.h — C-ABI of libnppc_b200.so: hand-written sm_100a CUDA kernels for the NPPC-over-FullSubNet+
 * speech-restoration hot path of kfirc1503/generative-audio.
 *
 * The reference has NO native code / FFI: its boundary for this path is a set of Python
 * functions that land in ATen library kernels.  Each entry point below replaces one of those call sites
 * (cited as path:line under /root/reference).  Conventions:
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - tensors are dense, row-major, fp32 unless stated; shapes are given in the comment of each function;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); kernels are enqueued
 *     asynchronously on it, no host synchronisation, no internal streams;
 *   - return value: 0 on success, <0 on error (NPPC_ERR_*); nppc_last_error() gives the message
 *     (thread-local).  Errors mirror the reference's Python `assert`/`raise` sites.
 *   - inputs are borrowed, outputs are caller-allocated; the library owns no tensor memory except
 *     the opaque plan objects created by *_create and released by *_destroy.
 */
#ifndef NPPC_B200_H_
#define NPPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPPC_OK 0
#define NPPC_ERR_INVALID_ARGUMENT (-1)
#define NPPC_ERR_CUDA (-2)
#define NPPC_ERR_UNSUPPORTED (-3)

const char* nppc_last_error(void);
/* Library / build identification: "nppc_b200 <version> sm_100a". */
const char* nppc_version(void);
/* Number of kernel launches issued by this library since the last reset (for bench.py's gpu_launches). */
long long nppc_launch_count(void);
void nppc_reset_launch_count(void);

/* ---- a1: STFT front end -------------------------------------------------------------------------
 * Replaces utils.prepare_input_from_waveform (utils.py:107-147): torch.hann_window + torch.stft(center=True,
 * reflect) + sqrt(re^2+im^2).  wave [B,L] -> mag, real, imag each [B,F,T], F = n_fft/2+1, T = 1 + L/hop.
 * Only n_fft == win == 512 with hop == 256 is compiled (every reference config: utils.py:14-17). */
int nppc_stft_mri(const float* wave, int B, int L, int n_fft, int hop, float* mag, float* real, float* imag,
                  void* stream);

/* ---- a15: iSTFT back end ------------------------------------------------------------------------
 * Replaces torch.istft(n_fft=512, hop=256, win=hann(512), center=True, length=length) as called at
 * utils.py:60-70, nppc_audio/validator.py:136-143.  real/imag [B,F,T] -> wave [B,length]. */
int nppc_istft(const float* real, const float* imag, int B, int T, int n_fft, int hop, int length, float* wave,
               void* stream);

/* ---- a9 + a10: cIRM decompress + mask apply -----------------------------------------------------
 * Replaces decompress_cIRM (audio_zen/acoustics/mask.py:57-60) followed by utils.crm_to_stft_components
 * (utils.py:241-249; conj != 0 reproduces its argument-order quirk: enhanced = conj(M)*N) or the correct
 * M*N of utils.noisy_to_enhanced / model_outputs_to_waveforms (utils.py:54,75-79; conj == 0).
 * crm [B,2,F,T] (compressed), real/imag [B,F,T] -> out_mag/out_real/out_imag [B,F,T] (out_mag may be NULL). */
int nppc_crm_decompress_apply(const float* crm, const float* real, const float* imag, int B, int FT, int conj,
                              float* out_mag, float* out_real, float* out_imag, void* stream);
/* decompress_cIRM alone (mask.py:57-60), elementwise over n values. */
int nppc_decompress_cirm(const float* m, long long n, float* out, void* stream);
/* build_complex_ideal_ratio_mask + compress_cIRM (mask.py:24-54): noisy/clean re,im [n] -> gt [n,2]->stored
 * planar as gt0[n], gt1[n] (the [B,2,F,T] layout the trainer permutes to, trainer.py:359). */
int nppc_build_cirm(const float* nr, const float* ni, const float* cr, const float* ci, int B, int FT,
                    float* gt /* [B,2,FT] */, void* stream);

/* ---- N1: PC directions -> spectrogram variations (NPPCAudioValidator._crm_directions_to_spectograms and the alpha
 * sweep of visualize_pc_spectrograms, nppc_audio/validator.py:55-102,246-290; utils.crm_to_spectogram utils.py:252-256):
 *   pc[b,d] = decompress_cIRM(w_mat[b,d]) * noisy[b]   (M*N),   var[b,d,a] = enhanced[b] + alphas[a] * pc[b,d]
 * w_mat [B,n,2,FT]; noisy_* / enh_* [B,FT]; alphas [A] (device); pc_* [B,n,FT] (optional, both or NULL); var_* [B,n,A,FT].
 * nppc_peak_normalize: save_audio_files' x /= max|x| + 1e-8 per waveform row (validator.py:118-134,281-282), in place. */
int nppc_pc_variations(const float* w_mat, const float* noisy_real, const float* noisy_imag, const float* enh_real,
                       const float* enh_imag, int B, int n, int FT, const float* alphas, int A, float* pc_real,
                       float* pc_imag, float* var_real, float* var_imag, void* stream);
int nppc_peak_normalize(float* x, int rows, int L, void* stream);

/* ---- N3: GPU data preparation (batched versions of the reference's per-item CPU dataset code) -------------------
 * nppc_mix_with_snr: AudioDataset._normalize_audio + _mix_with_snr (dataset/audio_dataset.py:92-158): clean' = clean
 *   RMS-normalised to target_db[b] dBFS, noisy = clean' + noise * sqrt(P_clean' / (10^(snr_db[b]/10) P_noise + 1e-8)),
 *   both scaled by 0.99 / max|noisy| when that peak exceeds 0.99.  clean / noise [B,L]; scratch: nppc_mix_scratch_bytes(B).
 * nppc_time_to_spec_mask: AudioInpaintingDataset.time_to_spec_mask (dataset/audio_dataset_inpainting.py:223-251):
 *   out[b,t] = 1 iff min(mask_time[b, window of frame t]) == 1 (window clipped to [0,L); empty window -> 0). */
size_t nppc_mix_scratch_bytes(int B);
int nppc_mix_with_snr(const float* clean, const float* noise, int B, int L, const float* snr_db, const float* target_db,
                      void* scratch, float* noisy, float* clean_out, void* stream);
int nppc_time_to_spec_mask(const float* mask_time, int B, int L, int T_frames, int win_length, int hop_length, int center,
                           float* out, void* stream);

/* ---- a2: laplace norms --------------------------------------------------------------------------
 * offline_laplace_norm (audio_zen/model/base_model.py:210-224): y = x / (mean_{per sample}(x) + 1e-5).
 * x [B, n] -> y [B, n] (n = C*F*T).  `sums` is a [B] fp64 scratch (device). In-place (y == x) allowed. */
int nppc_offline_laplace_norm(const float* x, int B, long long n, double* sums, float* y, void* stream);
/* Fused F.pad(x,[0,look_ahead]) (fullsubnet_plus.py:158-160) + offline_laplace_norm: x [B,F,T] -> y [B,F,T+la]. */
int nppc_pad_offline_laplace_norm(const float* x, int B, int F, int T, int look_ahead, double* sums, float* y,
                                  void* stream);
/* Conditioning diagnostic of the offline normaliser (base_model.py:219-222): depth[b] = max(depth[b], (mean|x_b| /
 * (|mean x_b| + 1e-5)) / sqrt(count)) for x [B, n]; count = the number of elements the reference's mean runs over (n plus the
 * zero look-ahead frames).  ~1 for a random-sign sum; errors of relative size eps in x leave the normaliser amplified by
 * ~eps * depth.  NPPCModel(lstm_impl="auto") re-runs utterances above a threshold through the split-precision path. */
int nppc_cancel_depth(const float* x, int B, long long n, double count, double* sums /* [2B] scratch */, float* depth /* [B], in/out */,
                      void* stream);
/* cumulative_laplace_norm (base_model.py:227-257): x [BC,F,T] -> y; y[f,t] = x[f,t]/(cumsum_t(sum_f x)/(F(t+1)) + eps). */
int nppc_cumulative_laplace_norm(const float* x, int BC, int F, int T, float* y, void* stream);

/* ---- a5 + a6: sub-band unfold, drop_band (bit-exact index kernels) ------------------------------
 * BaseModel.unfold (base_model.py:15-46): x [B,C,F,T] -> out [B,F,C,2n+1,T], out[b,f,c,k,t] = x[b,c,reflect(f+k-n),t]. */
int nppc_unfold(const float* x, int B, int C, int F, int T, int num_neighbor, float* out, void* stream);
/* drop_band (audio_zen/acoustics/feature.py:254-285): x [B,C,F,T] -> out [B,C,F/G,T]; requires B > G. */
int nppc_drop_band(const float* x, int B, int C, int F, int T, int groups, float* out, void* stream);

/* ---- a12: Gram-Schmidt --------------------------------------------------------------------------
 * gram_schmidt_to_crm (nppc_audio/pc_wrapper.py:8-44): x [B,n,2,P] -> out [B,n,2,P]; complex MGS with the
 * reference's conjugated coefficient.  gram_schmidt_to_spec_mag (nppc_audio/inpainting/nppc/pc_wrapper.py:43-59):
 * x [B,n,P] real.  `scratch` must hold nppc_gs_scratch_bytes(B,n) bytes (fp64 Gram + coefficients). */
/* Backward of Gram-Schmidt + objective (pc_wrapper.py:8-44 with its detach() at :37, trainer.py:259-298 with the detach()
 * at :295): out[b,i] = sum_{k<n} C[b,i,k] x[b,k] + C[b,i,n] (gt[b] - pred[b]) with complex coefficients C [B, n, n+1, 2] that
 * the host derives from the forward pass's Gram matrix (generative-audio_b200/gs_backward.py) — one streaming pass.
 * x, out [B, n, 2, P]; gt, pred [B, 2, P].  The scratch of nppc_gs_loss_fused holds, per sample, the Hermitian Gram matrix
 * of (x_0 .. x_{n-1}, gt - pred) as 13 x 13 complex doubles (upper triangle) followed by the 12 x 12 complex-float coefficient
 * matrix A of w_i = sum_k A[i][k] x_k. */
int nppc_complex_lincomb(const float* x, const float* gt, const float* pred, int B, int n, long long P, const float* coef,
                         float* out, void* stream);
size_t nppc_gs_scratch_bytes(int B, int n);
int nppc_gram_schmidt_complex(const float* x, int B, int n, long long P, void* scratch, float* out, void* stream);
int nppc_gram_schmidt_real(const float* x, int B, int n, long long P, void* scratch, float* out, void* stream);

/* ---- a12 + a14 fused: Gram-Schmidt + NPPC projection / second-moment loss (forward) --------------
 * head [B,n,2,P] (PC-head output before orthogonalisation), gt/pred [B,2,P] (compressed-cIRM domain, already
 * drop_band-ed) -> w_mat [B,n,2,P] and the per-sample statistics of NPPCAudioTrainer.base_step
 * (nppc_audio/trainer.py:259-298): err_norm [B], err_proj [B,n,2] (re,im), w_norms [B,n] (already divided by
 * err_norm), reconst_err [B], second_moment_mse [B,n]. One Gram pass over head+err, one apply pass. */
int nppc_gs_loss_fused(const float* head, const float* gt, const float* pred, int B, int n, long long P,
                       void* scratch, float* w_mat, float* err_norm, float* err_proj, float* w_norms,
                       float* reconst_err, float* second_moment_mse, void* stream);
/* Loss statistics for an explicit w_mat (trainer.py:259-298), same outputs as above. */
int nppc_projection_loss(const float* w_mat, const float* gt, const float* pred, int B, int n, long long P,
                         void* scratch, float* err_norm, float* err_proj, float* w_norms, float* reconst_err,
                         float* second_moment_mse, void* stream);

/* Real (inpainting) variants (nppc_audio/inpainting/trainer/nppc_trainer.py:338-385, 1e-6 added to both norms before the
 * divisions): x / w_mat [B,n,P], gt / pred [B,P]; err_proj [B,n] real, err_norm includes the 1e-6. */
int nppc_gs_loss_fused_real(const float* x, const float* gt, const float* pred, int B, int n, long long P,
                            void* scratch, float* w_mat, float* err_norm, float* err_proj, float* w_norms,
                            float* reconst_err, float* second_moment_mse, void* stream);
int nppc_projection_loss_real(const float* w_mat, const float* gt, const float* pred, int B, int n, long long P,
                              void* scratch, float* err_norm, float* err_proj, float* w_norms, float* reconst_err,
                              float* second_moment_mse, void* stream);

/* ---- a16: inpainting variant glue ---------------------------------------------------------------
 * utils.preprocess_data / preprocess_log_magnitude (utils.py:281-306): spec [B,2,P] (re, im) -> log(|S| + 1e-6);
 * nppc_logmag_stats accumulates (sum, sum of squares) over the whole batch tensor into sums[2] (fp64);
 * nppc_logmag_apply writes (log|S| - mean) / std with the UNBIASED std over n_stat elements (torch.std default).
 * nppc_mask_blend: RestorationWrapper.forward (nppc_audio/inpainting/networks/unet.py:298-313)
 *   out[b,c,p] = x_in[b,0,p] * mask[b,p] + x[b,c,p] * (1 - mask[b,p]);  x_in == NULL gives x * (1 - mask)
 *   (AudioInpaintingPCWrapper.forward, nppc_audio/inpainting/nppc/pc_wrapper.py:75-81). */
int nppc_logmag_stats(const float* spec, int B, long long P, double* sums, void* stream);
int nppc_logmag_apply(const float* spec, int B, long long P, const double* sums, long long n_stat, float* out, void* stream);
int nppc_mask_blend(const float* x_in, int Cin, const float* x, const float* mask, int B, int C, long long P, float* out,
                    void* stream);

/* ---- a5+a2 fused for the sub-band LSTM: feature packing ------------------------------------------
 * Builds the sub-band model input of fullsubnet_plus.py:203-223 / networks.py:133-151 without materialising
 * the [B,F,34,T'] tensor: unfold(nbr_src, N) ++ fb ++ fbr ++ fbi, offline_laplace_norm over (F,S,T') per
 * sample, drop_band(groups) row selection, written TIME-MAJOR as xs [T', R_stride, KP] (KP >= S, zero padded;
 * rows R..R_stride-1 of every step are zero: the tensor-core LSTM wants R_stride % 128 == 0),
 * R = B*F' rows ordered like the reference's reshape(B*F', S, T').  nbr_src/fb/fbr/fbi are [B,F,T'].
 * xs_f32 (fp32) and/or xs_f16 (IEEE fp16, saturating) may be NULL.  `sums` [B] fp64 scratch.
 * cumulative != 0: norm_type = "cumulative_laplace_norm" (base_model.py:227-257 on the [B,F,S,T'] tensor): every row is divided
 * by the running mean of its own S features, cumsum_t(sum_k x) / (S (t+1)) + EPSILON, instead of the per-sample mean. */
int nppc_subband_pack(const float* nbr_src, const float* fb, const float* fbr, const float* fbi, int B, int F,
                      int Tp, int num_neighbor, int groups, int KP, int R_stride, int cumulative, double* sums, float* xs_f32,
                      void* xs_f16, void* stream);

/* ---- a3: TSSE channel attention ----------------------------------------------------------------------
 * ChannelTimeSenseSELayer.forward (audio_zen/model/module/attention_model.py:78-98): x [B,C,T] -> y = x * gate[b,c].
 * kersize, conv_w, conv_b are HOST arrays of 3 entries (conv_w[i]: device [C,1,k_i] depthwise weights, conv_b[i]: [C]);
 * fcat_* = feature_concate_fc (3->1), fc1 [Cr,C], fc2 [C,Cr].  scratch: 2*B*C floats (device). */
int nppc_tsse(const float* x, int B, int C, int T, const int* kersize, const float* const* conv_w,
              const float* const* conv_b, const float* fcat_w, const float* fcat_b, const float* fc1_w,
              const float* fc1_b, const float* fc2_w, const float* fc2_b, int C_reduced, float* scratch, float* y,
              void* stream);

/* ---- a4: TCN block, normalisation / depthwise half (causal_conv.py:96-108) ------------------------------
 * stats buffers are [B,2] fp64 (sum, sum of squares) on the device; prelu_* point to the 1-element PReLU weight.
 * nppc_prelu_stats: stats = moments of PReLU(y + bias[c]) per sample, y [B,C,T] (bias = the 1x1 convolution's, or NULL).
 * nppc_tcn_mid:     z = PReLU2(depthwise_dilated(GroupNorm1(PReLU1(y1)))) [B,C,T] and stats2 = moments of z.
 * nppc_tcn_out:     xnew = x + o*rstd2[b] + vb[c] - mean2[b]*rstd2[b]*u[c], where o = conv1x1(z; W2*diag(gamma2)),
 *                   u = W2 gamma2, vb = W2 beta2 + b2  (GroupNorm2 folded into the second 1x1 convolution). */
int nppc_prelu_stats(const float* y, int B, int C, int T, const float* bias /* [C] or NULL: added to y first */, const float* prelu_a,
                     double* stats, void* stream);
int nppc_tcn_mid(const float* y1, int B, int C, int T, const float* bias1 /* [C] or NULL */, const float* prelu1_a, const double* stats1, const float* gamma1,
                 const float* beta1, const float* dw_w, const float* dw_b, int dilation, const float* prelu2_a, float* z,
                 double* stats2, void* stream);
int nppc_tcn_out(const float* o, const float* x, int B, int C, int T, int C_hidden, const double* stats2, const float* u,
                 const float* vb, float* xnew, void* stream);

/* ---- a7: sub-band LSTM (2 layers) + fc -----------------------------------------------------------
 * Replaces nn.LSTM(I->H, 2 layers, batch_first) + nn.Linear(H->O) of SequenceModel
 * (audio_zen/model/module/sequence_model.py:31-38,79,113-123).  Gate order i,f,g,o; z = W_ih x + b_ih + W_hh h + b_hh.
 * Plan objects own the re-packed weight caches (derived from the fp32 master weights, never serialised). */
typedef struct nppc_lstm_plan nppc_lstm_plan;
/* weights are DEVICE pointers in the nn.LSTM state_dict layout: w_ih0 [4H,I], w_hh* [4H,H], w_ih1 [4H,H],
 * b_* [4H], fc_w [O,H], fc_b [O]. */
int nppc_lstm_plan_create(nppc_lstm_plan** plan, int I, int H, int O, const float* w_ih0, const float* w_hh0,
                          const float* b_ih0, const float* b_hh0, const float* w_ih1, const float* w_hh1,
                          const float* b_ih1, const float* b_hh1, const float* fc_w, const float* fc_b,
                          void* stream);
void nppc_lstm_plan_destroy(nppc_lstm_plan* plan);
/* Workspace bytes needed by nppc_lstm_forward for R rows x Tp steps with the given implementation
 * (impl: 0 = fp32 SIMT reference-precision path, 1 = fp16-operand / fp32-accumulate tcgen05 tensor-core path). */
size_t nppc_lstm_workspace_bytes(const nppc_lstm_plan* plan, int R, int Tp, int impl);
/* xs: time-major input [Tp, R_stride, KP] (fp32 for impl 0, fp16 for impl 1; KP / R_stride as given to
 * nppc_subband_pack; impl 1 needs R_stride % 128 == 0).
 * y: [R, O, Tp] fp32 — the layout SequenceModel.forward returns (sequence_model.py:122). */
int nppc_lstm_forward(const nppc_lstm_plan* plan, const void* xs, int R, int R_stride, int Tp, int KP, int impl,
                      void* workspace, size_t workspace_bytes, float* y, void* stream);

/* ---- a7, implementation 2: stepwise tensor-core LSTM for generic (I, H), with saved state and BPTT ------------------------
 * One tcgen05 GEMM launch per (layer, time step) whose epilogue is the LSTM cell (csrc/lstm_step.cu).  Replaces, for the
 * TRAINED PC head, nn.LSTM forward + backward (sequence_model.py:113-123 under trainer.py:100-106) and serves shapes the
 * persistent kernel is not built for (FullSubNet's full-band LSTM 257 -> 512, fullsubnet.py:39-47).  fp32 master weights in
 * the nn.LSTM layout are passed directly (repacked on the device every call).
 *   xs          packed time-major input [Tp, R_stride, KP]: fp16 (fast mode) or fp32 (precise mode), features >= I zero
 *   train != 0  gates / c / h of every step stay in the workspace for nppc_lstm_step_backward (same workspace, untouched)
 *   precise!=0  split-precision operands (hi + lo fp16 halves of x, h and the weights: ~22 mantissa bits), exact tanh / exp:
 *               fp32-class accuracy on the tensor cores, for utterances whose normaliser means cancel
 *   y           [R, O, Tp] fp32 (the layout SequenceModel.forward returns)
 * backward: dy [R, O, Tp] -> gradients of the ten parameters (nn.LSTM layout, overwritten) and, if dxs != NULL, of the packed
 * input ([Tp, R_stride, KP] fp32).  Gradient tensors are fp16 with a power-of-two loss scale chosen on the device. */
typedef struct {
    const float* w_ih[2];   /* [4H, I] / [4H, H] */
    const float* w_hh[2];   /* [4H, H] */
    const float* b_ih[2];
    const float* b_hh[2];
    const float* fc_w;      /* [O, H] */
    const float* fc_b;      /* [O] */
    int I, H, O;
} nppc_lstm_weights;
typedef struct {
    float* w_ih[2];
    float* w_hh[2];
    float* b_ih[2];
    float* b_hh[2];
    float* fc_w;
    float* fc_b;
} nppc_lstm_grads;
size_t nppc_lstm_step_workspace_bytes(int I, int H, int O, int R_stride, int Tp, int KP, int train, int precise);
int nppc_lstm_step_forward(const nppc_lstm_weights* w, const void* xs, int xs_is_f32, int R, int R_stride, int Tp, int KP,
                           int train, int precise, void* workspace, size_t workspace_bytes, float* y, void* stream);
int nppc_lstm_step_backward(const nppc_lstm_weights* w, const void* xs, int R, int R_stride, int Tp, int KP, void* workspace,
                            size_t workspace_bytes, const float* dy, const nppc_lstm_grads* grads, float* dxs, void* stream);
/* C[Mo, No] fp32 = A^T B with A [rows, Mo], B [rows, No] fp16 row-major (both consumed MN-major by tcgen05; split-K over
 * `splits` CTAs per tile into `partials` [splits, Mo, No], summed in a fixed order).  The weight-gradient GEMM of the LSTM and
 * of the TCN's 1x1 convolutions (dW = dY^T X).  rows % 64 == 0, Mo % 128 == 0, No % 64 == 0. */
int nppc_gemm_f16_atb(const void* A, const void* B, long long rows, int Mo, int No, int splits, float* partials, float* C,
                      void* stream);

/* ---- N4: the inpainting UNet's convolutions on tcgen05 (implicit GEMM, NHWC fp16) -----------------------------------------
 * Replaces F.conv2d(3x3, padding=1) + eval-mode BatchNorm + LeakyReLU(0.2) of double_conv (nppc_audio/inpainting/networks/
 * tmp_utils.py:8-35) and the decoder's torch.cat([skip, up]) + conv (tmp_utils.py:76-88): the K loop walks 9 taps x channel
 * blocks of up to TWO input tensors, the halo comes from TMA out-of-bounds zero fill.
 *   nppc_conv3x3_pack_weights: w [Cout, C0 + C1, 3, 3] fp32 (BatchNorm folded) -> fp16 [Cout, 9, C0p + C1p] (Cxp = Cx rounded up to 64)
 *   nppc_conv3x3_tc:           y [B,H,W,Cout] fp16 = leaky_relu(conv3x3(cat(x0 [B,H,W,C0p], x1 [B,H,W,C1p] or NULL)) + bias)
 *   nppc_nchw_to_nhwc_f16:     x [B,C,H,W] fp32 -> [B,H,W,Cp] fp16, channels >= C zero (network input)
 *   nppc_conv1x1_out:          outc (unet.py:276): x [B,H*W,Cin] fp16 -> y [B,Cout,H*W] fp32, Cout <= 16 */
int nppc_conv3x3_pack_weights(const float* w, int Cout, int C0, int C1, void* out, void* stream);
int nppc_conv3x3_tc(const void* x0, int C0p, const void* x1, int C1p, const void* w_packed, const float* bias, void* y, int B,
                    int H, int W, int Cout, float negative_slope, void* stream);
int nppc_nchw_to_nhwc_f16(const float* x, int B, int C, int H, int W, int Cp, void* y, void* stream);
int nppc_conv1x1_out(const void* x, int B, int HW, int Cin, const float* w, const float* bias, int Cout, float* y, void* stream);
/* nn.MaxPool2d(2) (tmp_utils.py:45) and nn.Upsample(x2, bilinear, align_corners=True) + the zero pad to the skip tensor's
 * H x W (tmp_utils.py:59-82) on NHWC fp16: [B,H,W,C] -> [B,H/2,W/2,C];  [B,h,w,C] -> [B,H,W,C]. */
int nppc_maxpool2x2_nhwc(const void* x, int B, int H, int W, int C, void* y, void* stream);
int nppc_upsample2x_pad_nhwc(const void* x, int B, int h, int w, int C, int H, int W, void* y, void* stream);

/* ---- a8/a11 output assembly -----------------------------------------------------------------------
 * y [B*F', O, T'] -> out [B, O, F', T'-la] dropping the first `look_ahead` frames
 * (fullsubnet_plus.py:227-229; networks.py:156-161 is the same memory layout with O = 2*n_dirs). */
int nppc_assemble_mask(const float* y, int B, int Fp, int O, int Tp, int look_ahead, float* out, void* stream);

/* ---- tensor-core GEMM (used by the LSTM input projections) ----------------------------------------
 * C[M,N] (bf16, row-major) = A[M,K] (bf16, row-major) * W[N,K]^T (bf16, row-major) + bias[N] (fp32, may be NULL).
 * tcgen05.mma + TMA + TMEM. K % 64 == 0, N % 128 == 0. */
int nppc_gemm_bf16_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K,
                      void* stream);
/* same with IEEE fp16 operands and output */
int nppc_gemm_f16_tn(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, void* stream);
/* fp16 operands with a WRAPPED A operand and optional fp32 output: A is [M, KA] and K-block kb of the product reads A block
 * kb % (KA / 64), W is [N, K].  Used for the split-precision 1x1 convolutions of the TCN path (A = [hi | lo], W = [Whi | Whi |
 * Wlo], K = 3 Kp, KA = 2 Kp): same arithmetic as the reference's fp32 Conv1d to ~22 mantissa bits
 * (causal_conv.py:71,80; sequence_model.py:79).  out_f32 != 0: C is fp32 [M, N]. */
int nppc_gemm_f16_tn_ex(const void* A, const void* W, const float* bias, void* C, long long M, int N, int K, int KA, int out_f32,
                        void* stream);

/* ---- N2: TCN stack in channel-last layout with its 1x1 convolutions on the tcgen05 GEMM ------------------------
 * (audio_zen/model/module/causal_conv.py:96-108, sequence_model.py:47-58,106-112).  M = B*T' rows, row = b*T' + t.
 * fp16 range: xh holds x / scale[b] (scale[b] = max|x| of the sample on entry, inv_scale = 1/scale); y1 and the final Linear
 * are multiplied back by scale[b] in fp32 where they are read.
 * split != 0 (nppc_tcn_cl_pack / nppc_tcn_out_cl): xh is [M, 2 Kp] = [hi | lo] fp16 halves of x / scale[b] (see nppc_gemm_f16_tn_ex).
 * nppc_tcn_cl_scale:   scale[b] = max(max|x[b]|, 1), inv_scale[b] = 1/scale[b] of a [B, n_per_sample] f32 tensor (two launches).
 * nppc_tcn_cl_pack:    x [B,C,T'] f32 -> x32 [M,Kp] f32 (residual stream, rows padded like xh; padding never read) and xh [M,Kp]
 *                      fp16 (GEMM operand; the K-padding columns [C,Kp) are written as zeros here and stay zero).
 * nppc_prelu_stats_cl: stats[b] = (sum, sum^2) of PReLU(y1 + bias[c]) over the sample, y1 [M,512] fp16.
 * nppc_tcn_mid_cl:     z [M,512] fp16 = PReLU2(depthwise_dilated(GroupNorm1(PReLU1(y1 + bias1)))), stats2 = moments of z.
 * nppc_tcn_out_cl:     x32 += o*rstd2 + vb - mean2*rstd2*u (o [M,Np] fp16 = z*(W2 diag(gamma2))^T); xh = fp16(x32), or
 *                      fp16(relu(x32)) when relu_h (the stack's trailing ReLU before fc_output_layer).
 * nppc_tcn_cl_unpack:  o [M,Np] fp16 (+bias, optional ReLU) -> [B,C,T'] f32. */
int nppc_tcn_cl_scale(const float* x, int B, long long n_per_sample, float* scale, float* inv_scale, void* stream);
int nppc_tcn_cl_pack(const float* x, int B, int C, int T, int Kp, const float* inv_scale, float* x32, void* xh, int split,
                     void* stream);
int nppc_tcn_cl_unpack(const void* o, int o_f32 /* o is fp32 instead of fp16 */, int B, int C, int T, int Np,
                       const float* scale /* [B] or NULL */, const float* bias, int relu, float* out, void* stream);
int nppc_prelu_stats_cl(const void* y1, int B, int T, int H, const float* scale, const float* bias, const float* prelu_a,
                        double* stats, void* stream);
int nppc_tcn_mid_cl(const void* y1, int B, int T, int H, const float* scale, const float* bias1, const float* prelu1_a,
                    const double* stats1, const float* gamma1, const float* beta1, const float* dw_w, const float* dw_b,
                    int dilation, const float* prelu2_a, void* z, double* stats2, void* stream);
int nppc_tcn_out_cl(const void* o, float* x32, int B, int T, int C, int Np, int Kp, int H, const double* stats2, const float* u,
                    const float* vb, const float* inv_scale, void* xh, int relu_h, int split, void* stream);


#ifdef __cplusplus
}
#endif
#endif /* NPPC_B200_H_ */
