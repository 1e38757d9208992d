/*
 * sag.h -- C ABI of libsag.so: the B200-native (sm_100a) inference hot path of spatialaudiogen.
 *
 * The reference (pedro-morgado/spatialaudiogen) is pure Python + TensorFlow 1.4 and has no FFI of its
 * own; the boundary it exposes for this path is `sess.run(ambi_pred_t, feed_dict)` (deploy.py:141,
 * eval.py:145) over the graph built by model.py:SptAudioGen.inference_ops (model.py:356-434) with
 * variables restored by name (deploy.py:79-87).  Each entry point below cites the reference interface it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every function returns 0 on success or a negative SAG_E* code; sag_last_error() (thread local)
 *    holds the message.  No aborts, no exit().
 *  - all tensor pointers are DEVICE pointers to fp32 unless a parameter is called `host_*`;
 *    layouts are the reference's (NHWC activations, HWIO conv weights, [kh,kw,Cout,Cin] transposed-conv
 *    weights, [in,out] FC weights; complex64 as interleaved float pairs).
 *  - the caller owns inputs, outputs and the workspace; the handle owns packed weights and descriptors.
 *    No allocation happens inside sag_forward: sag_workspace_bytes(h, batch), called after sag_finalize_weights, plans
 *    the batch (it packs the tensor-core weight images whose tile width depends on the row count -- device allocation and a
 *    synchronising pack kernel); sag_forward fails with SAG_ESTATE if that has not happened for its batch size / precision.
 *  - `stream` is a cudaStream_t passed as void*; all launches go to it, nothing synchronises.
 *  - a handle is not thread safe: one handle per (device, stream).
 */
#ifndef SAG_H_
#define SAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAG_OK 0
#define SAG_EINVAL (-1)   /* bad argument / shape */
#define SAG_ECUDA (-2)    /* CUDA runtime or driver error */
#define SAG_ENOMEM (-3)   /* workspace too small */
#define SAG_ESTATE (-4)   /* missing weight / wrong call order */
#define SAG_EUNSUPPORTED (-5)

/* arithmetic type of the dense contractions (convs / transposed convs / FCs) */
#define SAG_PREC_FP32 0   /* FFMA, fp32 in / fp32 accumulate (parity path) */
                          /* (1 is unassigned: a tf32 mode was never built -- bf16x3 is both faster and more accurate) */
#define SAG_PREC_BF16 2   /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate in TMEM */
#define SAG_PREC_BF16X3 3 /* tcgen05 kind::f16, operands split hi+lo (3 MMAs per K step): fp32-grade result */
#define SAG_PREC_MIXED 4  /* per-layer plan: bf16x3 everywhere except the U-Net decoder's deconv5..deconv2, which take ONE bf16
                             product (their rounding reaches the waveform through the sigmoid mask only: measured <= 2e-4 of
                             max|y| at B=32, inside the 1e-3 parity tolerance -- DESIGN.md section 5).  The ResNet towers, the
                             FCs and the audio encoder need the split operands: plain bf16 there costs 3e-2 / 5e-3 / 3e-4. */

#define SAG_SEP_NONE 0      /* definitions.py NO_SEPARATION */
#define SAG_SEP_UNET_MASK 1 /* definitions.py FREQ_MASK */

typedef struct sag_handle sag_handle;

/* Mirrors the constructor arguments of model.py:SptAudioGen.__init__ (model.py:24-32) and
 * SptAudioGenParams (model.py:10-21). */
typedef struct sag_config {
  int32_t ambi_order;       /* 1 */
  int32_t audio_rate;       /* 48000 */
  int32_t video_rate;       /* 10 */
  double context;           /* 1.0 s  (python floats are doubles: the integer crops of model.py:166-172 depend on it) */
  double sample_duration;   /* 0.1 s */
  int32_t enc_audio;        /* encoders list membership (definitions.py:1-4) */
  int32_t enc_video;
  int32_t enc_flow;
  int32_t separation;       /* SAG_SEP_* */
  int32_t sep_num_tracks;   /* 32 */
  int32_t n_loc_fc;         /* len(loc_fc_units), <= 4 */
  int32_t loc_fc_units[4];  /* [512,512] */
  double sep_fft_window;    /* 0.025 s -> wind_size 1024 (model.py:59-60) */
  int32_t precision;        /* SAG_PREC_* */
  int32_t frame_h, frame_w; /* 224, 448 */
} sag_config;

/* Derived constants of model.py:36-60 (snd_contx, snd_dur, snd_size, wind_size, num_ambi_channels) and the
 * integer crops of model.py:166-172, 313-323, 344-347 -- the "bit-exact frame indexing" contract. */
typedef struct sag_dims {
  int32_t snd_contx, snd_dur, snd_size, wind_size, num_ambi_channels;
  int32_t n_stft_frames;            /* 200 */
  int32_t enc_ss, enc_tt;           /* 46, 173 */
  int32_t mask_ss, mask_tt, mask_skip; /* 89, 117, 46 */
  int32_t final_crop;               /* 448 */
  int32_t feat_dim;                 /* 1024 | 1536 | 2048 */
} sag_dims;

const char* sag_last_error(void);
const char* sag_version(void);

/* ---- lifetime ------------------------------------------------------------------------------ */
int sag_config_default(sag_config* cfg);
/* CRC-32C (Castagnoli) of host bytes, continuing from `crc` (0 to start): the tensor checksums of the TensorFlow V2 bundles
 * that tf.train.Saver writes and deploy.py:79-87 restores (tf_checkpoint.read_bundle verifies them). Host code. */
uint32_t sag_crc32c(const void* host_data, size_t size, uint32_t crc);
int sag_create(sag_handle** out, const sag_config* cfg);         /* model.py:24-60 + deploy.py:42-77 */
int sag_destroy(sag_handle* h);
int sag_get_dims(const sag_handle* h, sag_dims* out);

/* ---- weights: tf.train.Saver().restore by variable name (deploy.py:79-87, eval.py:98-118) -------- */
/* `host_data` is fp32 in the TF layout of the variable `tf_name`. */
int sag_load_weight(sag_handle* h, const char* tf_name, const float* host_data, const int64_t* shape, int rank);
int sag_num_weights_expected(const sag_handle* h);
int sag_weight_name(const sag_handle* h, int i, char* buf, int buflen, int64_t* shape4, int* rank);
int sag_finalize_weights(sag_handle* h, void* stream);   /* repack into kernel layouts; idempotent */

/* ---- the forward: sess.run(ambi_pred_t, feed_dict) (deploy.py:141, eval.py:145) ------------------ */
/* Bytes of workspace a forward of `batch` windows needs (0 on error) -- and, once the weights are finalized, the planning
 * step of that batch size (see "Conventions"): call it again after changing the precision or reloading weights. */
size_t sag_workspace_bytes(const sag_handle* h, int batch);
/* audio (B,snd_size,1); video / flow (B,1,H,W,3) or NULL when the encoder is absent;
 * ambix_out (B,snd_dur,3) = channels (Y,Z,X) (model.py:424-432). */
int sag_forward(sag_handle* h, const float* audio, const float* video, const float* flow, float* ambix_out,
                void* workspace, size_t workspace_bytes, int batch, void* stream);
/* The same forward fed with the frames as the reference's readers decode them from disk -- uint8 (B,1,H,W,3) -- instead of the
 * float32 tensors its feeder prepares on the host: the preparation runs inside the frame-ingest kernel (4x fewer bytes over
 * PCIe / HBM per frame).  SAG_FRAMES_U8 video: myutils.img_prep_fcn, x/255 - 0.5 (myutils.py:88-89), bit-identical to
 * feeding the prepared float32 frame.  SAG_FRAMES_U8 flow: FlowReader.get_by_index (feeder.py:147-161): channel 0 = angle,
 * channel 2 = magnitude quantised to the frame's (min, max) = flow_limits[b] (device doubles, (B,2), the rows of
 * flow_limits.npy) -> (mag cos, mag sin, mag).  SAG_FRAMES_F32 = what sag_forward takes. */
#define SAG_FRAMES_F32 0
#define SAG_FRAMES_U8 1
int sag_forward_frames(sag_handle* h, const float* audio, const void* video, int video_format, const void* flow, int flow_format,
                       const double* flow_limits, float* ambix_out, void* workspace, size_t workspace_bytes, int batch, void* stream);
/* Intermediate tensors of the last sag_forward (model.py `self.ends`, `sep_channels`, `loc_channels`).
 * The pointer aliases the workspace and is valid until the next forward.  shape has up to 5 entries. */
int sag_get_tensor(const sag_handle* h, const char* name, const float** dev_ptr, int64_t* shape5, int* rank,
                   int64_t* row_stride /* elements between consecutive indices of the second-to-last axis */);
/* Storage format of an intermediate tensor: SAG_FMT_F32, or SAG_FMT_BF16_SPLIT on the tensor-core path, where a
 * value x is kept as two bf16 planes hi = bf16(x) at dev_ptr and lo = bf16(x - hi) `plane_bytes` further
 * (plane_bytes == 0: hi only); shape and row_stride are in elements either way. */
#define SAG_FMT_F32 0
#define SAG_FMT_BF16_SPLIT 1
int sag_get_tensor_format(const sag_handle* h, const char* name, int* format, int64_t* plane_bytes);
int sag_num_tensors(const sag_handle* h);
int sag_tensor_name(const sag_handle* h, int i, char* buf, int buflen);
/* options: "skip_unused" (default 1: STFT frames / mask rows that cannot reach the cropped output are not
 * computed; the result is bit-identical), "precision" (SAG_PREC_*), "profile" (0/1, see sag_get_profile). */
int sag_set_option(sag_handle* h, const char* key, int value);
/* Tile plan of the tcgen05 contraction kernel for a [M x K] x [K x N] product (conv: M = B*OH*OW, K = kh*kw*Cin, N = Cout):
 * tile width (32 | 64 | 128 | 256 GEMM columns per CTA tile) and K split.  Pure host arithmetic; lets tests / profiles name
 * the kernel instantiation a layer runs at a given batch size. */
int sag_plan_contraction(int k, int n, int64_t m, int* tile_width, int* k_split);
/* Stream-K (layers whose tiles fill the last wave of the persistent grid badly: conv4_x / conv5_x at 32 windows): the
 * (tile, K chunk) units are dealt out evenly over the CTA clusters and the pieces of a tile meet in the epilogue of the CTA that
 * finishes it.  sag_plan_stream_k: 1 if the planner picks it for this shape on the forward's main stream.
 * sag_stream_k_schedule: the pieces of cluster `cluster` of `clusters` over `tiles` tiles of `k_chunks` chunks, in execution
 * order -- items[j] = {tile, first chunk, end chunk, leaves a partial (1) or finishes the tile (0), first cluster holding an
 * earlier piece of the tile (-1: none)}; returns the number of pieces (the kernel runs this same code).  Host arithmetic. */
int sag_plan_stream_k(int k, int n, int64_t m);
int sag_stream_k_schedule(int64_t tiles, int k_chunks, int clusters, int cluster, int* items, int max_items);
/* how many kernels the last sag_forward launched (bench.py gpu_launches) */
int sag_last_launch_count(const sag_handle* h);
/* With option "profile" = 1 every launch group of sag_forward is bracketed by CUDA events on the caller's stream.
 * Sums over the last forward for one category: 0 conv, 1 transposed conv, 2 fully connected (the three share the
 * contraction kernel), 3 STFT, 4 inverse STFT, 5 BN / pooling passes, 6 mixing.  flops / bytes are the algorithmic
 * figures of DESIGN.md (executed work only).  Synchronises on the recorded events. */
#define SAG_PROF_CONV 0
#define SAG_PROF_DECONV 1
#define SAG_PROF_FC 2
#define SAG_PROF_STFT 3
#define SAG_PROF_ISTFT 4
#define SAG_PROF_POINTWISE 5
#define SAG_PROF_MIX 6
#define SAG_PROF_NCAT 7
int sag_get_profile(sag_handle* h, int category, double* ms, double* flops, double* bytes, int* launches);
/* The individual records of the last profiled forward, in launch order: layer name (reference scope), category, CUDA-event
 * time in microseconds, useful flops (each product of the reference graph that can reach the output, once), issued flops
 * (what the kernel multiplies: zero taps / border cells of the sub-pixel transposed convs, conv1's K padded 147 -> 256),
 * algorithmic bytes, tile width and K split of the contraction kernel (0 for other kernels). */
int sag_num_profile_records(const sag_handle* h);
int sag_get_profile_record(sag_handle* h, int i, char* name, int name_len, int* category, double* us, double* flops,
                           double* flops_issued, double* bytes, int* tile_width, int* k_split);

/* ---- stage entry points (tests / ncu); each replaces the named reference op -------------------- */
/* myutils.stft (myutils.py:119-147): x (rows,n_samples) -> out complex (rows, n_frames_out, wind) for frames
 * [frame0, frame0+n_frames_out) of the n_overlap*n_winds total; mag_out (optional) gets |.| for frames
 * [mag0, mag0+n_mag).  wind must be a product of 2,3,4,5 radices and <= 4800. */
int sag_stft(const float* x, int rows, int n_samples, int wind, int n_overlap, int frame0, int n_frames_out,
             float* cplx_out, int mag0, int n_mag, float* mag_out, void* stream);
/* myutils.istft (myutils.py:181-211): in complex (rows,n_frames,wind) -> out (rows, (n_frames/n_overlap)*wind - (n_overlap-1)*wind/n_overlap) */
int sag_istft(const float* cplx_in, int rows, int n_frames, int wind, int n_overlap, float* out, void* stream);
/* tfw.conv_2d (core.py:156-220): x NHWC, w HWIO, padding 0=VALID 1=SAME(TF asymmetric), optional bias, relu */
int sag_conv2d(const float* x, int n, int h, int w, int cin, const float* w_hwio, int kh, int kw, int cout,
               int sh, int sw, int same_pad, const float* bias, int relu, float* y, int precision, void* stream);
/* tfw.deconv_2d VALID (core.py:96-153): w [kh,kw,Cout,Cin]; y (n,(h-1)*sh+kh,(w-1)*sw+kw,cout) */
int sag_deconv2d(const float* x, int n, int h, int w, int cin, const float* w_hwoi, int kh, int kw, int cout,
                 int sh, int sw, const float* bias, int relu, float* y, int precision, void* stream);
/* tfw.fully_connected (core.py:43-93): x (rows,in) @ w[in,out] + b, optional relu */
int sag_fc(const float* x, int rows, int in, const float* w, int out, const float* bias, int relu, float* y,
           int precision, void* stream);
/* contrib batch_norm(is_training=True) forward (core.py:209-210) + optional residual add + optional relu:
 * y = act(gamma*(x-mu_B)/sqrt(var_B+1e-3)+beta [+ residual]), statistics over (n,h,w). scratch >= 4*c doubles. */
int sag_batchnorm_train(const float* x, int64_t rows, int c, const float* gamma, const float* beta,
                        const float* residual, int relu, float* y, void* scratch, void* stream);
/* tf.nn.max_pool 3x3/2 SAME (resnet.py:135) */
int sag_maxpool_3x3s2_same(const float* x, int n, int h, int w, int c, float* y, void* stream);
/* ResNet18.inference_ops(truncate_at='conv5_2') with batch-statistics BN (resnet.py:123-190; model.py:189-201).
 * scope is 'video_encoder' or 'flow_encoder'; x (B,H,W,3) -> y (B,H/32,W/32,512). Uses the handle's weights. */
int sag_resnet18(sag_handle* h, const char* scope, const float* x, int batch, float* y, void* workspace,
                 size_t workspace_bytes, void* stream);
/* decode step of inference_ops (model.py:424-432): x_sep (B,K,T), loc (B,S,3*(K+1)) -> out (B,T,3) */
int sag_mix(const float* x_sep, const float* loc, int batch, int tracks, int t, int segments, float* out, void* stream);

/* ---- evaluation metrics (model.py:110-154, myutils.py:109-116, eval.py:147-198) ----------------- */
/* pred, gt (B,T,3).  Outputs (B,3) each: stft_ps, lsd_ps, mse_ps, snr_ps, env_ps; amp (B,2) = max|pred|, max|gt|.
 * scratch: sag_metrics_scratch_bytes(B,T). */
size_t sag_metrics_scratch_bytes(int batch, int t);
int sag_metrics(const float* pred, const float* gt, int batch, int t, int audio_rate, float* stft_ps, float* lsd_ps,
                float* mse_ps, float* snr_ps, float* env_ps, float* amp, void* scratch, void* stream);
/* AmbiDecoder.decode + RMS map (decoder.py:24-28, distance.py:41-52): ambi (B,T,4) [W,Y,Z,X] (already masked),
 * mesh of `ang_res` degrees -> rms (B, n_nu, n_phi), rows flipped like np.flipud. */
int sag_sh_rms_dims(float ang_res, int* n_nu, int* n_phi);
int sag_sh_rms(const float* ambi, int batch, int t, float ang_res, float* rms, void* stream);

/* Mel log-spectral distance (myutils.compute_lsd_dist, myutils.py:96-106: librosa melspectrogram n_fft 2048, hop 512,
 * 128 Slaney mels up to 12 kHz, power 2; 10*log10(|.| + 0.01); RMS over bands x frames): pred, gt (B,T,3) -> (B,3). */
int sag_mel_lsd(const float* pred, const float* gt, int batch, int t, int audio_rate, float* mel_lsd_ps, void* stream);
/* Earth mover's distance (EMD-hat, pyemd.emd semantics) of `count` pairs of n-bin histograms over one ground-distance
 * matrix: the last two eval-detailed.txt columns (eval.py:190-193 -> distance.py:100-143 ambix_emd / emd).  Host code
 * (the reference solves it on the host too); exact min-cost flow in doubles; extra_mass_penalty < 0 = max(dist). */
int sag_emd_hat(const double* first, const double* second, int n, const double* dist, double extra_mass_penalty, int count,
                double* out);

/* ---- JPEG frame decode ------------------------------------------------------------------------------
 * The reference's feeder reads every video / flow frame with scipy.misc.imread (feeder.py:120-127): PIL over libjpeg with
 * its default settings (JDCT_ISLOW inverse DCT, fancy chroma upsampling).  These entry points replace that call for a batch
 * of frames: the entropy-coded segments are decoded on the host (a pool of `threads` threads, one file per task) into pinned
 * staging, and dequantisation + inverse DCT + upsampling + YCbCr->RGB run as CUDA kernels that write the uint8
 * (n, height, width, 3) frames sag_forward_frames ingests.  Bit-identical to PIL's decode.  Baseline sequential files
 * (SOF0), 8 bit, 1 or 3 components in one interleaved scan, chroma sampled 1x1 / 2x1 / 2x2, restart intervals; anything else
 * fails with SAG_EUNSUPPORTED. */
typedef struct sag_jpeg sag_jpeg;
/* frame geometry of a file in host memory: h_samp / v_samp = the largest sampling factors (2,2 for 4:2:0) */
int sag_jpeg_info(const void* host_file, size_t size, int* width, int* height, int* components, int* h_samp, int* v_samp);
/* host only: the quantised coefficients of a file, component after component, each a (blocks_high, blocks_wide, 64) int16
 * array in natural (row-major) order on a block grid of whole MCUs, plus the components' quantisers (3 x 64, natural order) */
int sag_jpeg_coefficients(const void* host_file, size_t size, int16_t* host_coef, size_t capacity, int* blocks_wide, int* blocks_high,
                          uint16_t* host_qt);
/* a decoder for batches of at most max_frames frames of height x width pixels on the current device (owns pinned staging
 * and device scratch; not thread safe) */
int sag_jpeg_create(sag_jpeg** dec, int max_frames, int height, int width);
void sag_jpeg_destroy(sag_jpeg* dec);
/* n files in host memory -> frames (device, uint8, (n, height, width, 3) RGB) on `stream`.  Returns once the kernels are
 * queued; the files may be released on return.  threads <= 0: one per host core (at most 32). */
int sag_jpeg_decode(sag_jpeg* dec, const void* const* host_files, const size_t* sizes, int n, uint8_t* frames, int threads, void* stream);
/* "device_huffman" (default 1): the entropy-coded segments are decoded on the GPU too -- each frame's bit stream is cut into
 * subsequences decoded in parallel, whose decoders fall into step with the true one by themselves (a few rounds until every
 * subsequence starts where its predecessor ended) -- so only the compressed bytes cross PCIe; 0: host entropy decoding. */
int sag_jpeg_set_option(sag_jpeg* dec, const char* key, int value);
/* synchronisation rounds each of the first n frames of the last device-side decode took (synchronises with the device) */
int sag_jpeg_sync_rounds(sag_jpeg* dec, int* host_rounds, int n);
/* host only, for tests: the device's parallel entropy decoder run thread by thread on the host (`nthreads` = its CTA size);
 * same output as sag_jpeg_coefficients */
int sag_jpeg_coefficients_parallel(const void* host_file, size_t size, int nthreads, int16_t* host_coef, size_t capacity, int* rounds);

#ifdef __cplusplus
}
#endif
#endif /* SAG_H_ */
