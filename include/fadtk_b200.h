/* fadtk_b200 - C ABI of the B200-native Frechet-Audio-Distance hot path.
 *
 * The reference (microsoft/fadtk) has no native layer: its hot path is Python calling
 * third-party PyTorch models and numpy/scipy.  This header is the boundary a maintainer
 * binds instead (ctypes stub in INTEGRATION.md); every entry point names the reference code
 * it replaces.  Conventions:
 *   - extern "C", plain C types, no C++ exceptions cross the boundary;
 *   - every function returns 0 on success, non-zero on failure with a message available from
 *     fad_last_error() (thread-local);
 *   - all data buffers are CALLER-OWNED DEVICE pointers (e.g. torch.Tensor.data_ptr()) unless
 *     the parameter name ends in _host; `stream` is a cudaStream_t passed as void*;
 *   - a fad_handle belongs to one device and must not be used from two threads at once;
 *     distinct handles are independent;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef FADTK_B200_H
#define FADTK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fad_handle fad_handle;

/* ---- library ------------------------------------------------------------------------ */
int         fad_version(void);
const char* fad_last_error(void);

/* One handle per (process, device).  max_examples bounds the number of 0.96-s VGGish examples
 * processed per internal batch (workspace ~0.7 MB per example). */
int  fad_create(int device, int max_examples, fad_handle** out);
int  fad_destroy(fad_handle* h);

/* ---- VGGish embedder: replaces VGGishModel.load_model/_get_embedding ------------------
 * (fadtk/model_loader.py:89-108 -> torchvggish front-end + VGG stack) and the float32 ->
 * float16 conversion of ModelLoader.get_embedding (fadtk/model_loader.py:40-50). */
typedef struct {
    const float*    conv1_w_host;   /* [64, 9]  fp32                                      */
    const float*    conv1_b_host;   /* [64]                                               */
    const uint16_t* conv_w_host[5]; /* conv2..conv6: fp16 [Cout, 9*Cin], k=(kh*3+kw)*Cin+c */
    const float*    conv_b_host[5];
    const uint16_t* fc_w_host[3];   /* fc1..fc3: fp16 [out, in]                           */
    const float*    fc_b_host[3];
    /* bit i set (i = 0..7: conv2, conv3_1, conv3_2, conv4_1, conv4_2, fc1, fc2, fc3): that layer's
     * weights are an fp16 hi/lo pair W = Wh + Wl stored as [2*Cout, K] with the 128 hi rows of
     * each 128-channel tile followed by its 128 lo rows.  fp16-only weights are a fixed model
     * perturbation worth ~1.4e-4 relative on FAD, the split removes it. */
    uint32_t        split_mask;
} fad_vggish_weights;

int fad_vggish_load(fad_handle* h, const fad_vggish_weights* w);

/* Examples (rows of the embedding) produced by a clip of n_samples at 16 kHz:
 * 1 + floor((T - 96) / 96) with T = 1 + floor((n_samples - 400) / 160); 0 if too short. */
long long fad_vggish_num_examples(long long n_samples);

/* Host-side planning: clip_offsets_host[n_clips + 1] (sample offsets into one flat PCM
 * buffer) -> start sample of every example.  Returns the number of examples; writes at most
 * `capacity` entries to ex_start_host (pass NULL/0 to only count).  rows_per_clip_host (may be
 * NULL) receives the per-clip example counts. */
long long fad_vggish_plan(const long long* clip_offsets_host, long long n_clips,
                          long long* ex_start_host, long long capacity,
                          long long* rows_per_clip_host);

/* pcm: int16 mono 16 kHz (device).  ex_start: int64 [n_examples] (device).
 * emb_out: fp16 [n_examples, 128] (device) - exactly what the reference caches as .npy. */
int fad_vggish_forward(fad_handle* h, const int16_t* pcm, const long long* ex_start,
                       long long n_examples, void* emb_out_f16, void* stream);

/* Stage-level entry points (used by the parity tests and profiling). */
int fad_vggish_logmel(fad_handle* h, const int16_t* pcm, const long long* ex_start,
                      long long n_examples, float* logmel_out /* [n,96,64] */, int use_double,
                      void* stream);
/* conv1 of the VGG stack alone (3x3, 1 -> 64, pad 1, + bias, ReLU, 2x2 max-pool; CUDA-core fp32 stencil) with the
 * loaded weights: logmel fp32 [n, 96, 64] -> fp16 NHWC [n, 48, 32, 64]. */
int fad_vggish_conv1(fad_handle* h, const float* logmel, long long n_examples, void* out_f16, void* stream);
/* One tensor-core layer: 3x3 conv pad 1 (taps = 9) or fully connected (taps = 1, H = W = 1) on
 * NHWC fp16 input x[NB,H,W,Cin] with fp16 weights w[Cout, taps*Cin]; fused bias, optional ReLU,
 * optional 2x2 max-pool; fp16 NHWC output (and optional fp32 copy of the un-pooled output). */
int fad_umma_layer(fad_handle* h, const void* x_f16, int NB, int H, int W, int Cin,
                   const void* w_f16, const float* bias, int Cout, int taps, int relu, int pool,
                   int split_w /* weights are [2*Cout, K] hi/lo tiles */,
                   void* out_f16, float* out_f32_or_null, void* stream);

/* ---- CLAP-LAION audio embedder (HTSAT-tiny): replaces CLAPLaionModel.load_model/_get_embedding
 * (fadtk/model_loader.py:382-418 -> laion_clap.CLAP_Module + torchlibrosa front-end).
 * tensors_host: 180 (HTSAT-tiny = clap-laion-audio) or 258 (HTSAT-base = clap-laion-music,
 * model_loader.py:385) host pointers in the order documented at the top of csrc/clap_host.inc; the count
 * selects the variant
 * (produced by fadtk_b200/weights_clap.py).  max_chunks bounds the 10-s windows per internal batch
 * (~12 MB of workspace each). */
int fad_clap_load(fad_handle* h, const void* const* tensors_host, int n_tensors, int max_chunks);

/* Windows of a clip: one every 48 000 samples (range(0, T, sr), model_loader.py:396-398), each
 * covering up to 480 000 samples and zero padded.  Returns the number of windows; fills
 * chunk_start_host (sample offset into the flat PCM buffer) and chunk_valid_host (samples available). */
long long fad_clap_plan(const long long* clip_offsets_host, long long n_clips, long long* chunk_start_host,
                        int* chunk_valid_host, long long capacity, long long* rows_per_clip_host);

/* Frame pool: windows of one clip are 1-s shifts of the same audio, so every STFT frame that does
 * not touch a window edge is shared by up to ten windows.  fad_clap_plan_frames lists each DISTINCT
 * frame once - pool_start (sample offset of the window it is taken from), pool_valid, pool_frame
 * (frame index inside that window) - and fills frame_index [n_chunks][1001]: pool row of every
 * (window, frame).  Returns the pool size (pass NULL outputs to only count). */
long long fad_clap_plan_frames(const long long* clip_offsets_host, long long n_clips, long long* pool_start_host,
                               int* pool_valid_host, int* pool_frame_host, long long pool_capacity,
                               int* frame_index_host);

/* pcm: int16 mono 48 kHz; pool_* [n_pool] and frame_index [n_chunks*1001] from fad_clap_plan_frames (all
 * device).  emb_out: fp16 [n_chunks, 512], L2-normalised - what the reference caches as .npy. */
int fad_clap_forward(fad_handle* h, const int16_t* pcm, const long long* pool_start, const int* pool_valid,
                     const int* pool_frame, long long n_pool, const int* frame_index, long long n_chunks,
                     void* emb_out_f16, void* stream);
/* stage entry point: BatchNorm-ed log-mel rows [n_pool, 64] fp32 */
int fad_clap_logmel(fad_handle* h, const int16_t* pcm, const long long* pool_start, const int* pool_valid,
                    const int* pool_frame, long long n_pool, float* out, void* stream);

/* ---- Whisper: replaces WhisperModel.load_model / _get_embedding (fadtk/model_loader.py:657-669):
 * WhisperFeatureExtractor (clip padded / truncated to 30 s, log-mel 80 x 3000) and
 * transformers.WhisperModel(input_features, decoder_input_ids = [[sot, sot]]).last_hidden_state.
 * cfg: {d_model, heads (= d_model / 64), encoder layers, decoder layers, ffn dim}; tensors_host: host pointers in
 * the order documented at the top of csrc/whisper_host.inc (5 + 12 L_enc + 3 + 20 L_dec + 2), packed by
 * fadtk_b200/weights_whisper.py.  max_clips bounds the clips per launch sequence (45 MB of workspace each at
 * d_model = 768). */
int fad_whisper_load(fad_handle* h, const int* cfg, const void* const* tensors_host, int n_tensors, int max_clips);
/* pcm: int16 mono 16 kHz; clip_start int64 / clip_len int32 [n_clips] (all device).
 * emb_out: fp16 [n_clips][2][d_model]. */
int fad_whisper_forward(fad_handle* h, const int16_t* pcm, const long long* clip_start, const int* clip_len,
                        long long n_clips, void* emb_out_f16, void* stream);
/* stage entry point: out = fp32 [n_clips*3000*80] log10 mel (time-major) followed by [n_clips] per-clip maxima;
 * the features are (max(x, max - 8) + 4) / 4. */
int fad_whisper_logmel(fad_handle* h, const int16_t* pcm, const long long* clip_start, const int* clip_len,
                       long long n_clips, float* out, void* stream);

/* ---- Encodec: replaces EncodecEmbModel.load_model / _get_frame for the 24 kHz variant
 * (fadtk/model_loader.py:123-130, 155-166): EncodecModel.encodec_model_24khz().encoder(audio) -> [T/320, 128].
 * tensors_host: 78 host pointers in the order documented at the top of csrc/encodec_host.inc, packed by
 * fadtk_b200/weights_encodec.py (weight-norm folded, im2col column order, fp16 hi/lo tiles).
 * variant 0: encodec_model_24khz (causal, mono, whole file); 1: encodec_model_48khz (non-causal, GroupNorm(1, C)
 * after every conv, the mono file duplicated to stereo; the caller passes the 1-s segments of :139-152 as clips).
 * max_chunk_samples bounds clips x samples per convolution chunk (0.55 KB of workspace per sample). */
int fad_encodec_load(fad_handle* h, const void* const* tensors_host, int n_tensors, long long max_chunk_samples, int variant);
/* pcm: int16 mono 24 kHz [n_clips][T] (device), all clips of one call have the same length T.
 * emb_out: fp16 [n_clips][ceil(T/320)][128] (device). */
int fad_encodec_forward(fad_handle* h, const int16_t* pcm, long long n_clips, int T, void* emb_out_f16, void* stream);

/* ---- wav2vec 2.0 / HuBERT / MERT: replaces W2V2Model, HuBERTModel, MERTModel load_model / _get_embedding
 * (fadtk/model_loader.py:254-288, 525-596) for the "group-norm feature encoder + post-LN transformer" checkpoints
 * (wav2vec2-base-960h, hubert-base-ls960, MERT-v1-95M): processor normalisation, Wav2Vec2Model / HubertModel
 * forward with output_hidden_states, hidden_states[layer].
 * cfg: {d_model, heads (= d_model / 64), layers, ffn}; tensors_host in the order of csrc/wav2vec_host.inc
 * (39 + 12 layers), packed by fadtk_b200/weights_w2v.py. */
int fad_w2v_load(fad_handle* h, const int* cfg, const void* const* tensors_host, int n_tensors, int max_clips, int max_len);
/* pcm: int16 mono [n_clips][L] (device), equal lengths; emb_out: fp16 [n_clips][frames(L)][d_model]. */
int fad_w2v_forward(fad_handle* h, const int16_t* pcm, long long n_clips, int L, int layer, void* emb_out_f16, void* stream);

/* ---- statistics: replaces calc_embd_statistics (fadtk/fad.py:42-48) and
 * _process_file / calculate_embd_statistics_online (fadtk/utils.py:13-46) ----------------
 * Packed fp64 accumulator of length fad_stats_acc_len(d):
 *   acc[0] = n, acc[1..d] = sum(x - shift) (exact), acc[1+d..1+d+d*d) = sum y y^T (d x d),
 *   acc[1+d+d*d..] = sum y,   y = x - shift (exact modes) or its fp16 hi/lo pair (mode 1)
 * It is additive: accumulate batches into it, all-reduce (sum) it across GPUs, then finalize.
 * `shift` (fp16 [d], device) must be identical for every contribution to one accumulator. */
size_t fad_stats_acc_len(int d);
/* tensor_core = 0 (default everywhere in the product): E^T E on the FP64 TENSOR pipe (mma.sync
 * m8n8k4 f64 -> DMMA): the products (x - s)(x - s)^T of fp16 data are exact in fp64 and the
 * accumulation is fp64 in a fixed order, so the result is the Gram matrix of the data to ~1e-16,
 * positive semi-definite and bit-reproducible (rank-deficient per-song sets and covariances with
 * cond ~1e9 need that).
 * tensor_core = 1: tcgen05 kind::f16 hi/lo-split E^T E (fp32 accumulation in TMEM, ~1e-6 relative)
 * for well-conditioned, full-rank sets.
 * tensor_core = 2: the same exact arithmetic on the CUDA cores (DFMA + fp64 atomics): verification. */
int fad_stats_accumulate(fad_handle* h, const void* emb_f16, long long n_rows, int d,
                         const void* shift_f16, double* acc, int tensor_core, void* stream);
/* ---- multi-GPU: the ONE exchange step of the path ----------
 * Ranks embed disjoint shards of the clips and accumulate with the SAME shift vector; the packed accumulators are
 * then summed over NVLink and every rank finalises identical statistics.  The reference has no counterpart (it is
 * single-device); this replaces the pickled per-file scatter matrices of its process map (fadtk/utils.py:35-45).
 * NCCL is loaded at run time (dlopen of libnccl.so.2, or $FADTK_NCCL_LIB); nothing is linked.
 *   fad_comm_unique_id   rank 0 creates the 128-byte rendezvous id (ncclGetUniqueId); the host ships it to the
 *                        other ranks however it likes (file, MPI, torch.distributed, a socket)
 *   fad_comm_init        every rank joins; the communicator belongs to the handle
 *   fad_stats_allreduce  in-place sum of acc[fad_stats_acc_len(d)] on `stream`; nccl_comm_or_null = an existing
 *                        ncclComm_t of the host application, or NULL for the handle's own communicator
 *   fad_allreduce_sum_f64  the same for any fp64 device buffer (e.g. both datasets packed into one call) */
#define FAD_COMM_ID_BYTES 128
int fad_comm_unique_id(void* id_out_host);
int fad_comm_init(fad_handle* h, const void* id_host, int rank, int world);
int fad_comm_destroy(fad_handle* h);
int fad_stats_allreduce(fad_handle* h, void* nccl_comm_or_null, double* acc, int d, void* stream);
int fad_allreduce_sum_f64(fad_handle* h, void* nccl_comm_or_null, double* buf, long long n_values, void* stream);

/* The reference's DIRECTORY statistics (per-file np.mean rounded to fp16, per-file scatter, Chan merge:
 * fadtk/utils.py:13-46) for n_files files of rows_per_file rows each, without leaving the device:
 *   fad_file_means               m64[f], m16[f] (fp64 rows [n_files, d]): the exact mean of file f and its mean as the
 *                                reference's _process_file returns it (fp32 accumulation rounded to fp16)
 *   fad_stats_accumulate_f64     exact Gram statistics (DMMA) of fp64 rows, unshifted, into a packed accumulator
 *   fad_stats_finalize_mirrored  (mu, cov) exactly as calculate_embd_statistics_online returns them, from the packed
 *                                accumulators of the rows, of m64 and of m16 (all three additive: all-reduce them first);
 *                                rows_per_file == 1 gives the reference's all-NaN covariance (utils.py:16) */
int fad_file_means(fad_handle* h, const void* emb_f16, long long n_files, int rows_per_file, int d,
                   double* m64_out, double* m16_out, void* stream);
int fad_stats_accumulate_f64(fad_handle* h, const double* rows, long long n_rows, int d, double* acc, void* stream);
int fad_stats_finalize_mirrored(fad_handle* h, const double* acc, const double* acc_means64, const double* acc_means16,
                                const void* shift_f16, int rows_per_file, int d, double* mu_out, double* cov_out, void* stream);
/* rows emb[idx[i]] for i < n_idx (FAD-inf bootstrap, fadtk/fad.py:333-336) */
int fad_stats_accumulate_gather(fad_handle* h, const void* emb_f16, long long n_src_rows,
                                const long long* idx, long long n_idx, int d,
                                const void* shift_f16, double* acc, void* stream);
int fad_stats_finalize(fad_handle* h, const double* acc, const void* shift_f16, int d,
                       double* mu_out, double* cov_out, void* stream);

/* ---- Frechet distance: replaces calc_frechet_distance (fadtk/fad.py:51-120) ------------
 * mu/cov fp64 device arrays.  out (device, 8 doubles): [0] FAD, [1] tr sqrt(C1 C2),
 * [2] relative residual of the final square root, [3] iterations, [4] |mu1-mu2|^2,
 * [5] tr C1, [6] tr C2, [7] reserved.  iters <= 0 selects the default. */
int fad_frechet(fad_handle* h, const double* mu1, const double* cov1, const double* mu2,
                const double* cov2, int d, int iters, double* out, void* stream);

/* The baseline's square root can be computed once and reused (FAD-inf, per-song scoring):
 * fad_sqrt_psd -> sqrt_out (d*d doubles) and scal_out (2 doubles: |C|_F, tr C), both device. */
int fad_sqrt_psd(fad_handle* h, const double* cov, int d, int iters, double* sqrt_out, double* scal_out,
                 void* stream);
int fad_frechet_presqrt(fad_handle* h, const double* mu1, const double* sqrt1, const double* scal1,
                        const double* mu2, const double* cov2, int d, int iters, double* out, void* stream);

/* Ragged-batched form for per-file scoring: replaces the loop of FrechetAudioDistance.score_individual
 * (fadtk/fad.py:353-395; per file calc_embd_statistics fad.py:42-48 + calc_frechet_distance :51-120).
 * emb_f16: fp16 [N, d] (device); offsets: int64 [n_items + 1] (device), item z = rows
 * [offsets[z], offsets[z+1]).  Per item the mean is rounded to fp16 (np.mean dtype rule) and the
 * covariance is the exact fp64 ddof=1 Gram.  out: fp64 [n_items][8] in fad_frechet's layout with
 * [7] = row count; an item with fewer than two rows gets NaN in [0], [1] (the reference asserts). */
int fad_frechet_batched(fad_handle* h, const double* mu1, const double* sqrt1, const double* scal1,
                        const void* emb_f16, const long long* offsets, long long n_items, int d, int iters,
                        double* out, void* stream);

/* ---- audio conversion: replaces the torchaudio branch of FrechetAudioDistance.load_audio
 * (fadtk/fad.py:147-160): mono mix (:150), Resample(lowpass_filter_width=64, rolloff=0.9475937167399596,
 * sinc_interp_kaiser, beta=14.769656459379492) (:151-158), PCM16 quantisation (:160).
 * fad_resample_geometry / _length / _bank are host-only (no GPU): reduced rates orig/new, filter
 * half-width, taps = 2*width + orig; output length ceil(new*length/orig); the [new][taps] float32
 * filter bank exactly as torchaudio builds it.
 * fad_resample: exactly one of in_i16 (interleaved [length][channels], scaled by 1/32768) and in_f32
 * (planar [channels][length]) is given (device); out_pcm int16 [fad_resample_length] (device);
 * out_f32 optional un-quantised copy (device) or NULL. */
int fad_resample_geometry(int sr_in, int sr_out, int* orig, int* new_, int* width, int* taps);
long long fad_resample_length(int sr_in, int sr_out, long long length);
int fad_resample_bank(int sr_in, int sr_out, float* bank_host);
int fad_resample(fad_handle* h, const int16_t* in_i16, const float* in_f32, int channels, long long length,
                 int sr_in, int sr_out, int16_t* out_pcm, float* out_f32, void* stream);

/* ---- measurement ----------------------------------------------------------------------
 * When enabled, CUDA events are recorded on the launching stream around every kernel group;
 * fad_profile_collect synchronises the device and returns accumulated milliseconds and launch
 * counts per category (arrays of FAD_PROF_CATEGORIES entries). */
#define FAD_PROF_LOGMEL        0
#define FAD_PROF_CONV1         1
#define FAD_PROF_LAYER0        2   /* +i: conv2, conv3_1, conv3_2, conv4_1, conv4_2, fc1, fc2, fc3 */
#define FAD_PROF_STATS        10
#define FAD_PROF_STATS_REDUCE 11
#define FAD_PROF_FRECHET      12
#define FAD_PROF_CLAP_FRONT   13   /* log-mel + patch embedding */
#define FAD_PROF_CLAP_GEMM    14   /* tcgen05 GEMMs of the Swin blocks */
#define FAD_PROF_CLAP_ATTN    15   /* window attention */
#define FAD_PROF_CLAP_OTHER   16   /* LayerNorm, residual adds, head */
#define FAD_PROF_CATEGORIES   20
int fad_profile_enable(fad_handle* h, int on);
int fad_profile_collect(fad_handle* h, double* ms_out, long long* count_out, int reset);

/* Number of CUDA kernels this library has launched through `h` (bench.py's gpu_launches). */
long long fad_launch_count(fad_handle* h);

/* Stage entry (parity test / profiling): the encoder self-attention of the Whisper and wav2vec-family forwards alone.
 * qkv: fp16 [n_clips * S][3 d] (q | k | v, head i at columns i * 64), out: fp16 [n_clips * S][d], softmax(q k^T / 8) v per
 * head.  legacy = 0: tcgen05 kernel (csrc/attention_umma.cuh); 1: the mma.sync flash kernel it replaced. */
int fad_attention(fad_handle* h, const void* qkv_f16, long long n_clips, int S, int d, void* out_f16, int legacy, void* stream);

/* ---- measurement utility -------------------------------------------------------------
 * fp64 tensor-pipe (DMMA m8n8k4) rate of this GPU in TFLOP/s, measured with a register-only
 * issue loop: the roofline denominator of the exact-Gram and Newton-Schulz kernels, which
 * MEASURED_PEAKS.json (bf16 GEMM, HBM copy) does not carry.  Synchronous; iters <= 0 = default. */
int fad_bench_dmma_peak(fad_handle* h, int iters, double* tflops_out_host);
/* Milliseconds that `ksteps` K steps (128 x 128 x 64, hi/lo split weights) take on every SM at once under tcgen05 issue
 * pattern `mode` (0..5, csrc/umma_bench.cuh): operands resident in shared memory, no TMA, no epilogue - what the
 * tensor pipe itself (and the power cap) allows for each way of applying the low weight parts. */
int fad_bench_umma_mode(fad_handle* h, int mode, int ksteps, double* ms_out_host);

#ifdef __cplusplus
}
#endif
#endif /* FADTK_B200_H */
