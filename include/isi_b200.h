/*
 * isi_b200.h -- C ABI of the B200-native VQ-VAE-2 code-extraction hot path.
 *
 * The reference (SonyCSLParis/interactive-spectrogram-inpainting) has no FFI
 * layer: its boundary for this path is the Python nn.Module API.  Each entry
 * point below names the reference statement(s) it replaces (file:line relative
 * to the reference root); the Python modules in
 * interactive_spectrogram_inpainting_b200/ bind these with ctypes and mirror
 * the reference class/method names on top.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name starts with "h_";
 *  - the library never allocates, frees or keeps global state: callers own all
 *    buffers (outputs, workspaces, tables); calls are re-entrant and are enqueued
 *    on the given CUDA stream without synchronising it;
 *  - return value: 0 = ok, <0 = isi_status (bad argument / unsupported), >0 = a
 *    cudaError_t raised by the launch;
 *  - sm_100a only.  There is no CPU or other-architecture fallback.
 */
#ifndef ISI_B200_H_
#define ISI_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ISI_API __attribute__((visibility("default")))
#else
#define ISI_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* isi_stream_t; /* cudaStream_t */

enum isi_status {
  ISI_OK = 0,
  ISI_ERR_NULL = -1,        /* required pointer is NULL */
  ISI_ERR_SHAPE = -2,       /* non-positive or inconsistent sizes */
  ISI_ERR_UNSUPPORTED = -3, /* valid request this build has no kernel for */
  ISI_ERR_WORKSPACE = -4,   /* workspace too small / misaligned */
  ISI_ERR_ALIGN = -5        /* pointer or stride violates an alignment rule */
};

/* Which kernel isi_vq_assign uses. AUTO picks TCGEN05 when the shape allows. */
enum isi_assign_algo {
  ISI_ASSIGN_AUTO = 0,
  ISI_ASSIGN_SIMT_FP32 = 1, /* CUDA-core FP32 FMA, any D / K / layout          */
  ISI_ASSIGN_TCGEN05 = 2,   /* tcgen05.mma kind::tf32, 3xTF32 split, TMEM accum; */
                            /* codebook streamed in 64-code tiles (D = 64, any K)  */
  ISI_ASSIGN_TCGEN05_PAIR = 3, /* the same on a CTA pair (cta_group::2) with the   */
                            /* codebook resident in shared memory (D = 64, K<=512) */
  ISI_ASSIGN_TCGEN05_PAIR_STREAM = 4 /* CTA pair, codebook streamed in 128-code    */
                            /* tiles (D = 64 or 128, K <= 4096)                    */
};

/*
 * A logical [n_rows, dim] FP32 matrix whose rows are grouped in batches:
 *   element(row, d) = base[(row / rows_per_batch) * batch_stride
 *                          + (row % rows_per_batch) * row_stride + d * col_stride]
 * (strides in elements).  Contiguous [N, D]: {N, 0, D, 1}.  The permuted NCHW
 * view the reference feeds the quantiser (vqvae.py:260,272: conv output
 * [B, D, H, W].permute(0, 2, 3, 1)): {H*W, D*H*W, 1, H*W}.
 */
typedef struct isi_rows_layout {
  int64_t rows_per_batch;
  int64_t batch_stride;
  int64_t row_stride;
  int64_t col_stride;
} isi_rows_layout;

ISI_API int isi_version(void);
ISI_API const char* isi_status_string(int status);

/* ------------------------------------------------------------------ *
 *  (2) quantiser: distance / argmin / gather                          *
 * ------------------------------------------------------------------ */

/* Bytes of the "prepared codebook" scratch for a [dim, n_embed] codebook. */
ISI_API size_t isi_vq_prepared_bytes(int dim, int n_embed);

/*
 * Derive everything the kernels read from the codebook `embed` ([dim, n_embed]
 * FP32, code index contiguous -- the reference buffer `embed`,
 * bottleneck.py:47-49): ||e_k||^2 (bottleneck.py:59), the code-major copy E^T
 * that `embed_code` gathers from (bottleneck.py:103-104), and the pre-split,
 * pre-swizzled TF32 hi/lo operand tiles of -2E for the tensor-core kernel.
 * Must be re-run whenever `embed` changes (after every EMA update).
 */
ISI_API int isi_vq_prepare_codebook(const float* embed, int dim, int n_embed,
                            void* prepared, size_t prepared_bytes,
                            isi_stream_t stream);

/*
 * Nearest-code search.  Replaces bottleneck.py:55-61:
 *   dist = |x|^2 - 2 x.E + |E|^2 ;  _, ind = (-dist).max(1)
 * out_index[n] (int64) = argmin_k dist(n, k), lowest k on exact ties;
 * out_score[n] (optional) = min_k (|e_k|^2 - 2 x_n.e_k)  (= dist - |x_n|^2).
 */
ISI_API int isi_vq_assign(const float* x, const isi_rows_layout* x_layout, int64_t n_rows,
                  int dim, int n_embed, const void* prepared,
                  int64_t* out_index, float* out_score, int algo,
                  isi_stream_t stream);

/* Bytes of the per-call scratch of isi_vq_gather_stats. */
ISI_API size_t isi_vq_gather_workspace_bytes(int64_t n_rows, int dim);  /* 8-byte aligned */

/*
 * Code lookup + commitment partials + (training) EMA statistics.
 * Replaces bottleneck.py:75-77 (one_hot, embed_code), :81 and :83 (one-hot column
 * sums and x^T.onehot, here a segmented reduction), the numerator of :94 and the
 * forward value of :95.
 *   out_q(n, :)      = E^T[index[n]]       written with `q_layout` (nullable)
 *   stats[0:K]       += #{n : index[n]==k}                (FP32, nullable)
 *   stats[K + k*D+d] += sum_{n : index[n]==k} x(n, d)     (code-major, nullable)
 *   workspace        <- per-CTA partial sums of (q - x)^2 for isi_vq_finish
 * `stats` (K + K*D floats) must be zeroed by the caller before the first call of a
 * step; it is the buffer a data-parallel job all-reduces.
 * `counts_only`: accumulate stats[0:K] only (eval-mode perplexity).
 * Out-of-range indices set *status_flag (int32, nullable) to 1 and are skipped.
 */
ISI_API int isi_vq_gather_stats(const float* x, const isi_rows_layout* x_layout,
                        const int64_t* index, int64_t n_rows, int dim, int n_embed,
                        const void* prepared, float* out_q,
                        const isi_rows_layout* q_layout, float* stats,
                        int counts_only, void* workspace, size_t workspace_bytes,
                        int32_t* status_flag, isi_stream_t stream);

/*
 * Scalars of the forward: replaces bottleneck.py:94 (diff = mean((q-x)^2)) and
 * :97-100 (perplexity = exp(-sum p log max(p,1e-7)), p = counts / n_rows).
 * Deterministic (fixed-order FP64 reduction of the per-CTA partials).
 */
ISI_API int isi_vq_finish(const void* workspace, int64_t n_rows, int dim, int n_embed,
                  const float* stats, float* out_diff, float* out_perplexity,
                  isi_stream_t stream);

/*
 * EMA codebook update, in place.  Replaces bottleneck.py:80-92:
 *   cluster_size <- g*cluster_size + (1-g)*counts
 *   embed_avg    <- g*embed_avg    + (1-g)*embed_sum
 *   n = sum(cluster_size); cs' = (cluster_size+eps)/(n+K*eps)*n
 *   embed        <- embed_avg / cs'
 * `stats` as produced by isi_vq_gather_stats (optionally summed over ranks);
 * stats[0:K] is consumed and overwritten with cs'.  cluster_size [K],
 * embed_avg [D,K], embed [D,K] are the reference buffers.  decay / eps are the
 * Python doubles of the module; they are rounded to FP32 where the reference's
 * scalars meet FP32 tensors ((float)(1-decay), (float)(K*eps)).
 */
ISI_API int isi_vq_ema_update(float* stats, float* cluster_size, float* embed_avg,
                      float* embed, int dim, int n_embed, double decay, double eps,
                      isi_stream_t stream);

/*
 * embed_code: out(n, :) = E^T[index[n]].  Replaces bottleneck.py:103-104
 * (F.embedding(ids, embed.T)) and, with an NCHW `out_layout`, the permute of
 * vqvae.py:290,292.  Out-of-range ids set *status_flag and write zeros.
 */
ISI_API int isi_embed_code(const int64_t* index, int64_t n_rows, int dim, int n_embed,
                   const void* prepared, float* out,
                   const isi_rows_layout* out_layout, int32_t* status_flag,
                   isi_stream_t stream);

/* ------------------------------------------------------------------ *
 *  (2') pre-quantiser projection: concat + 1x1 conv + bias            *
 * ------------------------------------------------------------------ */

/* Bytes of the prepared weight image for `c_in` input channels (multiple of 64). */
ISI_API size_t isi_vq_project_prepared_bytes(int c_in);

/*
 * weight [c_out = 64, c_in] FP32 row-major -- the Conv2d(c_in, 64, 1) weight of
 * quantize_conv_t / quantize_conv_b (vqvae.py:149-150,175-177) -- -> TF32 hi/lo operand
 * chunks.  Re-run when the weight changes.  `prepared` 128-byte aligned.
 */
ISI_API int isi_vq_project_prepare(const float* weight, int c_in, int c_out, void* prepared,
                           size_t prepared_bytes, isi_stream_t stream);

/*
 * out[n, 0:64] = bias + W[:, 0:c0] src0[n, :] + W[:, c0:c0+c1] src1[n, :]
 * Replaces vqvae.py:260 (quantize_conv_t(enc_t)) with src1 = NULL, c1 = 0, and vqvae.py:271-272
 * (quantize_conv_b(torch.cat([dec_t, enc_b], 1))) with src0 = dec_t, src1 = enc_b: the
 * concatenation is never materialised.  Sources are channels-last rows: src_s[n, c] =
 * src_s[n * row_stride_s + c] (what a torch.channels_last [B, C, H, W] tensor is, with n = (b, h,
 * w)); c0, c1 multiples of 64, c0 + c1 <= 1024; row strides multiples of 4; pointers 16-byte
 * aligned.  `out` is contiguous [n_rows, 64]: the layout isi_vq_assign reads fastest.
 * 3xTF32 tensor-core contraction (FP32-equivalent; the reference's conv runs in TF32 on GPU).
 */
ISI_API int isi_vq_project(const float* src0, int c0, int64_t row_stride0, const float* src1, int c1,
                   int64_t row_stride1, int64_t n_rows, int c_out, const void* prepared,
                   const float* bias, float* out, isi_stream_t stream);

/* ------------------------------------------------------------------ *
 *  (1) front end: STFT -> (mel) -> log-magnitude + instantaneous freq *
 * ------------------------------------------------------------------ */

/*
 * Parameters of SpectrogramsHelper / MelSpectrogramsHelper.to_spectrogram
 * (external GANsynth_pytorch; built at utils/misc.py:10-29 from the kwargs
 * fs_hz, n_fft, hop_length, window_length, mel_*).  Tables are device arrays the
 * host side computes once in FP64 (see utils/spectrograms_helper.py).
 */
typedef struct isi_melif_params {
  int32_t n_fft;        /* 2048 (512 and 1024 also built)                       */
  int32_t hop;          /* hop_length                                           */
  int32_t pad_left;     /* zeros before the first sample (GANSynth: n_fft-hop)  */
  int32_t n_frames;     /* output time steps                                    */
  int32_t drop_dc;      /* 1: keep bins 1..n_fft/2, 0: keep bins 0..n_fft/2-1   */
  int32_t use_mel;      /* 0: linear log|X| + IF, 1: mel log-mag^2 + mel IF     */
  int32_t mel_width;    /* max non-zeros per mel bin (row pitch of mel_weight)  */
  float safelog_eps;    /* log(v + eps)                                         */
  const float* window;  /* [n_fft] analysis window                              */
  const float* twiddle; /* [n_fft, 2] cos,sin of -2*pi*j/n_fft, j<n_fft; 8-byte aligned */
  const int32_t* mel_start; /* [n_fft/2] first linear bin of each mel band      */
  const int32_t* mel_count; /* [n_fft/2] band length (0..mel_width)             */
  const float* mel_weight;  /* [n_fft/2, mel_width] band weights, zero beyond the band's
                               length; mel_width == 8 and 16-byte alignment select
                               two 16-byte loads per band, anything else scalar loads */
  int32_t channels_last;    /* 0: out is [B,2,F,T'] planes (the reference layout);   */
                            /* 1: the same logical tensor in torch channels_last     */
                            /*    storage [B,F,T',2] (what the cuDNN convs consume)  */
                            /* 2: 2x2 space-to-depth blocks, [B,F/2,T'/2,(f&1,t&1,c)]: */
                            /*    the encoder's first conv (4x4, stride 2, 2 input      */
                            /*    channels, encoder_decoder.py:66-70) becomes a 3x3      */
                            /*    stride-1 conv over 8 channels; needs even n_frames     */
                            /* 3: the same blocks with the frequency index fastest,      */
                            /*    [B,T'/2,F/2,(f&1,t&1,c)]: the layout the kernel's lanes  */
                            /*    (consecutive rows) write as whole 128-byte lines; the    */
                            /*    encoder then runs on the transposed plane with transposed */
                            /*    filters (vqvae.py: space_to_depth="transposed")          */
  /* fused epilogue (both live in GANsynth_pytorch in the reference):       */
  int32_t mask_phase;       /* 1: channel 1 := 0 where channel 0 < mask_threshold (the       */
  float mask_threshold;     /*    masked-phase transform, extract_code.py:178-181)           */
  float out_scale[2];       /* then channel c := c * out_scale[c] + out_bias[c] (the         */
  float out_bias[2];        /*    DataNormalizer affine of vqvae.py:254-255); 1 / 0 = off    */
  /* input samples: FP32, or 16-bit PCM as NSynth stores it (the reference's dataset class     */
  /* converts on the CPU and uploads FP32; reading PCM halves the host->device bytes)          */
  int32_t audio_format;     /* isi_audio_format                                                */
  float pcm_scale;          /* PCM16 only: sample = (float)pcm * pcm_scale (e.g. 1/32768)      */
} isi_melif_params;

typedef enum { ISI_AUDIO_F32 = 0, ISI_AUDIO_PCM16 = 1 } isi_audio_format;
typedef enum { ISI_SPEC_PLANAR = 0, ISI_SPEC_CHANNELS_LAST = 1, ISI_SPEC_SPACE_TO_DEPTH = 2,
               ISI_SPEC_SPACE_TO_DEPTH_T = 3 } isi_spec_layout;

/*
 * audio [n_notes, n_samples] (contiguous; FP32 or int16 per h_params->audio_format)
 * -> out [n_notes, 2, n_fft/2, n_frames] FP32: channel 0 log-magnitude, channel 1 IF,
 * frequency-major / time-contiguous like the reference tensors
 * (Inference.ipynb:71, flask_server.py:891-896), or channel-interleaved storage of the
 * same logical tensor when h_params->channels_last is set.
 */
ISI_API int isi_melif_forward(const void* audio, int64_t n_notes, int64_t n_samples,
                      const isi_melif_params* h_params, float* out,
                      isi_stream_t stream);

/* ------------------------------------------------------------------ *
 *  (1') inverse front end: (mel) log-magnitude + IF -> audio          *
 * ------------------------------------------------------------------ */

/*
 * Parameters of SpectrogramsHelper / MelSpectrogramsHelper.to_audio (external
 * GANsynth_pytorch; reference call sites flask_server.py:596,1016,1110, sample.py:599,
 * train_vqvae.py:392-394, utils/losses/spectral.py:122-126).  Same transform geometry as
 * isi_melif_params; the band tables are the TRANSPOSE of the analysis filterbank with
 * GANSynth's column normalisation (for every linear row the mel rows that feed it).
 */
typedef struct isi_imelif_params {
  int32_t n_fft;        /* 2048 (512 and 1024 also built)                            */
  int32_t hop;          /* hop_length, <= n_fft                                      */
  int32_t pad_left;     /* padded samples the forward transform put before sample 0  */
  int32_t n_frames;     /* input time steps                                          */
  int32_t drop_dc;      /* 1: rows are bins 1..n_fft/2, 0: bins 0..n_fft/2-1         */
  int32_t use_mel;      /* 0: linear log|X| + IF in, 1: mel log-mag^2 + mel IF in    */
  int32_t band_width;   /* row pitch of band_weight (<= 8)                           */
  float safelog_eps;    /* mel mode: |X| = sqrt(projected mag^2 + eps)               */
  const float* window;  /* [n_fft] synthesis window (the analysis window)            */
  const float* twiddle; /* [n_fft, 2] as isi_melif_params.twiddle                    */
  const int32_t* band_start; /* [n_fft/2] first mel row feeding each linear row      */
  const int32_t* band_count; /* [n_fft/2] how many (0..band_width)                   */
  const float* band_weight;  /* [n_fft/2, band_width], zero beyond the count         */
  const float* ola_scale;    /* [hop*(n_frames-1)+n_fft]: 1 / (n_fft * sum over the   */
                             /* frames covering the sample of window^2), 0 where none */
  float in_scale[2];    /* channel c is read as c * in_scale[c] + in_bias[c] (the     */
  float in_bias[2];     /* shape of DataNormalizer.denormalize); 1 / 0 = off          */
  int32_t seg_frames;   /* 0 = choose; else frames per CTA (testing / tuning)         */
} isi_imelif_params;

/*
 * spec [n_notes, 2, n_fft/2, n_frames] FP32 planes (channel 0 log-magnitude, channel 1 IF,
 * time contiguous) -> audio [n_notes, n_samples] FP32, sample j = padded position
 * j + pad_left of the overlap-added inverse STFT; n_samples <= hop*(n_frames-1) + n_fft
 * - pad_left (hop*n_frames - pad_left undoes isi_melif_forward's padding).
 */
ISI_API int isi_melif_inverse(const float* spec, int64_t n_notes, const isi_imelif_params* h_params,
                      float* audio, int64_t n_samples, isi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ISI_B200_H_ */
