// extern "C" surface of libisi_b200.so (declared in include/isi_b200.h).
// Argument validation lives here; the launchers live next to their kernels.
#include "common.cuh"

namespace isi {
int launch_assign_simt(const float*, const isi_rows_layout&, int64_t, int, int, const Prepared&,
                       int64_t*, float*, cudaStream_t);
bool assign_tc_supported(const isi_rows_layout&, int64_t, int, int);
int launch_assign_tc(const float*, const isi_rows_layout&, int64_t, int, int, const Prepared&,
                     int64_t*, float*, cudaStream_t);
int launch_prepare(const float*, int, int, const Prepared&, cudaStream_t);
int launch_prepare_tc(const float*, int, int, const Prepared&, cudaStream_t);
bool assign_pair_supported(const isi_rows_layout&, int64_t, int, int);
int launch_prepare_pair(const float*, int, int, const Prepared&, cudaStream_t);
int launch_assign_pair(const float*, const isi_rows_layout&, int64_t, int, int, const Prepared&,
                       int64_t*, float*, cudaStream_t);
bool assign_pstream_supported(const isi_rows_layout&, int64_t, int, int);
int launch_prepare_pstream(const float*, int, int, const Prepared&, cudaStream_t);
int launch_assign_pstream(const float*, const isi_rows_layout&, int64_t, int, int, const Prepared&,
                          int64_t*, float*, cudaStream_t);
size_t gather_workspace_bytes(int64_t);
int launch_gather_stats(const float*, const isi_rows_layout&, const int64_t*, int64_t, int, int,
                        const Prepared&, float*, const isi_rows_layout&, float*, int, void*,
                        int32_t*, cudaStream_t);
int launch_finish(const void*, int64_t, int, int, const float*, float*, float*, cudaStream_t);
int launch_ema_update(float*, float*, float*, float*, int, int, double, double, cudaStream_t);
int launch_embed_code(const int64_t*, int64_t, int, int, const Prepared&, float*,
                      const isi_rows_layout&, int32_t*, cudaStream_t);
size_t project_prepared_bytes(int);
int launch_project_prepare(const float*, int, void*, cudaStream_t);
int launch_project(const float*, int, int64_t, const float*, int, int64_t, int64_t, const void*, const float*,
                   float*, cudaStream_t);
int launch_melif(const void*, int64_t, int64_t, const isi_melif_params&, float*, cudaStream_t);
int launch_imelif(const float*, int64_t, const isi_imelif_params&, float*, int64_t, int, cudaStream_t);
}  // namespace isi

using namespace isi;

static bool layout_ok(const isi_rows_layout* l) { return l && l->rows_per_batch > 0; }

extern "C" {

ISI_API int isi_version(void) { return 100; }

ISI_API const char* isi_status_string(int s) {
  if (s > 0) return cudaGetErrorString((cudaError_t)s);
  switch (s) {
    case ISI_OK: return "ok";
    case ISI_ERR_NULL: return "required pointer is NULL";
    case ISI_ERR_SHAPE: return "bad shape";
    case ISI_ERR_UNSUPPORTED: return "unsupported configuration";
    case ISI_ERR_WORKSPACE: return "workspace too small or misaligned";
    case ISI_ERR_ALIGN: return "misaligned pointer or stride";
    default: return "unknown status";
  }
}

ISI_API size_t isi_vq_prepared_bytes(int dim, int n_embed) {
  if (dim <= 0 || n_embed <= 0) return 0;
  return prepared_view(nullptr, dim, n_embed).bytes;
}

ISI_API int isi_vq_prepare_codebook(const float* embed, int dim, int n_embed, void* prepared,
                            size_t prepared_bytes, isi_stream_t stream) {
  if (!embed || !prepared) return ISI_ERR_NULL;
  if (dim <= 0 || n_embed <= 0) return ISI_ERR_SHAPE;
  if (prepared_bytes < isi_vq_prepared_bytes(dim, n_embed)) return ISI_ERR_WORKSPACE;
  if ((uintptr_t)prepared % 256) return ISI_ERR_ALIGN;
  Prepared p = prepared_view(prepared, dim, n_embed);
  int rc = launch_prepare(embed, dim, n_embed, p, (cudaStream_t)stream);
  if (rc) return rc;
  rc = launch_prepare_tc(embed, dim, n_embed, p, (cudaStream_t)stream);
  if (rc) return rc;
  rc = launch_prepare_pair(embed, dim, n_embed, p, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_prepare_pstream(embed, dim, n_embed, p, (cudaStream_t)stream);
}

ISI_API int isi_vq_assign(const float* x, const isi_rows_layout* xl, int64_t n_rows, int dim, int n_embed,
                  const void* prepared, int64_t* out_index, float* out_score, int algo,
                  isi_stream_t stream) {
  if (!x || !prepared || !out_index) return ISI_ERR_NULL;
  if (n_rows < 0 || dim <= 0 || n_embed <= 0 || !layout_ok(xl)) return ISI_ERR_SHAPE;
  if (n_rows == 0) return ISI_OK;
  Prepared p = prepared_view(prepared, dim, n_embed);
  const bool tc_ok = assign_tc_supported(*xl, n_rows, dim, n_embed);
  const bool pair_ok = assign_pair_supported(*xl, n_rows, dim, n_embed);
  if (algo == ISI_ASSIGN_TCGEN05_PAIR && !pair_ok) return ISI_ERR_UNSUPPORTED;
  if (algo == ISI_ASSIGN_TCGEN05_PAIR || (algo == ISI_ASSIGN_AUTO && pair_ok))
    return launch_assign_pair(x, *xl, n_rows, dim, n_embed, p, out_index, out_score,
                              (cudaStream_t)stream);
  const bool pstream_ok = !pair_ok && assign_pstream_supported(*xl, n_rows, dim, n_embed);
  if (algo == ISI_ASSIGN_TCGEN05_PAIR_STREAM && !pstream_ok) return ISI_ERR_UNSUPPORTED;
  if (algo == ISI_ASSIGN_TCGEN05_PAIR_STREAM || (algo == ISI_ASSIGN_AUTO && pstream_ok))
    return launch_assign_pstream(x, *xl, n_rows, dim, n_embed, p, out_index, out_score,
                                 (cudaStream_t)stream);
  if (algo == ISI_ASSIGN_TCGEN05 && !tc_ok) return ISI_ERR_UNSUPPORTED;
  if (algo == ISI_ASSIGN_TCGEN05 || (algo == ISI_ASSIGN_AUTO && tc_ok))
    return launch_assign_tc(x, *xl, n_rows, dim, n_embed, p, out_index, out_score,
                            (cudaStream_t)stream);
  if (algo != ISI_ASSIGN_AUTO && algo != ISI_ASSIGN_SIMT_FP32) return ISI_ERR_UNSUPPORTED;
  return launch_assign_simt(x, *xl, n_rows, dim, n_embed, p, out_index, out_score,
                            (cudaStream_t)stream);
}

ISI_API size_t isi_vq_gather_workspace_bytes(int64_t n_rows, int dim) {
  (void)dim;
  return gather_workspace_bytes(n_rows < 0 ? 0 : n_rows);
}

ISI_API int isi_vq_gather_stats(const float* x, const isi_rows_layout* xl, const int64_t* index,
                        int64_t n_rows, int dim, int n_embed, const void* prepared, float* out_q,
                        const isi_rows_layout* ql, float* stats, int counts_only, void* workspace,
                        size_t workspace_bytes, int32_t* status_flag, isi_stream_t stream) {
  if (!index || !prepared || !workspace) return ISI_ERR_NULL;
  if (n_rows < 0 || dim <= 0 || n_embed <= 0) return ISI_ERR_SHAPE;
  if (x && !layout_ok(xl)) return ISI_ERR_SHAPE;
  if (out_q && !layout_ok(ql)) return ISI_ERR_SHAPE;
  if (stats && !counts_only && !x) return ISI_ERR_NULL;
  if (workspace_bytes < gather_workspace_bytes(n_rows) || (uintptr_t)workspace % 8)
    return ISI_ERR_WORKSPACE;
  if (n_rows == 0) return ISI_OK;
  isi_rows_layout none{1, 0, 0, 0};
  return launch_gather_stats(x, x ? *xl : none, index, n_rows, dim, n_embed,
                             prepared_view(prepared, dim, n_embed), out_q, out_q ? *ql : none,
                             stats, counts_only, workspace, status_flag, (cudaStream_t)stream);
}

ISI_API int isi_vq_finish(const void* workspace, int64_t n_rows, int dim, int n_embed, const float* stats,
                  float* out_diff, float* out_perplexity, isi_stream_t stream) {
  if (!workspace) return ISI_ERR_NULL;
  if (n_rows <= 0 || dim <= 0 || n_embed <= 0) return ISI_ERR_SHAPE;
  if (out_perplexity && !stats) return ISI_ERR_NULL;
  return launch_finish(workspace, n_rows, dim, n_embed, stats, out_diff, out_perplexity,
                       (cudaStream_t)stream);
}

ISI_API int isi_vq_ema_update(float* stats, float* cluster_size, float* embed_avg, float* embed, int dim,
                      int n_embed, double decay, double eps, isi_stream_t stream) {
  if (!stats || !cluster_size || !embed_avg || !embed) return ISI_ERR_NULL;
  if (dim <= 0 || n_embed <= 0) return ISI_ERR_SHAPE;
  return launch_ema_update(stats, cluster_size, embed_avg, embed, dim, n_embed, decay, eps,
                           (cudaStream_t)stream);
}

ISI_API int isi_embed_code(const int64_t* index, int64_t n_rows, int dim, int n_embed,
                   const void* prepared, float* out, const isi_rows_layout* ol,
                   int32_t* status_flag, isi_stream_t stream) {
  if (!index || !prepared || !out) return ISI_ERR_NULL;
  if (n_rows < 0 || dim <= 0 || n_embed <= 0 || !layout_ok(ol)) return ISI_ERR_SHAPE;
  if (n_rows == 0) return ISI_OK;
  return launch_embed_code(index, n_rows, dim, n_embed, prepared_view(prepared, dim, n_embed), out,
                           *ol, status_flag, (cudaStream_t)stream);
}

ISI_API size_t isi_vq_project_prepared_bytes(int c_in) {
  if (c_in <= 0 || c_in % 64 || c_in > 1024) return 0;
  return project_prepared_bytes(c_in);
}

ISI_API int isi_vq_project_prepare(const float* weight, int c_in, int c_out, void* prepared,
                           size_t prepared_bytes, isi_stream_t stream) {
  if (!weight || !prepared) return ISI_ERR_NULL;
  if (c_out != 64 || c_in <= 0 || c_in % 64 || c_in > 1024) return ISI_ERR_UNSUPPORTED;
  if (prepared_bytes < project_prepared_bytes(c_in)) return ISI_ERR_WORKSPACE;
  if ((uintptr_t)prepared % 128) return ISI_ERR_ALIGN;
  return launch_project_prepare(weight, c_in, prepared, (cudaStream_t)stream);
}

ISI_API int isi_vq_project(const float* src0, int c0, int64_t row_stride0, const float* src1, int c1,
                   int64_t row_stride1, int64_t n_rows, int c_out, const void* prepared,
                   const float* bias, float* out, isi_stream_t stream) {
  if (!src0 || !prepared || !out || (c1 > 0 && !src1)) return ISI_ERR_NULL;
  if (n_rows < 0 || c0 <= 0 || c1 < 0) return ISI_ERR_SHAPE;
  if (c_out != 64 || c0 % 64 || c1 % 64 || c0 + c1 > 1024) return ISI_ERR_UNSUPPORTED;
  if (row_stride0 < c0 || (c1 > 0 && row_stride1 < c1)) return ISI_ERR_SHAPE;
  if (row_stride0 % 4 || (c1 > 0 && row_stride1 % 4) || (uintptr_t)src0 % 16 || (c1 > 0 && (uintptr_t)src1 % 16) ||
      (uintptr_t)out % 16 || (uintptr_t)prepared % 128 || (bias && (uintptr_t)bias % 4))
    return ISI_ERR_ALIGN;
  if (n_rows == 0) return ISI_OK;
  return launch_project(src0, c0, row_stride0, c1 > 0 ? src1 : src0, c1, c1 > 0 ? row_stride1 : row_stride0,
                        n_rows, prepared, bias, out, (cudaStream_t)stream);
}

ISI_API int isi_melif_forward(const void* audio, int64_t n_notes, int64_t n_samples,
                      const isi_melif_params* hp, float* out, isi_stream_t stream) {
  if (!audio || !hp || !out || !hp->window || !hp->twiddle) return ISI_ERR_NULL;
  if (hp->use_mel && (!hp->mel_start || !hp->mel_count || !hp->mel_weight)) return ISI_ERR_NULL;
  if (n_notes < 0 || n_samples <= 0 || hp->hop <= 0 || hp->n_frames <= 0 || hp->pad_left < 0)
    return ISI_ERR_SHAPE;
  if (hp->use_mel && hp->mel_width <= 0) return ISI_ERR_SHAPE;
  if (hp->audio_format != ISI_AUDIO_F32 && hp->audio_format != ISI_AUDIO_PCM16) return ISI_ERR_UNSUPPORTED;
  if (hp->channels_last < ISI_SPEC_PLANAR || hp->channels_last > ISI_SPEC_SPACE_TO_DEPTH_T) return ISI_ERR_UNSUPPORTED;
  if (hp->channels_last >= ISI_SPEC_SPACE_TO_DEPTH && (hp->n_frames % 2 || (hp->n_fft / 2) % 2)) return ISI_ERR_SHAPE;
  if ((uintptr_t)audio % (hp->audio_format == ISI_AUDIO_PCM16 ? 2 : 4)) return ISI_ERR_ALIGN;
  if ((uintptr_t)out % 16 || (uintptr_t)hp->twiddle % 8) return ISI_ERR_ALIGN;
  if (n_notes == 0) return ISI_OK;
  return launch_melif(audio, n_notes, n_samples, *hp, out, (cudaStream_t)stream);
}

ISI_API int isi_melif_inverse(const float* spec, int64_t n_notes, const isi_imelif_params* hp, float* audio,
                      int64_t n_samples, isi_stream_t stream) {
  if (!spec || !hp || !audio || !hp->window || !hp->twiddle || !hp->ola_scale) return ISI_ERR_NULL;
  if (hp->use_mel && (!hp->band_start || !hp->band_count || !hp->band_weight)) return ISI_ERR_NULL;
  if (n_notes < 0 || n_samples <= 0 || hp->hop <= 0 || hp->n_frames <= 0 || hp->pad_left < 0 || hp->seg_frames < 0)
    return ISI_ERR_SHAPE;
  if (hp->hop > hp->n_fft) return ISI_ERR_UNSUPPORTED;
  if (hp->use_mel && hp->band_width <= 0) return ISI_ERR_SHAPE;
  if (n_samples > (int64_t)hp->hop * (hp->n_frames - 1) + hp->n_fft - hp->pad_left) return ISI_ERR_SHAPE;
  if ((uintptr_t)spec % 4 || (uintptr_t)audio % 4 || (uintptr_t)hp->twiddle % 8) return ISI_ERR_ALIGN;
  if (n_notes == 0) return ISI_OK;
  return launch_imelif(spec, n_notes, *hp, audio, n_samples, hp->seg_frames, (cudaStream_t)stream);
}

}  // extern "C"
