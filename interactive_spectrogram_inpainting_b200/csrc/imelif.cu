// Inverse front end: (mel) log-magnitude + instantaneous frequency -> audio.
//
// Replaces SpectrogramsHelper / MelSpectrogramsHelper.to_audio (external GANsynth_pytorch;
// reference call sites flask_server.py:596,1016,1110, sample.py:599, train_vqvae.py:392-394,
// utils/losses/spectral.py:122-126) -- SURVEY.md section 8(f) N4, the caller on the far side of
// decode_code in the interactive path.
//
// The mirror of melif.cu.  One CTA owns a run of frames of one note and walks it in batches
// of FB frames; per batch the [2][M][FB] slab of input values is brought to shared memory with
// TMA tensor copies (cp.async.bulk.tensor.2d: eight [256 rows x FB frames] boxes of the
// [notes*2*M, frames] view of the input, landing densely in the slab, completion on an mbarrier;
// 16-byte cp.async gathers when no tensor map can be had) issued a batch ahead, exponentiated in place, projected mel->linear
// row by row (banded transpose of the analysis filterbank) into FB natural-order spectra
// with the running phase of every row in a register, folded into the half-size complex
// spectrum, transformed by the forward FFT passes (conjugate trick), and overlap-added in
// shared memory: HBM sees each input value and each output sample once.  A CTA that does
// not start at frame 0 seeds its phases with one FP64 prefix sum over the earlier frames
// and re-synthesises the few frames whose windows reach into its first hop.
#include <stdlib.h>
#include <string.h>

#include <cuda.h>          // CUtensorMap and its enums only: the encoder is fetched at run time

#include "common.cuh"
#include "imelif_core.cuh"

namespace isi {
using namespace imelif;

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// one [box rows x box frames] tile of the 2-D view, coordinates (frame, row), dense in shared memory
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int frame, int row, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_addr(dst)), "l"(map), "r"(frame), "r"(row), "r"(smem_addr(bar)) : "memory");
}
constexpr int kTmaBoxRows = 256;            // a TMA box dimension is at most 256

struct ImelifSmem {
  int tw, win, slab, za, carry, bar, total;   // byte offsets into dynamic shared memory
};

template <int NFFT, int FB>
__host__ __device__ inline ImelifSmem imelif_smem_layout() {
  using P = Plan<NFFT>;
  ImelifSmem s;
  int off = 0;
  s.tw = off;    off += (NFFT / 2) * 8;              // FFT twiddles (fft_table_source)
  s.win = off;   off += NFFT * 4;                    // synthesis window
  s.slab = off;  off += 2 * (NFFT / 2) * FB * 4;     // input values of one batch
  s.za = off;    off += FB * P::kPitchA * 8;         // spectra / FFT workspace / frame samples
  s.carry = off; off += 2 * NFFT * 4;                // overlap-add carry, ping-pong
  s.bar = off;   off += 16;                          // mbarrier of the slab's tensor copies
  s.total = off;
  return s;
}

}  // namespace

template <int NFFT, int FB, int NT, bool MEL>
__global__ void __launch_bounds__(NT, 2)
imelif_kernel(const float* __restrict__ spec, isi_imelif_params p, float* __restrict__ audio,
              int64_t n_samples, int vec_in, int vec_out, int seg_frames, int n_segs,
              const __grid_constant__ CUtensorMap spec_map, int use_tma) {
  using P = Plan<NFFT>;
  constexpr int M = P::M;
  constexpr int IPT = (M / 2) / NT;           // tangle items per thread and frame
  constexpr int RPT = M / NT;                 // linear rows per thread
  constexpr int CPT = 2 * M / NT;             // slab chunks per thread
  constexpr int kGroups = NT / 64;            // frames transformed concurrently
  static_assert(IPT >= 1 && (M / 2) % NT == 0 && NT % 64 == 0, "bad thread count");
  extern __shared__ __align__(128) unsigned char smem[];
  const ImelifSmem L = imelif_smem_layout<NFFT, FB>();
  cpx* twm = reinterpret_cast<cpx*>(smem + L.tw);
  float* win = reinterpret_cast<float*>(smem + L.win);
  float* slab = reinterpret_cast<float*>(smem + L.slab);
  cpx* zA = reinterpret_cast<cpx*>(smem + L.za);
  float* carry = reinterpret_cast<float*>(smem + L.carry);
  uint64_t* slab_bar = reinterpret_cast<uint64_t*>(smem + L.bar);

  const int tid = threadIdx.x;
  const int note_idx = blockIdx.x / n_segs, seg = blockIdx.x - note_idx * n_segs;
  const int fs = seg * seg_frames;                          // first frame this CTA owns
  const int fe = min(p.n_frames, fs + seg_frames);          // one past its last frame
  const bool last_seg = fe == p.n_frames;
  // earlier frames whose windows reach the CTA's first sample are synthesised again, from a
  // batch boundary so that the slab copies stay aligned
  int fstart = fs - ola_lookback_frames(NFFT, p.hop);
  fstart = fstart <= 0 ? 0 : fstart / FB * FB;
  const int64_t emit_from = (int64_t)fs * p.hop;            // first padded position it writes
  const float* note0 = spec + (int64_t)note_idx * 2 * M * p.n_frames;
  const float* note1 = note0 + (int64_t)M * p.n_frames;
  float* out = audio + (int64_t)note_idx * n_samples;
  const int dc = p.drop_dc ? 1 : 0;
  const int real_row = dc ? M - 1 : 0;                      // the row that is X[M] / X[0]

  // ---- one-time setup ----
  const cpx* tw_global = reinterpret_cast<const cpx*>(p.twiddle);     // W_N^j, j < N
  for (int i = tid; i < M; i += NT) twm[i] = tw_global[fft_table_source<P>(i)];
  for (int i = tid; i < NFFT; i += NT) win[i] = p.window[i];
  for (int i = tid; i < 2 * NFFT; i += NT) carry[i] = 0.f;
  if (tid == 0) {
    mbar_init(slab_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t slab_phase = 0;
  cpx w_item[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) w_item[i] = tw_global[tid + i * NT];
  int band_start[RPT], band_count[RPT];
  float band_w[RPT][MEL ? kMaxMelWidth : 1];
  float phase[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int row = tid + r * NT;
    band_start[r] = row; band_count[r] = 0; phase[r] = 0.f;
    if (MEL) {
      band_start[r] = __ldg(p.band_start + row);
      band_count[r] = __ldg(p.band_count + row);
#pragma unroll
      for (int i = 0; i < kMaxMelWidth; ++i)
        band_w[r][i] = (i < p.band_width) ? __ldg(p.band_weight + (int64_t)row * p.band_width + i) : 0.f;
    }
  }

  // slab of frames [f0, f0 + nf): chunk q = tid + i NT is FB time steps of row q % M, channel q / M
  auto fill_slab = [&](int f0, int nf) {
    if (use_tma) {
      // (every batch of this path is a full one: n_frames and the segments are multiples of FB)
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slab was read / rewritten in place
        mbar_expect_tx(slab_bar, 2 * M * FB * 4);
        const int row0 = note_idx * 2 * M;
#pragma unroll
        for (int r = 0; r < 2 * M; r += kTmaBoxRows)
          tma_load_2d(slab + r * FB, &spec_map, f0, row0 + r, slab_bar);
      }
      return;
    }
    const bool vec = vec_in && nf == FB;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int q = tid + i * NT;
      const float* rows = (q < M ? note0 + (int64_t)q * p.n_frames : note1 + (int64_t)(q - M) * p.n_frames);
      if (vec) cp_async16(slab + q * FB, rows + f0);
      else slab_fill_chunk<FB>(slab, q, rows, f0, nf);
    }
    cp_async_commit();
  };

  fill_slab(fstart, min(FB, fe - fstart));
  __syncthreads();

  if (fstart > 0) {
    // running phase before frame fstart: FP64 prefix sums of channel 1, projected per row
    double* sums = reinterpret_cast<double*>(zA);
    double row_sum[RPT];
    lookback_rows_sum<RPT>(note1, p.n_frames, tid, NT, fstart, p.in_scale[1], p.in_bias[1], vec_in != 0, row_sum);
#pragma unroll
    for (int r = 0; r < RPT; ++r) sums[tid + r * NT] = row_sum[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; ++r) phase[r] = lookback_phase<MEL>(sums, band_start[r], band_count[r], band_w[r]);
    __syncthreads();
  }

  const uint32_t group_bar = 1 + (tid >> 6);
  int flip = 0;
  for (int f0 = fstart; f0 < fe; f0 += FB) {
    const int nf = min(FB, fe - f0);
    const bool last_batch = f0 + FB >= fe;
    float* carry_in = carry + flip * NFFT;
    float* carry_out = carry + (flip ^ 1) * NFFT;
    flip ^= 1;

    // ---- slab: the batch's values arrived; exponentiate / scale them in place ----
    if (use_tma) { mbar_wait(slab_bar, slab_phase & 1); ++slab_phase; }
    else cp_async_wait_all();
#pragma unroll
    for (int i = 0; i < CPT; ++i)
      slab_transform_chunk<FB>(slab, tid + i * NT, M, p.in_scale[0], p.in_bias[0], p.in_scale[1], p.in_bias[1]);
    __syncthreads();

    // ---- build the FB spectra ----
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int row = tid + r * NT;
      const int cu = MEL ? __reduce_max_sync(0xffffffffu, band_count[r]) : 0;
      build_row<FB, MEL>(slab, M, band_start[r], band_count[r], cu, band_w[r], p.safelog_eps,
                         row == real_row, phase[r], zA + row + dc, P::kPitchA);
    }
    if (tid < FB) zA[tid * P::kPitchA + (dc ? 0 : M)] = cpx{0.f, 0.f};      // the dropped bin
    __syncthreads();
    if (!last_batch) fill_slab(f0 + FB, min(FB, fe - f0 - FB));            // slab is free: next batch

    // ---- fold into the half-size spectrum ----
    for (int fb = 0; fb < nf; ++fb) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) tangle_item<P>(tid + i * NT, zA + fb * P::kPitchA, w_item[i]);
    }
    __syncthreads();

    // ---- transform: the 64 threads of a group take one frame at a time ----
    for (int fb = tid / 64; fb < nf; fb += kGroups) {
      cpx* z = zA + fb * P::kPitchA;
      {
        cpx v[P::R1];
        ifft_pass1_load<P>(tid & 63, z, v);
        asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
        ifft_pass1_store<P>(tid & 63, v, twm, z);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
      fft_pass2<P>(tid & 63, twm, z);
      asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
      Pass3Regs<P, cpx> regs;
      fft_pass3_load<P>(tid & 63, z, regs);
      asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
      fft_pass3_store<P>(tid & 63, regs, z);
    }
    __syncthreads();

    // ---- overlap-add ----
    const int span = (nf - 1) * p.hop + NFFT;
    const int emit_len = (last_batch && last_seg) ? span : nf * p.hop;
    const int64_t pos0 = (int64_t)f0 * p.hop;
    if (vec_out) {
      for (int s0 = 4 * tid; s0 < span; s0 += 4 * NT) {
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s0 < NFFT) c = *reinterpret_cast<const float4*>(carry_in + s0);  // zero beyond the carry
        float acc[4] = {c.x, c.y, c.z, c.w};
        ola_quad<FB>(zA, P::kPitchA, win, NFFT, p.hop, nf, s0, acc);
        if (s0 < emit_len) {
          const int64_t pos = pos0 + s0, j = pos - p.pad_left;
          if (pos >= emit_from && j >= 0 && j < n_samples) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.ola_scale + pos));
            st_stream4(out + j, make_float4(acc[0] * sc.x, acc[1] * sc.y, acc[2] * sc.z, acc[3] * sc.w));
          }
        } else {
          *reinterpret_cast<float4*>(carry_out + (s0 - emit_len)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
      }
    } else {
      for (int s = tid; s < span; s += NT) {
        const float acc = (s < NFFT ? carry_in[s] : 0.f) + ola_sample(zA, P::kPitchA, win, NFFT, p.hop, nf, s);
        if (s < emit_len) {
          const int64_t pos = pos0 + s, j = pos - p.pad_left;
          if (pos >= emit_from && j >= 0 && j < n_samples) out[j] = acc * __ldg(p.ola_scale + pos);
        } else {
          carry_out[s - emit_len] = acc;
        }
      }
    }
    // the carry only fills [0, span - emit_len); what the next batch reads beyond must be zero
    for (int s = span - emit_len + tid; s < NFFT; s += NT) carry_out[s] = 0.f;
    __syncthreads();
  }
}

// Frames per CTA: whole notes when the batch fills the GPU on its own, else the split that
// maximises (wave efficiency) x (useful / useful + re-synthesised frames).
static void choose_segments(int64_t n_notes, int n_frames, int fb, int lookback, int* seg_frames, int* n_segs) {
  const int frames_padded = (n_frames + fb - 1) / fb * fb;
  const int max_segs = frames_padded / fb > 1 ? frames_padded / fb : 1;   // one batch per CTA at most
  const double slots = 2.0 * kNumSms;   // __launch_bounds__(NT, 2)
  const double redo = (lookback + fb - 1) / fb * fb;
  double best = -1.0;
  *seg_frames = frames_padded; *n_segs = 1;
  for (int s = 1; s <= max_segs; ++s) {
    const int sf = ((frames_padded + s - 1) / s + fb - 1) / fb * fb;
    const int ns = (n_frames + sf - 1) / sf;
    const double waves = (double)n_notes * ns / slots;
    const double wave_eff = waves / (double)(int64_t)(waves + 0.999999);
    const double eff = wave_eff * (ns == 1 ? 1.0 : (double)sf / (sf + redo));
    if (eff > best + 1e-9) { best = eff; *seg_frames = sf; *n_segs = ns; }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return (EncodeTiledFn)ptr;
  }();
  return fn;
}

// [n_notes * 2 * M rows, n_frames] FP32 view of the input, boxes of kTmaBoxRows x fb
static bool make_spec_map(CUtensorMap* map, const float* spec, int64_t n_notes, int m, int n_frames, int fb) {
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode || getenv("ISI_IMELIF_NO_TMA")) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n_frames, (cuuint64_t)n_notes * 2 * (cuuint64_t)m};
  const cuuint64_t strides[1] = {(cuuint64_t)n_frames * 4};
  const cuuint32_t box[2] = {(cuuint32_t)fb, (cuuint32_t)kTmaBoxRows};
  const cuuint32_t elem[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(spec), dims, strides, box, elem,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NFFT, int FB, int NT, bool MEL>
static int launch_imelif_t(const float* spec, int64_t n_notes, const isi_imelif_params& p, float* audio,
                           int64_t n_samples, int seg_frames_override, cudaStream_t stream) {
  const ImelifSmem L = imelif_smem_layout<NFFT, FB>();
  cudaError_t e = cudaFuncSetAttribute(imelif_kernel<NFFT, FB, NT, MEL>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return (int)e;
  const int vec_in = (FB == 4) && (p.n_frames % 4 == 0) && ((uintptr_t)spec % 16 == 0);
  const int vec_out = (p.hop % 4 == 0) && (p.pad_left % 4 == 0) && (n_samples % 4 == 0) &&
                      ((uintptr_t)audio % 16 == 0) && ((uintptr_t)p.ola_scale % 16 == 0);
  int seg_frames, n_segs;
  choose_segments(n_notes, p.n_frames, FB, ola_lookback_frames(NFFT, p.hop), &seg_frames, &n_segs);
  if (seg_frames_override > 0) {
    seg_frames = (seg_frames_override + FB - 1) / FB * FB;
    n_segs = (p.n_frames + seg_frames - 1) / seg_frames;
  }
  if (n_notes * n_segs > 0x7fffffff) return ISI_ERR_SHAPE;
  // the slab by TMA: full batches only (vec_in), whole boxes (2 M rows is a multiple of 256), and
  // row coordinates that fit the instruction's 32-bit signed operands
  CUtensorMap spec_map;
  memset(&spec_map, 0, sizeof(spec_map));
  const int use_tma = vec_in && (2 * (NFFT / 2)) % kTmaBoxRows == 0 && n_notes * 2 * (NFFT / 2) < 0x7fffffff &&
                      make_spec_map(&spec_map, spec, n_notes, NFFT / 2, p.n_frames, FB);
  imelif_kernel<NFFT, FB, NT, MEL><<<(unsigned)(n_notes * n_segs), NT, L.total, stream>>>(
      spec, p, audio, n_samples, vec_in, vec_out, seg_frames, n_segs, spec_map, use_tma);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_imelif(const float* spec, int64_t n_notes, const isi_imelif_params& p, float* audio,
                  int64_t n_samples, int seg_frames_override, cudaStream_t stream) {
  if (p.use_mel && p.band_width > kMaxMelWidth) return ISI_ERR_UNSUPPORTED;
#define ISI_IMELIF_CASE(N, FB, NT)                                                                          \
  case N:                                                                                                  \
    return p.use_mel ? launch_imelif_t<N, FB, NT, true>(spec, n_notes, p, audio, n_samples, seg_frames_override, stream) \
                     : launch_imelif_t<N, FB, NT, false>(spec, n_notes, p, audio, n_samples, seg_frames_override, stream);
  switch (p.n_fft) {
    ISI_IMELIF_CASE(2048, 4, 256)
    ISI_IMELIF_CASE(1024, 4, 128)
    ISI_IMELIF_CASE(512, 4, 64)
    default: return ISI_ERR_UNSUPPORTED;
  }
#undef ISI_IMELIF_CASE
}

}  // namespace isi
