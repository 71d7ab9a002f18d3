// Fused front end: STFT -> (mel projection) -> log-magnitude + instantaneous frequency.
//
// Replaces SpectrogramsHelper / MelSpectrogramsHelper.to_spectrogram (external
// GANsynth_pytorch; reference call sites utils/misc.py:10-29, extract_code.py:199-206,
// train_vqvae.py:604-611).  One CTA owns one note and walks its frames in batches of
// FB: the audio samples are read once per overlapping frame (L1/L2 hits), the complex
// spectrum, magnitudes and phases never leave shared memory / registers, and the only
// HBM write is the final [2, F, T'] tensor, FB consecutive time steps per row at a time.
#include "common.cuh"
#include "melif_core.cuh"

namespace isi {
using namespace melif;

template <int NFFT, int FB, int NT>
__global__ void __launch_bounds__(NT)
melif_kernel(const float* __restrict__ audio, int64_t n_samples, isi_melif_params p,
             float* __restrict__ out) {
  using P = Plan<NFFT>;
  constexpr int M = P::M;
  constexpr int IPT = (M / 2) / NT;   // polar work items per thread
  constexpr int RPT = M / NT;         // output rows per thread
  static_assert(IPT >= 1 && (M / 2) % NT == 0, "thread count must divide the item count");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* zbuf = reinterpret_cast<cpx*>(smem_raw);   // [FB][M]

  const int tid = threadIdx.x;
  const float* note = audio + (int64_t)blockIdx.x * n_samples;
  float* out0 = out + (int64_t)blockIdx.x * 2 * M * p.n_frames;
  float* out1 = out0 + (int64_t)M * p.n_frames;
  const cpx* tw = reinterpret_cast<const cpx*>(p.twiddle);
  const bool use_mel = p.use_mel != 0, drop_dc = p.drop_dc != 0;

  BinState sa[IPT], sb[IPT];
  RowState rs[RPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) { sa[i] = BinState{0.f, 0.f}; sb[i] = BinState{0.f, 0.f}; }
#pragma unroll
  for (int r = 0; r < RPT; ++r) rs[r] = RowState{0.f};

  for (int f0 = 0; f0 < p.n_frames; f0 += FB) {
    const int nf = min(FB, p.n_frames - f0);
    // A: window + pack
    for (int fb = 0; fb < nf; ++fb)
      pack_frame<P>(tid, NT, zbuf + fb * M, note, n_samples,
                    (int64_t)(f0 + fb) * p.hop - p.pad_left, p.window);
    __syncthreads();
    // B: in-place FFT, 64 threads per frame
    for (int fb = tid / 64; fb < nf; fb += NT / 64) fft_pass1<P>(tid & 63, zbuf + fb * M, tw);
    __syncthreads();
    for (int fb = tid / 64; fb < nf; fb += NT / 64) fft_pass2<P>(tid & 63, zbuf + fb * M, tw);
    __syncthreads();
    for (int fb = tid / 64; fb < nf; fb += NT / 64) fft_pass3<P>(tid & 63, zbuf + fb * M);
    __syncthreads();
    // C: untangle + polar + time unwrap (state in registers, frames in order)
    for (int fb = 0; fb < nf; ++fb) {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
        polar_item<P>(tid + i * NT, zbuf + fb * M, tw, f0 + fb == 0, use_mel, drop_dc,
                      p.safelog_eps, sa[i], sb[i]);
    }
    __syncthreads();
    // D: project / copy rows and write FB consecutive time steps per row
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int row = tid + r * NT;
      int ms = 0, mc = 0;
      const float* mw = nullptr;
      if (use_mel) { ms = p.mel_start[row]; mc = p.mel_count[row]; mw = p.mel_weight + (int64_t)row * p.mel_width; }
      float v0[FB], v1[FB];
#pragma unroll
      for (int fb = 0; fb < FB; ++fb) {
        v0[fb] = 0.f; v1[fb] = 0.f;
        if (fb < nf)
          emit_row<P>(row, zbuf + fb * M, f0 + fb == 0, use_mel, drop_dc, p.safelog_eps, ms, mc,
                      mw, rs[r], v0[fb], v1[fb]);
      }
      float* d0 = out0 + (int64_t)row * p.n_frames + f0;
      float* d1 = out1 + (int64_t)row * p.n_frames + f0;
      if (nf == FB && (FB % 4 == 0) && (p.n_frames % 4 == 0)) {
#pragma unroll
        for (int q = 0; q < FB / 4; ++q) {
          reinterpret_cast<float4*>(d0)[q] = make_float4(v0[4 * q], v0[4 * q + 1], v0[4 * q + 2], v0[4 * q + 3]);
          reinterpret_cast<float4*>(d1)[q] = make_float4(v1[4 * q], v1[4 * q + 1], v1[4 * q + 2], v1[4 * q + 3]);
        }
      } else {
#pragma unroll
        for (int fb = 0; fb < FB; ++fb)
          if (fb < nf) { d0[fb] = v0[fb]; d1[fb] = v1[fb]; }
      }
    }
    __syncthreads();
  }
}

template <int NFFT, int FB, int NT>
static int launch_melif_t(const float* audio, int64_t n_notes, int64_t n_samples,
                          const isi_melif_params& p, float* out, cudaStream_t stream) {
  size_t smem = (size_t)FB * (NFFT / 2) * sizeof(cpx);
  cudaError_t e = cudaFuncSetAttribute(melif_kernel<NFFT, FB, NT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  melif_kernel<NFFT, FB, NT><<<(unsigned)n_notes, NT, smem, stream>>>(audio, n_samples, p, out);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_melif(const float* audio, int64_t n_notes, int64_t n_samples,
                 const isi_melif_params& p, float* out, cudaStream_t stream) {
  if (n_notes > 0x7fffffff) return ISI_ERR_SHAPE;
  switch (p.n_fft) {
    case 2048: return launch_melif_t<2048, 8, 512>(audio, n_notes, n_samples, p, out, stream);
    case 1024: return launch_melif_t<1024, 8, 256>(audio, n_notes, n_samples, p, out, stream);
    case 512:  return launch_melif_t<512, 8, 128>(audio, n_notes, n_samples, p, out, stream);
    default:   return ISI_ERR_UNSUPPORTED;
  }
}

}  // namespace isi
