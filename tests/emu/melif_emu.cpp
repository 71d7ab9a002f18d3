// CPU emulation of melif_kernel's phase sequence (TEST INFRASTRUCTURE ONLY).
// Compiles csrc/melif_core.cuh with g++ and runs the per-thread phases in the same
// order as the CUDA kernel, one "thread" after another between barriers, so the index
// arithmetic of the device code can be checked against the oracle without a GPU.
#include <cstdint>
#include <vector>

#include "melif_core.cuh"

using namespace isi::melif;

template <int NFFT, int FB, int NT>
static void emulate(const float* audio, int64_t n_notes, int64_t n_samples, int hop, int pad_left,
                    int n_frames, int drop_dc, int use_mel, int mel_width, float eps,
                    const float* window, const float* twiddle, const int32_t* mel_start,
                    const int32_t* mel_count, const float* mel_weight, float* out) {
  using P = Plan<NFFT>;
  constexpr int M = P::M, RPT = M / NT, kGroups = NT / 64;
  static_assert(NT == M / 2, "one polar item per thread");
  const cpx* tw = reinterpret_cast<const cpx*>(twiddle);
  const int span = (FB - 1) * hop + NFFT;
  const int dc = drop_dc ? 1 : 0;
  std::vector<float> stage(span);
  std::vector<cpx> zA((size_t)FB * P::kPitchA), zB((size_t)FB * P::kPitchB);
  for (int64_t n = 0; n < n_notes; ++n) {
    const float* note = audio + n * n_samples;
    float* out0 = out + n * 2 * M * n_frames;
    float* out1 = out0 + (int64_t)M * n_frames;
    std::vector<BinState> sa(NT, BinState{1.f, 0.f, 0.f}), sb = sa, sc = sa;
    std::vector<float> prev((size_t)NT * RPT, 0.f);
    for (int f0 = 0; f0 < n_frames; f0 += FB) {
      const int nf = (FB < n_frames - f0) ? FB : n_frames - f0;
      for (int tid = 0; tid < NT; ++tid)
        stage_fill(tid, NT, stage.data(), span, note, n_samples, (int64_t)f0 * hop - pad_left);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = tid / 64; fb < nf; fb += kGroups)
          fft_pass1<P>(tid & 63, stage.data() + fb * hop, window, tw, zA.data() + fb * P::kPitchA);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = tid / 64; fb < nf; fb += kGroups)
          fft_pass2<P>(tid & 63, tw, zA.data() + fb * P::kPitchA);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = tid / 64; fb < nf; fb += kGroups)
          fft_pass3<P>(tid & 63, zA.data() + fb * P::kPitchA, zB.data() + fb * P::kPitchB);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = 0; fb < nf; ++fb)
          polar_item<P>(tid, zB.data() + fb * P::kPitchB, tw[tid], f0 + fb == 0, use_mel != 0, eps,
                        sa[tid], sb[tid], sc[tid]);
      for (int tid = 0; tid < NT; ++tid)
        for (int r = 0; r < RPT; ++r) {
          const int row = tid + r * NT;
          float w[kMaxMelWidth] = {0};
          int bin0 = row + dc, cnt = 0;
          if (use_mel) {
            bin0 = mel_start[row] + dc; cnt = mel_count[row];
            for (int i = 0; i < mel_width && i < kMaxMelWidth; ++i) w[i] = mel_weight[(int64_t)row * mel_width + i];
          }
          for (int fb = 0; fb < nf; ++fb) {
            float v0, v1;
            if (use_mel)
              emit_mel(zB.data() + fb * P::kPitchB, bin0, cnt, w, f0 + fb == 0, eps,
                       prev[tid * RPT + r], v0, v1);
            else
              emit_linear(zB.data() + fb * P::kPitchB, bin0, v0, v1);
            out0[(int64_t)row * n_frames + f0 + fb] = v0;
            out1[(int64_t)row * n_frames + f0 + fb] = v1;
          }
        }
    }
  }
}

extern "C" int melif_emulate(const float* audio, int64_t n_notes, int64_t n_samples, int n_fft,
                             int hop, int pad_left, int n_frames, int drop_dc, int use_mel,
                             int mel_width, float eps, const float* window, const float* twiddle,
                             const int32_t* mel_start, const int32_t* mel_count,
                             const float* mel_weight, float* out) {
#define ARGS audio, n_notes, n_samples, hop, pad_left, n_frames, drop_dc, use_mel, mel_width, eps, \
             window, twiddle, mel_start, mel_count, mel_weight, out
  switch (n_fft) {
    case 2048: emulate<2048, 8, 512>(ARGS); return 0;
    case 1024: emulate<1024, 8, 256>(ARGS); return 0;
    case 512:  emulate<512, 8, 128>(ARGS); return 0;
    default: return -3;
  }
}
