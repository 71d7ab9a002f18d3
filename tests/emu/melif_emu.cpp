// CPU emulation of melif_kernel's phase sequence (TEST INFRASTRUCTURE ONLY).
// Compiles csrc/melif_core.cuh with g++ and runs the per-thread phases in the same
// order as the CUDA kernel (segments, look-back transform, batches of FB frames), one
// "thread" after another between barriers, so the index arithmetic of the device code
// can be checked against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "melif_core.cuh"

using namespace isi::melif;

template <int NFFT, int FB, int NT, bool MEL>
static void emulate(const float* audio, int64_t n_notes, int64_t n_samples, int hop, int pad_left,
                    int n_frames, int drop_dc, int use_mel, int mel_width, float eps,
                    const float* window, const float* twiddle, const int32_t* mel_start,
                    const int32_t* mel_count, const float* mel_weight, float* out, int seg_frames,
                    int mask_phase, float mask_threshold, const float* affine) {
  using P = Plan<NFFT>;
  constexpr int M = P::M, IPT = (M / 2) / NT, RPT = M / NT, kGroups = NT / 64;
  const cpx* tw = reinterpret_cast<const cpx*>(twiddle);
  std::vector<cpx> twm(M);
  for (int i = 0; i < M; ++i) twm[i] = tw[fft_table_source<P>(i)];
  const bool aligned8 = (hop % 2) == 0;
  const int dc = drop_dc ? 1 : 0;
  std::vector<float> stage((FB - 1) * hop + NFFT);
  std::vector<cpx> zA((size_t)FB * P::kPitchA);
  const int n_segs = (n_frames + seg_frames - 1) / seg_frames;

  auto transform = [&](const float* note, int frame, int nf) {
    const int span = (nf - 1) * hop + NFFT;
    for (int tid = 0; tid < NT; ++tid)
      stage_fill(tid, NT, stage.data(), span, note, n_samples, (int64_t)frame * hop - pad_left);
    for (int tid = 0; tid < NT; ++tid)
      for (int fb = tid / 64; fb < nf; fb += kGroups)
        fft_pass1<P>(tid & 63, stage.data() + fb * hop, aligned8, 1.f, window, twm.data(), zA.data() + fb * P::kPitchA);
    for (int tid = 0; tid < NT; ++tid)
      for (int fb = tid / 64; fb < nf; fb += kGroups)
        fft_pass2<P>(tid & 63, twm.data(), zA.data() + fb * P::kPitchA);
    std::vector<Pass3Regs<P>> regs((size_t)NT * FB);
    for (int tid = 0; tid < NT; ++tid)
      for (int fb = tid / 64; fb < nf; fb += kGroups)
        fft_pass3_load<P>(tid & 63, zA.data() + fb * P::kPitchA, regs[tid * FB + fb]);
    for (int tid = 0; tid < NT; ++tid)
      for (int fb = tid / 64; fb < nf; fb += kGroups)
        fft_pass3_store<P>(tid & 63, regs[tid * FB + fb], zA.data() + fb * P::kPitchA);
  };

  for (int64_t n = 0; n < n_notes; ++n)
    for (int seg = 0; seg < n_segs; ++seg) {
      const float* note = audio + n * n_samples;
      float* out0 = out + n * 2 * M * n_frames;
      float* out1 = out0 + (int64_t)M * n_frames;
      const int fs = seg * seg_frames, fe = std::min(n_frames, fs + seg_frames);
      std::vector<BinState> sa((size_t)NT * IPT, BinState{1.f, 0.f}), sb = sa;
      if (fs > 0) {
        transform(note, fs - 1, 1);
        for (int tid = 0; tid < NT; ++tid)
          for (int i = 0; i < IPT; ++i)
            polar_item<P, MEL>(tid + i * NT, zA.data(), tw[tid + i * NT], dc ? M : 0, true, eps,
                               sa[tid * IPT + i], sb[tid * IPT + i]);
      }
      for (int f0 = fs; f0 < fe; f0 += FB) {
        const int nf = std::min(FB, fe - f0);
        transform(note, f0, nf);
        for (int tid = 0; tid < NT; ++tid)
          for (int fb = 0; fb < nf; ++fb)
            for (int i = 0; i < IPT; ++i)
              polar_item<P, MEL>(tid + i * NT, zA.data() + fb * P::kPitchA, tw[tid + i * NT], dc ? M : 0,
                                 f0 + fb == 0, eps, sa[tid * IPT + i], sb[tid * IPT + i]);
        for (int tid = 0; tid < NT; ++tid)
          for (int r = 0; r < RPT; ++r) {
            const int row = tid + r * NT;
            float w[kMaxMelWidth] = {0};
            int bin0 = row + dc, cnt = 0;
            if (use_mel) {
              bin0 = mel_start[row] + dc; cnt = mel_count[row];
              for (int i = 0; i < mel_width && i < kMaxMelWidth; ++i) w[i] = mel_weight[(int64_t)row * mel_width + i];
            }
            float v0[FB], v1[FB];
            if (MEL) emit_mel<FB>(zA.data(), P::kPitchA, bin0, cnt, kMaxMelWidth, w, f0 == 0, eps, v0, v1);
            else     emit_linear<FB>(zA.data(), P::kPitchA, bin0, v0, v1);
            apply_epilogue<FB>(v0, v1, mask_phase != 0, mask_threshold, affine[0], affine[1], affine[2], affine[3]);
            for (int fb = 0; fb < nf; ++fb) {
              out0[(int64_t)row * n_frames + f0 + fb] = v0[fb];
              out1[(int64_t)row * n_frames + f0 + fb] = v1[fb];
            }
          }
      }
    }
}

extern "C" int melif_emulate(const float* audio, int64_t n_notes, int64_t n_samples, int n_fft,
                             int hop, int pad_left, int n_frames, int drop_dc, int use_mel,
                             int mel_width, float eps, const float* window, const float* twiddle,
                             const int32_t* mel_start, const int32_t* mel_count,
                             const float* mel_weight, float* out, int seg_frames, int mask_phase,
                             float mask_threshold, const float* affine) {
#define ARGS audio, n_notes, n_samples, hop, pad_left, n_frames, drop_dc, use_mel, mel_width, eps, \
             window, twiddle, mel_start, mel_count, mel_weight, out, seg_frames, mask_phase, mask_threshold, affine
  if (seg_frames <= 0) seg_frames = (n_frames + 3) / 4 * 4;
#define CASE(N, FB, NT) case N: if (use_mel) emulate<N, FB, NT, true>(ARGS); else emulate<N, FB, NT, false>(ARGS); return 0;
  switch (n_fft) {
    CASE(2048, 8, 512)
    CASE(1024, 4, 128)
    CASE(512, 4, 64)
    default: return -3;
  }
}
