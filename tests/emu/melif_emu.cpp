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
  constexpr int M = P::M, IPT = (M / 2) / NT, RPT = M / NT;
  const cpx* tw = reinterpret_cast<const cpx*>(twiddle);
  std::vector<cpx> zbuf((size_t)FB * M);
  for (int64_t n = 0; n < n_notes; ++n) {
    const float* note = audio + n * n_samples;
    float* out0 = out + n * 2 * M * n_frames;
    float* out1 = out0 + (int64_t)M * n_frames;
    std::vector<BinState> sa((size_t)NT * IPT, BinState{0.f, 0.f}), sb = sa;
    std::vector<RowState> rs((size_t)NT * RPT, RowState{0.f});
    for (int f0 = 0; f0 < n_frames; f0 += FB) {
      const int nf = (FB < n_frames - f0) ? FB : n_frames - f0;
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = 0; fb < nf; ++fb)
          pack_frame<P>(tid, NT, zbuf.data() + fb * M, note, n_samples,
                        (int64_t)(f0 + fb) * hop - pad_left, window);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = tid / 64; fb < nf; fb += NT / 64) fft_pass1<P>(tid & 63, zbuf.data() + fb * M, tw);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = tid / 64; fb < nf; fb += NT / 64) fft_pass2<P>(tid & 63, zbuf.data() + fb * M, tw);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = tid / 64; fb < nf; fb += NT / 64) fft_pass3<P>(tid & 63, zbuf.data() + fb * M);
      for (int tid = 0; tid < NT; ++tid)
        for (int fb = 0; fb < nf; ++fb)
          for (int i = 0; i < IPT; ++i)
            polar_item<P>(tid + i * NT, zbuf.data() + fb * M, tw, f0 + fb == 0, use_mel != 0,
                          drop_dc != 0, eps, sa[tid * IPT + i], sb[tid * IPT + i]);
      for (int tid = 0; tid < NT; ++tid)
        for (int r = 0; r < RPT; ++r) {
          const int row = tid + r * NT;
          int ms = 0, mc = 0;
          const float* mw = nullptr;
          if (use_mel) { ms = mel_start[row]; mc = mel_count[row]; mw = mel_weight + (int64_t)row * mel_width; }
          for (int fb = 0; fb < nf; ++fb) {
            float v0, v1;
            emit_row<P>(row, zbuf.data() + fb * M, f0 + fb == 0, use_mel != 0, drop_dc != 0, eps, ms,
                        mc, mw, rs[tid * RPT + r], v0, v1);
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
