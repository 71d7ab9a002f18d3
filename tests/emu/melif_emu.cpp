// CPU emulation of melif_kernel's phase sequence (TEST INFRASTRUCTURE ONLY).
// Compiles csrc/melif_core.cuh with g++ and runs the per-thread phases in the same
// order as the CUDA kernel (segments, look-back transform, batches of FB frames), one
// "thread" after another between barriers, so the index arithmetic of the device code
// can be checked against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <type_traits>
#include <vector>

#include "melif_core.cuh"

using namespace isi::melif;

// W32: the warp-specialised kernel's one-warp transform (PlanW32: 32 lanes x 32 points, two
// radix-32 passes) instead of the generic 16 x 16 x 4 plan; polar and emit are shared.
template <int NFFT, int FB, int NT, bool MEL, bool W32 = false>
static void emulate(const float* audio, int64_t n_notes, int64_t n_samples, int hop, int pad_left,
                    int n_frames, int drop_dc, int use_mel, int mel_width, float eps,
                    const float* window, const float* twiddle, const int32_t* mel_start,
                    const int32_t* mel_count, const float* mel_weight, float* out, int seg_frames,
                    int mask_phase, float mask_threshold, const float* affine) {
  using P = typename std::conditional<W32, PlanW32, Plan<NFFT>>::type;
  static_assert(!W32 || NFFT == 2048, "the one-warp plan is for n_fft 2048");
  constexpr int M = P::M, NP = FB / 2, IPT = (M / 2) / NT, RPT = M / NT, kGroups = NT / 64;
  const cpx* tw = reinterpret_cast<const cpx*>(twiddle);
  std::vector<cpx> twm(M);
  for (int i = 0; i < M; ++i) twm[i] = tw[W32 ? fft32_table_source(i) : fft_table_source<Plan<NFFT>>(i)];
  std::vector<float> win(NFFT);
  for (int i = 0; i < NFFT; ++i) win[i] = window[i] * 0.5f;   // the kernel's table: untangle's 1/2 folded in
  const bool aligned = (hop % 2) == 0;
  const int dc = drop_dc ? 1 : 0;
  std::vector<float> stage((FB - 1) * hop + NFFT, 0.f);
  std::vector<cpx2> zA((size_t)NP * P::kPitchA);
  for (auto& v : zA) { v.re = bc(0.f); v.im = bc(0.f); }
  const int n_segs = (n_frames + seg_frames - 1) / seg_frames;

  // the FFT passes of the batch staged at `frame`: pairs q = slots (q, q + NP)
  auto transform = [&](const float* note, int frame, int nf, bool lookback) {
    const int span = (nf - 1) * hop + NFFT;
    for (int tid = 0; tid < NT; ++tid)
      stage_fill(tid, NT, stage.data(), span, note, n_samples, (int64_t)frame * hop - pad_left);
    auto active = [&](int q) { return lookback ? (q == NP - 1) : (q < nf); };
    if constexpr (W32) {
      std::vector<PassB32Regs> regs32((size_t)NP * 32);
      for (int q = 0; q < NP; ++q) {
        if (!active(q)) continue;
        cpx2* z = zA.data() + q * P::kPitchA;
        for (int j = 0; j < 32; ++j)
          fft32_passA(j, stage.data() + (lookback ? 0 : q * hop), stage.data() + (lookback ? 0 : (q + NP) * hop),
                      aligned, 1.f, win.data(), twm.data(), z);
        for (int l = 0; l < 32; ++l) fft32_passB_load(l, z, regs32[q * 32 + l]);
        for (int l = 0; l < 32; ++l) fft32_passB_store(l, regs32[q * 32 + l], z);
      }
      return;
    }
    if constexpr (!W32) {
      using PG = Plan<NFFT>;
      for (int tid = 0; tid < NT; ++tid)
        for (int q = tid / 64; q < NP; q += kGroups)
          if (active(q))
            fft_pass1_pair<PG>(tid & 63, stage.data() + (lookback ? 0 : q * hop),
                               stage.data() + (lookback ? 0 : (q + NP) * hop), aligned, 1.f, win.data(),
                               twm.data(), zA.data() + q * P::kPitchA);
      for (int tid = 0; tid < NT; ++tid)
        for (int q = tid / 64; q < NP; q += kGroups)
          if (active(q)) fft_pass2<PG>(tid & 63, twm.data(), zA.data() + q * P::kPitchA);
      std::vector<Pass3Regs<PG, cpx2>> regs((size_t)NT * NP);
      for (int tid = 0; tid < NT; ++tid)
        for (int q = tid / 64; q < NP; q += kGroups)
          if (active(q)) fft_pass3_load<PG>(tid & 63, zA.data() + q * P::kPitchA, regs[tid * NP + q]);
      for (int tid = 0; tid < NT; ++tid)
        for (int q = tid / 64; q < NP; q += kGroups)
          if (active(q)) fft_pass3_store<PG>(tid & 63, regs[tid * NP + q], zA.data() + q * P::kPitchA);
    }
  };
  auto polar = [&](std::vector<BinState>& st, bool seed_only) {
    for (int tid = 0; tid < NT; ++tid)
      for (int i = 0; i < IPT; ++i)
        polar_item<P, MEL, NP, true>(tid + i * NT, zA.data(), P::kPitchA, tw[tid + i * NT], dc ? M : 0,
                                     seed_only, eps, st[tid * IPT + i]);
  };

  for (int64_t n = 0; n < n_notes; ++n)
    for (int seg = 0; seg < n_segs; ++seg) {
      const float* note = audio + n * n_samples;
      float* out0 = out + n * 2 * M * n_frames;
      float* out1 = out0 + (int64_t)M * n_frames;
      const int fs = seg * seg_frames, fe = std::min(n_frames, fs + seg_frames);
      std::vector<BinState> st((size_t)NT * IPT, bin_state_init());
      if (fs > 0) {
        transform(note, fs - 1, 1, true);
        polar(st, true);
      }
      for (int f0 = fs; f0 < fe; f0 += FB) {
        const int nf = std::min(FB, fe - f0);
        transform(note, f0, nf, false);
        polar(st, false);
        for (int tid = 0; tid < NT; ++tid)
          for (int r = 0; r < RPT; ++r) {
            const int row = tid + r * NT;
            float w[kMaxMelWidth] = {0};
            int bin0 = row + dc;
            if (use_mel) {
              bin0 = mel_start[row] + dc;
              for (int i = 0; i < mel_width && i < kMaxMelWidth; ++i) w[i] = mel_weight[(int64_t)row * mel_width + i];
            }
            f2 lg[NP], ph[NP];
            if (MEL) emit_mel<NP>(zA.data(), P::kPitchA, bin0, kMaxMelWidth, w, f0 == 0, eps, lg, ph);
            else     emit_linear<NP>(zA.data(), P::kPitchA, bin0, lg, ph);
            float v0[FB], v1[FB];
            finish_row<NP>(lg, ph, mask_phase != 0, mask_threshold, affine[0], affine[1], affine[2], affine[3], v0, v1);
            for (int s = 0; s < nf; ++s) {
              out0[(int64_t)row * n_frames + f0 + s] = v0[s];
              out1[(int64_t)row * n_frames + f0 + s] = v1[s];
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
  if (seg_frames <= 0) seg_frames = (n_frames + 7) / 8 * 8;
  if (n_fft == -2048) {      // the warp-specialised kernel's one-warp transform plan
    if (use_mel) emulate<2048, 8, 256, true, true>(ARGS); else emulate<2048, 8, 256, false, true>(ARGS);
    return 0;
  }
#define CASE(N, FB, NT) case N: if (use_mel) emulate<N, FB, NT, true>(ARGS); else emulate<N, FB, NT, false>(ARGS); return 0;
  switch (n_fft) {
    CASE(2048, 8, 256)
    CASE(1024, 4, 128)
    CASE(512, 4, 64)
    default: return -3;
  }
}
