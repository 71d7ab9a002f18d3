// CPU emulation of imelif_kernel's phase sequence (TEST INFRASTRUCTURE ONLY).
// Compiles csrc/imelif_core.cuh with g++ and runs the per-thread phases in the order of the
// CUDA kernel (segments, FP64 look-back, batches of FB frames, ping-pong carry), one
// "thread" after another between barriers, so the index arithmetic of the device code can
// be checked against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "imelif_core.cuh"

using namespace isi::imelif;

template <int NFFT, int FB, int NT, bool MEL>
static void emulate(const float* spec, int64_t n_notes, int hop, int pad_left, int n_frames, int drop_dc,
                    int band_width, float eps, const float* window, const float* twiddle,
                    const int32_t* band_start, const int32_t* band_count, const float* band_weight,
                    const float* ola_scale, const float* affine, float* audio, int64_t n_samples,
                    int seg_frames, int vec_out) {
  using P = Plan<NFFT>;
  constexpr int M = P::M, IPT = (M / 2) / NT, RPT = M / NT, CPT = 2 * M / NT, kGroups = NT / 64;
  const cpx* tw = reinterpret_cast<const cpx*>(twiddle);
  std::vector<cpx> twm(M);
  for (int i = 0; i < M; ++i) twm[i] = tw[fft_table_source<P>(i)];
  const int dc = drop_dc ? 1 : 0, real_row = dc ? M - 1 : 0;
  const int n_segs = (n_frames + seg_frames - 1) / seg_frames;
  std::vector<float> slab(2 * M * FB);
  std::vector<cpx> zA((size_t)FB * P::kPitchA);
  std::vector<float> carry(2 * NFFT);
  std::vector<double> sums(M);

  for (int64_t n = 0; n < n_notes; ++n)
    for (int seg = 0; seg < n_segs; ++seg) {
      const int fs = seg * seg_frames, fe = std::min(n_frames, fs + seg_frames);
      const bool last_seg = fe == n_frames;
      int fstart = fs - ola_lookback_frames(NFFT, hop);
      fstart = fstart <= 0 ? 0 : fstart / FB * FB;
      const int64_t emit_from = (int64_t)fs * hop;
      const float* note0 = spec + n * 2 * M * n_frames;
      const float* note1 = note0 + (int64_t)M * n_frames;
      float* out = audio + n * n_samples;
      std::fill(carry.begin(), carry.end(), 0.f);
      std::vector<float> phase((size_t)NT * RPT, 0.f);
      std::vector<float> bw((size_t)M * kMaxMelWidth, 0.f);
      std::vector<int> bs(M), bc(M, 0);
      for (int row = 0; row < M; ++row) {
        bs[row] = row;
        if (MEL) {
          bs[row] = band_start[row]; bc[row] = band_count[row];
          for (int i = 0; i < band_width && i < kMaxMelWidth; ++i) bw[row * kMaxMelWidth + i] = band_weight[(int64_t)row * band_width + i];
        }
      }
      if (fstart > 0) {
        for (int tid = 0; tid < NT; ++tid) {
          double row_sum[RPT];
          lookback_rows_sum<RPT>(note1, n_frames, tid, NT, fstart, affine[2], affine[3], n_frames % 4 == 0, row_sum);
          for (int r = 0; r < RPT; ++r) sums[tid + r * NT] = row_sum[r];
        }
        for (int tid = 0; tid < NT; ++tid)
          for (int r = 0; r < RPT; ++r) {
            const int row = tid + r * NT;
            phase[tid * RPT + r] = lookback_phase<MEL>(sums.data(), bs[row], bc[row], &bw[row * kMaxMelWidth]);
          }
      }
      int flip = 0;
      for (int f0 = fstart; f0 < fe; f0 += FB) {
        const int nf = std::min(FB, fe - f0);
        const bool last_batch = f0 + FB >= fe;
        float* carry_in = carry.data() + flip * NFFT;
        float* carry_out = carry.data() + (flip ^ 1) * NFFT;
        flip ^= 1;
        for (int tid = 0; tid < NT; ++tid)
          for (int i = 0; i < CPT; ++i) {
            const int q = tid + i * NT;
            const float* rows = q < M ? note0 + (int64_t)q * n_frames : note1 + (int64_t)(q - M) * n_frames;
            slab_fill_chunk<FB>(slab.data(), q, rows, f0, nf);
            slab_transform_chunk<FB>(slab.data(), q, M, affine[0], affine[1], affine[2], affine[3]);
          }
        for (int tid = 0; tid < NT; ++tid)
          for (int r = 0; r < RPT; ++r) {
            const int row = tid + r * NT;
            build_row<FB, MEL>(slab.data(), M, bs[row], bc[row], kMaxMelWidth, &bw[row * kMaxMelWidth], eps,
                               row == real_row, phase[tid * RPT + r], zA.data() + row + dc, P::kPitchA);
          }
        for (int fb = 0; fb < FB; ++fb) zA[fb * P::kPitchA + (dc ? 0 : M)] = cpx{0.f, 0.f};
        for (int tid = 0; tid < NT; ++tid)
          for (int fb = 0; fb < nf; ++fb)
            for (int i = 0; i < IPT; ++i) tangle_item<P>(tid + i * NT, zA.data() + fb * P::kPitchA, tw[tid + i * NT]);
        // transform, one frame at a time, barriers between the sweeps over the 64 threads
        for (int fb = 0; fb < nf; ++fb) {
          cpx* z = zA.data() + fb * P::kPitchA;
          std::vector<cpx> v((size_t)64 * P::R1);
          for (int j = 0; j < 64; ++j) ifft_pass1_load<P>(j, z, &v[j * P::R1]);
          for (int j = 0; j < 64; ++j) ifft_pass1_store<P>(j, &v[j * P::R1], twm.data(), z);
          for (int j = 0; j < 64; ++j) fft_pass2<P>(j, twm.data(), z);
          std::vector<Pass3Regs<P, cpx>> regs(64);
          for (int j = 0; j < 64; ++j) fft_pass3_load<P>(j, z, regs[j]);
          for (int j = 0; j < 64; ++j) fft_pass3_store<P>(j, regs[j], z);
        }
        (void)kGroups;
        const int span = (nf - 1) * hop + NFFT;
        const int emit_len = (last_batch && last_seg) ? span : nf * hop;
        const int64_t pos0 = (int64_t)f0 * hop;
        if (vec_out) {
          for (int tid = 0; tid < NT; ++tid)
            for (int s0 = 4 * tid; s0 < span; s0 += 4 * NT) {
              float acc[4] = {0.f, 0.f, 0.f, 0.f};
              if (s0 < NFFT) for (int i = 0; i < 4; ++i) acc[i] = carry_in[s0 + i];
              ola_quad<FB>(zA.data(), P::kPitchA, window, NFFT, hop, nf, s0, acc);
              if (s0 < emit_len) {
                const int64_t pos = pos0 + s0, j = pos - pad_left;
                if (pos >= emit_from && j >= 0 && j < n_samples)
                  for (int i = 0; i < 4; ++i) out[j + i] = acc[i] * ola_scale[pos + i];
              } else {
                for (int i = 0; i < 4; ++i) carry_out[s0 - emit_len + i] = acc[i];
              }
            }
        } else {
          for (int tid = 0; tid < NT; ++tid)
            for (int s = tid; s < span; s += NT) {
              const float acc = (s < NFFT ? carry_in[s] : 0.f) + ola_sample(zA.data(), P::kPitchA, window, NFFT, hop, nf, s);
              if (s < emit_len) {
                const int64_t pos = pos0 + s, j = pos - pad_left;
                if (pos >= emit_from && j >= 0 && j < n_samples) out[j] = acc * ola_scale[pos];
              } else {
                carry_out[s - emit_len] = acc;
              }
            }
        }
        for (int s = span - emit_len; s < NFFT; ++s) carry_out[s] = 0.f;
      }
    }
}

extern "C" int imelif_emulate(const float* spec, int64_t n_notes, int n_fft, int hop, int pad_left,
                              int n_frames, int drop_dc, int use_mel, int band_width, float eps,
                              const float* window, const float* twiddle, const int32_t* band_start,
                              const int32_t* band_count, const float* band_weight, const float* ola_scale,
                              const float* affine, float* audio, int64_t n_samples, int seg_frames,
                              int vec_out) {
#define ARGS spec, n_notes, hop, pad_left, n_frames, drop_dc, band_width, eps, window, twiddle, band_start, \
             band_count, band_weight, ola_scale, affine, audio, n_samples, seg_frames, vec_out
  if (seg_frames <= 0) seg_frames = (n_frames + 3) / 4 * 4;
  seg_frames = (seg_frames + 3) / 4 * 4;
  if (vec_out && (hop % 4 || pad_left % 4 || n_samples % 4)) return -5;
#define CASE(N, FB, NT) case N: if (use_mel) emulate<N, FB, NT, true>(ARGS); else emulate<N, FB, NT, false>(ARGS); return 0;
  switch (n_fft) {
    CASE(2048, 4, 256)
    CASE(1024, 4, 128)
    CASE(512, 4, 64)
    default: return -3;
  }
}
