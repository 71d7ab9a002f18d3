// Per-thread phases of the inverse front end: (mel) log-magnitude + IF -> audio.
//
// Mirror of melif_core.cuh, and written the same way: against an abstract "thread id +
// shared buffers", so nvcc compiles it into imelif.cu and g++ into the CPU emulation of
// tests/test_imelif_emulation.py (test infrastructure, not a product path).
//
// Plan for a batch of FB frames of one note (M = n_fft/2):
//   slab     the batch's [2][M][FB] input values: channel 0 -> exp(.) (mel: squared magnitude,
//            linear: magnitude), channel 1 -> phase advance per frame in half-turns (IF)
//   build    per linear row: banded mel->linear projection of both, running phase (wrapped,
//            kept in a register across batches), X = mag (cos, sin)(pi phase) into the frame's
//            natural-order spectrum z[0..M]
//   tangle   X[k], X[M-k] -> conj Z[k], conj Z[M-k] in place, Z the M-point spectrum of
//            z[n] = x[2n] + i x[2n+1] (scaled by 2; folded into the overlap-add scale)
//   fft      the forward transform's three passes (pass 1 reads the natural-order buffer in
//            place): Y = FFT_M(conj Z), so x[2n] = Re Y[n], x[2n+1] = -Im Y[n] (times 1/n_fft)
//   ola      per output sample: sum over the batch's frames of window x frame sample, plus
//            the carry of the previous batch; the first nf*hop samples are complete and go to
//            HBM scaled by 1 / (n_fft x summed squared windows), the rest is the next carry
#pragma once
#include "melif_core.cuh"

namespace isi {
namespace imelif {

using namespace melif;

struct alignas(16) f4 { float x, y, z, w; };   // one LDS.128 / LDG.128

ISI_HD float fast_exp(float x) {            // one MUFU.EX2 + one multiply
#ifdef __CUDA_ARCH__
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.44269504088896340736f));
  return r;
#else
  return expf(x);
#endif
}
ISI_HD float fast_sqrt(float x) {           // one MUFU.SQRT
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
// cos and sin of pi*h for h in [-1, 1] (MUFU.SIN / MUFU.COS: |err| < 5e-7 on that range)
ISI_HD cpx fast_cis_pi(float h) {
#ifdef __CUDA_ARCH__
  float s, c;
  __sincosf(h * kPi, &s, &c);
  return cpx{c, s};
#else
  return cpx{cosf(h * kPi), sinf(h * kPi)};
#endif
}
// half-turns folded into [-1, 1]: exact in FP32 (0.5 h and 2 rint() are exact)
ISI_HD float wrap_half_turns(float h) { return fmaf(-2.f, rintf(0.5f * h), h); }

// ---- slab: chunk q (0 .. 2M-1) is row m = q % M of channel c = q / M, FB time steps ----
// Synchronous fill (emulation; the device path when a chunk is not one aligned 16-byte run).
template <int FB>
ISI_HD void slab_fill_chunk(float* slab, int q, const float* rows /* spec + note offset + channel/row offset */,
                            int f0, int nf) {
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) slab[q * FB + fb] = fb < nf ? rows[f0 + fb] : 0.f;
}

// channel 0: v -> exp(s0 v + b0); channel 1: v -> s1 v + b1 (half-turns per frame)
template <int FB>
ISI_HD void slab_transform_chunk(float* slab, int q, int M, float s0, float b0, float s1, float b1) {
  float* v = slab + q * FB;
  if (q < M) {
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) v[fb] = fast_exp(fmaf(v[fb], s0, b0));
  } else {
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) v[fb] = fmaf(v[fb], s1, b1);
  }
}

// FB consecutive floats of a slab chunk (16-byte aligned when FB = 4)
template <int FB>
ISI_HD void load_chunk(const float* p, float* v) {
  if (FB == 4) {
    const f4 q = *reinterpret_cast<const f4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) v[fb] = p[fb];
  }
}

// ---- build: one linear row, all FB frames.  `phase` (half-turns) is the row's running phase
//      before frame f0 and is advanced.  Mel mode: band = rows start .. start+count of the slab
//      with weights w (zero beyond count); `count_uniform` >= count is warp-uniform.  Linear
//      mode: start = the row itself.  `real_only`: the bin is X[0] or X[n_fft/2]. ----
template <int FB, bool MEL>
ISI_HD void build_row(const float* slab, int M, int start, int count, int count_uniform, const float* w,
                      float eps, bool real_only, float& phase, cpx* zrow /* z + bin */, int pitch) {
  float a[FB], d[FB];
  if (MEL) {
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) { a[fb] = 0.f; d[fb] = 0.f; }
#pragma unroll
    for (int i = 0; i < kMaxMelWidth; ++i) {
      if (i < count_uniform) {
        const bool on = i < count;
        const float* s0 = slab + (start + (on ? i : 0)) * FB;
        const float wi = on ? w[i] : 0.f;
        float v0[FB], v1[FB];
        load_chunk<FB>(s0, v0);
        load_chunk<FB>(s0 + M * FB, v1);
#pragma unroll
        for (int fb = 0; fb < FB; ++fb) { a[fb] = fmaf(wi, v0[fb], a[fb]); d[fb] = fmaf(wi, v1[fb], d[fb]); }
      }
    }
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) a[fb] = fast_sqrt(fmaxf(a[fb], 0.f) + eps);
  } else {
#pragma unroll
    load_chunk<FB>(slab + start * FB, a);
    load_chunk<FB>(slab + (M + start) * FB, d);
  }
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) {
    phase = wrap_half_turns(phase + d[fb]);
    const cpx u = fast_cis_pi(phase);
    zrow[fb * pitch] = cpx{a[fb] * u.re, real_only ? 0.f : a[fb] * u.im};
  }
}

// ---- look-back (a segment that does not start at frame 0): the running phase before frame
//      `f_end` is the projection of sum_{t < f_end} of channel 1; sums and projection in FP64
//      (the sum reaches ~100 half-turns, where FP32 resolves 1e-5), folded, then FP32. ----
ISI_HD double lookback_row_sum(const float* row /* channel-1 row */, int f_end, float s1, float b1) {
  double acc = 0.0;
  for (int t = 0; t < f_end; ++t) acc += (double)fmaf(row[t], s1, b1);
  return acc;
}
// The same sums for R rows at once (rows first_row + i * row_step of a [.., pitch] plane), frames
// added in rising order per row.  `vec`: rows 16-byte aligned and f_end a multiple of 4 -- then
// every round issues 4 independent 16-byte loads per row before the first add, so R x 4 loads
// are in flight per thread instead of one row's.
template <int R>
ISI_HD void lookback_rows_sum(const float* plane, int64_t pitch, int first_row, int row_step, int f_end,
                              float s1, float b1, bool vec, double* acc /* [R] */) {
#pragma unroll
  for (int i = 0; i < R; ++i) acc[i] = 0.0;
  if (!vec) {
#pragma unroll
    for (int i = 0; i < R; ++i)
      acc[i] = lookback_row_sum(plane + (int64_t)(first_row + i * row_step) * pitch, f_end, s1, b1);
    return;
  }
  for (int t = 0; t < f_end; t += 16) {
    f4 q[R][4];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float* row = plane + (int64_t)(first_row + i * row_step) * pitch + t;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        q[i][j] = (t + 4 * j < f_end) ? *reinterpret_cast<const f4*>(row + 4 * j) : f4{0.f, 0.f, 0.f, 0.f};
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (t + 4 * j < f_end) {
          acc[i] += (double)fmaf(q[i][j].x, s1, b1); acc[i] += (double)fmaf(q[i][j].y, s1, b1);
          acc[i] += (double)fmaf(q[i][j].z, s1, b1); acc[i] += (double)fmaf(q[i][j].w, s1, b1);
        }
      }
    }
  }
}
template <bool MEL>
ISI_HD float lookback_phase(const double* sums, int start, int count, const float* w) {
  double ph = 0.0;
  if (MEL) {
#pragma unroll
    for (int i = 0; i < kMaxMelWidth; ++i) if (i < count) ph += (double)w[i] * sums[start + i];
  } else {
    ph = sums[start];
  }
  ph -= 2.0 * rint(0.5 * ph);
  return (float)ph;
}

// ---- tangle: work item `it` (0 .. M/2-1) of one frame, in place on z[0..M].  Item it>0 owns
//      bins it and M-it; item 0 owns bins 0 (with X[M]) and M/2.  Writes conj(2 Z[k]).
//      `w` = W_N^it (the forward twiddle; its conjugate is what the inverse needs). ----
template <typename P>
ISI_HD void tangle_item(int it, cpx* z, cpx w) {
  constexpr int M = P::M;
  if (it == 0) {
    const float a0 = z[0].re, am = z[M].re;
    z[0] = cpx{a0 + am, -(a0 - am)};                 // conj(E0 + i D0), both real
    const cpx h = z[M / 2];
    z[M / 2] = cpx{2.f * h.re, 2.f * h.im};           // conj Z[M/2] = X[M/2]
    return;
  }
  const cpx a = z[it], b = z[M - it];
  const cpx e = cpx{a.re + b.re, a.im - b.im};        // A + conj B = 2 E
  const cpx d = cpx{a.re - b.re, a.im + b.im};        // A - conj B = 2 D
  const cpx o = cmul(cpx{w.re, -w.im}, d);            // 2 O = conj(W^k) 2 D
  // conj Z[k] = conj E - i conj O ;  conj Z[M-k] = E - i O
  z[it] = cpx{e.re - o.im, -e.im - o.re};
  z[M - it] = cpx{e.re + o.im, e.im - o.re};
}

// ---- pass 1 of the transform reading the natural-order buffer in place: all of a thread's
//      loads happen before the frame group's barrier, all stores after it ----
template <typename P>
ISI_HD void ifft_pass1_load(int j, const cpx* z, cpx* v) {
#pragma unroll
  for (int r = 0; r < P::R1; ++r) v[r] = z[j + 64 * r];
}
template <typename P>
ISI_HD void ifft_pass1_store(int j, cpx* v, const cpx* tws, cpx* zA) {
  dft_small<P::R1>(v);
#pragma unroll
  for (int p0 = 1; p0 < P::R1; p0 += 4) {
    cpx t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) if (p0 + i < P::R1) t[i] = tws[64 * (p0 + i) + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) if (p0 + i < P::R1) v[p0 + i] = cmul(v[p0 + i], t[i]);
  }
#pragma unroll
  for (int p = 0; p < P::R1; ++p) zA[j + P::kBlockPitch * p] = v[p];
}

// ---- overlap-add: batch-relative sample s of a batch of nf frames; frame fb's transform
//      output sits at z[fb * pitch + n] ----
ISI_HD float ola_sample(const cpx* z, int pitch, const float* win, int n_fft, int hop, int nf, int s) {
  float acc = 0.f;
  for (int fb = 0; fb < nf; ++fb) {
    const int idx = s - fb * hop;
    if (idx >= 0 && idx < n_fft) {
      const cpx y = z[fb * pitch + (idx >> 1)];
      acc = fmaf((idx & 1) ? -y.im : y.re, win[idx], acc);
    }
  }
  return acc;
}
// four consecutive samples s0 .. s0+3, s0 and hop multiples of 4: whole groups are inside or
// outside a frame, and a group is two complex values
template <int FB>
ISI_HD void ola_quad(const cpx* z, int pitch, const float* win, int n_fft, int hop, int nf, int s0, float* acc) {
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) {
    const int idx = s0 - fb * hop;
    if (fb < nf && idx >= 0 && idx < n_fft) {
      const cpx y0 = z[fb * pitch + (idx >> 1)], y1 = z[fb * pitch + (idx >> 1) + 1];
      const f4 wq = *reinterpret_cast<const f4*>(win + idx);
      acc[0] = fmaf(y0.re, wq.x, acc[0]);
      acc[1] = fmaf(-y0.im, wq.y, acc[1]);
      acc[2] = fmaf(y1.re, wq.z, acc[2]);
      acc[3] = fmaf(-y1.im, wq.w, acc[3]);
    }
  }
}

// frames before `fs` whose windows reach sample fs*hop
ISI_HD int ola_lookback_frames(int n_fft, int hop) { return (n_fft - 1) / hop; }

}  // namespace imelif
}  // namespace isi
