// Per-thread phases of the fused STFT -> (mel) -> log-magnitude + IF kernel.
//
// Everything here is written against an abstract "thread id + shared buffers" so the
// very same code is compiled by nvcc into melif.cu and by g++ into the CPU emulation
// that tests/test_melif_emulation.py uses to check the index arithmetic against the
// oracle without a GPU (the emulation is test infrastructure, not a product path).
//
// Transform plan for an n_fft-point real frame (M = n_fft/2 complex points):
//   pass 1  window + pack z[m] = w[2m] a[2m] + i w[2m+1] a[2m+1] straight from the staged
//           audio, radix R1 = M/64 over stride 64, twiddle, into zA (64-blocks at pitch 65)
//   pass 2  radix 16 over stride 4 inside each 64-block of zA, twiddle, in place
//   pass 3  radix 4 on consecutive quadruples of zA, rewritten in place in natural bin order
//   polar   untangle X[k], X[M-k] from Z[k], Z[M-k]; |X|, phase step = arg(X_t conj X_t-1)
//           (previous spectrum value carried in registers); (v0, v1) overwrite Z in place
//   emit    banded mel projections of (|X|+eps)^2 and of the phase steps (or a copy in
//           linear mode) for all FB frames of a row at once, log and wrapped mel-IF
// One buffer of FB frames is the kernel's whole working set (65 KB at n_fft 2048 with the
// tables and the audio stage), which is what lets three CTAs share an SM.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define ISI_HD __host__ __device__ __forceinline__
#else
#define ISI_HD inline
#endif

namespace isi {
namespace melif {

struct alignas(8) cpx { float re, im; };   // 8-byte aligned: one LDS.64 / STS.64 per value

ISI_HD cpx cmul(cpx a, cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
ISI_HD cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
ISI_HD cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }
ISI_HD cpx mul_neg_i(cpx a) { return {a.im, -a.re}; }   // a * (-i)

constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfPi = 1.57079632679489661923f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvTwoPi = 0.15915494309189533577f;
constexpr float kInvPi = 0.31830988618379067154f;

// ---- fast scalar math (device: MUFU-based; host emulation: libm) ----
ISI_HD float fast_rcp(float x) {           // one MUFU.RCP, ~1 ulp
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
ISI_HD float fast_rsqrt(float x) {         // one MUFU.RSQ, ~2 ulp
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
ISI_HD float fast_log(float x) {           // one MUFU.LG2 + one multiply
#ifdef __CUDA_ARCH__
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.69314718055994530942f;
#else
  return logf(x);
#endif
}

// atan2 with a degree-7 (in a^2) minimax polynomial on [0,1]: |err| < 2e-7 rad.
// atan2(0, 0) = 0 like torch.angle(0).
ISI_HD float fast_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = (mx < 1.17549435e-38f) ? 0.f : mn * fast_rcp(mx);   // zero / denormal -> 0
  const float s = a * a;
  float p = -0.0040545654483139515f;
  p = fmaf(p, s, 0.021862952038645744f);
  p = fmaf(p, s, -0.0559123195707798f);
  p = fmaf(p, s, 0.0964219719171524f);
  p = fmaf(p, s, -0.1390862911939621f);
  p = fmaf(p, s, 0.19946566224098206f);
  p = fmaf(p, s, -0.33329859375953674f);
  p = fmaf(p, s, 0.9999993443489075f);
  float r = p * a;
  if (ay > ax) r = kHalfPi - r;
  if (x < 0.f) r = kPi - r;
  return copysignf(r, y);
}

// forward 4-point DFT, outputs in natural order
ISI_HD void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
  cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
  a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

// forward 16-point DFT in registers: v[n] -> v[k], natural order in and out
ISI_HD void dft16(cpx* v) {
  // n = a + 4b, k = c + 4d:  y_c[a] = W16^(a c) DFT4_b(v[a+4b])[c];  X[c+4d] = DFT4_a(y_c[a])[d]
  const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f;   // cos/sin(pi/8)
  const float h = 0.70710678118654752440f;
#pragma unroll
  for (int a = 0; a < 4; ++a) dft4(v[a], v[a + 4], v[a + 8], v[a + 12]);      // index a + 4c now
  v[1 + 4] = cmul(v[1 + 4], cpx{c1, -s1});      // a=1,c=1 : W^1
  v[1 + 8] = cmul(v[1 + 8], cpx{h, -h});        // a=1,c=2 : W^2
  v[1 + 12] = cmul(v[1 + 12], cpx{s1, -c1});    // a=1,c=3 : W^3
  v[2 + 4] = cmul(v[2 + 4], cpx{h, -h});        // a=2,c=1 : W^2
  v[2 + 8] = mul_neg_i(v[2 + 8]);               // a=2,c=2 : W^4
  v[2 + 12] = cmul(v[2 + 12], cpx{-h, -h});     // a=2,c=3 : W^6
  v[3 + 4] = cmul(v[3 + 4], cpx{s1, -c1});      // a=3,c=1 : W^3
  v[3 + 8] = cmul(v[3 + 8], cpx{-h, -h});       // a=3,c=2 : W^6
  v[3 + 12] = cmul(v[3 + 12], cpx{-c1, s1});    // a=3,c=3 : W^9
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);  // c + 4d at 4c+d
  cpx t;
#define ISI_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
  ISI_SWAP(1, 4) ISI_SWAP(2, 8) ISI_SWAP(3, 12) ISI_SWAP(6, 9) ISI_SWAP(7, 13) ISI_SWAP(11, 14)
#undef ISI_SWAP
}

ISI_HD void dft8(cpx* v) {
  // n = a + 2b (a<2, b<4), k = c + 4d (c<4, d<2)
  const float h = 0.70710678118654752440f;
  dft4(v[0], v[2], v[4], v[6]);
  dft4(v[1], v[3], v[5], v[7]);
  v[3] = cmul(v[3], cpx{h, -h});
  v[5] = mul_neg_i(v[5]);
  v[7] = cmul(v[7], cpx{-h, -h});
  cpx out[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) { out[c] = cadd(v[2 * c], v[2 * c + 1]); out[c + 4] = csub(v[2 * c], v[2 * c + 1]); }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = out[i];
}

template <int R> ISI_HD void dft_small(cpx* v);
template <> ISI_HD void dft_small<16>(cpx* v) { dft16(v); }
template <> ISI_HD void dft_small<8>(cpx* v) { dft8(v); }
template <> ISI_HD void dft_small<4>(cpx* v) { dft4(v[0], v[1], v[2], v[3]); }

// Geometry of one transform size.
template <int NFFT>
struct Plan {
  static constexpr int N = NFFT;
  static constexpr int M = NFFT / 2;        // complex points; bins 0..M
  static constexpr int R1 = M / 64;         // first radix: 16 (2048), 8 (1024), 4 (512)
  static constexpr int kFftThreads = 64;    // threads cooperating on one frame's FFT
  static constexpr int kBlockPitch = 65;      // 64-blocks one element apart: lanes that walk the
                                              // blocks (passes 2 and 3) hit 16 distinct bank pairs
  static constexpr int kPitchA = kBlockPitch * R1 + 1;   // >= M + 1: also holds bins 0..M
  static_assert(R1 == 16 || R1 == 8 || R1 == 4, "n_fft must be 2048, 1024 or 512");
};

// ---- synchronous staging of the audio span of one frame batch (emulation, and the
//      device path when the span is not 16-byte copyable) ----
template <typename S>
ISI_HD void stage_fill(int t, int nt, S* stage, int span, const S* audio,
                       int64_t n_samples, int64_t first_sample) {
  for (int i = t; i < span; i += nt) {
    int64_t s = first_sample + i;
    stage[i] = (s >= 0 && s < n_samples) ? audio[s] : S(0);
  }
}

// Two consecutive samples of a frame as floats.  S = float: the staged values; S = int16_t
// (PCM as the dataset stores it): float(x) * pcm_scale, the conversion a host loader would do
// before the upload.  `aligned` = the pair starts on a 2*sizeof(S) boundary (one load).
struct alignas(4) pcm_pair { int16_t a, b; };

template <bool ALIGNED>
ISI_HD cpx load_pair(const float* frame, int m, float) {
  if (ALIGNED) return reinterpret_cast<const cpx*>(frame)[m];
  return cpx{frame[2 * m], frame[2 * m + 1]};
}
template <bool ALIGNED>
ISI_HD cpx load_pair(const int16_t* frame, int m, float pcm_scale) {
  pcm_pair q;
  if (ALIGNED) q = reinterpret_cast<const pcm_pair*>(frame)[m];
  else q = pcm_pair{frame[2 * m], frame[2 * m + 1]};
  return cpx{(float)q.a * pcm_scale, (float)q.b * pcm_scale};
}

// Twiddle table in shared memory, M entries laid out for conflict-free reads:
//   tws[e]          = W_64^e           e in [0, 64)              (pass 2, warp-broadcast reads)
//   tws[64 p + j]   = W_M^(j p)        p in [1, R1), j in [0, 64) (pass 1, lanes along j)
// fft_table_source(i) is the index into the caller's W_N^k table (k < N) that slot i holds.
template <typename P>
ISI_HD int fft_table_source(int i) {
  const int p = i / 64, j = i % 64;
  return p == 0 ? (2 * P::M / 64) * j : 2 * j * p;
}

// ---- pass 1 (thread j of 64): window, pack, radix R1 over stride 64 ----
template <typename P, typename S>
ISI_HD void fft_pass1(int j, const S* frame /* stage + fb*hop */, bool pair_aligned, float pcm_scale,
                      const float* win, const cpx* tws, cpx* zA) {
  cpx v[P::R1];
  if (pair_aligned) {
#pragma unroll
    for (int r = 0; r < P::R1; ++r) {
      const int m = j + 64 * r;
      const cpx a = load_pair<true>(frame, m, pcm_scale);
      const cpx w = reinterpret_cast<const cpx*>(win)[m];
      v[r] = cpx{a.re * w.re, a.im * w.im};
    }
  } else {
#pragma unroll
    for (int r = 0; r < P::R1; ++r) {
      const int m = j + 64 * r;
      const cpx a = load_pair<false>(frame, m, pcm_scale);
      const cpx w = reinterpret_cast<const cpx*>(win)[m];
      v[r] = cpx{a.re * w.re, a.im * w.im};
    }
  }
  dft_small<P::R1>(v);
  // W_M^(j p): the twiddles are fetched four at a time so that their shared-memory latency
  // overlaps instead of serialising load -> multiply -> load
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

// ---- pass 2: inside each 64-block, radix 16 over stride 4 ----
template <typename P>
ISI_HD void fft_pass2(int t, const cpx* tws, cpx* zA) {
  for (int item = t; item < 4 * P::R1; item += P::kFftThreads) {
    const int b = item % P::R1, j = item / P::R1;       // lanes along b: bank = b (pitch 65)
    cpx* blk = zA + P::kBlockPitch * b;
    cpx v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = blk[j + 4 * r];
    dft16(v);
    if (j != 0) {
#pragma unroll
      for (int p0 = 1; p0 < 16; p0 += 4) {                                        // W_64^(j p)
        cpx t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) if (p0 + i < 16) t[i] = tws[j * (p0 + i)];
#pragma unroll
        for (int i = 0; i < 4; ++i) if (p0 + i < 16) v[p0 + i] = cmul(v[p0 + i], t[i]);
      }
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) blk[j + 4 * p] = v[p];
  }
}

// ---- pass 3: radix 4 on the quadruples (p1, p2) of zA, rewritten IN PLACE in natural bin
//      order (bin p1 + R1 p2 + 16 R1 p3).  Load + butterfly and store are two calls with the
//      frame group's barrier between them, because the natural-order slots overlap other
//      threads' quadruples.  Lanes run along p1, so the stores are contiguous. ----
template <typename P>
struct Pass3Regs { cpx q[P::R1 / 4][4]; };

template <typename P>
ISI_HD void fft_pass3_load(int t, const cpx* zA, Pass3Regs<P>& r) {
  constexpr int kStep = P::kFftThreads / P::R1;          // p2 values covered per sweep
  const int p1 = t % P::R1;
#pragma unroll
  for (int i = 0; i < P::R1 / 4; ++i) {
    const int p2 = t / P::R1 + kStep * i;
    const cpx* q = zA + P::kBlockPitch * p1 + 4 * p2;
    r.q[i][0] = q[0]; r.q[i][1] = q[1]; r.q[i][2] = q[2]; r.q[i][3] = q[3];
    dft4(r.q[i][0], r.q[i][1], r.q[i][2], r.q[i][3]);
  }
}

template <typename P>
ISI_HD void fft_pass3_store(int t, const Pass3Regs<P>& r, cpx* z) {
  constexpr int kStep = P::kFftThreads / P::R1;
  const int p1 = t % P::R1;
#pragma unroll
  for (int i = 0; i < P::R1 / 4; ++i) {
    const int p2 = t / P::R1 + kStep * i;
    cpx* o = z + p1 + P::R1 * p2;
    o[0] = r.q[i][0]; o[16 * P::R1] = r.q[i][1]; o[32 * P::R1] = r.q[i][2]; o[48 * P::R1] = r.q[i][3];
  }
}

// phase step folded into [-pi, pi] (round-to-nearest multiple of 2 pi; equals the
// numpy/magenta unwrap rule except on exact +-pi ties)
ISI_HD float wrap_step(float dd) {
  return fmaf(-kTwoPi, rintf(dd * kInvTwoPi), dd);
}

// State of one spectrogram bin across frames: the previous spectrum value (exact zeros
// replaced by 1+0i, whose phase is also 0).  The wrapped phase advance of frame t is
// arg(X_t conj X_{t-1}); nothing cumulative is carried, because the mel IF only needs
// differences of the projected unwrapped phase, and the projection is linear:
//   mel_phase_t - mel_phase_{t-1} = sum_k w_k (u_t[k] - u_{t-1}[k]) = sum_k w_k step_t[k].
struct BinState { float pre, pim; };

template <bool MEL>
ISI_HD cpx polar_bin(cpx x, bool first_frame, float eps, BinState& st) {
  const float m2 = fmaf(x.re, x.re, x.im * x.im);
  const float mag = (m2 > 1.17549435e-38f) ? m2 * fast_rsqrt(m2) : 0.f;
  if (x.re == 0.f && x.im == 0.f) x.re = 1.f;
  const float step = first_frame
      ? fast_atan2(x.im, x.re)
      : fast_atan2(x.im * st.pre - x.re * st.pim, x.re * st.pre + x.im * st.pim);
  st.pre = x.re; st.pim = x.im;
  if (MEL) { const float a = mag + eps; return cpx{a * a, step}; }
  return cpx{fast_log(mag + eps), step * kInvPi};
}

// ---- polar: work item `it` (0..M/2-1) of one frame, in place on its natural-order
//      spectrum z[0..M].  Item it>0 owns bins it and M-it.  Item 0 owns bin M/2 and ONE of the
//      two purely real bins: `real_bin` = M when the DC bin is the dropped one, else 0 (the
//      other one is never read by emit and keeps its raw FFT value).  Item 0 runs the same
//      instruction stream as every other item — only operands are selected — so the warp that
//      holds it does not execute a second, divergent path (that path used to make one warp
//      1.75x slower than the rest and stall the CTA at the barrier before emit).
//      z[k] <- (v0, v1): mel mode (|X|+eps)^2 and the phase step, linear mode log(|X|+eps), IF.
//      MAYBE_ZERO = false compiles the selects out for items that cannot be item 0. ----
template <typename P, bool MEL, bool MAYBE_ZERO = true>
ISI_HD void polar_item(int it, cpx* z, cpx w /* W_N^it */, int real_bin, bool first_frame, float eps,
                       BinState& sa, BinState& sb) {
  constexpr int M = P::M;
  const bool special = MAYBE_ZERO && it == 0;
  const int ka = special ? M / 2 : it, kb = special ? real_bin : M - it;
  const cpx a = z[it], b = z[special ? M / 2 : M - it];
  const cpx e = cpx{0.5f * (a.re + b.re), 0.5f * (a.im - b.im)};              // (A + conj B)/2
  const cpx d = cpx{0.5f * (a.re - b.re), 0.5f * (a.im + b.im)};              // (A - conj B)/2
  const cpx p = cmul(w, mul_neg_i(d));                                        // W_N^k (-i) d
  const cpx m = csub(e, p);
  cpx xa = cadd(e, p), xb = cpx{m.re, -m.im};
  if (special) {
    xa = cpx{b.re, -b.im};                                                    // X[M/2] = conj Z[M/2]
    xb = cpx{real_bin == 0 ? a.re + a.im : a.re - a.im, 0.f};                 // X[0] or X[M]
  }
  z[ka] = polar_bin<MEL>(xa, first_frame, eps, sa);
  z[kb] = polar_bin<MEL>(xb, first_frame, eps, sb);
}

// ---- emit: one output row, all FB frames of the batch at once; frame fb's values sit at
//      z[fb * pitch + bin].  `bin0` is the FFT bin of the row (linear mode) or of the first
//      band element (mel mode); mel weights live in registers, zero beyond `count`;
//      `count_uniform` >= count is uniform across the warp so whole taps are skipped without
//      divergence. ----
constexpr int kMaxMelWidth = 8;

template <int FB>
ISI_HD void emit_linear(const cpx* z, int pitch, int bin0, float* out0, float* out1) {
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) { const cpx v = z[fb * pitch + bin0]; out0[fb] = v.re; out1[fb] = v.im; }
}

template <int FB>
ISI_HD void emit_mel(const cpx* z, int pitch, int bin0, int count, int count_uniform, const float* w,
                     bool first_is_frame0, float eps, float* out0, float* out1) {
  float m2[FB], mp[FB];
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) { m2[fb] = 0.f; mp[fb] = 0.f; }
  // two taps per (warp-uniform) step: all 2*FB loads are issued before the first use
#pragma unroll
  for (int i = 0; i < kMaxMelWidth; i += 2) {
    if (i < count_uniform) {
      cpx va[FB], vb[FB];
#pragma unroll
      for (int fb = 0; fb < FB; ++fb) {
        va[fb] = (i < count) ? z[fb * pitch + bin0 + i] : cpx{0.f, 0.f};
        vb[fb] = (i + 1 < count) ? z[fb * pitch + bin0 + i + 1] : cpx{0.f, 0.f};
      }
#pragma unroll
      for (int fb = 0; fb < FB; ++fb) {
        m2[fb] = fmaf(w[i + 1], vb[fb].re, fmaf(w[i], va[fb].re, m2[fb]));
        mp[fb] = fmaf(w[i + 1], vb[fb].im, fmaf(w[i], va[fb].im, mp[fb]));
      }
    }
  }
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) {
    out0[fb] = fast_log(m2[fb] + eps);
    out1[fb] = ((fb == 0 && first_is_frame0) ? mp[fb] : wrap_step(mp[fb])) * kInvPi;
  }
}

// ---- fused epilogue: masked-phase transform, then the per-channel affine normalisation ----
template <int FB>
ISI_HD void apply_epilogue(float* v0, float* v1, bool mask_phase, float mask_threshold, float s0,
                           float b0, float s1, float b1) {
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) {
    const float ph = (mask_phase && v0[fb] < mask_threshold) ? 0.f : v1[fb];
    v0[fb] = fmaf(v0[fb], s0, b0);
    v1[fb] = fmaf(ph, s1, b1);
  }
}

}  // namespace melif
}  // namespace isi
