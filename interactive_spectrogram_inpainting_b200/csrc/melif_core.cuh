// Per-thread phases of the fused STFT -> (mel) -> log-magnitude + IF kernel.
//
// Everything here is written against an abstract "thread id + shared buffer" so the
// very same code is compiled by nvcc into melif.cu and by g++ into the CPU emulation
// that tests/test_melif_emulation.py uses to check the index arithmetic against the
// oracle without a GPU (the emulation is test infrastructure, not a product path).
//
// Transform plan for an n_fft-point real frame (M = n_fft/2 complex points):
//   pack    z[m] = w[2m] a[2m] + i w[2m+1] a[2m+1]
//   FFT     in place, decimation in frequency, radices (R1, 16, 4) with R1 = M/64;
//           bin k = p1 + R1*p2 + 16*R1*p3 ends up at slot (M/R1)*p1 + 4*p2 + p3
//   untangle X[k], X[M-k] from Z[k], Z[M-k] (one complex multiply per pair)
//   polar   |X|, angle(X); time-unwrapped phase carried in registers across frames
//   mel     banded projections of (|X|+eps)^2 and of the unwrapped phase
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

struct cpx { float re, im; };

ISI_HD cpx cmul(cpx a, cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
ISI_HD cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
ISI_HD cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }
ISI_HD cpx mul_neg_i(cpx a) { return {a.im, -a.re}; }   // a * (-i)

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvPi = 0.31830988618379067154f;

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
  // twiddles W16^(a c), W16 = exp(-i pi/8)
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
  // un-transpose: element for k = c + 4d sits at 4c + d
  cpx t;
#define ISI_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
  ISI_SWAP(1, 4) ISI_SWAP(2, 8) ISI_SWAP(3, 12) ISI_SWAP(6, 9) ISI_SWAP(7, 13) ISI_SWAP(11, 14)
#undef ISI_SWAP
}

// 8-point and smaller first radices are built from dft4 + one radix-2 layer
ISI_HD void dft8(cpx* v) {
  // n = a + 2b (a<2, b<4), k = c + 4d (c<4, d<2)
  const float h = 0.70710678118654752440f;
  dft4(v[0], v[2], v[4], v[6]);
  dft4(v[1], v[3], v[5], v[7]);
  // odd branch twiddles W8^c
  v[3] = cmul(v[3], cpx{h, -h});
  v[5] = mul_neg_i(v[5]);
  v[7] = cmul(v[7], cpx{-h, -h});
  // v[2c] = even[c], v[2c+1] = odd[c] ; X[c] = e+o, X[c+4] = e-o
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
  static constexpr int M = NFFT / 2;      // complex points / output bins
  static constexpr int R1 = M / 64;       // first radix: 16 (2048), 8 (1024), 4 (512)
  static constexpr int Q1 = 64;           // M / R1
  static constexpr int kFftThreads = 64;  // threads cooperating on one frame's FFT
  static_assert(R1 == 16 || R1 == 8 || R1 == 4, "n_fft must be 2048, 1024 or 512");
  // slot of bin k after the three in-place passes
  static ISI_HD int slot(int k) { return Q1 * (k % R1) + 4 * ((k / R1) % 16) + k / (16 * R1); }
};

// ---- phase A: window + pack one frame into z[0..M) (thread t of NT) ----
template <typename P>
ISI_HD void pack_frame(int t, int nt, cpx* z, const float* audio, int64_t n_samples,
                       int64_t first_sample, const float* window) {
  for (int m = t; m < P::M; m += nt) {
    int64_t i0 = first_sample + 2 * m;
    float a0 = (i0 >= 0 && i0 < n_samples) ? audio[i0] : 0.f;
    float a1 = (i0 + 1 >= 0 && i0 + 1 < n_samples) ? audio[i0 + 1] : 0.f;
    z[m] = cpx{a0 * window[2 * m], a1 * window[2 * m + 1]};
  }
}

// twiddle table: tw[j] = exp(-2 pi i j / N) for j in [0, N)
// ---- phase B1: first pass, radix R1 over stride Q1 (thread j of 64) ----
template <typename P>
ISI_HD void fft_pass1(int j, cpx* z, const cpx* tw) {
  cpx v[P::R1];
#pragma unroll
  for (int r = 0; r < P::R1; ++r) v[r] = z[j + P::Q1 * r];
  dft_small<P::R1>(v);
#pragma unroll
  for (int p = 1; p < P::R1; ++p) v[p] = cmul(v[p], tw[2 * j * p]);   // W_M^(j p) = W_N^(2 j p)
#pragma unroll
  for (int p = 0; p < P::R1; ++p) z[j + P::Q1 * p] = v[p];
}

// ---- phase B2: inside each block of 64, radix 16 over stride 4.  64 threads cover
//      R1 blocks x 4 columns = 4*R1 work items (one, or a half/quarter, per thread) ----
template <typename P>
ISI_HD void fft_pass2(int t, cpx* z, const cpx* tw) {
  for (int item = t; item < 4 * P::R1; item += P::kFftThreads) {
    const int b = item >> 2, j = item & 3;
    cpx* blk = z + 64 * b;
    cpx v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = blk[j + 4 * r];
    dft16(v);
#pragma unroll
    for (int p = 1; p < 16; ++p) v[p] = cmul(v[p], tw[(P::N / 64) * j * p]);   // W_64^(j p)
#pragma unroll
    for (int p = 0; p < 16; ++p) blk[j + 4 * p] = v[p];
  }
}

// ---- phase B3: radix-4 on consecutive quadruples ----
template <typename P>
ISI_HD void fft_pass3(int t, cpx* z) {
  for (int b = t; b < P::M / 4; b += P::kFftThreads) {
    cpx* q = z + 4 * b;
    cpx a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3];
    dft4(a0, a1, a2, a3);
    q[0] = a0; q[1] = a1; q[2] = a2; q[3] = a3;
  }
}

// numpy-style wrapped phase step (magenta spectral_ops.unwrap): the value that the
// time-unwrapped phase advances by, given the raw step dd
ISI_HD float wrapped_step(float dd) {
  if (fabsf(dd) < kPi) return dd;
  float m = dd + kPi;
  m = m - kTwoPi * floorf(m / kTwoPi) - kPi;       // python-style remainder into [-pi, pi)
  if (m == -kPi && dd > 0.f) m = kPi;
  return m;
}

// running state of one spectrogram bin across frames
struct BinState { float prev_phase; float unwrapped; };

// ---- phase C: work item `it` (0..M/2) of one frame: two bins in, two (v0, v1) out.
//      Item 0 owns bin M/2 and the real-only bin (Nyquist when drop_dc, else DC);
//      item it>0 owns bins it and M-it.  Results overwrite the FFT slots they came from.
//      mel mode : v0 = (|X|+eps)^2, v1 = unwrapped phase
//      linear   : v0 = log(|X|+eps), v1 = instantaneous frequency
template <typename P>
ISI_HD void polar_item(int it, cpx* z, const cpx* tw, bool first_frame, bool use_mel,
                       bool drop_dc, float eps, BinState& sa, BinState& sb) {
  const int M = P::M;
  cpx xa, xb;
  int slot_a, slot_b;
  if (it == 0) {
    cpx z0 = z[P::slot(0)], zh = z[P::slot(M / 2)];
    xa = cpx{zh.re, -zh.im};                                  // X[M/2] = conj(Z[M/2])
    xb = drop_dc ? cpx{z0.re - z0.im, 0.f} : cpx{z0.re + z0.im, 0.f};
    slot_a = P::slot(M / 2); slot_b = P::slot(0);
  } else {
    slot_a = P::slot(it); slot_b = P::slot(M - it);
    cpx a = z[slot_a], b = z[slot_b];
    cpx e = cpx{0.5f * (a.re + b.re), 0.5f * (a.im - b.im)};  // (A + conj B)/2
    cpx d = cpx{0.5f * (a.re - b.re), 0.5f * (a.im + b.im)};  // (A - conj B)/2
    cpx p = cmul(tw[it], mul_neg_i(d));                       // W_N^k * (-i) * d
    xa = cadd(e, p);
    cpx m = csub(e, p);
    xb = cpx{m.re, -m.im};
  }
  cpx xs[2] = {xa, xb};
  BinState* st[2] = {&sa, &sb};
  cpx out[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float mag = sqrtf(xs[i].re * xs[i].re + xs[i].im * xs[i].im);
    float ph = atan2f(xs[i].im, xs[i].re);
    float step = first_frame ? ph : wrapped_step(ph - st[i]->prev_phase);
    st[i]->prev_phase = ph;
    st[i]->unwrapped = first_frame ? ph : st[i]->unwrapped + step;
    if (use_mel) { float a = mag + eps; out[i] = cpx{a * a, st[i]->unwrapped}; }
    else         { out[i] = cpx{logf(mag + eps), step * kInvPi}; }
  }
  z[slot_a] = out[0];
  z[slot_b] = out[1];
}

// slot that holds output row `row` (0..M) after phase C
template <typename P>
ISI_HD int row_slot(int row, bool drop_dc) {
  int k = drop_dc ? row + 1 : row;            // FFT bin of this row
  return P::slot(k == P::M ? 0 : k);          // the Nyquist bin lives in DC's slot
}

// ---- phase D: one output row of one frame ----
struct RowState { float prev; };

template <typename P>
ISI_HD void emit_row(int row, const cpx* z, bool first_frame, bool use_mel, bool drop_dc,
                     float eps, int mel_start, int mel_count, const float* mel_w,
                     RowState& st, float& out0, float& out1) {
  if (!use_mel) {
    cpx v = z[row_slot<P>(row, drop_dc)];
    out0 = v.re; out1 = v.im;
    return;
  }
  float m2 = 0.f, mp = 0.f;
  for (int i = 0; i < mel_count; ++i) {
    cpx v = z[row_slot<P>(mel_start + i, drop_dc)];
    m2 = fmaf(mel_w[i], v.re, m2);
    mp = fmaf(mel_w[i], v.im, mp);
  }
  out0 = logf(m2 + eps);
  float step = first_frame ? mp : wrapped_step(mp - st.prev);
  st.prev = mp;
  out1 = step * kInvPi;
}

}  // namespace melif
}  // namespace isi
