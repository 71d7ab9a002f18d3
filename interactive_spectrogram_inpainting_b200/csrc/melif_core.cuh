// Per-thread phases of the fused STFT -> (mel) -> log-magnitude + IF kernel.
//
// Everything here is written against an abstract "thread id + shared buffers" so the
// very same code is compiled by nvcc into melif.cu and by g++ into the CPU emulation
// that tests/test_melif_emulation.py uses to check the index arithmetic against the
// oracle without a GPU (the emulation is test infrastructure, not a product path).
//
// FRAME PAIRS.  The forward kernel transforms two frames of a note at once: every value is
// an `f2` whose two lanes belong to the two frames of a pair, and every arithmetic
// instruction is one packed FP32 operation (sm_100 FADD2 / FMUL2 / FFMA2: add.f32x2,
// mul.f32x2, fma.rn.f32x2).  Constants (twiddles, window, mel weights) are the same for both
// frames and enter as the broadcast scalar operand those instructions accept, so no lane
// shuffles or duplicated tables are needed; a complex value of the pair (`cpx2`: re.A re.B
// im.A im.B) is one 16-byte shared-memory access.  The g++ build evaluates the lanes one
// after the other.  The scalar `cpx` versions of the FFT passes still serve the inverse
// kernel (imelif_core.cuh).
//
// Transform plan for an n_fft-point real frame (M = n_fft/2 complex points):
//   pass 1  window + pack z[m] = w[2m] a[2m] + i w[2m+1] a[2m+1] straight from the staged
//           audio, radix R1 = M/64 over stride 64, twiddle, into zA (64-blocks at pitch 65)
//   pass 2  radix 16 over stride 4 inside each 64-block of zA, twiddle, in place
//   pass 3  radix 4 on consecutive quadruples of zA, rewritten in place in natural bin order
//   polar   untangle X[k], X[M-k] from Z[k], Z[M-k]; |X| and the unit phasor X/|X|; phase step
//           = arg(u_t conj u_t-1) by a degree-7 arcsine of the smaller component (the
//           previous frame's phasor is the neighbouring pair's, or carried in registers
//           across batches); (v0, v1) overwrite Z in place
//   emit    banded mel projections of (|X|+eps)^2 and of the phase steps (or a copy in
//           linear mode) for all FB frames of a row at once, log and wrapped mel-IF
// One buffer of FB frames is the kernel's whole working set.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define ISI_HD __host__ __device__ __forceinline__
#else
#define ISI_HD inline
#endif

namespace isi {
namespace melif {

// ---- packed pairs: lane x = first frame of the pair, lane y = second ----
#ifdef __CUDACC__
typedef float2 f2;
#else
struct alignas(8) f2 { float x, y; };
#endif

ISI_HD f2 mk2(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
ISI_HD f2 bc(float s) { return mk2(s, s); }              // broadcast operand of FMUL2 / FFMA2
ISI_HD f2 neg2(f2 a) { return mk2(-a.x, -a.y); }         // folds into the consumer's operand modifier
ISI_HD f2 add2(f2 a, f2 b) {
#ifdef __CUDA_ARCH__
  return __fadd2_rn(a, b);
#else
  return mk2(a.x + b.x, a.y + b.y);
#endif
}
ISI_HD f2 sub2(f2 a, f2 b) { return add2(a, neg2(b)); }
ISI_HD f2 mul2(f2 a, f2 b) {
#ifdef __CUDA_ARCH__
  return __fmul2_rn(a, b);
#else
  return mk2(a.x * b.x, a.y * b.y);
#endif
}
ISI_HD f2 fma2(f2 a, f2 b, f2 c) {
#ifdef __CUDA_ARCH__
  return __ffma2_rn(a, b, c);
#else
  return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

struct alignas(8) cpx { float re, im; };    // 8-byte aligned: one LDS.64 / STS.64 per value
struct alignas(16) cpx2 { f2 re, im; };     // the same bin of both frames: one LDS.128 / STS.128

ISI_HD cpx cmul(cpx a, cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
ISI_HD cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
ISI_HD cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }
ISI_HD cpx mul_neg_i(cpx a) { return {a.im, -a.re}; }   // a * (-i)

// a (per frame) times w (shared by both frames): FMUL2 + FFMA2 per component
ISI_HD cpx2 cmul(cpx2 a, cpx w) {
  cpx2 r;
  r.re = fma2(a.im, bc(-w.im), mul2(a.re, bc(w.re)));
  r.im = fma2(a.re, bc(w.im), mul2(a.im, bc(w.re)));
  return r;
}
ISI_HD cpx2 cadd(cpx2 a, cpx2 b) { cpx2 r; r.re = add2(a.re, b.re); r.im = add2(a.im, b.im); return r; }
ISI_HD cpx2 csub(cpx2 a, cpx2 b) { cpx2 r; r.re = sub2(a.re, b.re); r.im = sub2(a.im, b.im); return r; }
ISI_HD cpx2 mul_neg_i(cpx2 a) { cpx2 r; r.re = a.im; r.im = neg2(a.re); return r; }

// Storing a cpx2 is ONE 16-byte store.  (ptxas then spends up to four MOVs lining the halves up
// in consecutive registers, because the packed instructions leave their results in arbitrary
// aligned pairs.  Measured alternative, profiles/r02_melif_ws_split_stores_r2e: two 8-byte
// stores need no MOVs but sit 16 bytes apart across lanes, a 2-way bank conflict each -- 65.3 M
// instead of 53.0 M shared-memory wavefronts per 444 notes and a slower kernel.)
struct SplitPtr { cpx2* p; };
ISI_HD SplitPtr split_ptr(cpx2* p) { SplitPtr s; s.p = p; return s; }
ISI_HD void put(SplitPtr s, int k, cpx2 v) { s.p[k] = v; }
// the scalar complex type stores as one 8-byte value
ISI_HD cpx* split_ptr(cpx* p) { return p; }
ISI_HD void put(cpx* s, int k, cpx v) { s[k] = v; }

constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfPi = 1.57079632679489661923f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvTwoPi = 0.15915494309189533577f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kLn2 = 0.69314718055994530942f;

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
ISI_HD float fast_log2(float x) {          // one MUFU.LG2
#ifdef __CUDA_ARCH__
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return log2f(x);
#endif
}
ISI_HD float fast_log(float x) { return fast_log2(x) * kLn2; }

// atan2 with a degree-7 (in a^2) minimax polynomial on [0,1]: |err| < 2e-7 rad.
// atan2(0, 0) = 0 like torch.angle(0).  (Scalar path: the inverse kernel's tests.)
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

// forward 4-point DFT, outputs in natural order (C = cpx or cpx2)
template <typename C>
ISI_HD void dft4(C& a0, C& a1, C& a2, C& a3) {
  C t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
  a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

// forward 16-point DFT in registers: v[n] -> v[k], natural order in and out
template <typename C>
ISI_HD void dft16(C* v) {
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
  C t;
#define ISI_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
  ISI_SWAP(1, 4) ISI_SWAP(2, 8) ISI_SWAP(3, 12) ISI_SWAP(6, 9) ISI_SWAP(7, 13) ISI_SWAP(11, 14)
#undef ISI_SWAP
}

template <typename C>
ISI_HD void dft8(C* v) {
  // n = a + 2b (a<2, b<4), k = c + 4d (c<4, d<2)
  const float h = 0.70710678118654752440f;
  dft4(v[0], v[2], v[4], v[6]);
  dft4(v[1], v[3], v[5], v[7]);
  v[3] = cmul(v[3], cpx{h, -h});
  v[5] = mul_neg_i(v[5]);
  v[7] = cmul(v[7], cpx{-h, -h});
  C out[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) { out[c] = cadd(v[2 * c], v[2 * c + 1]); out[c + 4] = csub(v[2 * c], v[2 * c + 1]); }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = out[i];
}

template <int R, typename C> struct DftSmall;
template <typename C> struct DftSmall<16, C> { static ISI_HD void run(C* v) { dft16(v); } };
template <typename C> struct DftSmall<8, C> { static ISI_HD void run(C* v) { dft8(v); } };
template <typename C> struct DftSmall<4, C> { static ISI_HD void run(C* v) { dft4(v[0], v[1], v[2], v[3]); } };
template <int R, typename C> ISI_HD void dft_small(C* v) { DftSmall<R, C>::run(v); }

// Geometry of one transform size.
template <int NFFT>
struct Plan {
  static constexpr int N = NFFT;
  static constexpr int M = NFFT / 2;        // complex points; bins 0..M
  static constexpr int R1 = M / 64;         // first radix: 16 (2048), 8 (1024), 4 (512)
  static constexpr int kFftThreads = 64;    // threads cooperating on one transform
  static constexpr int kBlockPitch = 65;      // 64-blocks one element apart: lanes that walk the
                                              // blocks (passes 2 and 3) hit distinct banks
  // one transform's region: the 64-blocks, bins 0..M, and the kMaxMelWidth - 1 elements a
  // band read may run past bin M (zero weights, but the values must be finite)
  static constexpr int kPitchA = (kBlockPitch * R1 + 1 >= M + 8) ? kBlockPitch * R1 + 1 : M + 8;
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

// Two consecutive samples of a frame as floats.  S = int16_t (PCM as the dataset stores it):
// the integer values -- the sample scale is folded into the window table.  `ALIGNED` = the
// pair starts on a 2*sizeof(S) boundary (one load).
struct alignas(4) pcm_pair { int16_t a, b; };

template <bool ALIGNED>
ISI_HD cpx load_pair(const float* frame, int m) {
  if (ALIGNED) return reinterpret_cast<const cpx*>(frame)[m];
  return cpx{frame[2 * m], frame[2 * m + 1]};
}
template <bool ALIGNED>
ISI_HD cpx load_pair(const int16_t* frame, int m) {
  pcm_pair q;
  if (ALIGNED) q = reinterpret_cast<const pcm_pair*>(frame)[m];
  else q = pcm_pair{frame[2 * m], frame[2 * m + 1]};
  return cpx{(float)q.a, (float)q.b};
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

// ---- pass 1 (thread j of 64), after the R1 windowed points are in v: radix R1 over stride
//      64, twiddle, store ----
template <typename P, typename C>
ISI_HD void fft_pass1_finish(int j, C* v, const cpx* tws, C* zA) {
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
  auto o = split_ptr(zA + j);
#pragma unroll
  for (int p = 0; p < P::R1; ++p) put(o, P::kBlockPitch * p, v[p]);
}

// Forward kernel: window + pack the two frames of a pair (frame_a / frame_b point at their
// first staged sample; `win` already carries the untangle's 1/2).  PCM samples are integers:
// a power-of-two sample scale is folded into `win` (exact), any other one is applied per
// sample first (`PRESCALE`), so that both give bit for bit what float(x) * scale on the host
// followed by the FP32 kernel gives.
template <typename P, typename S, bool ALIGNED, bool PRESCALE>
ISI_HD void fft_pass1_pair_t(int j, const S* frame_a, const S* frame_b, float sample_scale,
                             const float* win, const cpx* tws, cpx2* zA) {
  cpx2 v[P::R1];
#pragma unroll
  for (int r = 0; r < P::R1; ++r) {
    const int m = j + 64 * r;
    const cpx a = load_pair<ALIGNED>(frame_a, m), b = load_pair<ALIGNED>(frame_b, m);
    const cpx w = reinterpret_cast<const cpx*>(win)[m];
    f2 re = mk2(a.re, b.re), im = mk2(a.im, b.im);
    if (PRESCALE) { re = mul2(re, bc(sample_scale)); im = mul2(im, bc(sample_scale)); }
    v[r].re = mul2(re, bc(w.re));
    v[r].im = mul2(im, bc(w.im));
  }
  fft_pass1_finish<P>(j, v, tws, zA);
}
template <typename P, typename S>
ISI_HD void fft_pass1_pair(int j, const S* frame_a, const S* frame_b, bool pair_aligned,
                           float sample_scale, const float* win, const cpx* tws, cpx2* zA) {
  if (sample_scale != 1.f) {      // rare: a PCM scale that is not a power of two
    if (pair_aligned) fft_pass1_pair_t<P, S, true, true>(j, frame_a, frame_b, sample_scale, win, tws, zA);
    else fft_pass1_pair_t<P, S, false, true>(j, frame_a, frame_b, sample_scale, win, tws, zA);
  } else {
    if (pair_aligned) fft_pass1_pair_t<P, S, true, false>(j, frame_a, frame_b, 1.f, win, tws, zA);
    else fft_pass1_pair_t<P, S, false, false>(j, frame_a, frame_b, 1.f, win, tws, zA);
  }
}

// sample scale folded into the window table (exact) / left to pass 1
ISI_HD bool is_pow2_scale(float s) {
  union { float f; uint32_t u; } v;
  v.f = s;
  const uint32_t e = (v.u >> 23) & 0xffu;
  return s > 0.f && (v.u & 0x007fffffu) == 0 && e != 0 && e != 0xffu;
}

// ---- pass 2: inside each 64-block, radix 16 over stride 4 ----
template <typename P, typename C>
ISI_HD void fft_pass2(int t, const cpx* tws, C* zA) {
  for (int item = t; item < 4 * P::R1; item += P::kFftThreads) {
    const int b = item % P::R1, j = item / P::R1;       // lanes along b: bank = b (pitch 65)
    C* blk = zA + P::kBlockPitch * b;
    C v[16];
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
    auto ob = split_ptr(blk + j);
#pragma unroll
    for (int p = 0; p < 16; ++p) put(ob, 4 * p, v[p]);
  }
}

// ---- pass 3: radix 4 on the quadruples (p1, p2) of zA, rewritten IN PLACE in natural bin
//      order (bin p1 + R1 p2 + 16 R1 p3).  Load + butterfly and store are two calls with the
//      transform group's barrier between them, because the natural-order slots overlap other
//      threads' quadruples.  Lanes run along p1, so the stores are contiguous. ----
// NT3 = threads sharing one transform's quadruples (64: the transform group itself; 128: a
// quarter of the warp-specialised kernel's polar/emit role, which takes this pass over).
template <typename P, typename C, int NT3 = P::kFftThreads>
struct Pass3Regs { C q[(P::M / 4) / NT3][4]; };

template <typename P, typename C, int NT3>
ISI_HD void fft_pass3_load(int t, const C* zA, Pass3Regs<P, C, NT3>& r) {
  constexpr int kStep = NT3 / P::R1;                     // p2 values covered per sweep
  const int p1 = t % P::R1;
#pragma unroll
  for (int i = 0; i < (P::M / 4) / NT3; ++i) {
    const int p2 = t / P::R1 + kStep * i;
    const C* q = zA + P::kBlockPitch * p1 + 4 * p2;
    r.q[i][0] = q[0]; r.q[i][1] = q[1]; r.q[i][2] = q[2]; r.q[i][3] = q[3];
    dft4(r.q[i][0], r.q[i][1], r.q[i][2], r.q[i][3]);
  }
}

template <typename P, typename C, int NT3>
ISI_HD void fft_pass3_store(int t, const Pass3Regs<P, C, NT3>& r, C* z) {
  constexpr int kStep = NT3 / P::R1;
  const int p1 = t % P::R1;
#pragma unroll
  for (int i = 0; i < (P::M / 4) / NT3; ++i) {
    const int p2 = t / P::R1 + kStep * i;
    auto o = split_ptr(z + p1 + P::R1 * p2);
    put(o, 0, r.q[i][0]); put(o, 16 * P::R1, r.q[i][1]); put(o, 32 * P::R1, r.q[i][2]); put(o, 48 * P::R1, r.q[i][3]);
  }
}

// ------------------------------------------------------------------------------------------
// One-warp plan for n_fft = 2048 (the warp-specialised kernel): M = 1024 = 32 x 32.  A warp owns
// a frame pair; each lane holds 32 complex points of the pair (128 registers), so the transform
// is TWO radix-32 passes with ONE exchange through shared memory, inside the warp (no block or
// group barrier, __syncwarp only):
//   pass A  lane j: window + pack the points m = j + 32 r, DFT-32 over r, twiddle W_M^(j p),
//           store element (p, j) at p * 33 + j  (lanes along j: contiguous)
//   pass B  lane p: load (p, j) for all j (stride 33: a quarter warp hits 8 bank groups),
//           DFT-32 over j, store bin p + 32 q in natural order (lanes along p: contiguous)
//   X[p + 32 q] = sum_j W_32^(j q) [ W_M^(j p) sum_r W_32^(r p) x[j + 32 r] ]
// Against the 16 x 16 x 4 plan above: 3 instead of 5 sweeps of the pair's 16 KB through shared
// memory, half the threads (one warp instead of two per pair), no barriers.
struct PlanW32 {
  static constexpr int N = 2048, M = 1024;
  static constexpr int kFftThreads = 32;
  static constexpr int kBlockPitch = 33;
  static constexpr int kPitchA = kBlockPitch * 32;      // 1056 >= M + 8 (band overrun)
};

// slot i of the shared twiddle table holds W_M^(j p), i = 32 p + j: index into the W_N^k table
ISI_HD int fft32_table_source(int i) { return 2 * (i % 32) * (i / 32); }

// forward 32-point DFT in registers, natural order in and out: even / odd 16-point halves, then
// X[k] = E[k] + W_32^k O[k], X[k + 16] = E[k] - W_32^k O[k]
template <typename C>
ISI_HD void dft32(C* v) {
  C e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
  dft16(e);
  dft16(o);
  // cos / sin of 2 pi k / 32, k = 0..15
  const float c[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                       0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                       0.19509032201612826785f, 0.f, -0.19509032201612826785f, -0.38268343236508977173f,
                       -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                       -0.92387953251128675613f, -0.98078528040323044913f};
  const float s[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                       0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                       0.98078528040323044913f, 1.f, 0.98078528040323044913f, 0.92387953251128675613f,
                       0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                       0.38268343236508977173f, 0.19509032201612826785f};
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    C t;
    if (k == 0) t = o[0];
    else if (k == 8) t = mul_neg_i(o[8]);
    else t = cmul(o[k], cpx{c[k], -s[k]});
    v[k] = cadd(e[k], t);
    v[k + 16] = csub(e[k], t);
  }
}

template <typename S, bool ALIGNED, bool PRESCALE>
ISI_HD void fft32_passA_t(int j, const S* frame_a, const S* frame_b, float sample_scale,
                          const float* win, const cpx* tws, cpx2* zA) {
  cpx2 v[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const int m = j + 32 * r;
    const cpx a = load_pair<ALIGNED>(frame_a, m), b = load_pair<ALIGNED>(frame_b, m);
    const cpx w = reinterpret_cast<const cpx*>(win)[m];
    f2 re = mk2(a.re, b.re), im = mk2(a.im, b.im);
    if (PRESCALE) { re = mul2(re, bc(sample_scale)); im = mul2(im, bc(sample_scale)); }
    v[r].re = mul2(re, bc(w.re));
    v[r].im = mul2(im, bc(w.im));
  }
  dft32(v);
#pragma unroll
  for (int p0 = 1; p0 < 32; p0 += 4) {          // four twiddle loads in flight at a time
    cpx t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) if (p0 + i < 32) t[i] = tws[32 * (p0 + i) + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) if (p0 + i < 32) v[p0 + i] = cmul(v[p0 + i], t[i]);
  }
  // The exchange is stored as two planes of 8-byte values (re pairs, im pairs): a packed
  // instruction leaves its result in an aligned register PAIR, which is exactly what an 8-byte
  // store takes.  A 16-byte store of (re, im) needs an aligned QUAD: ptxas copied every value
  // into one staging quad, and each copy waited for the previous store to read it -- a fifth of
  // the transform warps' stall samples.  Same bytes, same wavefronts (lanes along j: 256
  // contiguous bytes per store; pass B reads at stride 33 x 8 bytes: conflict-free per half warp).
  f2* re_plane = reinterpret_cast<f2*>(zA);
  f2* im_plane = re_plane + PlanW32::kPitchA;
#pragma unroll
  for (int p = 0; p < 32; ++p) {
    re_plane[PlanW32::kBlockPitch * p + j] = v[p].re;
    im_plane[PlanW32::kBlockPitch * p + j] = v[p].im;
  }
}
template <typename S>
ISI_HD void fft32_passA(int j, const S* frame_a, const S* frame_b, bool pair_aligned,
                        float sample_scale, const float* win, const cpx* tws, cpx2* zA) {
  if (sample_scale != 1.f) {      // rare: a PCM scale that is not a power of two
    if (pair_aligned) fft32_passA_t<S, true, true>(j, frame_a, frame_b, sample_scale, win, tws, zA);
    else fft32_passA_t<S, false, true>(j, frame_a, frame_b, sample_scale, win, tws, zA);
  } else {
    if (pair_aligned) fft32_passA_t<S, true, false>(j, frame_a, frame_b, 1.f, win, tws, zA);
    else fft32_passA_t<S, false, false>(j, frame_a, frame_b, 1.f, win, tws, zA);
  }
}

// pass B in two calls with a __syncwarp between them: the natural-order slots overlap the
// (p, j) slots other lanes of the warp still have to read
struct PassB32Regs { cpx2 u[32]; };
ISI_HD void fft32_passB_load(int p, const cpx2* zA, PassB32Regs& r) {
  const f2* re_row = reinterpret_cast<const f2*>(zA) + PlanW32::kBlockPitch * p;
  const f2* im_row = re_row + PlanW32::kPitchA;
#pragma unroll
  for (int j = 0; j < 32; ++j) { r.u[j].re = re_row[j]; r.u[j].im = im_row[j]; }
  dft32(r.u);
}
ISI_HD void fft32_passB_store(int p, const PassB32Regs& r, cpx2* z) {
  auto o = split_ptr(z + p);
#pragma unroll
  for (int q = 0; q < 32; ++q) put(o, 32 * q, r.u[q]);
}

// phase step folded into [-pi, pi] (round-to-nearest multiple of 2 pi; equals the
// numpy/magenta unwrap rule except on exact +-pi ties)
ISI_HD float wrap_step(float dd) {
  return fmaf(-kTwoPi, rintf(dd * kInvTwoPi), dd);
}

// ---- polar ----
// The wrapped phase advance of frame t is arg(X_t conj X_{t-1}); nothing cumulative is
// carried, because the mel IF only needs differences of the projected unwrapped phase, and
// the projection is linear:
//   mel_phase_t - mel_phase_{t-1} = sum_k w_k (u_t[k] - u_{t-1}[k]) = sum_k w_k step_t[k].
// Each spectrum value is reduced to its magnitude and its unit phasor X/|X| (one MUFU.RSQ
// serves both); an exact zero counts as 1+0i, whose phase is 0 like torch.angle(0): a bias far
// below the rounding noise of any non-zero spectrum is added to the real part, and |X|^2 is
// clamped to the square of that bias so the reciprocal square root stays finite.
constexpr float kZeroBias = 2.168404344971009e-19f;     // 2^-62
constexpr float kMinNorm2 = 4.70197740328915e-38f;      // 2^-124

struct Unit2 { f2 re, im; };                            // unit phasors of one bin, both frames

// |X| is left as its two factors (|X|^2 clamped, 1/|X|): the consumer adds eps with one FFMA2.
// (ptxas contracts a packed multiply feeding a packed add into FFMA2 when it likes -- even
// mul.rn.f32x2 + add.rn.f32x2 -- so every such pair in this file is an explicit fma2: results
// then do not depend on which instantiation the code was inlined into.)
struct Mag2 { f2 m2, rs; };
ISI_HD void unit_mag(cpx2 x, Unit2& u, Mag2& mag) {
  const f2 re = add2(x.re, bc(kZeroBias));
  f2 m2 = fma2(re, re, mul2(x.im, x.im));
  m2 = mk2(fmaxf(m2.x, kMinNorm2), fmaxf(m2.y, kMinNorm2));
  const f2 rs = mk2(fast_rsqrt(m2.x), fast_rsqrt(m2.y));
  u.re = mul2(re, rs);
  u.im = mul2(x.im, rs);
  mag.m2 = m2; mag.rs = rs;
}

// arg(u conj p) for unit phasors: (c, s) = (cos, sin) of the angle; the smaller of |c|, |s| is
// at most sin(pi/4), where x P(x^2) with a degree-7 minimax P gives asin to 1e-7 rad; then
// the octant.  No division, no reciprocal; the polynomial runs on both frames at once.
ISI_HD float octant_fix(float r, float c, float s) {
  if (fabsf(s) > fabsf(c)) r = kHalfPi - r;
  if (c < 0.f) r = kPi - r;
  return copysignf(r, s);
}
ISI_HD f2 unit_step(Unit2 u, Unit2 p) {
  const f2 c = fma2(u.im, p.im, mul2(u.re, p.re));
  const f2 s = fma2(u.im, p.re, mul2(neg2(u.re), p.im));
  const f2 mn = mk2(fminf(fabsf(c.x), fabsf(s.x)), fminf(fabsf(c.y), fabsf(s.y)));
  const f2 t = mul2(mn, mn);
  f2 q = bc(0.1308550089597702f);
  q = fma2(q, t, bc(-0.1294308602809906f));
  q = fma2(q, t, bc(0.1047358289361f));
  q = fma2(q, t, bc(0.005715840496122837f));
  q = fma2(q, t, bc(0.04870210215449333f));
  q = fma2(q, t, bc(0.0746518149971962f));
  q = fma2(q, t, bc(0.16667994856834412f));
  q = fma2(q, t, bc(0.9999998807907104f));
  const f2 r = mul2(q, mn);
  return mk2(octant_fix(r.x, c.x, s.x), octant_fix(r.y, c.y, s.y));
}

// Untangle work item `it` (0..M/2-1) of one pair's natural-order half-scale spectrum z[0..M]
// (the window table carries the 1/2 of (A + conj B)/2).  Item it>0 owns bins it and M-it.
// Item 0 owns bin M/2 and ONE of the two purely real bins: `real_bin` = M when the DC bin is
// the dropped one, else 0 (the other one is never read by emit and keeps its raw FFT value).
// Item 0 runs the same instruction stream as every other item -- only operands are selected
// -- so the warp that holds it does not execute a second, divergent path.  MAYBE_ZERO = false
// compiles the selects out for items that cannot be item 0.
template <typename P, bool MAYBE_ZERO>
ISI_HD void untangle(int it, const cpx2* z, cpx w /* W_N^it */, int real_bin, cpx2& xa, cpx2& xb) {
  constexpr int M = P::M;
  const bool special = MAYBE_ZERO && it == 0;
  const cpx2 a = z[it], b = z[special ? M / 2 : M - it];
  cpx2 e, d;
  e.re = add2(a.re, b.re); e.im = sub2(a.im, b.im);                           // (A + conj B)/2
  d.re = sub2(a.re, b.re); d.im = add2(a.im, b.im);                           // (A - conj B)/2
  const cpx2 p = cmul(mul_neg_i(d), w);                                       // W_N^k (-i) d
  xa = cadd(e, p);
  xb.re = sub2(e.re, p.re); xb.im = sub2(p.im, e.im);                         // conj(e - p)
  if (special) {
    xa.re = add2(b.re, b.re); xa.im = neg2(add2(b.im, b.im));                 // X[M/2] = conj Z[M/2]
    const f2 t = real_bin == 0 ? add2(a.re, a.im) : sub2(a.re, a.im);         // X[0] or X[M]
    xb.re = add2(t, t); xb.im = bc(0.f);
  }
}

// What polar leaves per bin for emit: mel mode (|X|+eps)^2 and the phase step; linear mode
// log2(|X|+eps) and the phase step (finish_row turns log2 into log and radians into half-turns).
template <bool MEL>
ISI_HD cpx2 polar_value(Mag2 mag, f2 step, float eps) {
  cpx2 r;
  const f2 a = fma2(mag.m2, mag.rs, bc(eps));          // |X| + eps
  if (MEL) r.re = mul2(a, a);
  else r.re = mk2(fast_log2(a.x), fast_log2(a.y));
  r.im = step;
  return r;
}

// State of one bin across batches: the unit phasors of the batch's last pair (lane y = the
// batch's last frame).
struct BinState { Unit2 a, b; };
ISI_HD BinState bin_state_init() {
  BinState s;
  s.a.re = bc(1.f); s.a.im = bc(0.f); s.b = s.a;
  return s;
}

// ---- polar of one work item over a whole batch of NP pairs.  Frame slot s of the batch is
//      lane s / NP of pair s % NP (pair q = frames q and q + NP), so the predecessor of both
//      lanes of pair q > 0 is pair q - 1, and the predecessors of pair 0 are the previous
//      batch's last frame (state, lane y) and this batch's frame NP - 1 (last pair, lane x):
//      the only lane shuffle of the batch.  z[q * pitch + k] <- (v0, v1): mel mode
//      (|X|+eps)^2 and the phase step, linear mode log(|X|+eps) and the IF.
//      `seed_only`: a look-back transform sits in the last pair's lane y; only the state is
//      updated. ----
template <typename P, bool MEL, int NP, bool MAYBE_ZERO>
ISI_HD void polar_item(int it, cpx2* z, int pitch, cpx w, int real_bin, bool seed_only, float eps,
                       BinState& st) {
  constexpr int M = P::M;
  const bool special = MAYBE_ZERO && it == 0;
  const int ka = special ? M / 2 : it, kb = special ? real_bin : M - it;
  cpx2 xa, xb;
  Unit2 la, lb;          // last pair
  Mag2 mla, mlb;
  untangle<P, MAYBE_ZERO>(it, z + (NP - 1) * pitch, w, real_bin, xa, xb);
  unit_mag(xa, la, mla);
  unit_mag(xb, lb, mlb);
  if (!seed_only) {
    Unit2 pa, pb;        // predecessors of the pair being processed
    pa.re = mk2(st.a.re.y, la.re.x); pa.im = mk2(st.a.im.y, la.im.x);
    pb.re = mk2(st.b.re.y, lb.re.x); pb.im = mk2(st.b.im.y, lb.im.x);
#pragma unroll
    for (int q = 0; q < NP - 1; ++q) {
      Unit2 ua, ub;
      Mag2 ma, mb;
      untangle<P, MAYBE_ZERO>(it, z + q * pitch, w, real_bin, xa, xb);
      unit_mag(xa, ua, ma);
      unit_mag(xb, ub, mb);
      put(split_ptr(z + q * pitch + ka), 0, polar_value<MEL>(ma, unit_step(ua, pa), eps));
      put(split_ptr(z + q * pitch + kb), 0, polar_value<MEL>(mb, unit_step(ub, pb), eps));
      pa = ua; pb = ub;
    }
    put(split_ptr(z + (NP - 1) * pitch + ka), 0, polar_value<MEL>(mla, unit_step(la, pa), eps));
    put(split_ptr(z + (NP - 1) * pitch + kb), 0, polar_value<MEL>(mlb, unit_step(lb, pb), eps));
  }
  st.a = la; st.b = lb;
}

// ---- emit: one output row, all FB = 2 NP frames of the batch at once; pair q's values sit at
//      z[q * pitch + bin].  `bin0` is the FFT bin of the row (linear mode) or of the first band
//      element (mel mode); the mel weights are zero beyond the band's length and enter as
//      broadcast operands; `count_uniform` (the longest band of the warp) skips whole taps
//      without divergence.  Reads run up to kMaxMelWidth - 1 elements past the band (zero
//      weight): Plan::kPitchA keeps them inside the transform's region. ----
constexpr int kMaxMelWidth = 8;

// Both produce, per pair, log2 of the (mel) magnitude term and the (mel) phase step in radians.
template <int NP>
ISI_HD void emit_linear(const cpx2* z, int pitch, int bin0, f2* lg, f2* ph) {
#pragma unroll
  for (int q = 0; q < NP; ++q) { const cpx2 v = z[q * pitch + bin0]; lg[q] = v.re; ph[q] = v.im; }
}

template <int NP>
ISI_HD void emit_mel(const cpx2* z, int pitch, int bin0, int count_uniform, const float* w,
                     bool first_is_frame0, float eps, f2* lg, f2* ph) {
  f2 m2[NP], mp[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) { m2[q] = bc(0.f); mp[q] = bc(0.f); }
  // two taps per (warp-uniform) step: all 2*NP loads are issued before the first use; an odd
  // longest band ends with a single tap
#pragma unroll
  for (int i = 0; i < kMaxMelWidth; i += 2) {
    if (i + 1 < count_uniform) {
      cpx2 va[NP], vb[NP];
#pragma unroll
      for (int q = 0; q < NP; ++q) { va[q] = z[q * pitch + bin0 + i]; vb[q] = z[q * pitch + bin0 + i + 1]; }
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        m2[q] = fma2(vb[q].re, bc(w[i + 1]), fma2(va[q].re, bc(w[i]), m2[q]));
        mp[q] = fma2(vb[q].im, bc(w[i + 1]), fma2(va[q].im, bc(w[i]), mp[q]));
      }
    } else if (i < count_uniform) {
      cpx2 va[NP];
#pragma unroll
      for (int q = 0; q < NP; ++q) va[q] = z[q * pitch + bin0 + i];
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        m2[q] = fma2(va[q].re, bc(w[i]), m2[q]);
        mp[q] = fma2(va[q].im, bc(w[i]), mp[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const f2 a = add2(m2[q], bc(eps));
    lg[q] = mk2(fast_log2(a.x), fast_log2(a.y));
    // fold into [-pi, pi]: rint by the 1.5 * 2^23 trick (|step| / 2 pi is far below 2^22)
    const f2 turns = sub2(fma2(mp[q], bc(kInvTwoPi), bc(12582912.f)), bc(12582912.f));
    ph[q] = fma2(turns, bc(-kTwoPi), mp[q]);
    if (q == 0 && first_is_frame0) ph[q].x = mp[q].x;
  }
}

// value of frame slot s (pair s % NP, lane s / NP)
template <int NP>
ISI_HD float slot_value(const f2* v, int s) { return s < NP ? v[s].x : v[s - NP].y; }

// ---- row epilogue, per frame slot (scalar, so that each result is produced in the register
//      its 16-byte store needs): log2 -> log, radians -> half-turns, then the fused
//      masked-phase transform and the per-channel affine normalisation:
//        v0 = log * s0 + b0;  v1 = (mask && log < threshold ? 0 : IF) * s1 + b1 ----
template <int NP>
ISI_HD void finish_row(const f2* lg, const f2* ph, bool mask_phase, float mask_threshold, float s0,
                       float b0, float s1, float b1, float* v0, float* v1) {
  const float k0 = kLn2 * s0, k1 = kInvPi * s1;
#pragma unroll
  for (int s = 0; s < 2 * NP; ++s) {
    const float l = slot_value<NP>(lg, s);
    float p = slot_value<NP>(ph, s);
    if (mask_phase && l * kLn2 < mask_threshold) p = 0.f;
    v0[s] = fmaf(l, k0, b0);
    v1[s] = fmaf(p, k1, b1);
  }
}

}  // namespace melif
}  // namespace isi
