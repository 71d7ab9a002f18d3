// Fused front end: STFT -> (mel projection) -> log-magnitude + instantaneous frequency.
//
// Replaces SpectrogramsHelper / MelSpectrogramsHelper.to_spectrogram (external
// GANsynth_pytorch; reference call sites utils/misc.py:10-29, extract_code.py:199-206,
// train_vqvae.py:604-611).
//
// One CTA owns a run of frames of one note (a whole note when the batch alone fills the
// GPU, else a segment -- the only cross-frame state is the previous frame's spectrum, which
// a segment recomputes with one look-back transform) and walks it in batches of FB frames.
// Per batch the audio span (FB-1)*hop + n_fft is brought into shared memory by ONE bulk
// async copy (cp.async.bulk, the 1-D TMA path, completion on an mbarrier) issued a whole
// batch ahead, so HBM latency hides behind the previous batch's transform.  The audio is FP32
// or 16-bit PCM (converted in pass 1).  Twiddles and the window sit in shared memory; the mel
// band constants of a thread's rows are re-read from L1 each batch (kept in registers across
// the transform they spill); the complex spectrum, magnitudes and phase steps never leave
// shared memory / registers.  HBM traffic is the audio once (frame overlap is served from
// shared memory) and the final tensor, FB consecutive time steps per row at a time, in one of
// four layouts (planes, channels-last, 2x2 space-to-depth blocks time- or frequency-fastest),
// as 32-byte stores where the layout allows.  Two kernels: the generic one below (every thread
// runs every phase; one FB-frame buffer is its whole working set) and, for the NSynth shape, the
// warp-specialised one further down.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "melif_core.cuh"

namespace isi {
using namespace melif;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Waits of the warp-specialised kernel.  A plain try_wait loop polls (the polar/emit warps,
// which wait for the transform warps every batch, polled ~67 times per wait, each poll a SYNCS
// plus a reload of the spilled barrier address).  Here the barrier is picked by a uniform
// branch (address = uniform base + immediate, nothing to spill) and try_wait carries a
// suspend-time hint, which ptxas turns into TRYWAIT / NANOSLEEP.SYNCS / re-check: the warp
// sleeps until the barrier's phase flips.  A failed try backs off with nanosleep.
// cfg = hint_ns | sleep_ns << 16.  (Measured: neutral for the kernel time -- the polls only
// used idle issue slots -- but the ncu instruction and LSU counts no longer carry the noise.)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t cfg) {
  const uint32_t hint_ns = cfg & 0xffffu, sleep_ns = cfg >> 16;
  while (!mbar_try_wait_hint(bar, parity, hint_ns))
    if (sleep_ns) __nanosleep(sleep_ns);
}
__device__ __forceinline__ void mbar_wait_pair(uint64_t* bars, uint32_t which, uint32_t parity, uint32_t cfg) {
  if (which) mbar_wait_backoff(bars + 1, parity, cfg);
  else mbar_wait_backoff(bars, parity, cfg);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Cache policy: the band tables (40 KB, re-read by every CTA every batch) should stay in L1,
// the output stream (1 MB per note, written once) should not displace them.
__device__ __forceinline__ float4 ld_table4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_table(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.L1::evict_last.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// 32-byte store (sm_100: STG.256): a row's 8 time steps of one channel in ONE instruction.  The
// cost of a store instruction follows the number of 128-byte lines its lanes touch (measured:
// writing the same bytes into an L2-resident buffer costs the same, so it is SM-side), and in
// the time-fastest layouts every lane is in a line of its own: halving the instructions halved
// the store time (planar output: 0.387 -> 0.311 ms per 444 notes).
__device__ __forceinline__ void st_stream8(float* p, const float* v) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- pieces shared by the two kernels below ----

// Band constants of one output row (first bin incl. the dropped-DC offset, length, weights),
// read from the L1/L2-resident tables.
template <bool MEL>
struct RowBand { int row, bin, cnt; float w[MEL ? kMaxMelWidth : 1]; };

template <bool MEL>
__device__ __forceinline__ RowBand<MEL> load_row_band(const isi_melif_params& p, int row, int dc, bool w_vec) {
  RowBand<MEL> b;
  b.row = row;
  b.bin = row + dc;
  b.cnt = 0;
  if (MEL) {
    b.bin = ld_table(p.mel_start + row) + dc;
    b.cnt = ld_table(p.mel_count + row);
    if (w_vec) {
      const float4* wt = reinterpret_cast<const float4*>(p.mel_weight) + (int64_t)row * 2;
      const float4 a = ld_table4(wt), c = ld_table4(wt + 1);
      b.w[0] = a.x; b.w[1] = a.y; b.w[2] = a.z; b.w[3] = a.w;
      b.w[4 % (MEL ? 8 : 1)] = c.x; b.w[5 % (MEL ? 8 : 1)] = c.y;
      b.w[6 % (MEL ? 8 : 1)] = c.z; b.w[7 % (MEL ? 8 : 1)] = c.w;
    } else {
#pragma unroll
      for (int i = 0; i < kMaxMelWidth; ++i)
        b.w[i % (MEL ? 8 : 1)] = (i < p.mel_width) ? __ldg(p.mel_weight + (int64_t)row * p.mel_width + i) : 0.f;
    }
  }
  return b;
}

// One output row of a batch: projection, log / wrap, fused epilogue, and the stores of the FB
// time steps in the requested layout.  v0 / v1 are produced per frame slot by scalar FFMAs, so
// each 16-byte store finds its four values in consecutive registers.
template <typename P, int FB, bool MEL>
__device__ __forceinline__ void emit_row(const isi_melif_params& p, const cpx2* zA, const RowBand<MEL>& band,
                                         int note_idx, int f0, int nf, float eps, float* out) {
  constexpr int M = P::M, NP = FB / 2;
  const int row = band.row;
  f2 lg[NP], ph[NP];
  if (MEL) {
    const int cnt_warp = __reduce_max_sync(0xffffffffu, band.cnt);
    emit_mel<NP>(zA, P::kPitchA, band.bin, cnt_warp, band.w, f0 == 0, eps, lg, ph);
  } else {
    emit_linear<NP>(zA, P::kPitchA, band.bin, lg, ph);
  }
  float v0[FB], v1[FB];      // frame-slot order
  finish_row<NP>(lg, ph, p.mask_phase != 0, p.mask_threshold, p.out_scale[0], p.out_bias[0],
                 p.out_scale[1], p.out_bias[1], v0, v1);
  if (p.channels_last == ISI_SPEC_SPACE_TO_DEPTH_T) {
    // [B, T'/2, F/2, (f&1, t&1, channel)]: the same 2x2 blocks with the FREQUENCY index running
    // fastest.  Lanes are consecutive rows, so one store instruction writes 512 contiguous bytes
    // (4 LSU wavefronts); in the time-fastest layout below the same instruction touches 16
    // different 128-byte lines (16 wavefronts, a fifth of the kernel's time in stores).
    const int64_t blk = (int64_t)(M / 2) * 8;               // one time block of all row pairs
    float* d = out + (((int64_t)note_idx * (p.n_frames >> 1) + (f0 >> 1)) * (M / 2) + (row >> 1)) * 8 +
               (row & 1) * 4;
    if (nf == FB) {
#pragma unroll
      for (int k = 0; k < FB / 2; ++k)
        st_stream4(d + k * blk, make_float4(v0[2 * k], v1[2 * k], v0[2 * k + 1], v1[2 * k + 1]));
    } else {
#pragma unroll
      for (int s = 0; s < FB; ++s)
        if (s < nf) {
          float* e = d + (s >> 1) * blk + (s & 1) * 2;
          e[0] = v0[s]; e[1] = v1[s];
        }
    }
    return;
  }
  if (p.channels_last == ISI_SPEC_SPACE_TO_DEPTH) {
    // [B, F/2, T'/2, (f&1, t&1, channel)]: 2x2 spectrogram blocks as 8 channels.  A row
    // writes 16 bytes per block; the odd/even row pair (neighbouring lanes of the same
    // instruction) completes each 32-byte sector.
    float* d = out + ((((int64_t)note_idx * (M / 2) + (row >> 1)) * (p.n_frames >> 1) + (f0 >> 1)) * 8) +
               (row & 1) * 4;
    bool wide = false;
    if constexpr (FB == 8) {
      // A whole batch: the two lanes of a row pair (rows follow lanes, so lane parity = row
      // parity) swap half of their pieces -- the even lane ends up with blocks 0 and 1 of BOTH
      // rows, the odd lane with blocks 2 and 3 -- and each writes two 32-byte blocks: two store
      // instructions per row instead of four (every lane pair is in a 128-byte line of its own,
      // and a store instruction costs LSU cycles per line touched).
      wide = nf == FB && (reinterpret_cast<uintptr_t>(out) & 31) == 0;       // warp-uniform
      if (wide) {
        const bool odd = (row & 1) != 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float mine[4] = {v0[2 * j], v1[2 * j], v0[2 * j + 1], v1[2 * j + 1]};                  // block j
          const float far[4] = {v0[2 * j + 4], v1[2 * j + 4], v0[2 * j + 5], v1[2 * j + 5]};           // block j + 2
          float blk[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float got = __shfl_xor_sync(0xffffffffu, odd ? mine[e] : far[e], 1);
            blk[e] = odd ? got : mine[e];          // parity-0 row's piece
            blk[4 + e] = odd ? far[e] : got;       // parity-1 row's piece
          }
          st_stream8(d - (odd ? 4 : 0) + 8 * (odd ? j + 2 : j), blk);
        }
      }
    }
    if (wide) {
    } else if (nf == FB) {
#pragma unroll
      for (int k = 0; k < FB / 2; ++k)
        st_stream4(d + 8 * k, make_float4(v0[2 * k], v1[2 * k], v0[2 * k + 1], v1[2 * k + 1]));
    } else {
#pragma unroll
      for (int s = 0; s < FB; ++s)
        if (s < nf) {
          float* e = d + (s >> 1) * 8 + (s & 1) * 2;
          e[0] = v0[s]; e[1] = v1[s];
        }
    }
    return;
  }
  if (p.channels_last) {
    // [B, F, T', 2]: the FB time steps of both channels are one contiguous run
    float* d = out + (((int64_t)note_idx * M + row) * p.n_frames + f0) * 2;
    bool wide = false;
    if constexpr (FB == 8) {
      wide = nf == FB && (p.n_frames % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 31) == 0;
      if (wide) {
        const float a[8] = {v0[0], v1[0], v0[1], v1[1], v0[2], v1[2], v0[3], v1[3]};
        const float b[8] = {v0[4], v1[4], v0[5], v1[5], v0[6], v1[6], v0[7], v1[7]};
        st_stream8(d, a);
        st_stream8(d + 8, b);
      }
    }
    if (wide) {
    } else if (nf == FB && (p.n_frames % 2 == 0)) {
#pragma unroll
      for (int k = 0; k < FB / 2; ++k)
        st_stream4(d + 4 * k, make_float4(v0[2 * k], v1[2 * k], v0[2 * k + 1], v1[2 * k + 1]));
    } else {
#pragma unroll
      for (int s = 0; s < FB; ++s)
        if (s < nf) { d[2 * s] = v0[s]; d[2 * s + 1] = v1[s]; }
    }
    return;
  }
  float* d0 = out + (int64_t)note_idx * 2 * M * p.n_frames + (int64_t)row * p.n_frames + f0;
  float* d1 = d0 + (int64_t)M * p.n_frames;
  bool wide = false;
  if constexpr (FB == 8) {
    wide = nf == FB && (p.n_frames % 8 == 0) && (reinterpret_cast<uintptr_t>(out) & 31) == 0;
    if (wide) {
      st_stream8(d0, v0);
      st_stream8(d1, v1);
    }
  }
  if (wide) {
  } else if (nf == FB && (FB % 4 == 0) && (p.n_frames % 4 == 0)) {
#pragma unroll
    for (int k = 0; k < FB / 4; ++k) {
      st_stream4(d0 + 4 * k, make_float4(v0[4 * k], v0[4 * k + 1], v0[4 * k + 2], v0[4 * k + 3]));
      st_stream4(d1 + 4 * k, make_float4(v1[4 * k], v1[4 * k + 1], v1[4 * k + 2], v1[4 * k + 3]));
    }
  } else {
#pragma unroll
    for (int s = 0; s < FB; ++s)
      if (s < nf) { d0[s] = v0[s]; d1[s] = v1[s]; }
  }
}

// Stage the audio span of frames [frame, frame + nfr) of `note`: threads t of nt zero-fill what
// lies outside the note (`fill`), ONE thread sends one bulk copy for the rest (`issue`;
// completion on `bar`).  The zero-fill must be ordered before the issuer's arrive (a CTA / role
// barrier, or __syncwarp when the filling threads are the issuer's warp).
template <typename S>
struct StageSpan {
  int span; int64_t s0, lo, hi;
  __device__ StageSpan(int64_t n_samples, int hop, int pad_left, int n_fft, int frame, int nfr) {
    span = (nfr - 1) * hop + n_fft;
    s0 = (int64_t)frame * hop - pad_left;
    lo = s0 < 0 ? -s0 : 0;                                  // first valid index
    hi = n_samples - s0;                                    // one past the last valid index
    hi = hi < 0 ? 0 : (hi > span ? span : hi);
  }
  __device__ void fill(S* stage, int t, int nt) const {
    const int64_t vlo = lo < hi ? lo : hi;
    for (int i = t; i < vlo; i += nt) stage[i] = S(0);
    for (int i = (int)hi + t; i < span; i += nt) stage[i] = S(0);
  }
  __device__ void issue(S* stage, const S* note, uint64_t* bar) const {
    if (hi > lo) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t bytes = (uint32_t)(hi - lo) * (uint32_t)sizeof(S);
      mbar_expect_tx(bar, bytes);
      bulk_g2s(stage + lo, note + s0 + lo, bytes, bar);
    } else {
      mbar_arrive(bar);
    }
  }
};
template <typename S>
__device__ __forceinline__ void stage_span_bulk(S* stage, const S* note, int64_t n_samples, int hop, int pad_left,
                                                int n_fft, int frame, int nfr, int t, int nt, bool issuer,
                                                uint64_t* bar) {
  const StageSpan<S> sp(n_samples, hop, pad_left, n_fft, frame, nfr);
  sp.fill(stage, t, nt);
  if (issuer) sp.issue(stage, note, bar);
}

struct MelifSmem {
  int tw, win, stage, za, bar, total;   // byte offsets into dynamic shared memory
};

template <int NFFT, int FB>
__host__ __device__ inline MelifSmem melif_smem_layout(int hop, int sample_bytes, int n_buffers = 1,
                                                        int pitch_a = Plan<NFFT>::kPitchA) {
  MelifSmem s;
  int off = 0;
  s.tw = off;    off += (NFFT / 2) * 8;                     // FFT twiddles (fft_table_source)
  s.win = off;   off += NFFT * 4;
  s.stage = off; off += n_buffers * ((((FB - 1) * hop + NFFT) * sample_bytes + 15) / 16 * 16);
  s.za = off;    off += n_buffers * (FB / 2) * pitch_a * 16;      // FFT workspace, spectrum, polar values
  s.bar = off;   off += 64;
  s.total = off;
  return s;
}

// CTAs per SM each instantiation is built for (registers: 65536 / (NT * CTAS))
__host__ __device__ constexpr int melif_ctas_per_sm(int nt) { return nt >= 256 ? 2 : (nt >= 128 ? 4 : 8); }

// ------------------------------------------------------------------------------------------
// Generic kernel: every thread runs every phase (n_fft 512 / 1024 / 2048, any hop, any
// alignment; the synchronous staging path when the audio is not 16-byte copyable).
// ------------------------------------------------------------------------------------------
template <int NFFT, int FB, int NT, bool MEL, typename S>
__global__ void __launch_bounds__(NT, melif_ctas_per_sm(NT))
melif_kernel(const S* __restrict__ audio, int64_t n_samples, isi_melif_params p,
             float* __restrict__ out, int bulk_ok, int seg_frames, int n_segs) {
  using P = Plan<NFFT>;
  constexpr int M = P::M;
  constexpr int NP = FB / 2;                  // frame pairs per batch: pair q = slots q, q + NP
  constexpr int IPT = (M / 2) / NT;           // polar work items per thread
  constexpr int RPT = M / NT;                 // output rows per thread
  constexpr int kGroups = NT / 64;            // pairs transformed concurrently
  static_assert(FB % 2 == 0 && NP >= 2, "frames are processed in pairs");
  static_assert(IPT >= 1 && (M / 2) % NT == 0 && NT % 64 == 0, "bad thread count");
  static_assert(P::kPitchA >= M + kMaxMelWidth, "a pair's region must hold bins 0..M and the band overrun");
  extern __shared__ __align__(128) unsigned char smem[];
  const MelifSmem L = melif_smem_layout<NFFT, FB>(p.hop, (int)sizeof(S));
  cpx* twm = reinterpret_cast<cpx*>(smem + L.tw);
  float* win = reinterpret_cast<float*>(smem + L.win);
  S* stage = reinterpret_cast<S*>(smem + L.stage);
  cpx2* zA = reinterpret_cast<cpx2*>(smem + L.za);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);

  const int tid = threadIdx.x;
  const int note_idx = blockIdx.x / n_segs, seg = blockIdx.x - note_idx * n_segs;
  const int fs = seg * seg_frames;                          // first frame of this CTA
  const int fe = min(p.n_frames, fs + seg_frames);          // one past its last frame
  const S* note = audio + (int64_t)note_idx * n_samples;
  const int dc = p.drop_dc ? 1 : 0;
  const float eps = p.safelog_eps;
  const bool pairs_aligned = (p.hop % 2) == 0;   // a frame starts on a sample-pair boundary

  // ---- one-time setup: tables to shared memory, per-thread constants to registers ----
  const cpx* tw_global = reinterpret_cast<const cpx*>(p.twiddle);     // W_N^j, j < N
  for (int i = tid; i < M; i += NT) twm[i] = tw_global[fft_table_source<P>(i)];
  // the window carries the untangle's 1/2 and, for PCM input, a power-of-two sample scale
  const bool fold_scale = sizeof(S) == 2 && is_pow2_scale(p.pcm_scale);
  const float win_scale = 0.5f * (fold_scale ? p.pcm_scale : 1.f);
  const float sample_scale = (sizeof(S) == 2 && !fold_scale) ? p.pcm_scale : 1.f;
  for (int i = tid; i < NFFT; i += NT) win[i] = p.window[i] * win_scale;
  // band reads may run past a band into never-written padding: make it finite once
  for (int i = tid; i < NP * P::kPitchA; i += NT) { zA[i].re = bc(0.f); zA[i].im = bc(0.f); }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cpx w_item[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) w_item[i] = tw_global[tid + i * NT];
  const bool w_vec = MEL && p.mel_width == kMaxMelWidth && (reinterpret_cast<uintptr_t>(p.mel_weight) & 15) == 0;
  BinState st[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) st[i] = bin_state_init();
  __syncthreads();

  auto stage_span = [&](int frame, int nfr) {
    if (!bulk_ok) {
      stage_fill(tid, NT, stage, (nfr - 1) * p.hop + NFFT, note, n_samples, (int64_t)frame * p.hop - p.pad_left);
      return;
    }
    stage_span_bulk(stage, note, n_samples, p.hop, p.pad_left, NFFT, frame, nfr, tid, NT, tid == 0, bar);
  };

  // Batch -1 is the look-back transform of frame fs-1: it only seeds the phase-step state.
  const int n_batches = (fe - fs + FB - 1) / FB;
  const int b_begin = fs > 0 ? -1 : 0;
  stage_span(fs > 0 ? fs - 1 : fs, fs > 0 ? 1 : min(FB, fe - fs));
  uint32_t stage_phase = 0;

  for (int b = b_begin; b < n_batches; ++b) {
    const bool lookback = b < 0;
    const int f0 = lookback ? fs - 1 : fs + b * FB;
    const int nf = lookback ? 1 : min(FB, fe - f0);
    const int next_f0 = fs + (b + 1) * FB;
    const int next_nf = (b + 1 < n_batches) ? min(FB, fe - next_f0) : 0;

    if (bulk_ok) { mbar_wait(bar, stage_phase & 1); ++stage_phase; }
    __syncthreads();       // zero-filled pads / synchronous fills come from other threads;
                           // also fences the previous batch's emit from this pass 1
    // The three FFT passes of a pair only involve the 64 threads of its group: they meet
    // on a named barrier of their own instead of stalling the whole CTA.  Slots >= nf of a
    // ragged last batch transform whatever the stage holds (finite; never emitted).  The
    // look-back frame is transformed as lane y of the last pair, where the state lives.
    const uint32_t group_bar = 1 + (tid >> 6);
    for (int q = tid / 64; q < NP; q += kGroups) {
      if (lookback ? (q != NP - 1) : (q >= nf)) continue;
      const S* fa = stage + (lookback ? 0 : q * p.hop);
      const S* fb = stage + (lookback ? 0 : (q + NP) * p.hop);
      cpx2* z = zA + q * P::kPitchA;
      fft_pass1_pair<P>(tid & 63, fa, fb, pairs_aligned, sample_scale, win, twm, z);
      asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
      fft_pass2<P>(tid & 63, twm, z);
      asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
      Pass3Regs<P, cpx2> regs;
      fft_pass3_load<P>(tid & 63, z, regs);
      asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
      fft_pass3_store<P>(tid & 63, regs, z);
    }
    __syncthreads();
    if (next_nf > 0 && bulk_ok) stage_span(next_f0, next_nf);   // every group is done with the stage
    // polar: the whole batch of a work item, previous phasors in registers, in place
#pragma unroll
    for (int i = 0; i < IPT; ++i)
      if (i == 0)
        polar_item<P, MEL, NP, true>(tid, zA, P::kPitchA, w_item[0], dc ? M : 0, lookback, eps, st[0]);
      else
        polar_item<P, MEL, NP, false>(tid + i * NT, zA, P::kPitchA, w_item[i], 0, lookback, eps, st[i]);
    if (!lookback) {
      // Per-row band constants are NOT kept in registers across the transform and polar
      // (they would spill): each batch re-reads them from the L1-resident tables just
      // before the barrier that precedes emit.
      RowBand<MEL> band[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) band[r] = load_row_band<MEL>(p, tid + r * NT, dc, w_vec);
      __syncthreads();
      // emit: all FB time steps of a row at once
#pragma unroll
      for (int r = 0; r < RPT; ++r)
        emit_row<P, FB, MEL>(p, zA, band[r], note_idx, f0, nf, eps, out);
    }
    if (!bulk_ok && next_nf > 0) { __syncthreads(); stage_span(next_f0, next_nf); }
  }
}

// ------------------------------------------------------------------------------------------
// Warp-specialised kernel (n_fft 2048, bulk-copyable audio, even hop): the NSynth shape.
//
// One PERSISTENT CTA per SM (it walks its work items, see the kernel), two roles that run
// CONCURRENTLY on consecutive batches of FB = 8 frames through a double-buffered workspace:
//   * 4 transform warps, ONE WARP PER FRAME PAIR (melif_core.cuh, PlanW32): each lane holds 32
//     complex points of the pair (128 of its 224 registers, setmaxnreg.inc), so the 1024-point
//     transform is two radix-32 passes with one exchange through shared memory and no barrier
//     but __syncwarp;
//   * 16 polar/emit warps (512 threads, one untangle item and two output rows each): polar
//     of the batch the transform warps finished last, then the mel projection, log / wrap,
//     epilogue and stores; 64 registers (setmaxnreg.dec).
// The generic kernel gives every thread the transform's register budget, which caps an SM at
// 16 warps; here the registers are split by need.  The split must CONSERVE the CTA's launch
// allocation (640 threads x 96 registers = 60 Ki): the transform warps' setmaxnreg.inc spins
// until the pool holds what the polar/emit warps' setmaxnreg.dec released,
// 4 x 32 x (224 - 96) = 16 x 32 x (96 - 64).  Hand-off: full[buf] / empty[buf] mbarriers
// (transform -> polar/emit -> transform), suspended waits.
// The previous plan (16 x 16 x 4 in three passes, 8 transform warps in groups of 64 with named
// barriers, 112 / 64 registers) is still built: ISI_MELIF_WS_PLAN=3, and ISI_MELIF_WS_FB=4 (two
// CTAs of 384 threads per SM) -- measurement knobs.  Measured per 444 notes: 0.299 ms (three-pass)
// against 0.277 ms; moving the last pass to the polar/emit role had made THAT role the critical
// path (0.373 ms).
// ------------------------------------------------------------------------------------------
// Geometry of one instantiation.  (Three-pass plan: FB = 8 is one CTA of 768 threads per SM, FB = 4
// two CTAs of 384 -- the same 24 warps per SM; two CTAs fill each other's pipeline fill / drain and
// set-up bubbles, at twice the per-batch fixed work.)
// W32 = the one-warp transform plan (melif_core.cuh, PlanW32): 4 transform warps (one per frame
// pair, 32 points per lane: 224 registers) + 16 polar/emit warps at 64 registers, 640 threads
// launched at 96 registers: 4 x 32 x (224 - 96) = 16 x 32 x (96 - 64).  The transform warps are
// the critical path and schedule better with room to spare (measured per 444 notes, same box:
// 160 / 80 0.297 ms, 192 / 72 0.293, 224 / 64 0.283, 256 / 56 0.289).
template <int FB, bool W32>
struct WsGeometry {
  static constexpr int kPairs = FB / 2;
  static constexpr int kFftThreads = (W32 ? 32 : 64) * kPairs;    // a warp / a group of 64 per frame pair
  static constexpr int kPeThreads = 128 * kPairs;
  static constexpr int kThreads = kFftThreads + kPeThreads;
  static constexpr int kCtasPerSm = W32 ? 1 : 768 / kThreads;
  static constexpr int kFftRegs = W32 ? 224 : 112, kPeRegs = 64, kLaunchRegs = W32 ? 96 : 80;
  static_assert(W32 || kThreads * kCtasPerSm == 768, "24 warps per SM");
  static_assert(!W32 || FB == 8, "the one-warp plan is built for FB = 8");
  static_assert(kFftThreads * (kFftRegs - kLaunchRegs) <= kPeThreads * (kLaunchRegs - kPeRegs),
                "setmaxnreg.inc would wait forever for registers nobody releases");
  static_assert(kThreads * kCtasPerSm * kLaunchRegs <= 65536 && kThreads * kCtasPerSm * (kLaunchRegs + 8) > 65536,
                "kLaunchRegs must be what __launch_bounds__ gives this many threads per SM");
};
constexpr uint32_t kWsWaitHintNs = 1000, kWsWaitSleepNs = 0;

template <int FB, bool MEL, typename S, bool W32>
__global__ void __launch_bounds__(WsGeometry<FB, W32>::kThreads, WsGeometry<FB, W32>::kCtasPerSm)
melif_ws_kernel(const S* __restrict__ audio, int64_t n_samples, isi_melif_params p,
                float* __restrict__ out, int seg_frames, int n_segs, int n_items_total, uint32_t wait_cfg,
                int ablate) {
  constexpr int NFFT = 2048;
  using G = WsGeometry<FB, W32>;
  using P = typename std::conditional<W32, PlanW32, Plan<NFFT>>::type;
  constexpr int M = P::M, NP = FB / 2;
  constexpr int kWsFftThreads = G::kFftThreads, kWsPeThreads = G::kPeThreads;
  constexpr int kWsThreads = G::kThreads;
  static_assert(NP == kWsFftThreads / P::kFftThreads, "one transform group per frame pair");
  constexpr int IPT = (M / 2) / kWsPeThreads;         // untangle items per polar/emit thread
  constexpr int RPT = M / kWsPeThreads;               // output rows per polar/emit thread
  static_assert(IPT * kWsPeThreads == M / 2, "whole items per thread");
  extern __shared__ __align__(128) unsigned char smem[];
  const MelifSmem L = melif_smem_layout<NFFT, FB>(p.hop, (int)sizeof(S), 2, P::kPitchA);
  cpx* twm = reinterpret_cast<cpx*>(smem + L.tw);
  float* win = reinterpret_cast<float*>(smem + L.win);
  S* stage = reinterpret_cast<S*>(smem + L.stage);
  cpx2* zA = reinterpret_cast<cpx2*>(smem + L.za);         // [2][NP][kPitchA]
  uint64_t* bar_stage = reinterpret_cast<uint64_t*>(smem + L.bar);   // [2] audio of a batch has landed
  uint64_t* bar_full = bar_stage + 2;                      // [2] transform -> polar/emit
  uint64_t* bar_empty = bar_stage + 4;                     // [2] polar/emit -> transform
  uint64_t* bar_pass1 = bar_stage + 6;                     // [2] every transform thread is done with a stage
  const int stage_elems = ((((FB - 1) * p.hop + NFFT) * (int)sizeof(S) + 15) / 16 * 16) / (int)sizeof(S);
  constexpr int kBufElems = NP * P::kPitchA;

  const int tid = threadIdx.x;
  const int dc = p.drop_dc ? 1 : 0;
  const float eps = p.safelog_eps;
  // PERSISTENT: a CTA walks the work items blockIdx.x, blockIdx.x + gridDim.x, ... (an item = a
  // run of frames of one note).  The tables, the barriers and the two-deep pipeline live across
  // items: the first batch of the next item is staged and transformed while the polar/emit
  // warps finish the last batch of this one, so only the first fill and the last drain of a CTA
  // are exposed (with one CTA per item they were, for every note).
  const int n_items = n_items_total;
  const int item_step = (int)gridDim.x;
  struct Item { int note_idx, fs, fe, n_batches, b_begin; };
  auto make_item = [&](int item) {
    Item w;
    w.note_idx = item / n_segs;
    const int seg = item - w.note_idx * n_segs;
    w.fs = seg * seg_frames;
    w.fe = min(p.n_frames, w.fs + seg_frames);
    w.n_batches = (w.fe - w.fs + FB - 1) / FB;
    w.b_begin = w.fs > 0 ? -1 : 0;
    return w;
  };

  const cpx* tw_global = reinterpret_cast<const cpx*>(p.twiddle);
  for (int i = tid; i < M; i += kWsThreads)
    twm[i] = tw_global[W32 ? fft32_table_source(i) : fft_table_source<Plan<NFFT>>(i)];
  const bool fold_scale = sizeof(S) == 2 && is_pow2_scale(p.pcm_scale);
  const float win_scale = 0.5f * (fold_scale ? p.pcm_scale : 1.f);
  const float sample_scale = (sizeof(S) == 2 && !fold_scale) ? p.pcm_scale : 1.f;
  for (int i = tid; i < NFFT; i += kWsThreads) win[i] = p.window[i] * win_scale;
  for (int i = tid; i < 2 * kBufElems; i += kWsThreads) { zA[i].re = bc(0.f); zA[i].im = bc(0.f); }
  if (tid == 0) {
    mbar_init(bar_stage + 0, 1);
    mbar_init(bar_stage + 1, 1);
    mbar_init(bar_pass1 + 0, kWsFftThreads);
    mbar_init(bar_pass1 + 1, kWsFftThreads);
    mbar_init(bar_full + 0, kWsFftThreads);
    mbar_init(bar_full + 1, kWsFftThreads);
    mbar_init(bar_empty + 0, kWsPeThreads);
    mbar_init(bar_empty + 1, kWsPeThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // The transform role takes the HIGHEST warp ids: the SMSP arbiter prefers the highest eligible
  // warp, and the transform warps are the critical path (13 % of their stall samples were
  // "not selected" while they sat below the polar/emit warps).
  if (tid >= kWsPeThreads) {
    // =========================== transform warps ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::kFftRegs));
    const int ft = tid - kWsPeThreads;                     // 0 .. kWsFftThreads - 1
    const int q = ft / P::kFftThreads, j = ft % P::kFftThreads;   // frame pair of this group, lane in it
    const uint32_t group_bar = 3 + q;
    // The audio stage is double-buffered and fed by the first transform warp alone (its lanes
    // zero-fill what lies outside the note, lane 0 sends the bulk copy; the mbarrier's release /
    // acquire carries both to the readers), so the four groups never meet on a role-wide
    // barrier: they drift apart and their load bursts and butterfly phases interleave.
    const bool feeder = ft < 32;
    // first batch of an item: the look-back frame of a segment that starts mid-note, else frames fs..
    auto first_batch = [&](const Item& w, int* frame, int* nfr) {
      *frame = w.fs > 0 ? w.fs - 1 : w.fs;
      *nfr = w.fs > 0 ? 1 : min(FB, w.fe - w.fs);
    };
    if (feeder) {
      const Item w0 = make_item((int)blockIdx.x);
      int frame, nfr;
      first_batch(w0, &frame, &nfr);
      const StageSpan<S> sp(n_samples, p.hop, p.pad_left, NFFT, frame, nfr);
      sp.fill(stage, ft, 32);
      __syncwarp();
      if (ft == 0) sp.issue(stage, audio + (int64_t)w0.note_idx * n_samples, bar_stage);
    }
    uint32_t it = 0;
    for (int item = (int)blockIdx.x; item < n_items; item += item_step) {
    const Item w = make_item(item);
    const int fs = w.fs, fe = w.fe, n_batches = w.n_batches;
    const S* note = audio + (int64_t)w.note_idx * n_samples;
    for (int b = w.b_begin; b < n_batches; ++b, ++it) {
      const bool lookback = b < 0;
      const int f0 = lookback ? fs - 1 : fs + b * FB;
      const int nf = lookback ? 1 : min(FB, fe - f0);
      // the batch after this one: the item's next, or the first of the CTA's next item
      int next_f0 = fs + (b + 1) * FB, next_nf = 0;
      const S* next_note = note;
      if (b + 1 < n_batches) {
        next_nf = min(FB, fe - next_f0);
      } else if (item + item_step < n_items) {
        const Item wn = make_item(item + item_step);
        first_batch(wn, &next_f0, &next_nf);
        next_note = audio + (int64_t)wn.note_idx * n_samples;
      }
      const uint32_t buf = it & 1, use = it >> 1;
      cpx2* z = zA + buf * kBufElems + q * P::kPitchA;
      const S* st_cur = stage + buf * stage_elems;
      const bool active = (lookback ? (q == NP - 1) : (q < nf)) && !(ablate & 2);   // & 2: polar/emit role alone

      if (feeder && next_nf > 0) {
        // the other stage buffer was read by pass 1 of the previous batch: all transform threads
        // have arrived on its barrier by now (they are at most a pass or two behind)
        if (it > 0) mbar_wait_pair(bar_pass1, buf ^ 1, ((it - 1) >> 1) & 1, wait_cfg);
        S* st_next = stage + (buf ^ 1) * stage_elems;
        const StageSpan<S> sp(n_samples, p.hop, p.pad_left, NFFT, next_f0, next_nf);
        sp.fill(st_next, ft, 32);
        __syncwarp();
        if (ft == 0) sp.issue(st_next, next_note, bar_stage + (buf ^ 1));
      }
      mbar_wait_pair(bar_stage, buf, use & 1, wait_cfg);     // this batch's audio has landed
      mbar_wait_pair(bar_empty, buf, (use & 1) ^ 1, wait_cfg);   // polar/emit released the workspace
      const S* fr_a = st_cur + (lookback ? 0 : q * p.hop);
      const S* fr_b = st_cur + (lookback ? 0 : (q + NP) * p.hop);
      if constexpr (W32) {
        // one warp per pair: two radix-32 passes, the exchange stays inside the warp
        if (active) fft32_passA(j, fr_a, fr_b, true, sample_scale, win, twm, z);
        mbar_arrive(bar_pass1 + buf);                      // done with this stage buffer
        if (active) {                                      // warp-uniform
          __syncwarp();
          PassB32Regs regs;
          fft32_passB_load(j, z, regs);
          __syncwarp();
          fft32_passB_store(j, regs, z);
        }
      } else {
        if (active) fft_pass1_pair<Plan<NFFT>>(j, fr_a, fr_b, true, sample_scale, win, twm, z);
        mbar_arrive(bar_pass1 + buf);                      // done with this stage buffer
        if (active) asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
        if (active) {
          fft_pass2<Plan<NFFT>>(j, twm, z);
          asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
          Pass3Regs<Plan<NFFT>, cpx2> regs;
          fft_pass3_load<Plan<NFFT>>(j, z, regs);
          asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
          fft_pass3_store<Plan<NFFT>>(j, regs, z);
        }
      }
      mbar_arrive(bar_full + buf);                         // release: the spectrum is in place
    }
    }
  } else {
    // =========================== polar / emit warps ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::kPeRegs));
    const int t = tid;                                     // 0 .. 511: untangle item and first row
    cpx w_item[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) w_item[i] = tw_global[t + i * kWsPeThreads];
    const bool w_vec = MEL && p.mel_width == kMaxMelWidth && (reinterpret_cast<uintptr_t>(p.mel_weight) & 15) == 0;
    BinState st[IPT];
    uint32_t it = 0;
    for (int item = (int)blockIdx.x; item < n_items; item += item_step) {
    const Item w = make_item(item);
    const int fs = w.fs, fe = w.fe, note_idx = w.note_idx;
#pragma unroll
    for (int i = 0; i < IPT; ++i) st[i] = bin_state_init();     // no phase history across items
    for (int b = w.b_begin; b < w.n_batches; ++b, ++it) {
      const bool lookback = b < 0;
      const int f0 = lookback ? fs - 1 : fs + b * FB;
      const int nf = lookback ? 1 : min(FB, fe - f0);
      const uint32_t buf = it & 1, use = it >> 1;
      cpx2* z = zA + buf * kBufElems;
      mbar_wait_pair(bar_full, buf, use & 1, wait_cfg);
      if (ablate & 1) { mbar_arrive(bar_empty + buf); continue; }     // profiling: transform role alone
#pragma unroll
      for (int i = 0; i < IPT; ++i)
        if (i == 0)
          polar_item<P, MEL, NP, true>(t, z, P::kPitchA, w_item[0], dc ? M : 0, lookback, eps, st[0]);
        else
          polar_item<P, MEL, NP, false>(t + i * kWsPeThreads, z, P::kPitchA, w_item[i], 0, lookback, eps, st[i]);
      if (!lookback) {
        // band constants after polar (held across it they spill): their L1/L2 latency hides
        // behind the role's barrier
        RowBand<MEL> band[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) band[r] = load_row_band<MEL>(p, t + r * kWsPeThreads, dc, w_vec);
        asm volatile("bar.sync 2, %0;" ::"n"(kWsPeThreads) : "memory");  // every bin of the batch is polar
#pragma unroll
        for (int r = 0; r < RPT; ++r)
          emit_row<P, FB, MEL>(p, z, band[r], note_idx, f0, nf, eps, out);
      }
      mbar_arrive(bar_empty + buf);                        // release: the buffer may be overwritten
    }
    }
  }
}

// Frames per CTA: whole notes when the batch fills the GPU on its own, else the split that
// minimises the estimated makespan, waves x (batches per CTA + the look-back transform of a CTA
// that does not start at frame 0 + half a batch of set-up).  A single note (the interactive
// server's request) becomes 16 CTAs of one batch each instead of 4 CTAs of four.
static void choose_segments(int64_t n_notes, int n_frames, int fb, int ctas_per_sm, int* seg_frames, int* n_segs) {
  const int frames_padded = (n_frames + fb - 1) / fb * fb;
  const int max_segs = frames_padded / fb;
  const double slots = (double)ctas_per_sm * kNumSms;   // resident CTAs: __launch_bounds__(NT, ctas_per_sm)
  double best = 1e30;
  *seg_frames = frames_padded; *n_segs = 1;
  for (int s = 1; s <= max_segs; ++s) {
    const int sf = ((frames_padded + s - 1) / s + fb - 1) / fb * fb;
    const int ns = (n_frames + sf - 1) / sf;
    const double waves = (double)(int64_t)((double)n_notes * ns / slots + 0.999999);
    const double cost = waves * (sf / fb + (ns > 1 ? 1.0 : 0.0) + 0.5);
    if (cost < best - 1e-9) { best = cost; *seg_frames = sf; *n_segs = ns; }
  }
}

template <int NFFT, int FB, int NT, bool MEL, typename S>
static int launch_melif_t(const S* audio, int64_t n_notes, int64_t n_samples,
                          const isi_melif_params& p, float* out, cudaStream_t stream, int bulk_ok) {
  const MelifSmem L = melif_smem_layout<NFFT, FB>(p.hop, (int)sizeof(S));
  if (L.total > 227 * 1024) return ISI_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(melif_kernel<NFFT, FB, NT, MEL, S>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return (int)e;
  int seg_frames, n_segs;
  choose_segments(n_notes, p.n_frames, FB, melif_ctas_per_sm(NT), &seg_frames, &n_segs);
  if (n_notes * n_segs > 0x7fffffff) return ISI_ERR_SHAPE;
  melif_kernel<NFFT, FB, NT, MEL, S><<<(unsigned)(n_notes * n_segs), NT, L.total, stream>>>(
      audio, n_samples, p, out, bulk_ok, seg_frames, n_segs);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

static constexpr int ws_pitch(bool w32) { return w32 ? PlanW32::kPitchA : Plan<2048>::kPitchA; }
// ISI_MELIF_WAIT_HINT / ISI_MELIF_WAIT_SLEEP (ns; testing / profiling): see mbar_wait_backoff
static uint32_t ws_wait_cfg() {
  static const uint32_t cfg = [] {
    const char* h = getenv("ISI_MELIF_WAIT_HINT");
    const char* z = getenv("ISI_MELIF_WAIT_SLEEP");
    const uint32_t hint = h ? (uint32_t)atoi(h) : kWsWaitHintNs, sleep = z ? (uint32_t)atoi(z) : kWsWaitSleepNs;
    return (hint & 0xffffu) | (sleep & 0xffffu) << 16;
  }();
  return cfg;
}
// ISI_MELIF_ABLATE (profiling only, the output is WRONG): 1 = the polar/emit warps only hand the
// buffers back (times the transform role alone), 2 = the transform warps only hand them over
// (times the polar/emit role alone), 3 = both (the skeleton: staging, barriers, prologue).
static int ws_ablate() {
  static const int v = [] {
    const int a = getenv("ISI_MELIF_ABLATE") ? atoi(getenv("ISI_MELIF_ABLATE")) : 0;
    if (a) fprintf(stderr, "libisi_b200: ISI_MELIF_ABLATE=%d -- profiling mode, isi_melif_forward's output is WRONG\n", a);
    return a;
  }();
  return v;
}

template <int FB, bool MEL, typename S, bool W32>
static int launch_melif_ws(const S* audio, int64_t n_notes, int64_t n_samples,
                           const isi_melif_params& p, float* out, cudaStream_t stream) {
  using G = WsGeometry<FB, W32>;
  const MelifSmem L = melif_smem_layout<2048, FB>(p.hop, (int)sizeof(S), 2, ws_pitch(W32));
  cudaError_t e = cudaFuncSetAttribute(melif_ws_kernel<FB, MEL, S, W32>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return (int)e;
  int seg_frames, n_segs;
  choose_segments(n_notes, p.n_frames, FB, G::kCtasPerSm, &seg_frames, &n_segs);
  if (n_notes * n_segs > 0x7fffffff) return ISI_ERR_SHAPE;
  const int64_t n_items = n_notes * n_segs, slots = (int64_t)G::kCtasPerSm * kNumSms;
  melif_ws_kernel<FB, MEL, S, W32><<<(unsigned)(n_items < slots ? n_items : slots), G::kThreads, L.total, stream>>>(
      audio, n_samples, p, out, seg_frames, n_segs, (int)n_items, ws_wait_cfg(), ws_ablate());
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

template <typename S>
static int launch_melif_s(const S* audio, int64_t n_notes, int64_t n_samples,
                          const isi_melif_params& p, float* out, cudaStream_t stream) {
  // the bulk copy needs 16-byte aligned global addresses and sizes
  constexpr int kPer16 = 16 / (int)sizeof(S);
  const int bulk_ok = (n_samples % kPer16 == 0) && (p.hop % kPer16 == 0) &&
                      (p.pad_left % kPer16 == 0) && ((uintptr_t)audio % 16 == 0);
  // ISI_MELIF_GENERIC=1 (testing / profiling): keep the NSynth shape on the generic kernel too.
  // The choice must not depend on the sample format -- PCM and FP32 input of the same notes give
  // bit-identical spectrograms only through the same kernel (ptxas contracts differently in
  // different instantiations) -- so the geometry has to be bulk-copyable in BOTH formats.
  static const bool force_generic = getenv("ISI_MELIF_GENERIC") != nullptr;
  const bool geometry_ok = (n_samples % 8 == 0) && (p.hop % 8 == 0) && (p.pad_left % 8 == 0);
  if (!force_generic && p.n_fft == 2048 && bulk_ok && geometry_ok && p.hop <= 2048 &&
      melif_smem_layout<2048, 8>(p.hop, (int)sizeof(S), 2, ws_pitch(true)).total <= 227 * 1024) {
    // ISI_MELIF_WS_FB=4 (testing / profiling): two 384-thread CTAs per SM instead of one of 768
    static const bool fb4 = getenv("ISI_MELIF_WS_FB") != nullptr && atoi(getenv("ISI_MELIF_WS_FB")) == 4;
    if (fb4)
      return p.use_mel ? launch_melif_ws<4, true, S, false>(audio, n_notes, n_samples, p, out, stream)
                       : launch_melif_ws<4, false, S, false>(audio, n_notes, n_samples, p, out, stream);
    // ISI_MELIF_WS_PLAN=3 (testing / profiling): the 16 x 16 x 4 transform plan (8 transform warps)
    static const bool plan3 = getenv("ISI_MELIF_WS_PLAN") != nullptr && atoi(getenv("ISI_MELIF_WS_PLAN")) == 3;
    if (plan3)
      return p.use_mel ? launch_melif_ws<8, true, S, false>(audio, n_notes, n_samples, p, out, stream)
                       : launch_melif_ws<8, false, S, false>(audio, n_notes, n_samples, p, out, stream);
    return p.use_mel ? launch_melif_ws<8, true, S, true>(audio, n_notes, n_samples, p, out, stream)
                     : launch_melif_ws<8, false, S, true>(audio, n_notes, n_samples, p, out, stream);
  }
#define ISI_MELIF_CASE(N, FB, NT)                                                               \
  case N:                                                                                     \
    return p.use_mel ? launch_melif_t<N, FB, NT, true, S>(audio, n_notes, n_samples, p, out, stream, bulk_ok) \
                     : launch_melif_t<N, FB, NT, false, S>(audio, n_notes, n_samples, p, out, stream, bulk_ok);
  switch (p.n_fft) {
    ISI_MELIF_CASE(2048, 8, 256)
    ISI_MELIF_CASE(1024, 4, 128)
    ISI_MELIF_CASE(512, 4, 64)
    default: return ISI_ERR_UNSUPPORTED;
  }
#undef ISI_MELIF_CASE
}

int launch_melif(const void* audio, int64_t n_notes, int64_t n_samples,
                 const isi_melif_params& p, float* out, cudaStream_t stream) {
  if (p.use_mel && p.mel_width > kMaxMelWidth) return ISI_ERR_UNSUPPORTED;
  switch (p.audio_format) {
    case ISI_AUDIO_F32:
      return launch_melif_s(static_cast<const float*>(audio), n_notes, n_samples, p, out, stream);
    case ISI_AUDIO_PCM16:
      return launch_melif_s(static_cast<const int16_t*>(audio), n_notes, n_samples, p, out, stream);
    default: return ISI_ERR_UNSUPPORTED;
  }
}

}  // namespace isi
