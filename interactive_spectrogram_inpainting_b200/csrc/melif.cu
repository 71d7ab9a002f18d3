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
// three layouts (planes, channels-last, 2x2 space-to-depth blocks).  The whole working set is
// one FB-frame buffer (65 KB with tables and stage at n_fft 2048), so three CTAs share an SM
// and fill each other's barrier and latency stalls.
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
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct MelifSmem {
  int tw, win, stage, za, bar, total;   // byte offsets into dynamic shared memory
};

template <int NFFT, int FB>
__host__ __device__ inline MelifSmem melif_smem_layout(int hop, int sample_bytes) {
  using P = Plan<NFFT>;
  MelifSmem s;
  int off = 0;
  s.tw = off;    off += (NFFT / 2) * 8;                     // FFT twiddles (fft_table_source)
  s.win = off;   off += NFFT * 4;
  s.stage = off; off += (((FB - 1) * hop + NFFT) * sample_bytes + 15) / 16 * 16;
  s.za = off;    off += FB * P::kPitchA * 8;                // FFT workspace, spectrum, polar values
  s.bar = off;   off += 16;
  s.total = off;
  return s;
}

template <int NFFT, int FB, int NT, bool MEL, typename S>
__global__ void __launch_bounds__(NT, NT >= 512 ? 2 : 3)
melif_kernel(const S* __restrict__ audio, int64_t n_samples, isi_melif_params p,
             float* __restrict__ out, int bulk_ok, int seg_frames, int n_segs) {
  using P = Plan<NFFT>;
  constexpr int M = P::M;
  constexpr int IPT = (M / 2) / NT;           // polar work items per thread
  constexpr int RPT = M / NT;                 // output rows per thread
  constexpr int kGroups = NT / 64;            // frames transformed concurrently
  static_assert(IPT >= 1 && (M / 2) % NT == 0 && NT % 64 == 0, "bad thread count");
  static_assert(P::kPitchA >= M + 1, "a frame region must hold bins 0..M");
  extern __shared__ __align__(128) unsigned char smem[];
  const MelifSmem L = melif_smem_layout<NFFT, FB>(p.hop, (int)sizeof(S));
  cpx* twm = reinterpret_cast<cpx*>(smem + L.tw);
  float* win = reinterpret_cast<float*>(smem + L.win);
  S* stage = reinterpret_cast<S*>(smem + L.stage);
  cpx* zA = reinterpret_cast<cpx*>(smem + L.za);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);

  const int tid = threadIdx.x;
  const int note_idx = blockIdx.x / n_segs, seg = blockIdx.x - note_idx * n_segs;
  const int fs = seg * seg_frames;                          // first frame of this CTA
  const int fe = min(p.n_frames, fs + seg_frames);          // one past its last frame
  const S* note = audio + (int64_t)note_idx * n_samples;
  float* out0 = out + (int64_t)note_idx * 2 * M * p.n_frames;
  float* out1 = out0 + (int64_t)M * p.n_frames;
  const int dc = p.drop_dc ? 1 : 0;
  const float eps = p.safelog_eps;
  const bool pairs_aligned = (p.hop % 2) == 0;   // a frame starts on a sample-pair boundary

  // ---- one-time setup: tables to shared memory, per-thread constants to registers ----
  const cpx* tw_global = reinterpret_cast<const cpx*>(p.twiddle);     // W_N^j, j < N
  for (int i = tid; i < M; i += NT) twm[i] = tw_global[fft_table_source<P>(i)];
  for (int i = tid; i < NFFT; i += NT) win[i] = p.window[i];
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cpx w_item[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) w_item[i] = tw_global[tid + i * NT];
  // Per-row band constants (first bin, length, weights) are NOT kept in registers across the
  // transform (they would spill): each batch re-reads them from the L1-resident tables ahead
  // of the barrier that precedes emit.
  const bool w_vec = MEL && p.mel_width == kMaxMelWidth && (reinterpret_cast<uintptr_t>(p.mel_weight) & 15) == 0;
  BinState sa[IPT], sb[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) { sa[i] = BinState{1.f, 0.f}; sb[i] = BinState{1.f, 0.f}; }
  __syncthreads();

  // Stage the audio span of frames [frame, frame + nfr): zero-fill what lies outside the
  // note, one bulk copy for the rest (thread 0), completion signalled on `bar`.
  auto stage_span = [&](int frame, int nfr) {
    const int span = (nfr - 1) * p.hop + NFFT;
    const int64_t s0 = (int64_t)frame * p.hop - p.pad_left;
    if (!bulk_ok) {
      stage_fill(tid, NT, stage, span, note, n_samples, s0);
      return;
    }
    const int64_t lo = s0 < 0 ? -s0 : 0;                      // first valid index
    int64_t hi = n_samples - s0;                              // one past the last valid index
    hi = hi < 0 ? 0 : (hi > span ? span : hi);
    const int64_t vlo = lo < hi ? lo : hi;
    for (int i = tid; i < vlo; i += NT) stage[i] = S(0);
    for (int i = (int)hi + tid; i < span; i += NT) stage[i] = S(0);
    if (tid == 0) {
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
    // The three FFT passes of a frame only involve the 64 threads of its group: they meet
    // on a named barrier of their own instead of stalling the whole CTA.
    const uint32_t group_bar = 1 + (tid >> 6);
    for (int fb = tid / 64; fb < nf; fb += kGroups)
      fft_pass1<P>(tid & 63, stage + fb * p.hop, pairs_aligned, p.pcm_scale, win, twm, zA + fb * P::kPitchA);
    asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
    for (int fb = tid / 64; fb < nf; fb += kGroups) fft_pass2<P>(tid & 63, twm, zA + fb * P::kPitchA);
    asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
    for (int fb = tid / 64; fb < nf; fb += kGroups) {
      Pass3Regs<P> regs;
      fft_pass3_load<P>(tid & 63, zA + fb * P::kPitchA, regs);
      asm volatile("bar.sync %0, 64;" ::"r"(group_bar) : "memory");
      fft_pass3_store<P>(tid & 63, regs, zA + fb * P::kPitchA);
    }
    __syncthreads();
    if (next_nf > 0 && bulk_ok) stage_span(next_f0, next_nf);   // every group is done with the stage
    // polar: frames in order, previous spectrum value in registers, in place
#pragma unroll 2
    for (int fb = 0; fb < nf; ++fb) {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
        if (i == 0)
          polar_item<P, MEL, true>(tid, zA + fb * P::kPitchA, w_item[0], dc ? M : 0,
                                   lookback || (f0 + fb == 0), eps, sa[0], sb[0]);
        else
          polar_item<P, MEL, false>(tid + i * NT, zA + fb * P::kPitchA, w_item[i], 0,
                                    lookback || (f0 + fb == 0), eps, sa[i], sb[i]);
    }
    if (!lookback) {
      float row_w[RPT][MEL ? kMaxMelWidth : 1];
      int row_bin[RPT], row_cnt[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int row = tid + r * NT;
        row_bin[r] = row + dc;
        row_cnt[r] = 0;
        if (MEL) {
          row_bin[r] = ld_table(p.mel_start + row) + dc;
          row_cnt[r] = ld_table(p.mel_count + row);
          if (w_vec) {
            const float4* wt = reinterpret_cast<const float4*>(p.mel_weight) + (int64_t)row * 2;
            const float4 a = ld_table4(wt), c = ld_table4(wt + 1);
            row_w[r][0] = a.x; row_w[r][1] = a.y; row_w[r][2] = a.z; row_w[r][3] = a.w;
            row_w[r][4] = c.x; row_w[r][5] = c.y; row_w[r][6] = c.z; row_w[r][7] = c.w;
          } else {
#pragma unroll
            for (int i = 0; i < kMaxMelWidth; ++i)
              row_w[r][i] = (i < p.mel_width) ? __ldg(p.mel_weight + (int64_t)row * p.mel_width + i) : 0.f;
          }
        }
      }
      __syncthreads();
      // emit: all FB time steps of a row at once
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int row = tid + r * NT;
        const int row_cnt_warp = MEL ? __reduce_max_sync(0xffffffffu, row_cnt[r]) : 0;
        float v0[FB], v1[FB];
        if (MEL)
          emit_mel<FB>(zA, P::kPitchA, row_bin[r], row_cnt[r], row_cnt_warp, row_w[r], f0 == 0, eps, v0, v1);
        else
          emit_linear<FB>(zA, P::kPitchA, row_bin[r], v0, v1);
        apply_epilogue<FB>(v0, v1, p.mask_phase != 0, p.mask_threshold, p.out_scale[0], p.out_bias[0],
                           p.out_scale[1], p.out_bias[1]);
        if (p.channels_last == ISI_SPEC_SPACE_TO_DEPTH) {
          // [B, F/2, T'/2, (f&1, t&1, channel)]: 2x2 spectrogram blocks as 8 channels.  A row
          // writes 16 bytes per block; the odd/even row pair (neighbouring lanes of the same
          // instruction) completes each 32-byte sector.
          float* d = out + ((((int64_t)note_idx * (M / 2) + (row >> 1)) * (p.n_frames >> 1) + (f0 >> 1)) * 8) +
                     (row & 1) * 4;
          if (nf == FB && (FB % 2 == 0)) {
#pragma unroll
            for (int q = 0; q < FB / 2; ++q)
              st_stream4(d + 8 * q, make_float4(v0[2 * q], v1[2 * q], v0[2 * q + 1], v1[2 * q + 1]));
          } else {
#pragma unroll
            for (int fb = 0; fb < FB; ++fb)
              if (fb < nf) {
                float* e = d + (fb >> 1) * 8 + (fb & 1) * 2;
                e[0] = v0[fb]; e[1] = v1[fb];
              }
          }
          continue;
        }
        if (p.channels_last) {
          // [B, F, T', 2]: the FB time steps of both channels are one contiguous run
          float* d = out + (((int64_t)note_idx * M + row) * p.n_frames + f0) * 2;
          if (nf == FB && (FB % 2 == 0) && (p.n_frames % 2 == 0)) {
#pragma unroll
            for (int q = 0; q < FB / 2; ++q)
              st_stream4(d + 4 * q, make_float4(v0[2 * q], v1[2 * q], v0[2 * q + 1], v1[2 * q + 1]));
          } else {
#pragma unroll
            for (int fb = 0; fb < FB; ++fb)
              if (fb < nf) { d[2 * fb] = v0[fb]; d[2 * fb + 1] = v1[fb]; }
          }
          continue;
        }
        float* d0 = out0 + (int64_t)row * p.n_frames + f0;
        float* d1 = out1 + (int64_t)row * p.n_frames + f0;
        if (nf == FB && (FB % 4 == 0) && (p.n_frames % 4 == 0)) {
#pragma unroll
          for (int q = 0; q < FB / 4; ++q) {
            st_stream4(d0 + 4 * q, make_float4(v0[4 * q], v0[4 * q + 1], v0[4 * q + 2], v0[4 * q + 3]));
            st_stream4(d1 + 4 * q, make_float4(v1[4 * q], v1[4 * q + 1], v1[4 * q + 2], v1[4 * q + 3]));
          }
        } else {
#pragma unroll
          for (int fb = 0; fb < FB; ++fb)
            if (fb < nf) { d0[fb] = v0[fb]; d1[fb] = v1[fb]; }
        }
      }
    }
    if (!bulk_ok && next_nf > 0) { __syncthreads(); stage_span(next_f0, next_nf); }
  }
}

// Frames per CTA: whole notes when the batch fills the GPU on its own, else the split that
// maximises (wave efficiency) x (useful / useful + look-back work).
static void choose_segments(int64_t n_notes, int n_frames, int fb, int ctas_per_sm, int* seg_frames, int* n_segs) {
  const int frames_padded = (n_frames + fb - 1) / fb * fb;
  const int max_segs = frames_padded / (4 * fb) > 1 ? frames_padded / (4 * fb) : 1;
  const double slots = (double)ctas_per_sm * kNumSms;   // resident CTAs: __launch_bounds__(NT, ctas_per_sm)
  double best = -1.0;
  *seg_frames = frames_padded; *n_segs = 1;
  for (int s = 1; s <= max_segs; ++s) {
    const int sf = ((frames_padded + s - 1) / s + fb - 1) / fb * fb;
    const int ns = (n_frames + sf - 1) / sf;
    const double waves = (double)n_notes * ns / slots;
    const double wave_eff = waves / (double)(int64_t)(waves + 0.999999);
    const double eff = wave_eff * (ns == 1 ? 1.0 : (double)sf / (sf + 2.0));
    if (eff > best + 1e-9) { best = eff; *seg_frames = sf; *n_segs = ns; }
  }
}

template <int NFFT, int FB, int NT, bool MEL, typename S>
static int launch_melif_t(const S* audio, int64_t n_notes, int64_t n_samples,
                          const isi_melif_params& p, float* out, cudaStream_t stream) {
  const MelifSmem L = melif_smem_layout<NFFT, FB>(p.hop, (int)sizeof(S));
  if (L.total > 227 * 1024) return ISI_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(melif_kernel<NFFT, FB, NT, MEL, S>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return (int)e;
  // the bulk copy needs 16-byte aligned global addresses and sizes
  constexpr int kPer16 = 16 / (int)sizeof(S);
  const int bulk_ok = (n_samples % kPer16 == 0) && (p.hop % kPer16 == 0) &&
                      (p.pad_left % kPer16 == 0) && ((uintptr_t)audio % 16 == 0);
  int seg_frames, n_segs;
  choose_segments(n_notes, p.n_frames, FB, NT >= 512 ? 2 : 3, &seg_frames, &n_segs);
  if (n_notes * n_segs > 0x7fffffff) return ISI_ERR_SHAPE;
  melif_kernel<NFFT, FB, NT, MEL, S><<<(unsigned)(n_notes * n_segs), NT, L.total, stream>>>(
      audio, n_samples, p, out, bulk_ok, seg_frames, n_segs);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

template <typename S>
static int launch_melif_s(const S* audio, int64_t n_notes, int64_t n_samples,
                          const isi_melif_params& p, float* out, cudaStream_t stream) {
#define ISI_MELIF_CASE(N, FB, NT)                                                               \
  case N:                                                                                     \
    return p.use_mel ? launch_melif_t<N, FB, NT, true, S>(audio, n_notes, n_samples, p, out, stream) \
                     : launch_melif_t<N, FB, NT, false, S>(audio, n_notes, n_samples, p, out, stream);
  switch (p.n_fft) {
    ISI_MELIF_CASE(2048, 8, 512)
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
