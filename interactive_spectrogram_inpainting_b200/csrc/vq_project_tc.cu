// Pre-quantiser projection on the tensor cores: the 1x1 `quantize_conv_{t,b}` of the reference
// (vqvae.py:149-150,175-177, called at :260 and :272) together with the channel concatenation in
// front of the bottom one (vqvae.py:271), as one kernel -- SURVEY.md section 8(f) N3.
//
//   out[n, :] = bias + sum_s W_s f_s[n, :]        f_s = rows of up to two channels-last sources
//
// i.e. torch.cat([dec_t, enc_b], 1) -> Conv2d(C, 64, 1) without materialising the concatenation,
// without a separate bias pass, and with the result written as the contiguous [N, 64] rows the
// nearest-code search reads.  tcgen05.mma kind::tf32 with the same 3xTF32 split as the search
// (FP32-equivalent accuracy: the dropped lo*lo term is ~2^-22 relative), accumulators in TMEM.
//
// Same skeleton as vq_assign_tc.cu, with the roles of the streamed operand exchanged: the 64
// output channels are ONE operand tile, and the contraction runs over 64-channel chunks of the
// inputs, accumulating in the same TMEM columns.  Persistent CTAs, one per SM, 128-row tiles; a
// "step" is one (tile, chunk) pair and both operands go through 2-stage rings:
//   warps 0-3  epilogue   tcgen05.ld the 128 x 64 accumulator, add the bias, store rows
//   warps 4-11 loader     coalesced reads of steps k+1 and k+2 sit in registers while step k is
//                         split into TF32 hi/lo and stored in the UMMA layout of stage k & 1
//   warp  12   W producer streams the pre-split, pre-swizzled 64x64 weight chunks (cp.async.bulk)
//   warp  13   MMA issuer one elected lane; owns TMEM (2 accumulator stages x 64 columns)
// so the global loads, the split and the MMAs of consecutive steps overlap and the kernel runs
// at the rate of its HBM stream (768 B read + 256 B written per bottom row).
#include "common.cuh"
#include "umma.cuh"

namespace isi {

namespace proj {

constexpr int kRowsPerMma = 128;
constexpr int kTileRows = kRowsPerMma;                 // 128
constexpr int kOut = 64;                               // output channels (UMMA N)
constexpr int kChunk = 64;                             // input channels per contraction chunk
constexpr int kSlabs = kChunk / 32;
constexpr int kABytesPart = kSlabs * kRowsPerMma * 128;          // one hi|lo operand: 32 KB
constexpr int kAStageBytes = 2 * kABytesPart;                     // 64 KB
constexpr int kAStages = 2;
constexpr int kWBytesPart = kSlabs * kOut * 128;                 // one (hi|lo) weight chunk: 16 KB
constexpr int kWStageBytes = 2 * kWBytesPart;                     // 32 KB
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kOut;                      // 128
constexpr int kFirstLoaderWarp = 4;
constexpr int kLoaderWarps = 8;
constexpr int kLoaderThreads = kLoaderWarps * 32;                                   // 256
constexpr int kChunksPerThread = kTileRows * (kChunk / 4) / kLoaderThreads;         // 8
constexpr int kProducerWarp = kFirstLoaderWarp + kLoaderWarps;                      // 12
constexpr int kMmaWarp = kProducerWarp + 1;                                         // 13
constexpr int kThreads = (kMmaWarp + 1) * 32;                                       // 448
constexpr int kMaxChunks = 16;                                                      // C <= 1024

constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kOut >> 3) << 17) |
                            ((uint32_t)(kRowsPerMma >> 4) << 24);

struct Smem {
  static constexpr int a = 0;                                   // 1024-aligned operand tiles
  static constexpr int w = a + kAStages * kAStageBytes;
  static constexpr int bias = w + 2 * kWStageBytes;
  static constexpr int bars = bias + kOut * 4;
  static constexpr int total = bars + 128;
};

}  // namespace proj

using namespace umma;

// weight [64, C] (the Conv2d weight [64, C, 1, 1]) -> per 64-channel chunk the TF32 hi and lo
// parts in exactly the bytes the kernel's shared-memory stage expects
__global__ void __launch_bounds__(256)
vq_project_prepare_kernel(const float* __restrict__ weight, int c_total, char* __restrict__ prepared) {
  using namespace proj;
  const int total = c_total * kOut;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e / c_total, c = e % c_total;
    const float v = weight[e];
    const float hi = to_tf32(v);
    const float lo = to_tf32(v - hi);
    const size_t off = (size_t)(c / kChunk) * kWStageBytes + operand_offset(kOut, n, c % kChunk);
    *reinterpret_cast<float*>(prepared + off) = hi;
    *reinterpret_cast<float*>(prepared + off + kWBytesPart) = lo;
  }
}

__global__ void __launch_bounds__(proj::kThreads, 1)
vq_project_tc_kernel(const float* __restrict__ src0, int chunks0, int64_t stride0,
                     const float* __restrict__ src1, int chunks1, int64_t stride1, int64_t n_rows,
                     const char* __restrict__ w_tiles, const float* __restrict__ bias_global,
                     float* __restrict__ out) {
  using namespace proj;
  const bool wide_out = (reinterpret_cast<uintptr_t>(out) & 31) == 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = s32(smem);
  float* bias = reinterpret_cast<float*>(smem + Smem::bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  const uint32_t bar_w_full = s32(bars + 0);     // [2]
  const uint32_t bar_w_empty = s32(bars + 2);    // [2]
  const uint32_t bar_acc_full = s32(bars + 4);   // [2]
  const uint32_t bar_acc_empty = s32(bars + 6);  // [2]
  const uint32_t bar_a_full = s32(bars + 8);     // [2]
  const uint32_t bar_a_empty = s32(bars + 10);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_chunks = (uint32_t)(chunks0 + chunks1);
  const int64_t n_row_tiles = (n_rows + kTileRows - 1) / kTileRows;
  // tiles blockIdx.x, blockIdx.x + gridDim.x, ...; step k = (k / n_chunks)-th of them, chunk k % n_chunks
  const uint32_t my_tiles = (int64_t)blockIdx.x < n_row_tiles
      ? (uint32_t)((n_row_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0u;
  const uint32_t total_steps = my_tiles * n_chunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_w_full + 8 * i, 1);
      mbar_init(bar_w_empty + 8 * i, 1);
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 128);
      mbar_init(bar_a_full + 8 * i, kLoaderThreads);
      mbar_init(bar_a_empty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x < kOut) bias[threadIdx.x] = bias_global ? bias_global[threadIdx.x] : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kFirstLoaderWarp && warp < kFirstLoaderWarp + kLoaderWarps) {
    // ===================== input loader / splitter =====================
    // Thread t owns 16-byte piece c4 = t % 16 of rows t / 16 + 16 i: the 16 lanes of a row read
    // its 256 contiguous bytes of the chunk.  Two register buffers: steps k+1 and k+2 are in
    // flight while step k is split.
    const int t = threadIdx.x - kFirstLoaderWarp * 32;
    const int c4 = t % (kChunk / 4), r0 = t / (kChunk / 4);
    constexpr int kRowStep = kLoaderThreads / (kChunk / 4);    // 16
    auto fetch = [&](uint32_t step, float4* buf) {
      if (step >= total_steps) return;
      const uint32_t i_tile = step / n_chunks, chunk = step - i_tile * n_chunks;
      const bool first = chunk < (uint32_t)chunks0;
      const float* base = first ? src0 + chunk * kChunk : src1 + (chunk - chunks0) * kChunk;
      const int64_t stride = first ? stride0 : stride1;
      const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)i_tile * gridDim.x) * kTileRows;
#pragma unroll
      for (int i = 0; i < kChunksPerThread; ++i) {
        const int64_t row = row0 + r0 + i * kRowStep;
        buf[i] = row < n_rows ? __ldg(reinterpret_cast<const float4*>(base + row * stride) + c4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto split_store = [&](uint32_t step, uint32_t s, const float4* buf) {
      mbar_wait(bar_a_empty + 8 * s, ((step >> 1) & 1) ^ 1);     // the MMAs of step - 2 are done
      const uint32_t stage = Smem::a + s * kAStageBytes;
#pragma unroll
      for (int i = 0; i < kChunksPerThread; ++i) {
        const int r = r0 + i * kRowStep;
        const float4 v = buf[i];
        const float4 hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        const float4 lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z),
                                      to_tf32(v.w - hi.w));
        const uint32_t off = stage + operand_offset(kRowsPerMma, r, 4 * c4);
        *reinterpret_cast<float4*>(smem + off) = hi;
        *reinterpret_cast<float4*>(smem + off + kABytesPart) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy
      mbar_arrive(bar_a_full + 8 * s);
    };
    float4 buf0[kChunksPerThread], buf1[kChunksPerThread];
    fetch(0, buf0);
    fetch(1, buf1);
    for (uint32_t step = 0; step < total_steps; step += 2) {
      split_store(step, 0, buf0);
      fetch(step + 2, buf0);
      if (step + 1 < total_steps) {
        split_store(step + 1, 1, buf1);
        fetch(step + 3, buf1);
      }
    }
  } else if (warp == kProducerWarp) {
    // ===================== weight producer =====================
    if (lane == 0) {
      for (uint32_t step = 0; step < total_steps; ++step) {
        const uint32_t s = step & 1, ph = (step >> 1) & 1, chunk = step % n_chunks;
        mbar_wait(bar_w_empty + 8 * s, ph ^ 1);
        mbar_expect_tx(bar_w_full + 8 * s, kWStageBytes);
        bulk_g2s(smem_base + Smem::w + s * kWStageBytes, w_tiles + (size_t)chunk * kWStageBytes,
                 kWStageBytes, bar_w_full + 8 * s);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      for (uint32_t step = 0; step < total_steps; ++step) {
        const uint32_t s = step & 1, ph = (step >> 1) & 1;
        const uint32_t it = step / n_chunks, chunk = step - it * n_chunks, sa = it & 1;
        if (chunk == 0) mbar_wait(bar_acc_empty + 8 * sa, ((it >> 1) & 1) ^ 1);   // epilogue drained it
        mbar_wait(bar_a_full + 8 * s, ph);
        mbar_wait(bar_w_full + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t w_base = smem_base + Smem::w + s * kWStageBytes;
        const uint32_t a_base = smem_base + Smem::a + s * kAStageBytes;
        const uint32_t d_tmem = tmem_base + sa * kOut;
        uint32_t acc = chunk > 0 ? 1u : 0u;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          // (f_lo, w_hi), (f_hi, w_lo), (f_hi, w_hi): small terms first
          const uint32_t a_part = a_base + (term == 0 ? kABytesPart : 0);
          const uint32_t w_part = w_base + (term == 1 ? kWBytesPart : 0);
#pragma unroll
          for (int slab = 0; slab < kSlabs; ++slab) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = umma_desc(a_part + slab * (kRowsPerMma * 128) + kk * 32);
              const uint64_t wd = umma_desc(w_part + slab * (kOut * 128) + kk * 32);
              umma_tf32(d_tmem, ad, wd, kIdesc, acc);
              acc = 1;
            }
          }
        }
        umma_commit(bar_w_empty + 8 * s);          // the weight stage may be refilled
        umma_commit(bar_a_empty + 8 * s);          // the input stage may be overwritten
        if (chunk + 1 == n_chunks) umma_commit(bar_acc_full + 8 * sa);   // every chunk accumulated
      }
    }
  } else {
    // ===================== epilogue: bias, rows out =====================
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t sa = it & 1;
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      mbar_wait(bar_acc_full + 8 * sa, (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[kOut];
      tmem_ld64(tmem_base + lane_field + sa * kOut, v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_acc_empty + 8 * sa);                 // the values are in registers
      const int64_t row = tile * kTileRows + warp * 32 + lane;
      if (row < n_rows) {
        // a lane owns a whole 256-byte output row, so every store instruction touches 32
        // different lines and the LSU pays per line: 32-byte stores (STG.256) halve the count
        float* dst = out + row * kOut;
        if (wide_out) {
#pragma unroll
          for (int c = 0; c < kOut; c += 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + c);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + c + 4);
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         :: "l"(dst + c), "f"(v[c] + b0.x), "f"(v[c + 1] + b0.y), "f"(v[c + 2] + b0.z),
                            "f"(v[c + 3] + b0.w), "f"(v[c + 4] + b1.x), "f"(v[c + 5] + b1.y),
                            "f"(v[c + 6] + b1.z), "f"(v[c + 7] + b1.w)
                         : "memory");
          }
        } else {
#pragma unroll
          for (int c = 0; c < kOut; c += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + c);
            reinterpret_cast<float4*>(dst)[c >> 2] =
                make_float4(v[c] + b.x, v[c + 1] + b.y, v[c + 2] + b.z, v[c + 3] + b.w);
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t project_prepared_bytes(int c_total) {
  return (size_t)(c_total / proj::kChunk) * proj::kWStageBytes;
}

int launch_project_prepare(const float* weight, int c_total, void* prepared, cudaStream_t stream) {
  const int total = c_total * proj::kOut;
  vq_project_prepare_kernel<<<(total + 255) / 256, 256, 0, stream>>>(weight, c_total,
                                                                     reinterpret_cast<char*>(prepared));
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_project(const float* src0, int c0, int64_t stride0, const float* src1, int c1, int64_t stride1,
                   int64_t n_rows, const void* prepared, const float* bias, float* out, cudaStream_t stream) {
  using namespace proj;
  cudaError_t e = cudaFuncSetAttribute(vq_project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Smem::total);
  if (e != cudaSuccess) return (int)e;
  const int64_t n_row_tiles = (n_rows + kTileRows - 1) / kTileRows;
  const int grid = (int)(n_row_tiles < kNumSms ? n_row_tiles : kNumSms);
  vq_project_tc_kernel<<<grid, kThreads, Smem::total, stream>>>(
      src0, c0 / kChunk, stride0, src1, c1 / kChunk, stride1, n_rows,
      reinterpret_cast<const char*>(prepared), bias, out);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

}  // namespace isi
