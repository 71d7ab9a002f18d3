// Nearest-code search on the 5th-generation tensor cores (tcgen05, kind::tf32) with a
// 3xTF32 split for FP32-equivalent accuracy.  Replaces bottleneck.py:55-61.
//
//   score(n, k) = |e_k|^2 - 2 x_n . e_k      (the |x_n|^2 term cannot change the argmin)
//   x = x_hi + x_lo,  B = -2E = b_hi + b_lo  (each part exactly representable in TF32)
//   acc = x_hi b_hi + x_hi b_lo + x_lo b_hi  (the dropped x_lo b_lo term is ~2^-22 relative)
//
// Persistent CTAs, one per SM, each owning tiles of 256 rows (two M=128 accumulators).
// Warp roles:
//   warps 0-3  epilogue   tcgen05.ld the 128 x 64 accumulators, add |e|^2, running argmin
//   warps 4-11 x loader   coalesced global reads of the NEXT row tile into registers (in the
//                         caller's layout) while the MMAs run; then hi/lo split and
//                         st.shared into the SWIZZLE_128B K-major UMMA layout
//   warp  12   B producer streams pre-split, pre-swizzled 64-code operand tiles of the
//                         codebook through a 2-stage ring with cp.async.bulk + mbarrier
//   warp  13   MMA issuer one elected lane issues the tcgen05.mma chain; owns TMEM
// The accumulators live in TMEM (2 stages x 2 row tiles x 64 columns), so the argmin of
// tile j overlaps the MMAs of tile j+1.
#include "common.cuh"
#include "umma.cuh"

namespace isi {

namespace tc {

constexpr int kRowsPerMma = 128;
constexpr int kMmaPerTile = 2;                       // row tiles per CTA tile
constexpr int kTileRows = kRowsPerMma * kMmaPerTile;  // 256
constexpr int kTileCodes = 64;                       // codes per B stage (UMMA N)
constexpr int kDim = 64;                             // feature dimension this kernel is built for
constexpr int kSlabs = kDim / 32;                    // 128-byte K slabs per row
constexpr int kABytesPart = kSlabs * kRowsPerMma * 128;        // one (m, hi|lo) operand: 32 KB
constexpr int kABytes = kMmaPerTile * 2 * kABytesPart;          // 128 KB
constexpr int kBBytesPart = kSlabs * kTileCodes * 128;         // one (hi|lo) code tile: 16 KB
constexpr int kBStageBytes = 2 * kBBytesPart;                   // 32 KB
constexpr int kBStages = 2;
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kMmaPerTile * kTileCodes;  // 256
constexpr int kFirstLoaderWarp = 4;                  // warps 0-3: epilogue (TMEM lane quarters)
constexpr int kLoaderWarps = 8;
constexpr int kLoaderThreads = kLoaderWarps * 32;                                  // 256
constexpr int kChunksPerThread = kTileRows * (kDim / 4) / kLoaderThreads;          // 16
constexpr int kProducerWarp = kFirstLoaderWarp + kLoaderWarps;                     // 12
constexpr int kMmaWarp = kProducerWarp + 1;                                        // 13
constexpr int kThreads = (kMmaWarp + 1) * 32;                                      // 448
constexpr int kMaxCodes = 4096;                      // |e|^2 table held in shared memory

// instruction descriptor: D=F32, A=B=TF32, both K-major, N=64, M=128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileCodes >> 3) << 17) |
                            ((uint32_t)(kRowsPerMma >> 4) << 24);

struct Smem {
  static constexpr int a = 0;                                   // 1024-aligned operand tiles
  static constexpr int b = a + kABytes;
  static constexpr int e2 = b + kBStages * kBStageBytes;
  static constexpr int bars = e2 + kMaxCodes * 4;
  static constexpr int total = bars + 128;
};

}  // namespace tc

using namespace tc;
using namespace umma;

// ---------------------------------------------------------------------------
// codebook -> operand tiles:  b_hi / b_lo hold, per 64-code tile, the TF32 hi and lo
// parts of -2E in exactly the bytes the kernel's shared-memory stage expects
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_prepare_tc_kernel(const float* __restrict__ embed, int n_embed, Prepared p) {
  const int n_tiles = (n_embed + kTileCodes - 1) / kTileCodes;
  const int total = n_tiles * kTileCodes * kDim;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int code = e % (n_tiles * kTileCodes), d = e / (n_tiles * kTileCodes);
    const float v = code < n_embed ? -2.f * embed[(int64_t)d * n_embed + code] : 0.f;
    const float hi = to_tf32(v);
    const float lo = to_tf32(v - hi);
    const int tile = code / kTileCodes, n = code % kTileCodes;
    const size_t off = (size_t)tile * kBStageBytes + operand_offset(kTileCodes, n, d);
    *reinterpret_cast<float*>(reinterpret_cast<char*>(p.b_hi) + off) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(p.b_hi) + off + kBBytesPart) = lo;
  }
}

// ---------------------------------------------------------------------------
// the search kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
vq_assign_tc_kernel(const float* __restrict__ x, isi_rows_layout lay, int64_t n_rows, int n_embed,
                    const char* __restrict__ b_tiles, const float* __restrict__ e2_global,
                    int64_t* __restrict__ out_index, float* __restrict__ out_score) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = s32(smem);
  float* e2s = reinterpret_cast<float*>(smem + Smem::e2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  // barrier slots
  const uint32_t bar_b_full = s32(bars + 0);     // [2]
  const uint32_t bar_b_empty = s32(bars + 2);    // [2]
  const uint32_t bar_acc_full = s32(bars + 4);   // [2]
  const uint32_t bar_acc_empty = s32(bars + 6);  // [2]
  const uint32_t bar_a_full = s32(bars + 8);
  const uint32_t bar_a_empty = s32(bars + 9);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_code_tiles = (n_embed + kTileCodes - 1) / kTileCodes;
  const int64_t n_row_tiles = (n_rows + kTileRows - 1) / kTileRows;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_b_full + 8 * i, 1);
      mbar_init(bar_b_empty + 8 * i, 1);
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 128);
    }
    mbar_init(bar_a_full, kLoaderThreads);
    mbar_init(bar_a_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int k = threadIdx.x; k < n_code_tiles * kTileCodes; k += kThreads) e2s[k] = e2_global[k];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kFirstLoaderWarp && warp < kFirstLoaderWarp + kLoaderWarps) {
    // ===================== x loader / splitter =====================
    // Each thread owns kChunksPerThread 16-byte chunks (4 consecutive features of one row).
    // The chunks of the NEXT tile are fetched into registers while the tensor cores still
    // work on the current one; when the MMAs release the operand buffer only the hi/lo
    // split and the swizzled st.shared remain on the critical path.
    const int t = threadIdx.x - kFirstLoaderWarp * 32;    // 0 .. kLoaderThreads-1
    const bool rows_contiguous = (lay.row_stride == 1 && lay.col_stride != 1);
    const bool vec_ok = (lay.col_stride == 1) && ((lay.row_stride & 3) == 0) &&
                        ((lay.batch_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    float4 buf[kChunksPerThread];
    auto fetch = [&](int64_t tile) {
      const int64_t row0 = tile * kTileRows;
#pragma unroll
      for (int i = 0; i < kChunksPerThread; ++i) {
        const int e = t + i * kLoaderThreads;
        int r, c;
        if (rows_contiguous) { r = e % kTileRows; c = e / kTileRows; }
        else                 { c = e % (kDim / 4); r = e / (kDim / 4); }
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t row = row0 + r;
        if (row < n_rows) {
          const float* src = x + row_offset(lay, row) + (int64_t)(4 * c) * lay.col_stride;
          if (vec_ok) {
            v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            v.x = __ldg(src); v.y = __ldg(src + lay.col_stride); v.z = __ldg(src + 2 * lay.col_stride);
            v.w = __ldg(src + 3 * lay.col_stride);
          }
        }
        buf[i] = v;
      }
    };
    uint32_t it = 0;
    if ((int64_t)blockIdx.x < n_row_tiles) fetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x, ++it) {
      mbar_wait(bar_a_empty, (it & 1) ^ 1);               // MMAs of the previous tile are done
#pragma unroll
      for (int i = 0; i < kChunksPerThread; ++i) {
        const int e = t + i * kLoaderThreads;
        int r, c;
        if (rows_contiguous) { r = e % kTileRows; c = e / kTileRows; }
        else                 { c = e % (kDim / 4); r = e / (kDim / 4); }
        const float4 v = buf[i];
        const float4 hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        const float4 lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z),
                                      to_tf32(v.w - hi.w));
        const int m = r >> 7, rr = r & 127;
        const uint32_t off = Smem::a + (uint32_t)(m * 2) * kABytesPart + operand_offset(kRowsPerMma, rr, 4 * c);
        *reinterpret_cast<float4*>(smem + off) = hi;
        *reinterpret_cast<float4*>(smem + off + kABytesPart) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy
      mbar_arrive(bar_a_full);
      if (tile + gridDim.x < n_row_tiles) fetch(tile + gridDim.x);
    }
  } else if (warp == kProducerWarp) {
    // ===================== B producer =====================
    if (lane == 0) {
      uint32_t step = 0;
      for (int64_t tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
        for (int j = 0; j < n_code_tiles; ++j, ++step) {
          const uint32_t s = step & 1, ph = (step >> 1) & 1;
          mbar_wait(bar_b_empty + 8 * s, ph ^ 1);
          mbar_expect_tx(bar_b_full + 8 * s, kBStageBytes);
          bulk_g2s(smem_base + Smem::b + s * kBStageBytes, b_tiles + (size_t)j * kBStageBytes,
                   kBStageBytes, bar_b_full + 8 * s);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t step = 0, it = 0;
      for (int64_t tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x, ++it) {
        mbar_wait(bar_a_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < n_code_tiles; ++j, ++step) {
          const uint32_t s = step & 1, ph = (step >> 1) & 1;
          mbar_wait(bar_b_full + 8 * s, ph);
          mbar_wait(bar_acc_empty + 8 * s, ph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b_base = smem_base + Smem::b + s * kBStageBytes;
#pragma unroll
          for (int m = 0; m < kMmaPerTile; ++m) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((s * kMmaPerTile + m) * kTileCodes);
            const uint32_t a_base = smem_base + Smem::a + (uint32_t)(m * 2) * kABytesPart;
            uint32_t acc = 0;
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              // (x_lo, b_hi), (x_hi, b_lo), (x_hi, b_hi): small terms first
              const uint32_t a_part = a_base + (term == 0 ? kABytesPart : 0);
              const uint32_t b_part = b_base + (term == 1 ? kBBytesPart : 0);
#pragma unroll
              for (int slab = 0; slab < kSlabs; ++slab) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t ad = umma_desc(a_part + slab * (kRowsPerMma * 128) + kk * 32);
                  const uint64_t bd = umma_desc(b_part + slab * (kTileCodes * 128) + kk * 32);
                  umma_tf32(d_tmem, ad, bd, kIdesc, acc);
                  acc = 1;
                }
              }
            }
          }
          umma_commit(bar_b_empty + 8 * s);       // the B stage may be refilled
          umma_commit(bar_acc_full + 8 * s);      // the accumulators may be read
        }
        umma_commit(bar_a_empty);                  // the row tile may be overwritten
      }
    }
  } else {
    // ===================== epilogue: argmin =====================
    uint32_t step = 0;
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    for (int64_t tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
      float best_s[kMmaPerTile];
      int best_i[kMmaPerTile];
#pragma unroll
      for (int m = 0; m < kMmaPerTile; ++m) { best_s[m] = INFINITY; best_i[m] = 0; }
      for (int j = 0; j < n_code_tiles; ++j, ++step) {
        const uint32_t s = step & 1, ph = (step >> 1) & 1;
        mbar_wait(bar_acc_full + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* e2t = e2s + j * kTileCodes;
#pragma unroll
        for (int m = 0; m < kMmaPerTile; ++m) {
          float v[kTileCodes];
          tmem_ld64(tmem_base + lane_field + (uint32_t)((s * kMmaPerTile + m) * kTileCodes), v);
#pragma unroll
          for (int c = 0; c < kTileCodes; c += 4) {
            const float4 ee = *reinterpret_cast<const float4*>(e2t + c);
            const float sc[4] = {v[c] + ee.x, v[c + 1] + ee.y, v[c + 2] + ee.z, v[c + 3] + ee.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // codes are visited in rising order: strict '<' keeps the lowest index on ties
              if (sc[q] < best_s[m]) { best_s[m] = sc[q]; best_i[m] = j * kTileCodes + c + q; }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_acc_empty + 8 * s);
      }
      const int64_t row0 = tile * kTileRows + warp * 32 + lane;
#pragma unroll
      for (int m = 0; m < kMmaPerTile; ++m) {
        const int64_t row = row0 + m * kRowsPerMma;
        if (row < n_rows) {
          out_index[row] = best_i[m];
          if (out_score) out_score[row] = best_s[m];
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
bool assign_tc_supported(const isi_rows_layout& lay, int64_t n_rows, int dim, int n_embed) {
  (void)lay;
  // small problems stay on the SIMT kernel: a 256-row tile per SM needs ~38k rows to fill
  // the machine once, and the launch is latency-bound below a few thousand rows
  return dim == kDim && n_embed <= kMaxCodes && n_rows >= 4096;
}

int launch_prepare_tc(const float* embed, int dim, int n_embed, const Prepared& p, cudaStream_t stream) {
  if (dim != kDim) return ISI_OK;   // nothing to prepare: the SIMT kernel serves this shape
  const int n_tiles = (n_embed + kTileCodes - 1) / kTileCodes;
  const int total = n_tiles * kTileCodes * kDim;
  int grid = (total + 255) / 256;
  if (grid > 4 * kNumSms) grid = 4 * kNumSms;
  vq_prepare_tc_kernel<<<grid, 256, 0, stream>>>(embed, n_embed, p);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_assign_tc(const float* x, const isi_rows_layout& lay, int64_t n_rows, int dim, int n_embed,
                     const Prepared& p, int64_t* out_index, float* out_score, cudaStream_t stream) {
  if (dim != kDim || n_embed > kMaxCodes) return ISI_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(vq_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Smem::total);
  if (e != cudaSuccess) return (int)e;
  const int64_t n_row_tiles = (n_rows + kTileRows - 1) / kTileRows;
  const int grid = (int)(n_row_tiles < kNumSms ? n_row_tiles : kNumSms);
  vq_assign_tc_kernel<<<grid, kThreads, Smem::total, stream>>>(
      x, lay, n_rows, n_embed, reinterpret_cast<const char*>(p.b_hi), p.e2, out_index, out_score);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

}  // namespace isi
