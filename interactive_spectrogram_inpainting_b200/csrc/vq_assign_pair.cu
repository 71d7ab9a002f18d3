// Nearest-code search on a CTA PAIR (tcgen05 cta_group::2) with the whole codebook resident
// in shared memory.  Replaces bottleneck.py:55-61 for the deployed shape (D = 64, K <= 512).
//
// Same arithmetic as vq_assign_tc.cu (3xTF32 split of score = |e|^2 - 2 x.e, accumulators in
// TMEM, argmin in the epilogue); what changes is the data movement:
//  * two CTAs of a cluster share one MMA: M = 256 (128 rows per CTA), N = 256 codes per
//    instruction, of which each CTA holds 128 in ITS shared memory.  The pre-split,
//    pre-swizzled hi/lo image of a CTA's half of the codebook (128 KB at K = 512) is loaded
//    ONCE per CTA with bulk copies and stays resident: no codebook streaming at all;
//  * N = 256 MMAs run at the full 2048 MAC/clk/SM (tools/umma_probe.cu: N = 64 reaches 66 %);
//  * each CTA's 8 loader warps prefetch their next 128 rows into registers while the MMAs
//    run, so only the hi/lo split + st.shared sits between two tiles;
//  * TMEM holds 2 stages x 256 columns per CTA: the argmin of one 256-code tile overlaps the
//    MMAs of the next.
// Synchronisation: operands-ready and accumulator-free barriers live in the leader CTA (rank
// 0) and receive remote arrivals from the peer; MMA completion is multicast to both CTAs.
#include "common.cuh"
#include "umma.cuh"

namespace isi {
namespace pair {

using namespace umma;

constexpr int kDim = 64;
constexpr int kSlabs = kDim / 32;
constexpr int kRowsPerCta = 128;
constexpr int kPairRows = 2 * kRowsPerCta;                 // 256 rows per MMA (M = 256)
constexpr int kTileCodes = 256;                            // codes per MMA (N = 256)
constexpr int kCodesPerCta = kTileCodes / 2;               // 128 of them in each CTA's smem
constexpr int kMaxTiles = 2;                               // K <= 512
constexpr int kPartBytes = kSlabs * 128 * 128;             // one (hi|lo) 128-row operand: 32 KB
constexpr int kABytes = 2 * kPartBytes;                    // 64 KB
constexpr int kBTileBytes = 2 * kPartBytes;                // 64 KB per 256-code tile per CTA
constexpr int kTmemCols = 512;                             // 2 stages x 256 columns
constexpr int kFirstLoaderWarp = 4;
constexpr int kLoaderWarps = 8;
constexpr int kLoaderThreads = kLoaderWarps * 32;
constexpr int kChunksPerThread = kRowsPerCta * (kDim / 4) / kLoaderThreads;   // 8
constexpr int kProducerWarp = kFirstLoaderWarp + kLoaderWarps;
constexpr int kMmaWarp = kProducerWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;              // 448

// instruction descriptor: D=F32, A=B=TF32, both K-major, N=256, M=256 (cta_group::2)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileCodes >> 3) << 17) |
                            ((uint32_t)(kPairRows >> 4) << 24);

struct Smem {
  static constexpr int a = 0;
  static constexpr int b = a + kABytes;
  static constexpr int e2 = b + kMaxTiles * kBTileBytes;
  static constexpr int bars = e2 + kMaxTiles * kTileCodes * 4;
  static constexpr int total = bars + 128;
};

}  // namespace pair

using namespace pair;

// per CTA rank r, tile n, part p (hi, lo): codes n*256 + r*128 + [0,128) as a K-major
// SWIZZLE_128B operand of -2E -- the bytes the kernel bulk-copies into shared memory
__global__ void __launch_bounds__(256)
vq_prepare_pair_kernel(const float* __restrict__ embed, int n_embed, int n_tiles, float* __restrict__ image) {
  const int total = n_tiles * kTileCodes * kDim;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int code = e % (n_tiles * kTileCodes), d = e / (n_tiles * kTileCodes);
    const float v = code < n_embed ? -2.f * embed[(int64_t)d * n_embed + code] : 0.f;
    const float hi = to_tf32(v);
    const float lo = to_tf32(v - hi);
    const int n = code / kTileCodes, r = (code % kTileCodes) / kCodesPerCta, i = code % kCodesPerCta;
    const size_t off = (size_t)(r * n_tiles + n) * kBTileBytes + operand_offset(kCodesPerCta, i, d);
    *reinterpret_cast<float*>(reinterpret_cast<char*>(image) + off) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(image) + off + kPartBytes) = lo;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
vq_assign_pair_kernel(const float* __restrict__ x, isi_rows_layout lay, int64_t n_rows, int n_tiles,
                      const char* __restrict__ b_image, const float* __restrict__ e2_global,
                      int64_t* __restrict__ out_index, float* __restrict__ out_score) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = s32(smem);
  float* e2s = reinterpret_cast<float*>(smem + Smem::e2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  const uint32_t bar_a_full = s32(bars + 0);      // leader: 2 x 256 loader arrivals
  const uint32_t bar_a_empty = s32(bars + 1);     // each CTA: MMA commit (multicast)
  const uint32_t bar_acc_full = s32(bars + 2);    // [2] each CTA: MMA commit (multicast)
  const uint32_t bar_acc_empty = s32(bars + 4);   // [2] leader: 2 x 128 epilogue arrivals
  const uint32_t bar_b_ready = s32(bars + 6);     // each CTA: its codebook image has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t n_pair_tiles = (n_rows + kPairRows - 1) / kPairRows;
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 2 * kLoaderThreads);
    mbar_init(bar_a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 2 * 128);
    }
    mbar_init(bar_b_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  for (int k = threadIdx.x; k < n_tiles * kTileCodes; k += kThreads) e2s[k] = e2_global[k];
  __syncthreads();
  if (warp == kProducerWarp && lane == 0) {
    // the resident half-codebook of this CTA: n_tiles x 64 KB, in 32 KB bulk copies
    const uint32_t bytes = (uint32_t)n_tiles * kBTileBytes;
    mbar_expect_tx(bar_b_ready, bytes);
    const char* src = b_image + (size_t)rank * n_tiles * kBTileBytes;
    for (uint32_t off = 0; off < bytes; off += kPartBytes)
      bulk_g2s(smem_base + Smem::b + off, src + off, kPartBytes, bar_b_ready);
  }
  mbar_wait(bar_b_ready, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();          // barriers initialised, TMEM allocated, both codebook halves resident
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kFirstLoaderWarp && warp < kFirstLoaderWarp + kLoaderWarps) {
    // ===================== x loader / splitter (both CTAs) =====================
    const int t = threadIdx.x - kFirstLoaderWarp * 32;
    const bool rows_contiguous = (lay.row_stride == 1 && lay.col_stride != 1);
    const bool vec_ok = (lay.col_stride == 1) && ((lay.row_stride & 3) == 0) &&
                        ((lay.batch_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    float4 buf[kChunksPerThread];
    auto fetch = [&](int64_t pt) {
      const int64_t row0 = pt * kPairRows + (int64_t)rank * kRowsPerCta;
#pragma unroll
      for (int i = 0; i < kChunksPerThread; ++i) {
        const int e = t + i * kLoaderThreads;
        int r, c;
        if (rows_contiguous) { r = e % kRowsPerCta; c = e / kRowsPerCta; }
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
    if (cluster_id < n_pair_tiles) fetch(cluster_id);
    for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters, ++it) {
      mbar_wait_cluster(bar_a_empty, (it & 1) ^ 1);       // the pair's MMAs on the old tile are done
#pragma unroll
      for (int i = 0; i < kChunksPerThread; ++i) {
        const int e = t + i * kLoaderThreads;
        int r, c;
        if (rows_contiguous) { r = e % kRowsPerCta; c = e / kRowsPerCta; }
        else                 { c = e % (kDim / 4); r = e / (kDim / 4); }
        const float4 v = buf[i];
        const float4 hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        const float4 lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z),
                                      to_tf32(v.w - hi.w));
        const uint32_t off = Smem::a + operand_offset(kRowsPerCta, r, 4 * c);
        *reinterpret_cast<float4*>(smem + off) = hi;
        *reinterpret_cast<float4*>(smem + off + kPartBytes) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_cluster(bar_a_full, 0);                 // tell the leader's MMA issuer
      if (pt + n_clusters < n_pair_tiles) fetch(pt + n_clusters);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      uint32_t step = 0, it = 0;
      for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters, ++it) {
        mbar_wait_cluster(bar_a_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int n = 0; n < n_tiles; ++n, ++step) {
          const uint32_t ts = step & 1, ph = (step >> 1) & 1;
          mbar_wait_cluster(bar_acc_empty + 8 * ts, ph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem_base + ts * kTileCodes;
          const uint32_t a_base = smem_base + Smem::a;
          const uint32_t b_base = smem_base + Smem::b + (uint32_t)n * kBTileBytes;
          uint32_t acc = 0;
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            // (x_lo, b_hi), (x_hi, b_lo), (x_hi, b_hi): small terms first
            const uint32_t a_part = a_base + (term == 0 ? kPartBytes : 0);
            const uint32_t b_part = b_base + (term == 1 ? kPartBytes : 0);
#pragma unroll
            for (int slab = 0; slab < kSlabs; ++slab) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = umma_desc(a_part + slab * (kRowsPerCta * 128) + kk * 32);
                const uint64_t bd = umma_desc(b_part + slab * (kCodesPerCta * 128) + kk * 32);
                umma_tf32_2cta(d_tmem, ad, bd, kIdesc, acc);
                acc = 1;
              }
            }
          }
          umma_commit_2cta(bar_acc_full + 8 * ts);     // both CTAs may read their accumulators
        }
        umma_commit_2cta(bar_a_empty);                  // both CTAs may overwrite their rows
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue: argmin over this CTA's 128 rows =====================
    uint32_t step = 0;
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters) {
      float best_s = INFINITY;
      int best_i = 0;
      for (int n = 0; n < n_tiles; ++n, ++step) {
        const uint32_t ts = step & 1, ph = (step >> 1) & 1;
        mbar_wait_cluster(bar_acc_full + 8 * ts, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* e2t = e2s + n * kTileCodes;
#pragma unroll 1
        for (int q4 = 0; q4 < kTileCodes / 64; ++q4) {
          float v[64];
          tmem_ld64(tmem_base + lane_field + ts * kTileCodes + q4 * 64, v);
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            const float4 ee = *reinterpret_cast<const float4*>(e2t + q4 * 64 + c);
            const float sc[4] = {v[c] + ee.x, v[c + 1] + ee.y, v[c + 2] + ee.z, v[c + 3] + ee.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // rising code order + strict '<': the lowest index wins exact ties
              if (sc[q] < best_s) { best_s = sc[q]; best_i = n * kTileCodes + q4 * 64 + c + q; }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_cluster(bar_acc_empty + 8 * ts, 0);
      }
      const int64_t row = pt * kPairRows + (int64_t)rank * kRowsPerCta + warp * 32 + lane;
      if (row < n_rows) {
        out_index[row] = best_i;
        if (out_score) out_score[row] = best_s;
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();          // nobody leaves while the peer can still touch its smem / TMEM
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool assign_pair_supported(const isi_rows_layout& lay, int64_t n_rows, int dim, int n_embed) {
  (void)lay;
  return dim == kDim && n_embed <= kMaxTiles * kTileCodes && n_rows >= 4096;
}

int launch_prepare_pair(const float* embed, int dim, int n_embed, const Prepared& p, cudaStream_t stream) {
  if (dim != kDim || n_embed > kMaxTiles * kTileCodes) return ISI_OK;
  const int n_tiles = (n_embed + kTileCodes - 1) / kTileCodes;
  const int total = n_tiles * kTileCodes * kDim;
  int grid = (total + 255) / 256;
  if (grid > 4 * kNumSms) grid = 4 * kNumSms;
  vq_prepare_pair_kernel<<<grid, 256, 0, stream>>>(embed, n_embed, n_tiles, p.b_pair);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_assign_pair(const float* x, const isi_rows_layout& lay, int64_t n_rows, int dim, int n_embed,
                       const Prepared& p, int64_t* out_index, float* out_score, cudaStream_t stream) {
  if (dim != kDim || n_embed > kMaxTiles * kTileCodes) return ISI_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(vq_assign_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Smem::total);
  if (e != cudaSuccess) return (int)e;
  const int n_tiles = (n_embed + kTileCodes - 1) / kTileCodes;
  const int64_t n_pair_tiles = (n_rows + kPairRows - 1) / kPairRows;
  const int clusters = (int)(n_pair_tiles < kNumSms / 2 ? n_pair_tiles : kNumSms / 2);
  vq_assign_pair_kernel<<<2 * clusters, kThreads, Smem::total, stream>>>(
      x, lay, n_rows, n_tiles, reinterpret_cast<const char*>(p.b_pair), p.e2, out_index, out_score);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

}  // namespace isi
