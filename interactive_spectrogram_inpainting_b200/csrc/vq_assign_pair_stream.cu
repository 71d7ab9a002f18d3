// Nearest-code search on a CTA pair (tcgen05 cta_group::2) for codebooks that do NOT fit in
// shared memory: D in {64, 128}, any K (BASELINE config 4: 4096 x 128).  Replaces
// bottleneck.py:55-61.
//
// Same arithmetic as the other tensor-core kernels (3xTF32 split of |e|^2 - 2 x.e, TMEM
// accumulators, argmin epilogue).  The pair issues M = 256 x N = 128 MMAs; each CTA keeps its
// 128 rows (hi + lo operands) resident for a whole sweep over the codebook and streams ITS
// half (64 codes) of every 128-code tile through a ring of shared-memory slots with
// cp.async.bulk.  A slot holds one PART (hi or lo) of a tile: the MMA terms that need b_hi
// (x_lo.b_hi, x_hi.b_hi) and the one that needs b_lo (x_hi.b_lo) consume different slots, so
// two slots already overlap copy and math even when the row operands take 128 KB (D = 128).
//
// Barriers: slot-full is local to each CTA (its own bulk copy); CTA 1 relays it to the leader
// with a remote arrive; slot-empty, accumulator-full and rows-free come from tcgen05.commit
// multicast to both CTAs; rows-ready and accumulator-free live in the leader and take remote
// arrivals.
#include "common.cuh"
#include "umma.cuh"

namespace isi {
namespace pstream {

using namespace umma;

constexpr int kRowsPerCta = 128;
constexpr int kPairRows = 256;                             // M of the pair MMA
constexpr int kTileCodes = 128;                            // N of the pair MMA
constexpr int kCodesPerCta = kTileCodes / 2;               // 64 codes of each tile per CTA
constexpr int kMaxCodes = 4096;                            // |e|^2 table in shared memory
constexpr int kFirstLoaderWarp = 4;
constexpr int kLoaderWarps = 8;
constexpr int kLoaderThreads = kLoaderWarps * 32;
constexpr int kProducerWarp = kFirstLoaderWarp + kLoaderWarps;   // 12 (lane 0 copies, lane 1 relays)
constexpr int kMmaWarp = kProducerWarp + 1;                       // 13
constexpr int kThreads = (kMmaWarp + 1) * 32;                     // 448
constexpr int kTmemCols = 256;                                    // 2 stages x 128 columns

constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileCodes >> 3) << 17) |
                            ((uint32_t)(kPairRows >> 4) << 24);

template <int D>
struct Cfg {
  static constexpr int kSlabs = D / 32;
  static constexpr int kAPart = kSlabs * kRowsPerCta * 128;          // one (hi|lo) row operand
  static constexpr int kABytes = 2 * kAPart;                          // 64 KB / 128 KB
  static constexpr int kSlotBytes = kSlabs * kCodesPerCta * 128;     // one part of a tile: 16 / 32 KB
  static constexpr int kSlots = D == 64 ? 4 : 2;
  static constexpr int kChunksPerThread = kRowsPerCta * (D / 4) / kLoaderThreads;   // 8 / 16
  static constexpr int a = 0;
  static constexpr int b = a + kABytes;
  static constexpr int e2 = b + kSlots * kSlotBytes;
  static constexpr int bars = e2 + kMaxCodes * 4;
  static constexpr int total = bars + 256;
};

}  // namespace pstream

using namespace pstream;

// image: for tile j, CTA rank r, part p (hi, lo): codes j*128 + r*64 + [0,64) as a K-major
// SWIZZLE_128B operand of -2E; parts are contiguous slots the kernel bulk-copies in order
template <int D>
__global__ void __launch_bounds__(256)
vq_prepare_pstream_kernel(const float* __restrict__ embed, int n_embed, int n_tiles, float* __restrict__ image) {
  using C = Cfg<D>;
  const int total = n_tiles * kTileCodes * D;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int code = e % (n_tiles * kTileCodes), d = e / (n_tiles * kTileCodes);
    const float v = code < n_embed ? -2.f * embed[(int64_t)d * n_embed + code] : 0.f;
    const float hi = to_tf32(v);
    const float lo = to_tf32(v - hi);
    const int j = code / kTileCodes, r = (code % kTileCodes) / kCodesPerCta, i = code % kCodesPerCta;
    // operand_offset() is written for 64-wide rows; wider rows add whole slabs
    const int slab = d >> 5, dd = d & 31;
    const size_t within = (size_t)slab * kCodesPerCta * 128 + operand_offset(kCodesPerCta, i, dd);
    const size_t base = ((size_t)(j * 2 + r) * 2) * C::kSlotBytes;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(image) + base + within) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(image) + base + C::kSlotBytes + within) = lo;
  }
}

template <int D>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
vq_assign_pstream_kernel(const float* __restrict__ x, isi_rows_layout lay, int64_t n_rows, int n_tiles,
                         const char* __restrict__ b_image, const float* __restrict__ e2_global,
                         int64_t* __restrict__ out_index, float* __restrict__ out_score) {
  using C = Cfg<D>;
  constexpr int S = C::kSlots;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = s32(smem);
  float* e2s = reinterpret_cast<float*>(smem + C::e2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::bars);
  const uint32_t bar_a_full = s32(bars + 0);        // leader: 2 x 256 loader arrivals
  const uint32_t bar_a_empty = s32(bars + 1);       // each CTA: commit multicast
  const uint32_t bar_acc_full = s32(bars + 2);      // [2] each CTA: commit multicast
  const uint32_t bar_acc_empty = s32(bars + 4);     // [2] leader: 2 x 128 epilogue arrivals
  const uint32_t bar_slot_full = s32(bars + 6);     // [S] each CTA: its bulk copy landed
  const uint32_t bar_slot_empty = s32(bars + 10);   // [S] each CTA: commit multicast
  const uint32_t bar_peer_full = s32(bars + 14);    // [S] leader: CTA 1's slot landed (relayed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

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
    for (int i = 0; i < S; ++i) {
      mbar_init(bar_slot_full + 8 * i, 1);
      mbar_init(bar_slot_empty + 8 * i, 1);
      mbar_init(bar_peer_full + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  for (int k = threadIdx.x; k < n_tiles * kTileCodes; k += kThreads) e2s[k] = e2_global[k];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // per pair tile: n_tiles code tiles x 2 parts = 2*n_tiles slot fills, in this order:
  // fill f = 2*j + p  (p = 0: hi part of tile j, p = 1: lo part)
  if (warp >= kFirstLoaderWarp && warp < kFirstLoaderWarp + kLoaderWarps) {
    // ===================== x loader / splitter (both CTAs) =====================
    const int t = threadIdx.x - kFirstLoaderWarp * 32;
    const bool rows_contiguous = (lay.row_stride == 1 && lay.col_stride != 1);
    const bool vec_ok = (lay.col_stride == 1) && ((lay.row_stride & 3) == 0) &&
                        ((lay.batch_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    // D = 128 needs 16 chunks per thread: prefetch half of them (the second half is loaded
    // after the split of the first) to stay inside the register budget
    constexpr int kPre = C::kChunksPerThread > 8 ? 8 : C::kChunksPerThread;
    float4 buf[kPre];
    auto chunk_pos = [&](int i, int& r, int& c) {
      const int e = t + i * kLoaderThreads;
      if (rows_contiguous) { r = e % kRowsPerCta; c = e / kRowsPerCta; }
      else                 { c = e % (D / 4); r = e / (D / 4); }
    };
    auto load_chunk = [&](int64_t row0, int i) {
      int r, c;
      chunk_pos(i, r, c);
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
      return v;
    };
    auto store_chunk = [&](int i, float4 v) {
      int r, c;
      chunk_pos(i, r, c);
      const float4 hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      const float4 lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z),
                                    to_tf32(v.w - hi.w));
      const int d = 4 * c;
      const uint32_t off = C::a + (uint32_t)(d >> 5) * (kRowsPerCta * 128) + operand_offset(kRowsPerCta, r, d & 31);
      *reinterpret_cast<float4*>(smem + off) = hi;
      *reinterpret_cast<float4*>(smem + off + C::kAPart) = lo;
    };
    uint32_t it = 0;
    auto row0_of = [&](int64_t pt) { return pt * kPairRows + (int64_t)rank * kRowsPerCta; };
    if (cluster_id < n_pair_tiles) {
#pragma unroll
      for (int i = 0; i < kPre; ++i) buf[i] = load_chunk(row0_of(cluster_id), i);
    }
    for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters, ++it) {
      mbar_wait_cluster(bar_a_empty, (it & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < kPre; ++i) store_chunk(i, buf[i]);
#pragma unroll
      for (int i = kPre; i < C::kChunksPerThread; ++i) store_chunk(i, load_chunk(row0_of(pt), i));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_cluster(bar_a_full, 0);
      if (pt + n_clusters < n_pair_tiles) {
#pragma unroll
        for (int i = 0; i < kPre; ++i) buf[i] = load_chunk(row0_of(pt + n_clusters), i);
      }
    }
  } else if (warp == kProducerWarp) {
    if (lane == 0) {
      // ===================== B producer: this CTA's half of every tile part =====================
      uint32_t fill = 0;
      for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters) {
        for (int f = 0; f < 2 * n_tiles; ++f, ++fill) {
          const uint32_t s = fill % S, ph = (fill / S) & 1;
          mbar_wait_cluster(bar_slot_empty + 8 * s, ph ^ 1);
          mbar_expect_tx(bar_slot_full + 8 * s, C::kSlotBytes);
          const int j = f >> 1, p = f & 1;
          const char* src = b_image + ((size_t)(j * 2 + rank) * 2 + p) * C::kSlotBytes;
          bulk_g2s(smem_base + C::b + s * C::kSlotBytes, src, C::kSlotBytes, bar_slot_full + 8 * s);
        }
      }
    } else if (lane == 1 && rank == 1) {
      // ===================== relay: tell the leader that CTA 1's slot has landed =====================
      uint32_t fill = 0;
      for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters) {
        for (int f = 0; f < 2 * n_tiles; ++f, ++fill) {
          const uint32_t s = fill % S, ph = (fill / S) & 1;
          mbar_wait(bar_slot_full + 8 * s, ph);
          mbar_arrive_cluster(bar_peer_full + 8 * s, 0);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      uint32_t step = 0, it = 0, fill = 0;
      const uint32_t a_base = smem_base + C::a;
      auto issue_term = [&](uint32_t d_tmem, uint32_t a_part, uint32_t b_slot, uint32_t& acc) {
#pragma unroll
        for (int slab = 0; slab < C::kSlabs; ++slab) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = umma_desc(a_part + slab * (kRowsPerCta * 128) + kk * 32);
            const uint64_t bd = umma_desc(b_slot + slab * (kCodesPerCta * 128) + kk * 32);
            umma_tf32_2cta(d_tmem, ad, bd, kIdesc, acc);
            acc = 1;
          }
        }
      };
      for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters, ++it) {
        mbar_wait_cluster(bar_a_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < n_tiles; ++j, ++step) {
          const uint32_t ts = step & 1, aph = (step >> 1) & 1;
          mbar_wait_cluster(bar_acc_empty + 8 * ts, aph ^ 1);
          const uint32_t d_tmem = tmem_base + ts * kTileCodes;
          uint32_t acc = 0;
          {   // hi part of the tile: x_lo.b_hi then x_hi.b_hi
            const uint32_t s = fill % S, ph = (fill / S) & 1;
            mbar_wait(bar_slot_full + 8 * s, ph);
            mbar_wait_cluster(bar_peer_full + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t b_slot = smem_base + C::b + s * C::kSlotBytes;
            issue_term(d_tmem, a_base + C::kAPart, b_slot, acc);
            issue_term(d_tmem, a_base, b_slot, acc);
            umma_commit_2cta(bar_slot_empty + 8 * s);
            ++fill;
          }
          {   // lo part: x_hi.b_lo
            const uint32_t s = fill % S, ph = (fill / S) & 1;
            mbar_wait(bar_slot_full + 8 * s, ph);
            mbar_wait_cluster(bar_peer_full + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t b_slot = smem_base + C::b + s * C::kSlotBytes;
            issue_term(d_tmem, a_base, b_slot, acc);
            umma_commit_2cta(bar_slot_empty + 8 * s);
            ++fill;
          }
          umma_commit_2cta(bar_acc_full + 8 * ts);
        }
        umma_commit_2cta(bar_a_empty);
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue: argmin over this CTA's 128 rows =====================
    uint32_t step = 0;
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    for (int64_t pt = cluster_id; pt < n_pair_tiles; pt += n_clusters) {
      float best_s = INFINITY;
      int best_i = 0;
      for (int j = 0; j < n_tiles; ++j, ++step) {
        const uint32_t ts = step & 1, ph = (step >> 1) & 1;
        mbar_wait_cluster(bar_acc_full + 8 * ts, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* e2t = e2s + j * kTileCodes;
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
              if (sc[q] < best_s) { best_s = sc[q]; best_i = j * kTileCodes + q4 * 64 + c + q; }
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
  cluster_sync();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool assign_pstream_supported(const isi_rows_layout& lay, int64_t n_rows, int dim, int n_embed) {
  (void)lay;
  return (dim == 64 || dim == 128) && n_embed <= kMaxCodes && n_rows >= 4096;
}

int launch_prepare_pstream(const float* embed, int dim, int n_embed, const Prepared& p, cudaStream_t stream) {
  if (!(dim == 64 || dim == 128) || n_embed > kMaxCodes) return ISI_OK;
  if (dim == 64 && n_embed <= 512) return ISI_OK;          // the resident pair kernel owns b_pair
  const int n_tiles = (n_embed + kTileCodes - 1) / kTileCodes;
  const int total = n_tiles * kTileCodes * dim;
  int grid = (total + 255) / 256;
  if (grid > 8 * kNumSms) grid = 8 * kNumSms;
  if (dim == 64) vq_prepare_pstream_kernel<64><<<grid, 256, 0, stream>>>(embed, n_embed, n_tiles, p.b_pair);
  else           vq_prepare_pstream_kernel<128><<<grid, 256, 0, stream>>>(embed, n_embed, n_tiles, p.b_pair);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

template <int D>
static int launch_pstream_t(const float* x, const isi_rows_layout& lay, int64_t n_rows, int n_embed,
                            const Prepared& p, int64_t* out_index, float* out_score, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(vq_assign_pstream_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg<D>::total);
  if (e != cudaSuccess) return (int)e;
  const int n_tiles = (n_embed + kTileCodes - 1) / kTileCodes;
  const int64_t n_pair_tiles = (n_rows + kPairRows - 1) / kPairRows;
  const int clusters = (int)(n_pair_tiles < kNumSms / 2 ? n_pair_tiles : kNumSms / 2);
  vq_assign_pstream_kernel<D><<<2 * clusters, kThreads, Cfg<D>::total, stream>>>(
      x, lay, n_rows, n_tiles, reinterpret_cast<const char*>(p.b_pair), p.e2, out_index, out_score);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_assign_pstream(const float* x, const isi_rows_layout& lay, int64_t n_rows, int dim, int n_embed,
                          const Prepared& p, int64_t* out_index, float* out_score, cudaStream_t stream) {
  if (n_embed > kMaxCodes || (dim == 64 && n_embed <= 512)) return ISI_ERR_UNSUPPORTED;
  if (dim == 64) return launch_pstream_t<64>(x, lay, n_rows, n_embed, p, out_index, out_score, stream);
  if (dim == 128) return launch_pstream_t<128>(x, lay, n_rows, n_embed, p, out_index, out_score, stream);
  return ISI_ERR_UNSUPPORTED;
}

}  // namespace isi
