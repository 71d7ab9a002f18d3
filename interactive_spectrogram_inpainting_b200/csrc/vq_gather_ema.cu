// Codebook preparation, code lookup, commitment partials, EMA statistics (a
// warp-aggregated segmented reduction) and the EMA codebook update.
//
// Replaces bottleneck.py:75-100,103-104 of the reference; see include/isi_b200.h
// for the statement-by-statement mapping.
#include "common.cuh"
#include "umma.cuh"

namespace isi {

// ---------------------------------------------------------------------------
// prepare: e2[k] = sum_d E[d][k]^2 (d ascending, like embed.pow(2).sum(0)),
//          et[k][d] = ed[d][k] = E[d][k]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_prepare_kernel(const float* __restrict__ embed, int dim, int n_embed, Prepared p) {
  __shared__ float tile[32][33];
  const int tiles_k = (n_embed + 31) / 32, tiles_d = (dim + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int t = blockIdx.x; t < tiles_k * tiles_d; t += gridDim.x) {
    const int k0 = (t % tiles_k) * 32, d0 = (t / tiles_k) * 32;
    for (int j = ty; j < 32; j += 8) {
      int d = d0 + j, k = k0 + tx;
      float v = (d < dim && k < n_embed) ? embed[(int64_t)d * n_embed + k] : 0.f;
      tile[j][tx] = v;
      if (d < dim && k < n_embed) p.ed[(int64_t)d * n_embed + k] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      int k = k0 + j, d = d0 + tx;
      if (d < dim && k < n_embed) p.et[(int64_t)k * dim + d] = tile[tx][j];
    }
    __syncthreads();
  }
  const int kp = padded_codes(n_embed);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < kp; k += gridDim.x * blockDim.x) {
    float s = 0.f;
    if (k < n_embed) {
      for (int d = 0; d < dim; ++d) { float v = embed[(int64_t)d * n_embed + k]; s = fmaf(v, v, s); }
    } else {
      s = INFINITY;  // padded codes can never win the argmin
    }
    p.e2[k] = s;
  }
}

// ---------------------------------------------------------------------------
// gather + diff partials + EMA statistics
// ---------------------------------------------------------------------------
constexpr int kGatherRows = 128;   // rows per CTA tile, 32 per warp
constexpr int kGatherThreads = 128;

__global__ void __launch_bounds__(kGatherThreads)
vq_gather_stats_kernel(const float* __restrict__ x, isi_rows_layout xl,
                       const int64_t* __restrict__ index, int64_t n_rows, int dim, int n_embed,
                       const float* __restrict__ et, float* __restrict__ out_q,
                       isi_rows_layout ql, float* __restrict__ stats, int counts_only,
                       double* __restrict__ partials, int32_t* __restrict__ status_flag) {
  extern __shared__ __align__(16) float smem[];
  const int pitch = dim + 1;
  float* xs = smem;                                   // [kGatherRows][pitch]
  int* codes = (int*)(smem + kGatherRows * pitch);    // [kGatherRows]
  __shared__ double warp_part[kGatherThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kGatherRows;
  const int rows_here = (int)min((int64_t)kGatherRows, n_rows - row0);

  // 1. indices of the tile
  for (int r = tid; r < kGatherRows; r += kGatherThreads) {
    int c = -1;
    if (r < rows_here) {
      int64_t v = index[row0 + r];
      if (v >= 0 && v < n_embed) c = (int)v;
      else if (status_flag) atomicExch(status_flag, 1);
    }
    codes[r] = c;
  }
  // 2. x tile, global reads coalesced along whichever axis is contiguous
  if (x) {
    const bool rows_contiguous = (xl.row_stride == 1 && xl.col_stride != 1);
    for (int e = tid; e < kGatherRows * dim; e += kGatherThreads) {
      int r, d;
      if (rows_contiguous) { r = e % kGatherRows; d = e / kGatherRows; }
      else                 { d = e % dim; r = e / dim; }
      float v = 0.f;
      if (r < rows_here) v = x[row_offset(xl, row0 + r) + (int64_t)d * xl.col_stride];
      xs[r * pitch + d] = v;
    }
  }
  __syncthreads();

  // 3. EMA statistics: each warp owns 32 rows; rows that share a code are summed
  //    in registers first, so one reduction per (distinct code, d) reaches memory.
  const int my_code = codes[warp * 32 + lane];
  if (stats) {
    const unsigned peers = __match_any_sync(0xffffffffu, my_code);
    const bool leader = (my_code >= 0) && (lane == __ffs(peers) - 1);
    unsigned todo = __ballot_sync(0xffffffffu, leader);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned members = __shfl_sync(0xffffffffu, peers, src);
      const int c = __shfl_sync(0xffffffffu, my_code, src);
      if (lane == 0) atomicAdd(&stats[c], (float)__popc(members));
      if (!counts_only) {
        float* dst = stats + n_embed + (int64_t)c * dim;
        for (int d = lane; d < dim; d += 32) {
          float s = 0.f;
          for (unsigned m = members; m; m &= m - 1)
            s += xs[(warp * 32 + __ffs(m) - 1) * pitch + d];
          atomicAdd(&dst[d], s);
        }
      }
    }
  }

  // 4. lookup, commitment partial; q overwrites the x tile in shared memory
  float sq = 0.f;
  for (int i = 0; i < 32; ++i) {
    const int r = warp * 32 + i;
    const int c = __shfl_sync(0xffffffffu, my_code, i);
    if (r < rows_here) {
      for (int d = lane; d < dim; d += 32) {
        float q = (c >= 0) ? et[(int64_t)c * dim + d] : 0.f;
        if (x) {
          float xv = xs[r * pitch + d];
          float t = q - xv;
          sq = fmaf(t, t, sq);
          q = xv + t;                      // forward value of bottleneck.py:95
        }
        xs[r * pitch + d] = q;
      }
    }
  }
  sq = warp_sum(sq);
  if (lane == 0) warp_part[warp] = (double)sq;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kGatherThreads / 32; ++w) s += warp_part[w];
    partials[blockIdx.x] = s;
  }

  // 5. write the q tile with the output's own contiguous axis fastest
  if (out_q) {
    const bool rows_contiguous = (ql.row_stride == 1 && ql.col_stride != 1);
    for (int e = tid; e < kGatherRows * dim; e += kGatherThreads) {
      int r, d;
      if (rows_contiguous) { r = e % kGatherRows; d = e / kGatherRows; }
      else                 { d = e % dim; r = e / dim; }
      if (r < rows_here)
        out_q[row_offset(ql, row0 + r) + (int64_t)d * ql.col_stride] = xs[r * pitch + d];
    }
  }
}

// ---------------------------------------------------------------------------
// Row-major fast path (the layout of channels_last activations and of [N, D] tensors):
// LPR = D/4 lanes hold one row as float4s, so every global access is a 16-byte vector and a
// warp moves 512 B per instruction.  Each lane group owns kRowsPerGroup CONSECUTIVE rows and
// keeps the running sum of the current run of equal codes in registers: one vector reduction
// (red.global.add.v4.f32) per run reaches memory, not one per row.  Neighbouring
// spectrogram positions often share a code, and a collapsed codebook (all rows -> one code)
// degrades to one reduction per group instead of serialising a million atomics on one line.
// ---------------------------------------------------------------------------
template <int LPR>   // lanes per row = D / 4
__global__ void __launch_bounds__(256)
vq_gather_rowmajor_kernel(const float* __restrict__ x, int64_t x_row_stride,
                          const int64_t* __restrict__ index, int64_t n_rows, int n_embed,
                          const float* __restrict__ et, float* __restrict__ out_q,
                          int64_t q_row_stride, float* __restrict__ stats, int counts_only,
                          double* __restrict__ partials, int32_t* __restrict__ status_flag) {
  constexpr int D = LPR * 4;
  constexpr int kGroups = 256 / LPR;
  constexpr int kRows = kGatherRows / kGroups;          // consecutive rows per lane group
  static_assert(kGatherRows % kGroups == 0, "tile must split evenly over the lane groups");
  __shared__ double warp_part[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = tid / LPR, gl = tid % LPR;          // lane group, lane within the group
  const int64_t row0 = (int64_t)blockIdx.x * kGatherRows + (int64_t)group * kRows;

  float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
  int run_code = -1, run_len = 0;
  float sq = 0.f;
  auto flush = [&]() {
    if (run_code >= 0 && stats) {
      if (gl == 0) atomicAdd(&stats[run_code], (float)run_len);
      if (!counts_only)
        atomicAdd(reinterpret_cast<float4*>(stats + n_embed + (int64_t)run_code * D) + gl, run);
    }
  };
#pragma unroll 4
  for (int i = 0; i < kRows; ++i) {
    const int64_t row = row0 + i;
    if (row >= n_rows) break;
    const int64_t v = __ldg(index + row);
    const bool ok = (v >= 0 && v < n_embed);
    if (!ok && gl == 0 && status_flag) atomicExch(status_flag, 1);
    const int c = ok ? (int)v : -1;
    float4 q = ok ? __ldg(reinterpret_cast<const float4*>(et + (int64_t)c * D) + gl)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
    if (x) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + row * x_row_stride) + gl);
      const float4 t = make_float4(q.x - xv.x, q.y - xv.y, q.z - xv.z, q.w - xv.w);
      sq = fmaf(t.x, t.x, sq); sq = fmaf(t.y, t.y, sq); sq = fmaf(t.z, t.z, sq); sq = fmaf(t.w, t.w, sq);
      q = make_float4(xv.x + t.x, xv.y + t.y, xv.z + t.z, xv.w + t.w);   // bottleneck.py:95
      if (stats && !counts_only) {
        if (c != run_code) { flush(); run = make_float4(0.f, 0.f, 0.f, 0.f); run_code = c; run_len = 0; }
        run.x += xv.x; run.y += xv.y; run.z += xv.z; run.w += xv.w;
        ++run_len;
      }
    }
    if (stats && (counts_only || !x)) {
      if (c != run_code) { flush(); run_code = c; run_len = 0; }
      ++run_len;
    }
    if (out_q) reinterpret_cast<float4*>(out_q + row * q_row_stride)[gl] = q;
  }
  flush();
  sq = warp_sum(sq);
  if (lane == 0) warp_part[warp] = (double)sq;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += warp_part[w];
    partials[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------
// Training statistics with CTA-private accumulators (row-major input, K*D*4 <= ~140 KB):
// the segmented reduction proper.  Persistent CTAs each keep a full [K, D] accumulator in
// shared memory; inside a CTA every code is OWNED by one warp (code % 16), so rows are
// added with plain ld/add/st -- no atomics, no conflicts.  Global memory sees one vector
// reduction per (CTA, used code) at the very end: 148 x 128 KB at most, instead of 256 B of
// atomics per row.
// ---------------------------------------------------------------------------
constexpr int kStatsWarps = 16;                     // consumer warps (+ producer + dispatcher)
constexpr int kStatsThreads = (kStatsWarps + 2) * 32;
constexpr int kGranuleRows = 64;                    // rows per pipeline stage
constexpr int kHotElectionGranule = 1;              // the hot code is elected when this granule is listed

// per-stage work lists: for every consumer warp the granule rows whose code it owns
struct StageLists {
  unsigned char row[kStatsWarps][kGranuleRows];     // row ids, grouped by owner
  int count[kStatsWarps];
};

// Rows must be fully contiguous ([N, D], row stride D).  Three roles, no CTA-wide barrier:
//  * a producer warp streams 64-row granules (rows AND their int64 codes) into a ring of
//    shared-memory stages with cp.async.bulk;
//  * a dispatcher warp turns each granule's codes into per-owner row lists (owner = code % 16;
//    two match.any + prefix popcounts per granule);
//  * 16 consumer warps each take 4 rows of the granule for lookup / commitment / output (the
//    codewords are requested from L2 one granule ahead), then walk THEIR list two rows at a
//    time and add those rows to the CTA-private [K, D] accumulator with plain ld/add/st -- a
//    code is only ever touched by its owner, so there are no atomics and no conflicts -- and
//    finally signal the stage empty;
//  * ownership by code is unbalanced when one code is very popular (silence; a collapsed
//    codebook): the dispatcher elects the CTA's hot code from the first stages, keeps its rows
//    out of the lists, and every consumer warp sums the hot rows among its own 4 rows of each
//    granule in registers (merged into the accumulator once, at the end).
// Global memory sees one red.global.add.v4.f32 per (CTA, used code) at the very end.
template <int D, int STAGES>
__global__ void __launch_bounds__(kStatsThreads, 1)
vq_gather_stats_smem_kernel(const float* __restrict__ x, const int64_t* __restrict__ index,
                            int64_t n_rows, int n_embed, const float* __restrict__ et,
                            float* __restrict__ out_q, int64_t q_row_stride,
                            float* __restrict__ stats, double* __restrict__ partials,
                            int32_t* __restrict__ status_flag) {
  constexpr int VPL = D / 32;                       // accumulator floats per lane
  constexpr int C4 = D / 4;                         // float4 chunks (= lanes) per row
  constexpr int kRowsPerInstr = 32 / C4;            // rows one warp instruction covers
  constexpr int kRowsPerWarp = kGranuleRows / kStatsWarps;   // 4
  extern __shared__ __align__(16) float smem[];     // stage offsets below are 128-byte multiples
  float* acc = smem;                                // [K][D]
  float* cnt = acc + (size_t)n_embed * D;           // [K]
  float* xs = cnt + ((n_embed + 31) & ~31);         // [STAGES][kGranuleRows][D]
  long long* idx = reinterpret_cast<long long*>(xs + STAGES * kGranuleRows * D);   // [STAGES][kGranuleRows]
  StageLists* lists = reinterpret_cast<StageLists*>(idx + STAGES * kGranuleRows);  // [STAGES]
  uint64_t* full = reinterpret_cast<uint64_t*>(lists + STAGES);
  uint64_t* ready = full + STAGES;
  uint64_t* empty = ready + STAGES;
  int* hot_slot = reinterpret_cast<int*>(empty + STAGES);   // the CTA's hot code (-1: none)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n_embed * D + ((n_embed + 31) & ~31); i += kStatsThreads) smem[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma::s32(full + s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma::s32(ready + s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(umma::s32(empty + s)), "r"(kStatsWarps));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_granules = (n_rows + kGranuleRows - 1) / kGranuleRows;
  const int64_t total = (int64_t)blockIdx.x < n_granules
                            ? (n_granules - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto rows_in = [&](int64_t g, int64_t& row0) {
    row0 = ((int64_t)blockIdx.x + g * gridDim.x) * kGranuleRows;
    const int64_t left = n_rows - row0;
    return (int)(left > kGranuleRows ? kGranuleRows : left);
  };

  if (warp == kStatsWarps) {
    // ===================== producer =====================
    for (int64_t g = 0; g < total; ++g) {
      const int s = (int)(g % STAGES);
      int64_t row0;
      const int rows = rows_in(g, row0);
      umma::mbar_wait(umma::s32(empty + s), (uint32_t)(((g / STAGES) & 1) ^ 1));
      float* xt = xs + (size_t)s * kGranuleRows * D;
      long long* it = idx + s * kGranuleRows;
      if (rows == kGranuleRows) {
        if (lane == 0) {
          const uint32_t bar = umma::s32(full + s);
          const uint32_t bytes = kGranuleRows * D * 4u + kGranuleRows * 8u;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                           "r"(umma::s32(xt)), "l"(x + row0 * D), "r"(kGranuleRows * D * 4u), "r"(bar) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                           "r"(umma::s32(it)), "l"(index + row0), "r"(kGranuleRows * 8u), "r"(bar) : "memory");
        }
      } else {
        // ragged last granule: plain copies by the whole warp (sizes need not be 16-byte multiples)
        for (int e = lane; e < rows * C4; e += 32)
          reinterpret_cast<float4*>(xt)[e] = __ldg(reinterpret_cast<const float4*>(x + row0 * D) + e);
        for (int r = lane; r < kGranuleRows; r += 32) it[r] = r < rows ? index[row0 + r] : -1;
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(umma::s32(full + s));
      }
    }
  } else if (warp == kStatsWarps + 1) {
    // ===================== dispatcher =====================
    int hot = -1;
    for (int64_t g = 0; g < total; ++g) {
      const int s = (int)(g % STAGES);
      int64_t row0;
      const int rows = rows_in(g, row0);
      umma::mbar_wait(umma::s32(full + s), (uint32_t)((g / STAGES) & 1));
      const long long* it = idx + s * kGranuleRows;
      StageLists& L = lists[s];
      if (g == kHotElectionGranule) {
        // A code that takes >= 1/12 of the rows in the ring at this point is this CTA's hot
        // code for the rest of the launch (silence in a spectrogram, a collapsed codebook): its
        // rows stay out of the owner lists -- one warp would have to add them all -- and every
        // consumer warp sums the ones among its OWN rows in registers instead.  The sample is
        // this granule and the first fills of the later stages (stage 0 may already be being
        // refilled, so it is left out).
        const int n_sample = ((int)min((int64_t)STAGES, total) - 1) * kGranuleRows;
        for (int64_t ahead = g + 1; ahead < min((int64_t)STAGES, total); ++ahead)   // first fills
          umma::mbar_wait(umma::s32(full + (int)ahead), 0u);
        // candidates: the codes of this granule (two per lane); votes: every sampled row
        const int* low = reinterpret_cast<const int*>(idx + kGranuleRows);   // low words of the int64 codes, stages 1..
        int mine[kGranuleRows / 32], votes[kGranuleRows / 32];
#pragma unroll
        for (int i = 0; i < kGranuleRows / 32; ++i) {
          const long long v = it[i * 32 + lane];
          mine[i] = (i * 32 + lane < rows && v >= 0 && v < n_embed) ? (int)v : -2;
          votes[i] = 0;
        }
#pragma unroll 4
        for (int j = 0; j < n_sample; ++j) {
          const int cj = low[2 * j];
#pragma unroll
          for (int i = 0; i < kGranuleRows / 32; ++i) votes[i] += (cj == mine[i]);
        }
        int best = 0;
#pragma unroll
        for (int i = 0; i < kGranuleRows / 32; ++i)
          if (mine[i] >= 0) best = max(best, (votes[i] << 16) | mine[i]);
        best = __reduce_max_sync(0xffffffffu, best);
        hot = (best >> 16) * 12 >= n_sample ? (best & 0xffff) : -1;
        if (lane == 0) *hot_slot = hot;
        __syncwarp();
      }
      int base = 0;        // rows of the first half already listed for this lane's owner
#pragma unroll
      for (int half = 0; half < kGranuleRows / 32; ++half) {
        const int r = half * 32 + lane;
        const long long v = it[r];
        const bool valid = r < rows && v >= 0 && v < n_embed;
        if (r < rows && !valid && status_flag) atomicExch(status_flag, 1);
        const bool ok = valid && v != hot;
        const int owner = ok ? (int)(v % kStatsWarps) : -1 - lane;      // unlisted rows match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, owner);
        const int pos = __popc(peers & ((1u << lane) - 1));
        if (half == 1 && ok) base = L.count[owner];
        if (ok) L.row[owner][base + pos] = (unsigned char)r;
        __syncwarp();
        if (half == 0) {
          // counts of the first half: written by group leaders, zero for owners without rows
          if (lane < kStatsWarps) L.count[lane] = 0;
          __syncwarp();
          if (ok && pos == 0) L.count[owner] = __popc(peers);
        } else if (ok && pos == 0) {
          L.count[owner] = base + __popc(peers);
        }
        __syncwarp();
      }
      if (lane == 0) umma::mbar_arrive(umma::s32(ready + s));
    }
  } else {
    // ===================== consumers =====================
    float sq = 0.f;
    float hot_sum[VPL];
    int hot = -1, hot_len = 0;
#pragma unroll
    for (int v2 = 0; v2 < VPL; ++v2) hot_sum[v2] = 0.f;
    // The codebook rows of this warp's kRowsPerWarp rows of a granule come from L2 (~600
    // cycles); they are requested one granule ahead, so the latency hides behind the previous
    // granule's list walk.
    constexpr int kIters = kRowsPerWarp / kRowsPerInstr;
    float4 q_cur[kIters], q_next[kIters];
    auto request_codewords = [&](int64_t g, float4* q) {
      const int s = (int)(g % STAGES);
      const long long* it = idx + s * kGranuleRows;
      int64_t row0;
      const int rows_here = rows_in(g, row0);
      umma::mbar_wait(umma::s32(full + s), (uint32_t)((g / STAGES) & 1));
#pragma unroll
      for (int k = 0; k < kIters; ++k) {
        const int r = warp * kRowsPerWarp + k * kRowsPerInstr + lane / C4, j = lane % C4;
        const long long v = r < rows_here ? it[r] : -1;
        q[k] = (v >= 0 && v < n_embed) ? __ldg(reinterpret_cast<const float4*>(et + v * D) + j)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (total > 0) request_codewords(0, q_cur);
    for (int64_t g = 0; g < total; ++g) {
      const int s = (int)(g % STAGES);
      const float* xt = xs + (size_t)s * kGranuleRows * D;
      const long long* it = idx + s * kGranuleRows;
      int64_t row0;
      const int rows_here = rows_in(g, row0);
      // lookup, commitment term, output for this warp's rows of the granule (stage g is full:
      // request_codewords(g) waited for it)
#pragma unroll
      for (int k = 0; k < kIters; ++k) {
        const int r = warp * kRowsPerWarp + k * kRowsPerInstr + lane / C4, j = lane % C4;
        if (r < rows_here) {
          const float4 xv = reinterpret_cast<const float4*>(xt + r * D)[j];
          float4 q = q_cur[k];
          const float4 t = make_float4(q.x - xv.x, q.y - xv.y, q.z - xv.z, q.w - xv.w);
          sq = fmaf(t.x, t.x, sq); sq = fmaf(t.y, t.y, sq); sq = fmaf(t.z, t.z, sq); sq = fmaf(t.w, t.w, sq);
          q = make_float4(xv.x + t.x, xv.y + t.y, xv.z + t.z, xv.w + t.w);     // bottleneck.py:95
          if (out_q) reinterpret_cast<float4*>(out_q + (row0 + r) * q_row_stride)[j] = q;
        }
      }
      if (g + 1 < total) request_codewords(g + 1, q_next);
      // statistics: walk this warp's list of owned rows and add them to the accumulator rows
      // with plain ld/add/st (nobody else touches the codes this warp owns)
      umma::mbar_wait(umma::s32(ready + s), (uint32_t)((g / STAGES) & 1));
      const StageLists& L = lists[s];
      const int n_mine = L.count[warp];
      if (g == kHotElectionGranule) hot = *hot_slot;
      if (hot >= 0) {
        // the hot code's rows among this warp's own rows of the granule
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
          const int r = warp * kRowsPerWarp + k;
          if (r < rows_here && it[r] == hot) {
#pragma unroll
            for (int v2 = 0; v2 < VPL; ++v2) hot_sum[v2] += xt[r * D + lane * VPL + v2];
            ++hot_len;
          }
        }
      }
      // the list entries (row, code) are read lane-parallel once and broadcast by shuffles, two
      // rows per step, so the row loads of a step are independent of each other
      for (int base = 0; base < n_mine; base += 32) {
        const int left = min(32, n_mine - base);
        const int my_r = lane < left ? L.row[warp][base + lane] : 0;
        const int my_c = lane < left ? (int)it[my_r] : -1;
        for (int k = 0; k < left; k += 2) {
          const int ra = __shfl_sync(0xffffffffu, my_r, k), ca = __shfl_sync(0xffffffffu, my_c, k);
          const int kb = k + 1 < left ? k + 1 : k;
          const int rb = __shfl_sync(0xffffffffu, my_r, kb), cb = __shfl_sync(0xffffffffu, my_c, kb);
          float xa[VPL], xb[VPL];
#pragma unroll
          for (int v2 = 0; v2 < VPL; ++v2) {
            xa[v2] = xt[ra * D + lane * VPL + v2];
            xb[v2] = xt[rb * D + lane * VPL + v2];
          }
          float* pa = acc + (size_t)ca * D + lane * VPL;
          float* pb = acc + (size_t)cb * D + lane * VPL;
          if (k + 1 >= left || ca == cb) {
            const bool two = k + 1 < left;
#pragma unroll
            for (int v2 = 0; v2 < VPL; ++v2) pa[v2] += two ? xa[v2] + xb[v2] : xa[v2];
            if (lane == 0) cnt[ca] += two ? 2.f : 1.f;
          } else {
            // two different accumulator rows: independent read-modify-writes, loads first
            float va[VPL], vb[VPL];
#pragma unroll
            for (int v2 = 0; v2 < VPL; ++v2) { va[v2] = pa[v2]; vb[v2] = pb[v2]; }
#pragma unroll
            for (int v2 = 0; v2 < VPL; ++v2) { pa[v2] = va[v2] + xa[v2]; pb[v2] = vb[v2] + xb[v2]; }
            if (lane == 0) { const float na = cnt[ca], nb = cnt[cb]; cnt[ca] = na + 1.f; cnt[cb] = nb + 1.f; }
          }
        }
      }
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(umma::s32(empty + s));       // this warp is done with the stage
#pragma unroll
      for (int k = 0; k < kIters; ++k) q_cur[k] = q_next[k];
    }
    sq = warp_sum(sq);
    if (lane == 0 && sq != 0.f) atomicAdd(&partials[blockIdx.x], (double)sq);
    asm volatile("bar.sync 1, %0;" ::"n"(kStatsWarps * 32) : "memory");   // consumers only
    if (hot_len > 0) {        // 16 warps at most, once per launch: shared-memory atomics are fine
#pragma unroll
      for (int v2 = 0; v2 < VPL; ++v2) atomicAdd(acc + (size_t)hot * D + lane * VPL + v2, hot_sum[v2]);
      if (lane == 0) atomicAdd(cnt + hot, (float)hot_len);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kStatsWarps * 32) : "memory");
    // one vector reduction per used code of this CTA
    for (int k = warp; k < n_embed; k += kStatsWarps) {
      const float n = cnt[k];
      if (n > 0.f) {
        if (lane == 0) atomicAdd(&stats[k], n);
        if (lane < C4) atomicAdd(reinterpret_cast<float4*>(stats + n_embed + (int64_t)k * D) + lane,
                                 reinterpret_cast<const float4*>(acc + (size_t)k * D)[lane]);
      }
    }
  }
}

// diff = sum(partials) / (N*D) ; perplexity from the usage histogram
__global__ void __launch_bounds__(256)
vq_finish_kernel(const double* __restrict__ partials, int64_t n_partials, int64_t n_rows, int dim,
                 int n_embed, const float* __restrict__ stats, float* __restrict__ out_diff,
                 float* __restrict__ out_perplexity) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  double s = 0.0;
  if (out_diff)
    for (int64_t i = tid; i < n_partials; i += 256) s += partials[i];
  red[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  if (tid == 0 && out_diff) *out_diff = (float)(red[0] / ((double)n_rows * (double)dim));
  __syncthreads();
  if (out_perplexity && stats) {
    double h = 0.0;
    for (int k = tid; k < n_embed; k += 256) {
      float p = stats[k] / (float)n_rows;
      h += (double)(p * logf(fmaxf(p, 1e-7f)));
    }
    red[tid] = h;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
    if (tid == 0) *out_perplexity = expf((float)(-red[0]));
  }
}

// ---------------------------------------------------------------------------
// EMA update (bottleneck.py:80-92)
// ---------------------------------------------------------------------------
// phase 1 (one CTA): cluster_size in place; smoothed sizes -> smoothed[K]
__global__ void __launch_bounds__(1024)
vq_ema_cluster_kernel(const float* counts /* aliases smoothed */, float* cluster_size,
                      float* smoothed, int n_embed, float decay,
                      float one_minus_decay, float eps, float k_eps) {
  __shared__ float red[1024];
  const int tid = threadIdx.x;
  float s = 0.f;
  for (int k = tid; k < n_embed; k += 1024) {
    // mul_(decay) rounds, then add_(alpha, counts) is one fused multiply-add
    float cs = fmaf(counts[k], one_minus_decay, __fmul_rn(cluster_size[k], decay));
    cluster_size[k] = cs;
    s += cs;
  }
  red[tid] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  const float n = red[0];
  for (int k = tid; k < n_embed; k += 1024)
    smoothed[k] = (cluster_size[k] + eps) / (n + k_eps) * n;
}

// phase 2: embed_avg, embed [D,K] from the code-major embed_sum [K,D]
__global__ void __launch_bounds__(256)
vq_ema_embed_kernel(const float* __restrict__ embed_sum, const float* __restrict__ smoothed,
                    float* __restrict__ embed_avg, float* __restrict__ embed, int dim, int n_embed,
                    float decay, float one_minus_decay) {
  __shared__ float tile[32][33];
  const int tiles_k = (n_embed + 31) / 32;
  const int k0 = (blockIdx.x % tiles_k) * 32, d0 = (blockIdx.x / tiles_k) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    int k = k0 + j, d = d0 + tx;
    tile[j][tx] = (k < n_embed && d < dim) ? embed_sum[(int64_t)k * dim + d] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int d = d0 + j, k = k0 + tx;
    if (d < dim && k < n_embed) {
      int64_t o = (int64_t)d * n_embed + k;
      float ea = fmaf(tile[tx][j], one_minus_decay, __fmul_rn(embed_avg[o], decay));
      embed_avg[o] = ea;
      embed[o] = ea / smoothed[k];
    }
  }
}

// ---------------------------------------------------------------------------
// embed_code (bottleneck.py:103-104), optionally straight into NCHW
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_code_kernel(const int64_t* __restrict__ index, int64_t n_rows, int dim, int n_embed,
                  const float* __restrict__ et, float* __restrict__ out, isi_rows_layout ol,
                  int32_t* __restrict__ status_flag) {
  extern __shared__ __align__(16) float smem[];
  const int pitch = dim + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  const int rows_here = (int)min((int64_t)32, n_rows - row0);
  const bool rows_contiguous = (ol.row_stride == 1 && ol.col_stride != 1);
  for (int r = warp; r < rows_here; r += 8) {
    int64_t v = index[row0 + r];
    bool ok = (v >= 0 && v < n_embed);
    if (!ok && lane == 0 && status_flag) atomicExch(status_flag, 1);
    for (int d = lane; d < dim; d += 32) {
      float q = ok ? et[v * dim + d] : 0.f;
      if (rows_contiguous) smem[r * pitch + d] = q;
      else out[row_offset(ol, row0 + r) + (int64_t)d * ol.col_stride] = q;
    }
  }
  if (rows_contiguous) {
    __syncthreads();
    for (int e = tid; e < 32 * dim; e += 256) {
      int r = e & 31, d = e >> 5;
      if (r < rows_here)
        out[row_offset(ol, row0 + r) + (int64_t)d * ol.col_stride] = smem[r * pitch + d];
    }
  }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
int launch_prepare(const float* embed, int dim, int n_embed, const Prepared& p,
                   cudaStream_t stream) {
  int tiles = ((n_embed + 31) / 32) * ((dim + 31) / 32);
  int grid = tiles < 4 * kNumSms ? tiles : 4 * kNumSms;
  vq_prepare_kernel<<<grid, 256, 0, stream>>>(embed, dim, n_embed, p);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

size_t gather_workspace_bytes(int64_t n_rows) {
  int64_t grid = (n_rows + kGatherRows - 1) / kGatherRows;
  return (size_t)(grid > 0 ? grid : 1) * sizeof(double);
}

int launch_gather_stats(const float* x, const isi_rows_layout& xl, const int64_t* index,
                        int64_t n_rows, int dim, int n_embed, const Prepared& p, float* out_q,
                        const isi_rows_layout& ql, float* stats, int counts_only, void* workspace,
                        int32_t* status_flag, cudaStream_t stream) {
  int64_t grid = (n_rows + kGatherRows - 1) / kGatherRows;
  if (grid > 0x7fffffff) return ISI_ERR_SHAPE;
  double* partials = (double*)workspace;
  // row-major fast path: uniform row stride, 16-byte aligned float4 rows on both sides
  auto uniform = [&](const isi_rows_layout& l) {
    return l.col_stride == 1 && (l.row_stride & 3) == 0 &&
           (l.rows_per_batch >= n_rows || l.batch_stride == l.rows_per_batch * l.row_stride);
  };
  const bool fast = (dim == 16 || dim == 32 || dim == 64 || dim == 128) &&
                    (!x || (uniform(xl) && ((uintptr_t)x & 15) == 0)) &&
                    (!out_q || (uniform(ql) && ((uintptr_t)out_q & 15) == 0)) &&
                    (!stats || counts_only || ((n_embed & 3) == 0 && ((uintptr_t)stats & 15) == 0));
  // training statistics: CTA-private shared-memory accumulators when the codebook fits and the
  // rows are fully contiguous (one bulk copy per 64-row granule)
  const int stages = dim <= 64 ? 4 : 2;
  const size_t stats_smem = ((size_t)n_embed * dim + ((n_embed + 31) & ~31) +
                             (size_t)stages * kGranuleRows * dim) * 4 + stages * kGranuleRows * 8 +
                            stages * sizeof(StageLists) + 3 * stages * 8 + 64;
  if (fast && x && stats && !counts_only && (dim == 32 || dim == 64 || dim == 128) &&
      xl.row_stride == dim && ((uintptr_t)index & 15) == 0 && stats_smem <= 220 * 1024) {
    const int64_t granules = (n_rows + kGranuleRows - 1) / kGranuleRows;
    const int64_t want = (granules + 7) / 8;
    const unsigned ctas = (unsigned)(want < kNumSms ? (want < 1 ? 1 : want) : kNumSms);
    // per-CTA commitment sums are accumulated into the first `ctas` partial slots
    cudaError_t em = cudaMemsetAsync(partials, 0, (size_t)grid * sizeof(double), stream);
    if (em != cudaSuccess) return (int)em;
#define ISI_STATS_CASE(DD, ST)                                                                  \
    case DD: {                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(vq_gather_stats_smem_kernel<DD, ST>,                 \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                           (int)stats_smem);                                    \
      if (e != cudaSuccess) return (int)e;                                                      \
      vq_gather_stats_smem_kernel<DD, ST><<<ctas, kStatsThreads, stats_smem, stream>>>(         \
          x, index, n_rows, n_embed, p.et, out_q, ql.row_stride, stats, partials, status_flag); \
      break;                                                                                    \
    }
    switch (dim) { ISI_STATS_CASE(32, 4) ISI_STATS_CASE(64, 4) ISI_STATS_CASE(128, 2) }
#undef ISI_STATS_CASE
    ISI_LAUNCH_CHECK();
    return ISI_OK;
  }
  if (fast) {
#define ISI_GATHER_CASE(LPR)                                                                  \
    case LPR * 4:                                                                              \
      vq_gather_rowmajor_kernel<LPR><<<(unsigned)grid, 256, 0, stream>>>(                      \
          x, xl.row_stride, index, n_rows, n_embed, p.et, out_q, ql.row_stride, stats,         \
          counts_only, partials, status_flag);                                                 \
      break;
    switch (dim) {
      ISI_GATHER_CASE(4) ISI_GATHER_CASE(8) ISI_GATHER_CASE(16) ISI_GATHER_CASE(32)
    }
#undef ISI_GATHER_CASE
    ISI_LAUNCH_CHECK();
    return ISI_OK;
  }
  size_t smem = (size_t)kGatherRows * (dim + 1) * 4 + kGatherRows * 4;
  if (smem > 200 * 1024) return ISI_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(vq_gather_stats_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  vq_gather_stats_kernel<<<(unsigned)grid, kGatherThreads, smem, stream>>>(
      x, xl, index, n_rows, dim, n_embed, p.et, out_q, ql, stats, counts_only, partials,
      status_flag);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_finish(const void* workspace, int64_t n_rows, int dim, int n_embed, const float* stats,
                  float* out_diff, float* out_perplexity, cudaStream_t stream) {
  int64_t n_partials = (n_rows + kGatherRows - 1) / kGatherRows;
  const double* partials = (const double*)workspace;
  vq_finish_kernel<<<1, 256, 0, stream>>>(partials, n_partials, n_rows, dim, n_embed, stats,
                                          out_diff, out_perplexity);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_ema_update(float* stats, float* cluster_size, float* embed_avg, float* embed, int dim,
                      int n_embed, double decay_d, double eps_d, cudaStream_t stream) {
  // scalars are rounded to FP32 exactly where the reference's Python doubles meet tensors
  const float decay = (float)decay_d, one_minus_decay = (float)(1.0 - decay_d);
  const float eps = (float)eps_d, k_eps = (float)(n_embed * eps_d);
  // the smoothed cluster sizes overwrite the (already consumed) counts in stats[0:K]
  vq_ema_cluster_kernel<<<1, 1024, 0, stream>>>(stats, cluster_size, stats, n_embed, decay,
                                                  one_minus_decay, eps, k_eps);
  ISI_LAUNCH_CHECK();
  int tiles = ((n_embed + 31) / 32) * ((dim + 31) / 32);
  vq_ema_embed_kernel<<<tiles, 256, 0, stream>>>(stats + n_embed, stats, embed_avg, embed, dim,
                                                 n_embed, decay, one_minus_decay);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

int launch_embed_code(const int64_t* index, int64_t n_rows, int dim, int n_embed,
                      const Prepared& p, float* out, const isi_rows_layout& ol,
                      int32_t* status_flag, cudaStream_t stream) {
  int64_t grid = (n_rows + 31) / 32;
  if (grid > 0x7fffffff) return ISI_ERR_SHAPE;
  size_t smem = (size_t)32 * (dim + 1) * 4;
  if (smem > 48 * 1024) return ISI_ERR_UNSUPPORTED;
  embed_code_kernel<<<(unsigned)grid, 256, smem, stream>>>(index, n_rows, dim, n_embed, p.et, out,
                                                           ol, status_flag);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

}  // namespace isi
