// Nearest-code search on the CUDA cores in plain FP32 (any dim / n_embed / layout).
//
// Replaces bottleneck.py:55-61 of the reference.  This is the shape-generic kernel;
// the tcgen05 kernel in vq_assign_tc.cu takes the shapes the deployed models use.
//
// One CTA owns a tile of 64 rows and walks the codebook in chunks of 128 codes and
// the feature dimension in chunks of 32, keeping a 4x8 register tile of dot
// products per thread.  score(n,k) = |e_k|^2 - 2 x_n.e_k (the |x_n|^2 term of the
// reference is constant per row and cannot change the argmin); ties resolve to the
// lowest code index like torch's max (bottleneck.py:61).
#include "common.cuh"

namespace isi {

constexpr int kTileRows = 64;
constexpr int kChunkCodes = 128;
constexpr int kChunkDim = 32;
constexpr int kRowPitch = kTileRows + 4;  // keeps float4 alignment, spreads banks

__device__ __forceinline__ void keep_better(float& s, int& i, float s2, int i2) {
  if (s2 < s || (s2 == s && i2 < i)) { s = s2; i = i2; }
}

__global__ void __launch_bounds__(256)
vq_assign_simt_kernel(const float* __restrict__ x, isi_rows_layout lay, int64_t n_rows,
                      int dim, int n_embed, const float* __restrict__ embed,
                      const float* __restrict__ e2, int64_t* __restrict__ out_index,
                      float* __restrict__ out_score) {
  __shared__ __align__(16) float xs[kChunkDim][kRowPitch];     // [d][row]
  __shared__ __align__(16) float es[kChunkDim][kChunkCodes];   // [d][code]

  const int tid = threadIdx.x;
  const int rg = tid >> 4;   // rows 4*rg .. 4*rg+3
  const int cg = tid & 15;   // codes 4*cg..+3 and 64+4*cg..+3 of the chunk
  const int64_t row0 = (int64_t)blockIdx.x * kTileRows;
  const bool rows_contiguous = (lay.row_stride == 1 && lay.col_stride != 1);

  float best_s[4];
  int best_i[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best_s[i] = INFINITY; best_i[i] = 0x7fffffff; }

  for (int k0 = 0; k0 < n_embed; k0 += kChunkCodes) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int d0 = 0; d0 < dim; d0 += kChunkDim) {
      // ---- stage the x tile (transposed) ----
      for (int e = tid; e < kTileRows * kChunkDim; e += 256) {
        int r, d;
        if (rows_contiguous) { r = e % kTileRows; d = e / kTileRows; }
        else                 { d = e % kChunkDim; r = e / kChunkDim; }
        float v = 0.f;
        if (row0 + r < n_rows && d0 + d < dim)
          v = x[row_offset(lay, row0 + r) + (int64_t)(d0 + d) * lay.col_stride];
        xs[d][r] = v;
      }
      // ---- stage the codebook chunk: es[d][c] = E[d0+d][k0+c] ----
      for (int e = tid; e < kChunkDim * kChunkCodes; e += 256) {
        int c = e % kChunkCodes, d = e / kChunkCodes;
        float v = 0.f;
        if (k0 + c < n_embed && d0 + d < dim) v = embed[(int64_t)(d0 + d) * n_embed + k0 + c];
        es[d][c] = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int d = 0; d < kChunkDim; ++d) {
        float4 xv = *reinterpret_cast<const float4*>(&xs[d][4 * rg]);
        float4 ea = *reinterpret_cast<const float4*>(&es[d][4 * cg]);
        float4 eb = *reinterpret_cast<const float4*>(&es[d][64 + 4 * cg]);
        const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
        const float er[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xr[i], er[j], acc[i][j]);
      }
      __syncthreads();
    }
    // ---- fold this chunk into the running best (codes visited in rising order) ----
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int code = k0 + (j < 4 ? 4 * cg + j : 64 + 4 * cg + (j - 4));
      if (code < n_embed) {
        float ee = e2[code];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float s = fmaf(-2.f, acc[i][j], ee);
          keep_better(best_s[i], best_i[i], s, code);
        }
      }
    }
  }
  // ---- merge the 16 lanes that share a row group ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      float s2 = __shfl_xor_sync(0xffffffffu, best_s[i], o);
      int i2 = __shfl_xor_sync(0xffffffffu, best_i[i], o);
      keep_better(best_s[i], best_i[i], s2, i2);
    }
    int64_t row = row0 + 4 * rg + i;
    if (cg == 0 && row < n_rows) {
      // a row of NaN / Inf never wins a comparison: code 0, like torch's (-dist).max(1) on an
      // all-NaN row (bottleneck.py:61), never an out-of-range index
      out_index[row] = best_i[i] == 0x7fffffff ? 0 : best_i[i];
      if (out_score) out_score[row] = best_s[i];
    }
  }
}

int launch_assign_simt(const float* x, const isi_rows_layout& lay, int64_t n_rows, int dim,
                       int n_embed, const Prepared& p, int64_t* out_index, float* out_score,
                       cudaStream_t stream) {
  int64_t grid = (n_rows + kTileRows - 1) / kTileRows;
  if (grid > 0x7fffffff) return ISI_ERR_SHAPE;
  vq_assign_simt_kernel<<<(unsigned)grid, 256, 0, stream>>>(x, lay, n_rows, dim, n_embed, p.ed,
                                                            p.e2, out_index, out_score);
  ISI_LAUNCH_CHECK();
  return ISI_OK;
}

}  // namespace isi
