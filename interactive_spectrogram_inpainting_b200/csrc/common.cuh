// Shared device/host helpers for the isi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "isi_b200.h"

#define ISI_LAUNCH_CHECK()                         \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

namespace isi {

constexpr int kNumSms = 148;  // B200

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// codes are padded to a multiple of 256 for the tensor-core operand tiles
__host__ __device__ inline int padded_codes(int n_embed) { return (n_embed + 255) / 256 * 256; }

// Layout of the "prepared codebook" scratch (isi_vq_prepare_codebook).
struct Prepared {
  float* e2;      // [K]        ||e_k||^2
  float* et;      // [K, D]     code-major copy of the codebook
  float* ed;      // [D, K]     snapshot of the codebook as the caller holds it
  float* b_hi;    // tensor-core operand: TF32 "hi" part of -2E, canonical UMMA tiles
  float* b_lo;    // tensor-core operand: TF32 "lo" part of -2E
  float* b_pair;  // 2-CTA kernel: per CTA rank the resident image of its half of every 256-code tile
  size_t bytes;
};

__host__ __device__ inline Prepared prepared_view(const void* base, int dim, int n_embed) {
  Prepared p;
  char* c = (char*)base;
  size_t off = 0;
  p.e2 = (float*)(c + off);
  off += align_up((size_t)padded_codes(n_embed) * 4, 1024);
  p.et = (float*)(c + off);
  off += align_up((size_t)n_embed * dim * 4, 1024);
  p.ed = (float*)(c + off);
  off += align_up((size_t)n_embed * dim * 4, 1024);
  p.b_hi = (float*)(c + off);
  off += align_up((size_t)padded_codes(n_embed) * dim * 4, 1024);
  p.b_lo = (float*)(c + off);
  off += align_up((size_t)padded_codes(n_embed) * dim * 4, 1024);
  p.b_pair = (float*)(c + off);
  off += align_up((size_t)padded_codes(n_embed) * dim * 8, 1024);
  p.bytes = off;
  return p;
}

__device__ __forceinline__ int64_t row_offset(const isi_rows_layout& L, int64_t row) {
  int64_t b = row / L.rows_per_batch;
  int64_t r = row - b * L.rows_per_batch;
  return b * L.batch_stride + r * L.row_stride;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace isi
