// GPU probe: how many small cp.async.bulk shared->global stores an SM sustains.  The front end's
// polar/emit warps would hand each 128-byte output line (one row pair x 4 time blocks) to the
// TMA unit instead of writing it as eight 16-byte pieces from eight different instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulk_store_probe.bin tools/bulk_store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES, int OPS_PER_WARP>
__global__ void __launch_bounds__(512, 1) probe(float* out, int iters, int line_stride_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* tile = smem + warp * (OPS_PER_WARP * (BYTES + 32));
  for (int i = lane; i < OPS_PER_WARP * (BYTES + 32) / 4; i += 32) reinterpret_cast<float*>(tile)[i] = (float)i;
  __syncwarp();
  unsigned char* base = reinterpret_cast<unsigned char*>(out) +
                        ((size_t)blockIdx.x * 16 + warp) * OPS_PER_WARP * (size_t)line_stride_bytes;
  for (int it = 0; it < iters; ++it) {
    // the tile is rewritten every iteration (4 x 16-byte stores per lane, as the kernel would)
    if (BYTES == 128) {
      float4 v = make_float4(it, lane, warp, 1.f);
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4*>(tile + (lane >> 1) * (BYTES + 32) + (2 * k + (lane & 1)) * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < OPS_PER_WARP) {
      unsigned char* g = base + (size_t)lane * line_stride_bytes + (size_t)(it & 15) * BYTES;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
                   "r"(s32(tile + lane * (BYTES + 32))), "r"(BYTES)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane < OPS_PER_WARP) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the same bytes written the way the kernel does now: 4 x STG.128 per lane, 32-byte runs
__global__ void __launch_bounds__(512, 1) probe_stg(float* out, int iters, int line_stride_bytes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = reinterpret_cast<unsigned char*>(out) +
                        ((size_t)blockIdx.x * 16 + warp) * 16 * (size_t)line_stride_bytes;
  for (int it = 0; it < iters; ++it) {
    float4 v = make_float4(it, lane, warp, 1.f);
    unsigned char* g = base + (size_t)(lane >> 1) * line_stride_bytes + (size_t)(it & 15) * 128 + (lane & 1) * 16;
    for (int k = 0; k < 4; ++k)
      asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + 32 * k), "f"(v.x),
                   "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
  }
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  const int iters = 2000, stride = 2048;
  float* out;
  const size_t bytes = (size_t)148 * 16 * 16 * stride + (1 << 20);
  cudaMalloc(&out, bytes);
  cudaMemset(out, 0, bytes);
  const int smem128 = 16 * 16 * (128 + 32);
  cudaFuncSetAttribute(probe<128, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128);
  float ms = time_ms([&] { probe<128, 16><<<148, 512, smem128>>>(out, iters, stride); });
  double ops = 148.0 * 16 * 16 * iters;
  printf("{\"bulk128\": {\"ms\": %.4f, \"ops_per_sm_per_us\": %.2f, \"GBps\": %.1f, \"cycles_per_op_per_sm_at_1.9GHz\": %.2f},\n", ms,
         ops / 148 / (ms * 1e3), ops * 128 / (ms * 1e6), (ms * 1e-3 * 1.9e9) / (ops / 148));
  ms = time_ms([&] { probe_stg<<<148, 512>>>(out, iters, stride); });
  printf(" \"stg_4x16B_per_lane\": {\"ms\": %.4f, \"GBps\": %.1f},\n", ms, ops * 128 / (ms * 1e6));
  const int smem512 = 16 * 4 * (512 + 32);
  cudaFuncSetAttribute(probe<512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem512);
  ms = time_ms([&] { probe<512, 4><<<148, 512, smem512>>>(out, iters, stride); });
  ops = 148.0 * 16 * 4 * iters;
  printf(" \"bulk512\": {\"ms\": %.4f, \"ops_per_sm_per_us\": %.2f, \"GBps\": %.1f}}\n", ms, ops / 148 / (ms * 1e3),
         ops * 512 / (ms * 1e6));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
