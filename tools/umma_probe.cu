// Microbenchmark (GPU box only): cycles per tcgen05.mma instruction for kind::tf32 and
// kind::f16 (bf16) at M=128, N in {64,128,256}, operands in shared memory (SWIZZLE_128B,
// K-major), accumulators in TMEM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o umma_probe tools/umma_probe.cu ; run: ./umma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
               ::"r"(bar), "r"(parity) : "memory");
}

template <int KIND>  // 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void probe(int n_dim, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;   // tf32 : bf16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n_dim >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad = umma_desc(s32(smem)), bd = umma_desc(s32(smem) + 128 * 128);
    for (int pass = 0; pass < 2; ++pass) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) mma<KIND>(tmem + (r & 1) * 256, ad + 2 * (r & 3), bd + 2 * (r & 3), idesc, 1);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      mbar_wait(s32(&bar), pass & 1);
      long long t1 = clock64();
      out[pass] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int smem = (128 + 256) * 128 + 1024;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2048;
  for (int kind = 0; kind < 2; ++kind)
    for (int n : {64, 128, 256}) {
      if (kind == 0) probe<0><<<1, 128, smem>>>(n, reps, d); else probe<1><<<1, 128, smem>>>(n, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      double cyc = (double)h[1] / reps;
      double k_elems = kind == 0 ? 8 : 16;
      printf("%s M=128 N=%3d: %.1f cycles/MMA  -> %.0f MAC/clk/SM (%s)\n", kind == 0 ? "tf32" : "bf16", n, cyc,
             128.0 * n * k_elems / cyc, cudaGetErrorString(e));
    }
  return 0;
}
