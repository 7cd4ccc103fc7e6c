// Micro-probe (sm_100a): latencies of the hand-off primitives used by the tcgen05 conv kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe latency_probe.cu && ./latency_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void probe(long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]), b2 = smem_u32(&bars[2]);
  if (tid == 0) { mbar_init(b0, 1); mbar_init(b1, 1); mbar_init(b2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  const uint32_t sb = smem_u32(smem);
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    uint32_t par = 0;
    // (a) commit with nothing outstanding
    for (int rep = 0; rep < 4; ++rep) {
      long long t0 = clock64();
      tc_commit(b0);
      mbar_wait(b0, par); par ^= 1;
      out[rep] = clock64() - t0;
    }
    // (b) 6 MMAs (M128 N96 K16) + commit
    for (int rep = 0; rep < 4; ++rep) {
      long long t0 = clock64();
      for (int j = 0; j < 6; ++j) tc_mma(tm, make_desc(sb, 2112, 128), make_desc(sb + 32768, 1552, 128), IDESC, j > 0);
      long long t1 = clock64();
      tc_commit(b0);
      mbar_wait(b0, par); par ^= 1;
      out[4 + rep] = clock64() - t0;
      out[8 + rep] = t1 - t0;          // issue time of the 6 MMAs
    }
    // (c) plain arrive + wait by the same thread
    uint32_t p1 = 0;
    for (int rep = 0; rep < 4; ++rep) {
      long long t0 = clock64();
      mbar_arrive(b1);
      mbar_wait(b1, p1); p1 ^= 1;
      out[12 + rep] = clock64() - t0;
    }
    // (e) fence.proxy.async after one 16-byte shared store
    for (int rep = 0; rep < 4; ++rep) {
      long long t0 = clock64();
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sb + 1024), "r"(rep) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      out[20 + rep] = clock64() - t0;
    }
  }
  __syncthreads();
  // (d) cross-warp ping-pong: warp 0 lane 0 <-> warp 1 lane 0, 64 round trips
  if (lane == 0 && warp < 2) {
    uint32_t pa = 0, pb = 0;
    long long t0 = clock64();
    for (int r = 0; r < 64; ++r) {
      if (warp == 0) { mbar_arrive(b1 /*reuse*/ + 0 * 8 + 8 /* = b2 */); mbar_wait(b1, pa); pa ^= 1; }
      else { mbar_wait(b2, pb); pb ^= 1; mbar_arrive(b1); }
    }
    if (warp == 0) out[16] = (clock64() - t0) / 64;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128));
}

int main() {
  long long* d; cudaMalloc(&d, 32 * sizeof(long long)); cudaMemset(d, 0, 32 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<<<1, 64, 65536>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  printf("(a) tcgen05.commit -> mbarrier, idle pipe      : %lld %lld %lld %lld cycles\n", h[0], h[1], h[2], h[3]);
  printf("(b) 6 x tcgen05.mma(M128,N96,K16) + commit     : %lld %lld %lld %lld cycles\n", h[4], h[5], h[6], h[7]);
  printf("    issue time of the 6 MMAs (one thread)      : %lld %lld %lld %lld cycles\n", h[8], h[9], h[10], h[11]);
  printf("(c) mbarrier arrive + try_wait, same thread    : %lld %lld %lld %lld cycles\n", h[12], h[13], h[14], h[15]);
  printf("(d) cross-warp mbarrier round trip             : %lld cycles\n", h[16]);
  printf("(e) st.shared.v4 + fence.proxy.async           : %lld %lld %lld %lld cycles\n", h[20], h[21], h[22], h[23]);
  return 0;
}
