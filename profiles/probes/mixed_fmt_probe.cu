// Probe: does tcgen05.mma.kind::f16 accept MIXED operand formats (A = fp16, B = bf16) in one instruction?
// (The instruction descriptor has separate a_format / b_format fields; the guides do not say whether they may differ.)
// One CTA, one M128 x N32 x K16 MMA per format pair, K-major no-swizzle core-matrix layout as in conv_tc5.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mixed_fmt_probe mixed_fmt_probe.cu && ./mixed_fmt_probe
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}

// A: [128 rows][16 k] 16-bit, B: [32 rows (n)][16 k] 16-bit, both K-major; D[m][n] = sum_k A[m][k] B[n][k]
__global__ void probe(const uint16_t* A, const uint16_t* B, float* D, int a_fmt, int b_fmt) {
  __shared__ __align__(128) unsigned char sA[128 * 32];     // 2 k8-groups x 16 row-groups x 128 B
  __shared__ __align__(128) unsigned char sB[32 * 32];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int A_LBO = 16 * 128, B_LBO = 4 * 128, SBO = 128;
  {   // thread = row of A
    const int r = tid;
    for (int k8 = 0; k8 < 2; ++k8)
      *reinterpret_cast<uint4*>(sA + k8 * A_LBO + (r >> 3) * SBO + (r & 7) * 16) = *reinterpret_cast<const uint4*>(A + r * 16 + k8 * 8);
    if (r < 32)
      for (int k8 = 0; k8 < 2; ++k8)
        *reinterpret_cast<uint4*>(sB + k8 * B_LBO + (r >> 3) * SBO + (r & 7) * 16) = *reinterpret_cast<const uint4*>(B + r * 16 + k8 * 8);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;\n" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(sA), A_LBO, SBO), db = make_desc(smem_u32(sB), B_LBO, SBO);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!ok);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t r[16];
  for (int c0 = 0; c0 < 32; c0 += 16) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int e = 0; e < 16; ++e) D[(warp * 32 + lane) * 32 + c0 + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;\n" ::"r"(tmem));
}

static uint16_t enc(float v, int fmt) {      // 0 = fp16, 1 = bf16
  if (fmt == 0) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&b);
}
static float dec(uint16_t u, int fmt) {
  if (fmt == 0) { __half h = *reinterpret_cast<__half*>(&u); return __half2float(h); }
  __nv_bfloat16 b = *reinterpret_cast<__nv_bfloat16*>(&u); return __bfloat162float(b);
}

int main() {
  uint16_t hA[128 * 16], hB[32 * 16];
  float hD[128 * 32];
  uint16_t *dA, *dB; float* dD;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
  const int pairs[4][2] = {{1, 1}, {0, 0}, {0, 1}, {1, 0}};
  const char* nm[2] = {"fp16", "bf16"};
  for (auto& pr : pairs) {
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f; };
    // values with more than 8 mantissa bits: a wrong format interpretation cannot pass by accident
    for (int i = 0; i < 128 * 16; ++i) hA[i] = enc(rnd() * 3.1f, pr[0]);
    for (int i = 0; i < 32 * 16; ++i) hB[i] = enc(rnd() * 0.7f, pr[1]);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, sizeof hD);
    probe<<<1, 128>>>(dA, dB, dD, pr[0], pr[1]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("A=%s B=%s: CUDA error %s\n", nm[pr[0]], nm[pr[1]], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 16; ++k) ref += (double)dec(hA[m * 16 + k], pr[0]) * (double)dec(hB[n * 16 + k], pr[1]);
        worst = fmax(worst, fabs(ref - hD[m * 32 + n])); scale = fmax(scale, fabs(ref));
      }
    printf("A=%s B=%s: max |D - ref| = %.3e (max |ref| %.3f) -> %s\n", nm[pr[0]], nm[pr[1]], worst, scale, worst < 1e-5 * scale ? "OK" : "MISMATCH");
  }
  return 0;
}
