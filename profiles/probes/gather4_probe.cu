// TMA tile::gather4 probe (sm_100a).  Stand-alone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_probe gather4_probe.cu
//
//   gather4_probe layout <box_rows> <swizzle: 0 none | 64>     one gather4 of rows {5, 17, 1023, OOB} x 32 fp16 columns starting at column 32
//                                                              of X[1024][96]; dumps where every element landed in shared memory
//   gather4_probe rate <box_rows> <swizzle> <rows_per_stage>   every CTA (2 per SM) issues gather4's of random rows, 64-byte segments,
//                                                              into a 3-slot ring with an mbarrier round trip per stage: sustained GB/s
//
// Purpose (profiles/r2_results.md, "what is left"): the forward kernel's producers generate one address per 16 bytes in software; the
// TMA gather path would take that off the SM's issue slots.  What has to be known before rewriting the kernel around it: the landing
// layout of the four rows (to write the UMMA descriptor), the zero fill of absent rows, and whether 64-byte rows sustain the ~7 TB/s
// of L2 -> shared gathers the LDG/STS producers reach.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
               ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

constexpr int ROWS = 1024, C = 96, BOXC = 32;

__global__ void layout_kernel(const __grid_constant__ CUtensorMap tm, __half* out, int n_out, int oob_row) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) reinterpret_cast<__half*>(smem)[i] = __float2half(-1.f);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_expect(smem_u32(&bar), 4 * BOXC * 2);
    gather4(smem_u32(smem), &tm, smem_u32(&bar), 32, 5, 17, 1023, oob_row);
  }
  mbar_wait(smem_u32(&bar), 0);
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}

// Every CTA: `stages` stages of `rps` rows (rps / 4 gather4's of 64 bytes per row) into a 3-slot ring; one thread issues, waits for the
// slot's barrier before reusing it (the consumer is absent: this is the producer-side ceiling).
// NOTE on the round-2 run (profiles/reports/r2c27_gather4_probe.txt, 0.54-0.62 TB/s): that version read the four row indices of every
// gather4 from GLOBAL memory inside the issue loop, so the figure is bounded by that load's latency, not by the TMA unit.  The indices
// are now staged in shared memory first (as a convolution kernel would have them); re-run before drawing a conclusion on the rate.
constexpr int IDX_SMEM = 4096;
__global__ void rate_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ rows_g, int n_rows_list, int rps, int stages, int* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[3];
  __shared__ int rows[IDX_SMEM];
  const int slot_bytes = rps * BOXC * 2;
  for (int i = threadIdx.x; i < IDX_SMEM; i += blockDim.x) rows[i] = rows_g[((blockIdx.x * 977) % (n_rows_list - IDX_SMEM)) + i];
  n_rows_list = IDX_SMEM;
  if (threadIdx.x == 0) { for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    int base = 0;
    for (int s = 0; s < stages; ++s) {
      const int slot = s % 3;
      if (s >= 3) mbar_wait(smem_u32(&bar[slot]), ((s / 3) - 1) & 1);
      mbar_expect(smem_u32(&bar[slot]), slot_bytes);
      const uint32_t dst = smem_u32(smem) + slot * slot_bytes;
      const int col = (s % 3) * 32;
      for (int g = 0; g < rps / 4; ++g) {
        const int* r = rows + base + 4 * g;
        gather4(dst + g * 4 * BOXC * 2, &tm, smem_u32(&bar[slot]), col, r[0], r[1], r[2], r[3]);
      }
      base = (base + rps) % (n_rows_list - rps);
    }
    for (int s = stages < 3 ? 0 : stages - 3; s < stages; ++s) mbar_wait(smem_u32(&bar[s % 3]), (s / 3) & 1);
    if (sink) sink[blockIdx.x] = reinterpret_cast<int*>(smem)[0];
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* tm, void* base, int rows, int box_rows, int swz) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)BOXC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(box {%d,%d}, swizzle %d) -> CUresult %d\n", BOXC, box_rows, swz, (int)r); return 1; }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) { printf("usage: %s layout|rate <box_rows> <swizzle 0|64> [rows_per_stage]\n", argv[0]); return 2; }
  const bool rate = !strcmp(argv[1], "rate");
  const int box_rows = atoi(argv[2]), swz = atoi(argv[3]);
  printf("== %s box {%d,%d} swizzle %d\n", argv[1], BOXC, box_rows, swz);
  if (!rate) {
    std::vector<__half> h((size_t)ROWS * C);
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = __float2half((float)((r % 64) * 8 + c / 16) + (c % 16) / 16.0f);
    __half* X; CK(cudaMalloc(&X, h.size() * 2)); CK(cudaMemcpy(X, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tm; if (make_map(&tm, X, ROWS, box_rows, swz)) return 1;
    const int n_out = 4 * BOXC * 2;                 // twice the expected footprint, to see where things land
    __half* out; CK(cudaMalloc(&out, n_out * 2));
    layout_kernel<<<1, 128, 4096>>>(tm, out, n_out, 2000);
    CK(cudaDeviceSynchronize());
    std::vector<__half> o(n_out); CK(cudaMemcpy(o.data(), out, n_out * 2, cudaMemcpyDeviceToHost));
    // decode: value = (row % 64) * 8 + col / 16 + (col % 16) / 16  ->  for the requested rows 5, 17, 1023 (-> 63), OOB (-> expect 0)
    for (int chunk = 0; chunk < n_out / 8; ++chunk) {       // 16-byte chunks
      float v = __half2float(o[chunk * 8]);
      if (v < 0) { printf("chunk %2d: untouched\n", chunk); continue; }
      int rowm = (int)(v / 8), colb = (int)(v - rowm * 8);
      float frac = v - (int)v;
      printf("chunk %2d (byte %3d): row%%64 = %2d, col = %2d   [first = %.4f, last = %.4f]\n", chunk, chunk * 16, rowm, colb * 16 + (int)(frac * 16 + 0.5f), v,
             __half2float(o[chunk * 8 + 7]));
    }
    return 0;
  }
  const int rps = argc > 4 ? atoi(argv[4]) : 128;
  const int big_rows = 300000;
  __half* X; CK(cudaMalloc(&X, (size_t)big_rows * C * 2)); CK(cudaMemset(X, 0, (size_t)big_rows * C * 2));
  CUtensorMap tm; if (make_map(&tm, X, big_rows, box_rows, swz)) return 1;
  // neighbour-like row lists: runs of nearby rows with jumps (a kernel map's column), a few absent (-> out of bounds)
  const int n_list = 1 << 20;
  std::vector<int> rl(n_list);
  uint32_t st = 12345u; int cur = 1000;
  for (int i = 0; i < n_list; ++i) {
    st = st * 1664525u + 1013904223u;
    if ((st >> 28) == 0) cur = (int)((st >> 4) % (uint32_t)(big_rows - 64));
    cur += 1 + (int)((st >> 20) & 3);
    if (cur >= big_rows) cur = 17;
    rl[i] = ((st >> 12) & 15) == 0 ? big_rows + 5 : cur;
  }
  int* rows_d; CK(cudaMalloc(&rows_d, n_list * 4)); CK(cudaMemcpy(rows_d, rl.data(), n_list * 4, cudaMemcpyHostToDevice));
  int dev = 0, sms = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ctas = 2 * sms, stages = 2000;
  const size_t smem = (size_t)3 * rps * BOXC * 2 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int* sink; CK(cudaMalloc(&sink, ctas * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  rate_kernel<<<ctas, 32, smem>>>(tm, rows_d, n_list, rps, 50, sink);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  rate_kernel<<<ctas, 32, smem>>>(tm, rows_d, n_list, rps, stages, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double bytes = (double)ctas * stages * rps * BOXC * 2;
  printf("rate: %d CTAs x %d stages x %d rows x 64 B = %.2f GB in %.3f ms = %.0f GB/s (%.1f ns per gather4 per CTA)\n", ctas, stages, rps, bytes / 1e9, ms,
         bytes / ms / 1e6, ms * 1e6 / ((double)stages * rps / 4));
  return 0;
}
