// Sparse convolution on a dense neighbour table: output-stationary gather-GEMM (no scatter, no atomics).
//
//   forward / data-gradient :  Y[j,:] = sum_k X[tbl[kmap[k]][j], :] . W[k]
//   weight-gradient         :  dW[k]  = sum_j A[tbl[k][j], :]^T . B[j, :]
//
// Replaces MinkowskiEngine 0.4.3 ConvolutionForwardGPU / ConvolutionBackwardGPU (+Transpose): per-offset
// gather -> SIMT matmul -> atomicAdd scatter, K launches per layer (reference call sites in include/pcb200.h).
//
// Numerics: the contraction runs on the tensor cores as a 3-term bf16 split (x = hi + lo):
//   x.w ~= hi.hi + lo.hi + hi.lo,  fp32 accumulate  ->  per-product relative error <= ~2^-16, i.e. fp32-class
// parity (tests: 1e-3 relative against the fp64 oracle after 63 layers).  An exact fp32 SIMT kernel with the
// same interface covers channel counts the tensor-core tiling does not (Cin = 3) and is the in-library
// cross-check (PCB_CONV_FORCE_SIMT).
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"

using namespace pcb;

namespace {

// --------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// fp32 x4 -> bf16 hi x4 (8 bytes) + bf16 lo x4 (8 bytes)
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 h1 = __floats2bfloat162_rn(v.z, v.w);
  float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
  __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
  hi.x = *reinterpret_cast<uint32_t*>(&h0); hi.y = *reinterpret_cast<uint32_t*>(&h1);
  lo.x = *reinterpret_cast<uint32_t*>(&l0); lo.y = *reinterpret_cast<uint32_t*>(&l1);
}

struct KMap { int v[PCB_MAX_KERNEL_VOLUME]; };

// --------------------------------------------------------------------------------------------- forward (tensor cores)
constexpr int BM = 128;        // output rows per CTA
constexpr int BK = 32;         // input channels per pipeline step (one 128-byte line of every gathered row)
constexpr int NTHR = 256;      // 8 warps: 4 (rows) x 2 (cols)
constexpr int A_STRIDE = BK * 2 + 16;   // bytes per smem row, +16 keeps ldmatrix conflict-free

struct ConvArgs {
  const float* X; int ldx;
  const int32_t* tbl; int64_t tbl_stride;
  KMap kmap; int K;
  int64_t n_out; int Cin; int Cout;
  const __nv_bfloat16* w_hi; const __nv_bfloat16* w_lo;
  const float* bias;
  float* Y; int ldy;
  float* partial;     // non-NULL: gridDim.z CTAs split the (offset, channel-chunk) loop; tile sums go to partial[z][row][Cout]
};

template <int BN>
struct ConvSmem {
  static constexpr int B_STRIDE = BN * 2 + 16;
  static constexpr int A_PLANE = BM * A_STRIDE;
  static constexpr int B_PLANE = BK * B_STRIDE;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int IDX_OFF = 2 * STAGE;
  static constexpr int META_OFF = IDX_OFF + PCB_MAX_KERNEL_VOLUME * BM * 4;
  static constexpr int TOTAL = META_OFF + 3 * 32 * 4;
};

template <int BN>
__global__ void __launch_bounds__(NTHR, 2) conv_mma_kernel(const ConvArgs p) {
  using S = ConvSmem<BN>;
  constexpr int WN = BN / 2;       // columns per warp
  constexpr int NT = WN / 8;       // n8 tiles per warp (even)
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  int* s_idx = reinterpret_cast<int*>(smem + S::IDX_OFF);
  int* s_flag = reinterpret_cast<int*>(smem + S::META_OFF);
  int* s_klist = s_flag + 32;
  int* s_nk = s_klist + 32;

  // ---- neighbour rows of this tile for every kernel offset
  for (int e = tid; e < p.K * BM; e += NTHR) {
    int k = e / BM, r = e - k * BM;
    int64_t row = row0 + r;
    int v = -1;
    if (row < p.n_out) v = p.tbl[(int64_t)p.kmap.v[k] * p.tbl_stride + row];
    s_idx[e] = v;
  }
  __syncthreads();
  for (int k = warp; k < p.K; k += NTHR / 32) {      // which 32-row slabs have any neighbour for offset k
    unsigned m = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      unsigned b = __ballot_sync(0xffffffffu, s_idx[k * BM + s * 32 + lane] >= 0);
      if (b) m |= 1u << s;
    }
    if (lane == 0) s_flag[k] = (int)m;
  }
  __syncthreads();
  if (tid == 0) {
    int nk = 0;
    for (int k = 0; k < p.K; ++k) if (s_flag[k]) s_klist[nk++] = k;
    *s_nk = nk;
  }
  __syncthreads();
  const int nk = *s_nk;
  const int nkc = p.Cin / BK;
  const int T = nk * nkc;

  float acc[2][NT][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

  const uint32_t smem_base = smem_u32(smem);
  const int a_chunk = tid & 7;      // which 16-byte piece of the 128-byte row segment
  const int a_row = tid >> 3;       // + 32*i

  auto load_B = [&](int stage, int it) {
    const int k = s_klist[it / nkc], kc = it % nkc;
    constexpr int CH = BN / 8;                       // 16-byte chunks per weight row
    constexpr int PER_PLANE = BK * CH;
    for (int c = tid; c < 2 * PER_PLANE; c += NTHR) {
      int plane = c / PER_PLANE, rem = c - plane * PER_PLANE;
      int r = rem / CH, ch = rem - r * CH;
      const __nv_bfloat16* src = (plane ? p.w_lo : p.w_hi) + ((int64_t)k * p.Cin + kc * BK + r) * p.Cout + n0 + ch * 8;
      uint32_t dst = smem_base + stage * S::STAGE + 2 * S::A_PLANE + plane * S::B_PLANE + r * S::B_STRIDE + ch * 16;
      cp_async16(dst, src);
    }
    cp_async_commit();
  };
  auto load_A = [&](int it, float4 (&v)[4]) {
    const int k = s_klist[it / nkc], kc = it % nkc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = s_idx[k * BM + a_row + 32 * i];
      if (idx >= 0) v[i] = __ldg(reinterpret_cast<const float4*>(p.X + (int64_t)idx * p.ldx + kc * BK) + a_chunk);
      else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_A = [&](int stage, const float4 (&v)[4]) {
    unsigned char* base = smem + stage * S::STAGE;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint2 hi, lo;
      split4(v[i], hi, lo);
      int off = (a_row + 32 * i) * A_STRIDE + a_chunk * 8;
      *reinterpret_cast<uint2*>(base + off) = hi;
      *reinterpret_cast<uint2*>(base + S::A_PLANE + off) = lo;
    }
  };
  auto compute = [&](int stage, int it) {
    const int k = s_klist[it / nkc];
    if (!((s_flag[k] >> wm) & 1)) return;            // this warp's 32 rows have no neighbour at offset k
    const uint32_t a_base = smem_base + stage * S::STAGE;
    const uint32_t b_base = a_base + 2 * S::A_PLANE;
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        uint32_t addr = a_base + (wm * 32 + mi * 16 + (lane & 15)) * A_STRIDE + (ks * 16 + (lane >> 4) * 8) * 2;
        ldsm_x4(ah[mi], addr);
        ldsm_x4(al[mi], addr + S::A_PLANE);
      }
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t bh[4], bl[4];
        uint32_t addr = b_base + (ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * S::B_STRIDE +
                        (wn * WN + np * 16 + (lane >> 4) * 8) * 2;
        ldsm_x4_t(bh, addr);
        ldsm_x4_t(bl, addr + S::B_PLANE);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            mma_bf16(acc[mi][2 * np + j], al[mi], bh[2 * j], bh[2 * j + 1]);
            mma_bf16(acc[mi][2 * np + j], ah[mi], bl[2 * j], bl[2 * j + 1]);
            mma_bf16(acc[mi][2 * np + j], ah[mi], bh[2 * j], bh[2 * j + 1]);
          }
      }
    }
  };

  const int it0 = (int)((int64_t)T * blockIdx.z / gridDim.z);
  const int it1 = (int)((int64_t)T * (blockIdx.z + 1) / gridDim.z);
  if (it1 > it0) {
    float4 v[4];
    load_B(0, it0);
    load_A(it0, v);
    store_A(0, v);
    cp_async_wait_all();
    __syncthreads();
    for (int it = it0; it < it1; ++it) {
      const int s = (it - it0) & 1;
      const bool more = (it + 1 < it1);
      if (more) { load_B(s ^ 1, it + 1); load_A(it + 1, v); }
      compute(s, it);
      if (more) store_A(s ^ 1, v);
      cp_async_wait_all();
      __syncthreads();
    }
  }

  // ---- epilogue: fp32 accumulators -> Y (each quad writes 32 contiguous bytes per row)
  const int g = lane >> 2, t = lane & 3;
  float* outp = p.partial ? p.partial + (int64_t)blockIdx.z * p.n_out * p.Cout : p.Y;
  const int ldo = p.partial ? p.Cout : p.ldy;
  const float* bias = p.partial ? nullptr : p.bias;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      int col = n0 + wn * WN + nt * 8 + 2 * t;
      float b0 = 0.f, b1 = 0.f;
      if (bias) { b0 = bias[col]; b1 = bias[col + 1]; }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int64_t row = row0 + wm * 32 + mi * 16 + g + h * 8;
        if (row < p.n_out) {
          float2 o = make_float2(acc[mi][nt][2 * h] + b0, acc[mi][nt][2 * h + 1] + b1);
          *reinterpret_cast<float2*>(outp + row * ldo + col) = o;
        }
      }
    }
}

// Y[row, c] = bias[c] + sum_z partial[z][row][c]   (fixed order: deterministic)
__global__ void conv_split_reduce_kernel(const float* __restrict__ partial, int nsplit, int64_t n_out, int Cout,
                                         const float* __restrict__ bias, float* __restrict__ Y, int ldy, int accumulate) {
  pdl_wait(); pdl_trigger();
  const int cv = Cout / 4;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_out * cv) return;
  int64_t row = i / cv;
  int c4 = (int)(i - row * cv);
  float4 s = bias ? reinterpret_cast<const float4*>(bias)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t plane = n_out * cv;
  for (int z = 0; z < nsplit; ++z) {
    float4 v = __ldg(reinterpret_cast<const float4*>(partial) + z * plane + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4* dst = reinterpret_cast<float4*>(Y + row * ldy + c4 * 4);
  if (accumulate) { float4 o = *dst; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
  *dst = s;
}

// --------------------------------------------------------------------------------------------- pooling (table gather-sum)
// Y[j, :] = sum_k X[tbl[kmap[k]][j], :]  (cnt[j] = number of present neighbours): the kernel of MinkowskiSumPooling / AvgPooling /
// PoolingTranspose / AvgUnpooling and of their backward passes (the same sum over the transposed table) -- the sibling models'
// pooling layers (`model/resnet.py:63`, `model/modules/common.py:170-214`).  One thread per (row, 4 channels).
__global__ void gather_sum_kernel(const float* __restrict__ X, int ldx, const int32_t* __restrict__ tbl, int64_t tbl_stride, KMap kmap, int K,
                                  int64_t n_out, int C, float* __restrict__ Y, int ldy, float* __restrict__ cnt) {
  pdl_wait(); pdl_trigger();
  const int cv = C / 4;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_out * cv) return;
  const int64_t j = e / cv;
  const int c4 = (int)(e - j * cv);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int present = 0;
  for (int k = 0; k < K; ++k) {
    const int idx = __ldg(tbl + (int64_t)kmap.v[k] * tbl_stride + j);
    if (idx < 0) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(X + (int64_t)idx * ldx) + c4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    ++present;
  }
  *reinterpret_cast<float4*>(Y + j * ldy + c4 * 4) = acc;
  if (cnt && c4 == 0) cnt[j] = (float)present;
}

// --------------------------------------------------------------------------------------------- forward (exact fp32 SIMT)
__global__ void conv_simt_kernel(const float* __restrict__ X, int ldx, const int32_t* __restrict__ tbl, int64_t tbl_stride,
                                 KMap kmap, int K, int64_t n_out, int Cin, int Cout, const float* __restrict__ W,
                                 const float* __restrict__ bias, float* __restrict__ Y, int ldy) {
  pdl_wait(); pdl_trigger();
  // one thread per (row, cout); consecutive threads -> consecutive cout (W reads coalesced, X reads broadcast)
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_out * Cout) return;
  int64_t j = e / Cout;
  int co = (int)(e - j * Cout);
  float acc = bias ? bias[co] : 0.f;
  for (int k = 0; k < K; ++k) {
    int idx = tbl[(int64_t)kmap.v[k] * tbl_stride + j];
    if (idx < 0) continue;
    const float* x = X + (int64_t)idx * ldx;
    const float* w = W + (int64_t)k * Cin * Cout + co;
    for (int ci = 0; ci < Cin; ++ci) acc = fmaf(x[ci], w[(int64_t)ci * Cout], acc);
  }
  Y[j * ldy + co] = acc;
}

// The stem layer (CIN = 3 -> 32 channels, exact fp32).  Thread = output row (consecutive threads -> consecutive rows: the table reads
// tbl[k][j] are coalesced), all 32 output channels of the row in registers, weights [K][CIN][32] broadcast from shared memory.
template <int CIN>
__global__ void __launch_bounds__(128) conv_stem_kernel(const float* __restrict__ X, int ldx, const int32_t* __restrict__ tbl,
                                                        int64_t tbl_stride, KMap kmap, int K, int64_t n_out, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ Y, int ldy) {
  pdl_wait(); pdl_trigger();
  __shared__ __align__(16) float s_w[PCB_MAX_KERNEL_VOLUME * CIN * 32];
  for (int e = threadIdx.x; e < K * CIN * 32; e += 128) s_w[e] = W[e];
  __syncthreads();
  const int64_t j = blockIdx.x * 128ll + threadIdx.x;
  if (j >= n_out) return;
  float acc[32];
#pragma unroll
  for (int co = 0; co < 32; ++co) acc[co] = bias ? bias[co] : 0.f;
  for (int k = 0; k < K; ++k) {
    const int idx = __ldg(tbl + (int64_t)kmap.v[k] * tbl_stride + j);
    if (idx < 0) continue;
    const float* xr = X + (int64_t)idx * ldx;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float x = __ldg(xr + c);
      const float4* w = reinterpret_cast<const float4*>(s_w + (k * CIN + c) * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = w[q];
        acc[q * 4 + 0] = fmaf(x, v.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(x, v.y, acc[q * 4 + 1]);
        acc[q * 4 + 2] = fmaf(x, v.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(x, v.w, acc[q * 4 + 3]);
      }
    }
  }
  float4* yo = reinterpret_cast<float4*>(Y + j * ldy);
#pragma unroll
  for (int q = 0; q < 8; ++q) yo[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
}

// --------------------------------------------------------------------------------------------- weight gradient (tensor cores)
constexpr int WK = 32;    // rows (reduction dim) per pipeline step

struct WgradArgs {
  const float* A; int lda;
  const float* B; int ldb;
  const int32_t* tbl; int64_t tbl_stride;
  int K; int64_t n_out;
  int Ca; int Cb;
  int rows_per_split;
  float* partial;       // [splits][K][Ca][Cb] (or transposed)
  int transpose_out;
};

template <int TM, int TN>
struct WgradSmem {
  static constexpr int AS = TM * 2 + 16;
  static constexpr int BS = TN * 2 + 16;
  static constexpr int A_PLANE = WK * AS;
  static constexpr int B_PLANE = WK * BS;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int TOTAL = 2 * STAGE;
};

// grid: x = K * mblocks * nblocks, y = splits.  8 warps as 2 (M) x 4 (N).
template <int TM, int TN>
__global__ void __launch_bounds__(NTHR, 2) wgrad_mma_kernel(const WgradArgs p) {
  using S = WgradSmem<TM, TN>;
  constexpr int MT = TM / 2 / 16;      // m16 tiles per warp
  constexpr int NT = TN / 4 / 8;       // n8 tiles per warp
  constexpr int ACH = TM / 4;          // float4 chunks per gathered A row
  constexpr int BCH = TN / 4;
  constexpr int AV = (WK * ACH + NTHR - 1) / NTHR;
  constexpr int BV = (WK * BCH + NTHR - 1) / NTHR;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int mblocks = p.Ca / TM, nblocks = p.Cb / TN;
  int bx = blockIdx.x;
  const int nb = bx % nblocks; bx /= nblocks;
  const int mb = bx % mblocks; bx /= mblocks;
  const int k = bx;
  const int split = blockIdx.y;
  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  const int64_t r_end = min(p.n_out, r_begin + p.rows_per_split);
  const int m0 = mb * TM, n0 = nb * TN;
  const int32_t* trow = p.tbl + (int64_t)k * p.tbl_stride;

  float acc[MT][NT][4];
#pragma unroll
  for (int a = 0; a < MT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

  const uint32_t smem_base = smem_u32(smem);
  const int nsteps = (int)((r_end - r_begin + WK - 1) / WK);

  auto load = [&](int step, float4 (&va)[AV], float4 (&vb)[BV], bool& any) {
    const int64_t rbase = r_begin + (int64_t)step * WK;
    int my = -1;
    if (lane < WK) { int64_t r = rbase + lane; if (r < r_end) my = trow[r]; }
    unsigned bal = __ballot_sync(0xffffffffu, my >= 0);
    any = bal != 0;
#pragma unroll
    for (int i = 0; i < AV; ++i) {
      int e = tid + i * NTHR;
      const bool ok = e < WK * ACH;
      int r = ok ? e / ACH : 0, c = e - r * ACH;
      int idx = __shfl_sync(0xffffffffu, my, r);      // executed by the full warp
      va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && idx >= 0) va[i] = __ldg(reinterpret_cast<const float4*>(p.A + (int64_t)idx * p.lda + m0) + c);
    }
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      int e = tid + i * NTHR;
      vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < WK * BCH) {
        int r = e / BCH, c = e - r * BCH;
        int64_t row = rbase + r;
        if (any && row < r_end && ((bal >> r) & 1))
          vb[i] = __ldg(reinterpret_cast<const float4*>(p.B + row * p.ldb + n0) + c);
      }
    }
  };
  auto store = [&](int stage, const float4 (&va)[AV], const float4 (&vb)[BV]) {
    unsigned char* base = smem + stage * S::STAGE;
#pragma unroll
    for (int i = 0; i < AV; ++i) {
      int e = tid + i * NTHR;
      if (e < WK * ACH) {
        int r = e / ACH, c = e - r * ACH;
        uint2 hi, lo; split4(va[i], hi, lo);
        *reinterpret_cast<uint2*>(base + r * S::AS + c * 8) = hi;
        *reinterpret_cast<uint2*>(base + S::A_PLANE + r * S::AS + c * 8) = lo;
      }
    }
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      int e = tid + i * NTHR;
      if (e < WK * BCH) {
        int r = e / BCH, c = e - r * BCH;
        uint2 hi, lo; split4(vb[i], hi, lo);
        *reinterpret_cast<uint2*>(base + 2 * S::A_PLANE + r * S::BS + c * 8) = hi;
        *reinterpret_cast<uint2*>(base + 2 * S::A_PLANE + S::B_PLANE + r * S::BS + c * 8) = lo;
      }
    }
  };
  auto compute = [&](int stage) {
    const uint32_t a_base = smem_base + stage * S::STAGE;
    const uint32_t b_base = a_base + 2 * S::A_PLANE;
#pragma unroll
    for (int ks = 0; ks < WK / 16; ++ks) {
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) {
        // A operand = gathered rows transposed: smem is [row j (k-dim)][channel (m-dim)] -> ldmatrix.trans
        uint32_t addr = a_base + (ks * 16 + (lane >> 4) * 8 + (lane & 7)) * S::AS +
                        (wm * (TM / 2) + mi * 16 + ((lane >> 3) & 1) * 8) * 2;
        ldsm_x4_t(ah[mi], addr);
        ldsm_x4_t(al[mi], addr + S::A_PLANE);
      }
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        uint32_t bh[2], bl[2];
        uint32_t addr = b_base + (ks * 16 + (lane & 15)) * S::BS + (wn * (TN / 4) + ni * 8) * 2;
        ldsm_x2_t(bh, addr);
        ldsm_x2_t(bl, addr + S::B_PLANE);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          mma_bf16(acc[mi][ni], al[mi], bh[0], bh[1]);
          mma_bf16(acc[mi][ni], ah[mi], bl[0], bl[1]);
          mma_bf16(acc[mi][ni], ah[mi], bh[0], bh[1]);
        }
      }
    }
  };

  // software pipeline; steps whose 32 rows have no neighbour are skipped (block-uniform decision via smem flag)
  __shared__ int s_any[2];
  if (nsteps > 0) {
    float4 va[AV], vb[BV];
    bool any;
    load(0, va, vb, any);
    store(0, va, vb);
    if (tid == 0) s_any[0] = any ? 1 : 0;
    __syncthreads();
    for (int step = 0; step < nsteps; ++step) {
      const int s = step & 1;
      const bool more = step + 1 < nsteps;
      bool any_next = false;
      if (more) load(step + 1, va, vb, any_next);
      if (s_any[s]) compute(s);
      if (more) { store(s ^ 1, va, vb); if (tid == 0) s_any[s ^ 1] = any_next ? 1 : 0; }
      __syncthreads();
    }
  }

  // ---- write the partial tile
  const int g = lane >> 2, t = lane & 3;
  float* out = p.partial + ((int64_t)split * p.K + k) * (int64_t)p.Ca * p.Cb;
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int m = m0 + wm * (TM / 2) + mi * 16 + g + h * 8;
        int n = n0 + wn * (TN / 4) + ni * 8 + 2 * t;
        float v0 = acc[mi][ni][2 * h], v1 = acc[mi][ni][2 * h + 1];
        if (!p.transpose_out) {
          *reinterpret_cast<float2*>(out + (int64_t)m * p.Cb + n) = make_float2(v0, v1);
        } else {
          out[(int64_t)n * p.Ca + m] = v0;
          out[(int64_t)(n + 1) * p.Ca + m] = v1;
        }
      }
}

// exact fp32 SIMT weight gradient: block = (k, split), threads stride over the Ca*Cb outputs
__global__ void wgrad_simt_kernel(const WgradArgs p) {
  const int k = blockIdx.x, split = blockIdx.y;
  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  const int64_t r_end = min(p.n_out, r_begin + p.rows_per_split);
  const int32_t* trow = p.tbl + (int64_t)k * p.tbl_stride;
  float* out = p.partial + ((int64_t)split * p.K + k) * (int64_t)p.Ca * p.Cb;
  for (int e = threadIdx.x; e < p.Ca * p.Cb; e += blockDim.x) {
    int m = e / p.Cb, n = e - m * p.Cb;
    float acc = 0.f;
    for (int64_t r = r_begin; r < r_end; ++r) {
      int idx = trow[r];
      if (idx >= 0) acc = fmaf(p.A[(int64_t)idx * p.lda + m], p.B[r * p.ldb + n], acc);
    }
    if (!p.transpose_out) out[(int64_t)m * p.Cb + n] = acc; else out[(int64_t)n * p.Ca + m] = acc;
  }
}

// Stem layer (CA = 3 input channels -> 32), exact fp32, deterministic.  A CTA walks tiles of 64 table rows: the gathered inputs of
// the tile, xs[row][k][c] (table reads coalesced along the rows), and the 64 x 32 dY tile are staged in shared memory; thread
// (kc, co) then owns output dW[k][c][co] for ~K*CA/8 values of kc and accumulates over the tile's rows from shared memory
// (xs broadcast within a warp, dY conflict-free).  One partial tile [K*CA*32] per CTA, summed by wgrad_reduce_kernel.
template <int CA>
__global__ void __launch_bounds__(256) wgrad_stem_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                         const int32_t* __restrict__ tbl, int64_t tbl_stride, int K, int64_t n_out,
                                                         int rows_per_block, float* __restrict__ partial) {
  pdl_wait(); pdl_trigger();
  constexpr int TR = 64, KC = PCB_MAX_KERNEL_VOLUME * CA, PER = (KC + 7) / 8;
  __shared__ float s_x[TR][KC + 1];
  __shared__ float s_dy[TR][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nkc = K * CA;
  float acc[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) acc[i] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(n_out, r0 + rows_per_block);
  for (int64_t t0 = r0; t0 < r1; t0 += TR) {
    const int rows = (int)min((int64_t)TR, r1 - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < K * TR; e += 256) {          // (k, row): consecutive threads -> consecutive rows of one table row
      const int k = e / TR, r = e - k * TR;
      int idx = -1;
      if (r < rows) idx = __ldg(tbl + (int64_t)k * tbl_stride + t0 + r);
      const float* ar = A + (int64_t)(idx >= 0 ? idx : 0) * lda;
#pragma unroll
      for (int c = 0; c < CA; ++c) s_x[r][k * CA + c] = idx >= 0 ? __ldg(ar + c) : 0.f;
    }
    for (int e = threadIdx.x; e < TR * 32; e += 256) {
      const int r = e >> 5, co = e & 31;
      s_dy[r][co] = r < rows ? __ldg(B + (t0 + r) * ldb + co) : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
      const float dy = s_dy[r][lane];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int kc = warp + 8 * i;
        if (kc < nkc) acc[i] = fmaf(s_x[r][kc], dy, acc[i]);
      }
    }
  }
  float* out = partial + (int64_t)blockIdx.x * nkc * 32;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int kc = warp + 8 * i;
    if (kc < nkc) out[kc * 32 + lane] = acc[i];
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ dW, int accumulate) {
  pdl_wait(); pdl_trigger();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += partial[(int64_t)sp * n + i];
  dW[i] = accumulate ? dW[i] + s : s;
}

// --------------------------------------------------------------------------------------------- weight preparation
__global__ void weight_prep_kernel(const float* __restrict__ W, int K, int Cin, int Cout, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ thi,
                                   __nv_bfloat16* __restrict__ tlo) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t n = (int64_t)K * Cin * Cout;
  if (e >= n) return;
  float w = W[e];
  __nv_bfloat16 h = __float2bfloat16_rn(w);
  __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
  hi[e] = h; lo[e] = l;
  int co = (int)(e % Cout);
  int64_t r = e / Cout;
  int ci = (int)(r % Cin);
  int64_t k = r / Cin;
  int64_t te = (k * Cout + co) * Cin + ci;
  thi[te] = h; tlo[te] = l;
}

int pick_tile(int C) {      // largest of {128, 96, 64, 32} dividing C
  if (C % 128 == 0) return 128;
  if (C % 96 == 0) return 96;
  if (C % 64 == 0) return 64;
  if (C % 32 == 0) return 32;
  return 0;
}

// Small levels (a few hundred rows at 256 channels) would otherwise be a handful of CTAs each walking 27 x Cin/32
// pipeline steps serially: split that loop over gridDim.z and reduce.
int conv_splits(int K, int64_t n_out, int Cin, int Cout) {
  int bn = pick_tile(Cout);
  int64_t base = ((n_out + BM - 1) / BM) * (Cout / bn);
  const int64_t one_wave = 2ll * num_sms();
  if (base >= one_wave) return 1;
  static double waves = 0.0;        // CTAs to aim for on a small level, in units of one resident wave (measured on C1: 0.25-0.5 best; 2 costs 4 ms/step)
  if (waves == 0.0) { const char* e = getenv("PCB_CONV_SPLIT_WAVES"); waves = e ? atof(e) : 0.5; if (waves < 0.05) waves = 0.05; }
  int64_t s = ((int64_t)(waves * one_wave) + base - 1) / base;
  int64_t T = (int64_t)K * (Cin / BK);
  if (s > T) s = T;
  if (s > 64) s = 64;
  return s < 2 ? 1 : (int)s;
}

template <int BN>
int launch_conv(ConvArgs a, int nsplit, float* ws, cudaStream_t st) {
  using S = ConvSmem<BN>;
  static bool attr_set[64] = {};          // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    PCB_CUDA(cudaFuncSetAttribute(conv_mma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set[dev_] = true;
  }
  a.partial = nsplit > 1 ? ws : nullptr;
  dim3 grid((unsigned)((a.n_out + BM - 1) / BM), a.Cout / BN, nsplit);
  conv_mma_kernel<BN><<<grid, NTHR, S::TOTAL, st>>>(a);
  if (int e = check_launch("conv_mma_kernel")) return e;
  if (nsplit > 1) {
    int64_t n4 = a.n_out * (a.Cout / 4);
    launch_kernel(conv_split_reduce_kernel, (unsigned)((n4 + 255) / 256), 256, 0, st, ws, nsplit, a.n_out, a.Cout, a.bias, a.Y, a.ldy, 0);
    return check_launch("conv_split_reduce_kernel");
  }
  return PCB_OK;
}

template <int TM, int TN>
int launch_wgrad(const WgradArgs& a, int splits, cudaStream_t st) {
  using S = WgradSmem<TM, TN>;
  static bool attr_set[64] = {};          // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    PCB_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set[dev_] = true;
  }
  dim3 grid((unsigned)(a.K * (a.Ca / TM) * (a.Cb / TN)), splits);
  wgrad_mma_kernel<TM, TN><<<grid, NTHR, S::TOTAL, st>>>(a);
  return check_launch("wgrad_mma_kernel");
}

int wgrad_splits(int K, int64_t n_out, int Ca, int Cb, int tm, int tn) {
  int64_t base = (int64_t)K * (tm ? Ca / tm : 1) * (tn ? Cb / tn : 1);
  int64_t target = 4ll * num_sms();
  int64_t s = (target + base - 1) / base;
  int64_t max_s = (n_out + 255) / 256;      // at least 256 rows per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

}  // namespace

extern "C" int pcb_weight_prep(const float* W, int K, int Cin, int Cout, uint16_t* w_hi, uint16_t* w_lo,
                               uint16_t* wt_hi, uint16_t* wt_lo, void* stream) {
  PCB_ARG(W && w_hi && w_lo && wt_hi && wt_lo && K >= 1 && Cin >= 1 && Cout >= 1);
  int64_t n = (int64_t)K * Cin * Cout;
  weight_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      W, K, Cin, Cout, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo, (__nv_bfloat16*)wt_hi, (__nv_bfloat16*)wt_lo);
  return check_launch("weight_prep_kernel");
}

namespace pcb {
int launch_wgrad_tcgen05(const uint16_t* Ahi, const uint16_t* Alo, int lda, const uint16_t* Bhi, const uint16_t* Blo, int ldb,
                         const int32_t* tbl, int64_t tbl_stride, int K, int64_t n_out, int Ca, int Cb, int rows_per_split, int splits,
                         float* partial, int transpose_out, int tn, cudaStream_t st, int a_fp16 = 0, int b_fp16 = 0);
int launch_conv_tcgen05(const float* X, int ldx, const uint16_t* Xhi, const uint16_t* Xlo, int lds, const void* wt, const int32_t* tbl,
                        int64_t tbl_stride, const int* kmap, int K, int64_t n_out,
                        int Cin, int Cout, const uint16_t* wk_hi, const uint16_t* wk_lo, const float* bias, float* Y, int ldy,
                        float* partial, int nsplit, int bn, int accumulate, cudaStream_t st, int x_fp16 = 0, int w_fp16 = 0);
}

extern "C" size_t pcb_conv_forward_ws_bytes(int K, int64_t n_out, int Cin, int Cout) {
  if (Cin % 32 || Cout % 32 || n_out <= 0) return 256;
  int s = conv_splits(K, n_out, Cin, Cout);
  return s > 1 ? (size_t)s * n_out * Cout * sizeof(float) + 256 : 256;
}

extern "C" int pcb_conv_forward(const float* X, int ldx, const int32_t* tbl, int64_t tbl_stride, const int32_t* kmap, int K,
                                int64_t n_out, int Cin, int Cout, const uint16_t* w_hi, const uint16_t* w_lo,
                                const uint16_t* wk_hi, const uint16_t* wk_lo, const float* w_f32, const float* bias, float* Y,
                                int ldy, void* ws, size_t ws_bytes, int flags, void* stream) {
  PCB_ARG(K >= 1 && K <= PCB_MAX_KERNEL_VOLUME && n_out >= 0 && Cin >= 1 && Cout >= 1 && ldx >= Cin && ldy >= Cout);
  if (n_out == 0) return PCB_OK;
  PCB_ARG(X && tbl && Y && tbl_stride >= n_out);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, 0);
  KMap km;
  for (int k = 0; k < K; ++k) { km.v[k] = kmap ? kmap[k] : k; PCB_ARG(km.v[k] >= 0 && km.v[k] < PCB_MAX_KERNEL_VOLUME); }
  const bool tc_ok = (Cin % 32 == 0) && (Cout % 32 == 0) && (ldx % 4 == 0) && (ldy % 2 == 0) && w_hi && w_lo &&
                     !(flags & PCB_CONV_FORCE_SIMT);
  if (!tc_ok) {
    if (flags & PCB_CONV_ACCUMULATE) { set_error("PCB_CONV_ACCUMULATE needs the tcgen05 path"); return PCB_ERR_ARG; }
    if (!w_f32) { set_error("pcb_conv_forward: SIMT path needs w_f32 (Cin=%d Cout=%d)", Cin, Cout); return PCB_ERR_ARG; }
    if (Cin == 3 && Cout == 32) {         // the stem layer
      launch_kernel(conv_stem_kernel<3>, (unsigned)((n_out + 127) / 128), 128, 0, st, X, ldx, tbl, tbl_stride, km, K, n_out, w_f32, bias, Y, ldy);
      return check_launch("conv_stem_kernel");
    }
    int64_t total = n_out * Cout;
    launch_kernel(conv_simt_kernel, (unsigned)((total + 255) / 256), 256, 0, st, X, ldx, tbl, tbl_stride, km, K, n_out, Cin, Cout,
                                                                      w_f32, bias, Y, ldy);
    return check_launch("conv_simt_kernel");
  }
  ConvArgs a;
  a.X = X; a.ldx = ldx; a.tbl = tbl; a.tbl_stride = tbl_stride; a.kmap = km; a.K = K; a.n_out = n_out; a.Cin = Cin;
  a.Cout = Cout; a.w_hi = (const __nv_bfloat16*)w_hi; a.w_lo = (const __nv_bfloat16*)w_lo; a.bias = bias; a.Y = Y; a.ldy = ldy;
  a.partial = nullptr;
  int nsplit = conv_splits(K, n_out, Cin, Cout);
  if (nsplit > 1) {
    PCB_ARG(ws && ws_bytes >= (size_t)nsplit * n_out * Cout * sizeof(float));
    PCB_ARG(ldy % 4 == 0);
  }
  const int accumulate = (flags & PCB_CONV_ACCUMULATE) ? 1 : 0;
  if ((flags & PCB_CONV_TCGEN05) && wk_hi && wk_lo && ldy % 4 == 0) {
    if (int e = launch_conv_tcgen05(X, ldx, nullptr, nullptr, 0, nullptr, tbl, tbl_stride, km.v, K, n_out, Cin, Cout, wk_hi, wk_lo, bias, Y, ldy,
                                    nsplit > 1 ? (float*)ws : nullptr, nsplit, pick_tile(Cout), accumulate, st)) return e;
    if (nsplit > 1) {
      int64_t n4 = n_out * (Cout / 4);
      launch_kernel(conv_split_reduce_kernel, (unsigned)((n4 + 255) / 256), 256, 0, st, (const float*)ws, nsplit, n_out, Cout, bias, Y, ldy,
                                                                             accumulate);
      return check_launch("conv_split_reduce_kernel");
    }
    return PCB_OK;
  }
  if (accumulate) { set_error("PCB_CONV_ACCUMULATE needs the tcgen05 path"); return PCB_ERR_ARG; }
  switch (pick_tile(Cout)) {
    case 128: return launch_conv<128>(a, nsplit, (float*)ws, st);
    case 96: return launch_conv<96>(a, nsplit, (float*)ws, st);
    case 64: return launch_conv<64>(a, nsplit, (float*)ws, st);
    default: return launch_conv<32>(a, nsplit, (float*)ws, st);
  }
}

extern "C" int pcb_gather_sum(const float* X, int ldx, const int32_t* tbl, int64_t tbl_stride, const int32_t* kmap, int K, int64_t n_out, int C,
                              float* Y, int ldy, float* cnt, void* stream) {
  PCB_ARG(K >= 1 && K <= PCB_MAX_KERNEL_VOLUME && n_out >= 0 && C >= 4 && C % 4 == 0 && ldx >= C && ldy >= C && ldx % 4 == 0 && ldy % 4 == 0);
  if (n_out == 0) return PCB_OK;
  PCB_ARG(X && tbl && Y && tbl_stride >= n_out);
  KMap km;
  for (int k = 0; k < K; ++k) { km.v[k] = kmap ? kmap[k] : k; PCB_ARG(km.v[k] >= 0 && km.v[k] < PCB_MAX_KERNEL_VOLUME); }
  const int64_t total = n_out * (C / 4);
  launch_kernel(gather_sum_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, X, ldx, tbl, tbl_stride, km, K, n_out, C, Y, ldy,
                cnt);
  return check_launch("gather_sum_kernel");
}

extern "C" size_t pcb_conv_wgrad_ws_bytes(int K, int64_t n_out, int Ca, int Cb) {
  if (Ca == 3 && Cb == 32) return (size_t)2 * num_sms() * K * Ca * Cb * sizeof(float) + 256;
  int tm = pick_tile(Ca), tn = pick_tile(Cb);
  if (!tm || !tn) { tm = 0; tn = 0; }
  int s = wgrad_splits(K, n_out, Ca, Cb, tm, tn);
  return (size_t)s * K * Ca * Cb * sizeof(float) + 256;
}

#define WG_CASE(TM_, TN_) if (tm == TM_ && tn == TN_) rc = launch_wgrad<TM_, TN_>(a, splits, st)

extern "C" int pcb_conv_wgrad(const float* A, int lda, const float* B, int ldb, const int32_t* tbl, int64_t tbl_stride, int K,
                              int64_t n_out, int Ca, int Cb, float* dW, int transpose_out, void* ws, size_t ws_bytes,
                              int flags, void* stream) {
  PCB_ARG(K >= 1 && K <= PCB_MAX_KERNEL_VOLUME && n_out >= 0 && Ca >= 1 && Cb >= 1 && dW && lda >= Ca && ldb >= Cb);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nW = (int64_t)K * Ca * Cb;
  if (n_out == 0) {
    if (!(flags & PCB_CONV_ACCUMULATE)) PCB_CUDA(cudaMemsetAsync(dW, 0, nW * sizeof(float), st));
    return PCB_OK;
  }
  PCB_ARG(A && B && tbl && ws && tbl_stride >= n_out);
  ProfScope prof(st, 1);
  if (Ca == 3 && Cb == 32 && !transpose_out && !(flags & PCB_CONV_FORCE_SIMT)) {      // the stem layer: dedicated exact-fp32 kernel
    const int blocks = (int)((size_t)ws_bytes / ((size_t)nW * sizeof(float)));
    PCB_ARG(blocks >= 1);
    int nb = blocks < 2 * num_sms() ? blocks : 2 * num_sms();
    int64_t rpb = (n_out + nb - 1) / nb;
    rpb = (rpb + 63) / 64 * 64;
    nb = (int)((n_out + rpb - 1) / rpb);
    launch_kernel(wgrad_stem_kernel<3>, nb, 256, 0, st, A, lda, B, ldb, tbl, tbl_stride, K, n_out, (int)rpb, (float*)ws);
    if (int e = check_launch("wgrad_stem_kernel")) return e;
    launch_kernel(wgrad_reduce_kernel, (unsigned)((nW + 255) / 256), 256, 0, st, (const float*)ws, nb, nW, dW, (flags & PCB_CONV_ACCUMULATE) ? 1 : 0);
    return check_launch("wgrad_reduce_kernel");
  }
  int tm = pick_tile(Ca), tn = pick_tile(Cb);
  const bool tc_ok = tm && tn && (lda % 4 == 0) && (ldb % 4 == 0) && !(flags & PCB_CONV_FORCE_SIMT);
  if (!tc_ok) { tm = 0; tn = 0; }
  const int splits = wgrad_splits(K, n_out, Ca, Cb, tm, tn);
  PCB_ARG(ws_bytes >= (size_t)splits * nW * sizeof(float));
  WgradArgs a;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.tbl = tbl; a.tbl_stride = tbl_stride; a.K = K; a.n_out = n_out;
  a.Ca = Ca; a.Cb = Cb; a.partial = (float*)ws; a.transpose_out = transpose_out;
  int64_t rps = (n_out + splits - 1) / splits;
  rps = (rps + WK - 1) / WK * WK;
  a.rows_per_split = (int)rps;
  int rc = PCB_OK;
  if (!tc_ok) {
    dim3 grid(K, splits);
    wgrad_simt_kernel<<<grid, 256, 0, st>>>(a);
    rc = check_launch("wgrad_simt_kernel");
  } else {
    rc = PCB_ERR_ARG;
    WG_CASE(128, 128); WG_CASE(128, 96); WG_CASE(128, 64); WG_CASE(128, 32);
    WG_CASE(96, 128);  WG_CASE(96, 96);  WG_CASE(96, 64);  WG_CASE(96, 32);
    WG_CASE(64, 128);  WG_CASE(64, 96);  WG_CASE(64, 64);  WG_CASE(64, 32);
    WG_CASE(32, 128);  WG_CASE(32, 96);  WG_CASE(32, 64);  WG_CASE(32, 32);
  }
  if (rc) return rc;
  launch_kernel(wgrad_reduce_kernel, (unsigned)((nW + 255) / 256), 256, 0, st, (const float*)ws, splits, nW, dW,
                                                                    (flags & PCB_CONV_ACCUMULATE) ? 1 : 0);
  return check_launch("wgrad_reduce_kernel");
}


// ------------------------------------------------------------------------------------------------ split-operand entry points
// Weights as shared-memory images for the split conv kernel: per (offset k, 32-channel chunk kc, BN-column block nb) one blob
//   [hi plane | lo plane], plane = 4 k8-groups x (BN/8 core matrices x 128 B + 16 B pad)   (UMMA K-major, no swizzle)
// so that a pipeline stage's weight tile is ONE contiguous TMA bulk copy.
namespace {
__host__ __device__ inline int64_t tile_plane_bytes(int bn) { return 4ll * ((bn / 8) * 128 + 16); }

__global__ void weight_tile_kernel(const float* __restrict__ W, int K, int Cin, int Cout, int bn_f, int bn_d,
                                   unsigned char* __restrict__ fwd, unsigned char* __restrict__ dg, int fwd_fp16) {
  pdl_wait(); pdl_trigger();
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)K * Cin * Cout) return;
  const float w = W[e];
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
  const int co = (int)(e % Cout);
  const int64_t r = e / Cout;
  const int ci = (int)(r % Cin);
  const int k = (int)(r / Cin);
  {   // forward roles: N = Cout, contraction = Cin
    const int64_t plane = tile_plane_bytes(bn_f);
    const int64_t blob = ((int64_t)(k * (Cin / 32) + ci / 32) * (Cout / bn_f) + co / bn_f) * 2 * plane;
    const int n = co % bn_f, c = ci % 32;
    const int64_t off = blob + (c / 8) * (plane / 4) + (n / 8) * 128 + (n % 8) * 16 + (c % 8) * 2;
    if (fwd_fp16) {          // fp16 hi/lo of W * 2^10: the lo plane stays in fp16's normal range for every weight that matters
      const float ws = fminf(fmaxf(w * 1024.0f, -65000.f), 65000.f);
      const __half fh = __float2half_rn(ws);
      const __half fl = __float2half_rn(ws - __half2float(fh));
      *reinterpret_cast<__half*>(fwd + off) = fh;
      *reinterpret_cast<__half*>(fwd + off + plane) = fl;
    } else {
      *reinterpret_cast<__nv_bfloat16*>(fwd + off) = h;
      *reinterpret_cast<__nv_bfloat16*>(fwd + off + plane) = l;
    }
  }
  {   // data-gradient roles: N = Cin, contraction = Cout
    const int64_t plane = tile_plane_bytes(bn_d);
    const int64_t blob = ((int64_t)(k * (Cout / 32) + co / 32) * (Cin / bn_d) + ci / bn_d) * 2 * plane;
    const int n = ci % bn_d, c = co % 32;
    const int64_t off = blob + (c / 8) * (plane / 4) + (n / 8) * 128 + (n % 8) * 16 + (c % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(dg + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(dg + off + plane) = l;
  }
}
}  // namespace

extern "C" size_t pcb_weight_tile_bytes(int K, int Cin, int Cout, int dgrad_roles) {
  if (Cin % 32 || Cout % 32) return 0;
  const int N = dgrad_roles ? Cin : Cout, Kc = dgrad_roles ? Cout : Cin;
  const int bn = pick_tile(N);
  return (size_t)K * (Kc / 32) * (N / bn) * 2 * tile_plane_bytes(bn);
}

extern "C" int pcb_weight_tile(const float* W, int K, int Cin, int Cout, void* fwd_tiles, void* dgrad_tiles, int flags, void* stream) {
  PCB_ARG(W && fwd_tiles && dgrad_tiles && K >= 1 && Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32);
  int64_t n = (int64_t)K * Cin * Cout;
  launch_kernel(weight_tile_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, W, K, Cin, Cout, pick_tile(Cout), pick_tile(Cin),
                (unsigned char*)fwd_tiles, (unsigned char*)dgrad_tiles, (flags & PCB_PLANES_B_FP16) ? 1 : 0);
  return check_launch("weight_tile_kernel");
}

namespace pcb {
int bn_reduce_stats_launch(const float* P, int nsplit, float* Y, int ldy, int64_t n, int64_t n0, int C, float eps, float momentum,
                           float* mean, float* invstd, float* running_mean, float* running_var, void* ws, size_t ws_bytes, cudaStream_t st);

// bn != NULL: a BatchNorm follows this convolution (no bias, no accumulate).  When the convolution runs offset-split (small levels)
// its reduction pass also produces the BatchNorm statistics (one read of the partial planes instead of reduce + a column-sum pass
// over Y) and *bn_done = 1; in direct mode nothing changes and *bn_done = 0 (the caller runs pcb_bn_stats_seg: fusing the column
// sums into the TMEM epilogue was measured SLOWER than the separate pass, profiles/r2_results.md).
struct BnFuse { int64_t n0; float eps, momentum; float* mean; float* invstd; float* running_mean; float* running_var; void* ws; size_t ws_bytes; };
int conv_forward_split_impl(const uint16_t* Xhi, const uint16_t* Xlo, int lds, const int32_t* tbl, int64_t tbl_stride, const int32_t* kmap,
                            int K, int64_t n_out, int Cin, int Cout, const void* w_tiles, const float* bias, float* Y, int ldy, void* ws,
                            size_t ws_bytes, int flags, cudaStream_t st, const BnFuse* bn, int* bn_done) {
  PCB_ARG(K >= 1 && K <= PCB_MAX_KERNEL_VOLUME && n_out >= 0 && Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32);
  PCB_ARG(lds >= Cin && lds % 8 == 0 && ldy >= Cout && ldy % 4 == 0);
  if (bn_done) *bn_done = 0;
  if (n_out == 0) return PCB_OK;
  PCB_ARG(Xhi && Xlo && tbl && Y && w_tiles && tbl_stride >= n_out);
  ProfScope prof(st, 0);
  int km[PCB_MAX_KERNEL_VOLUME];
  for (int k = 0; k < K; ++k) { km[k] = kmap ? kmap[k] : k; PCB_ARG(km[k] >= 0 && km[k] < PCB_MAX_KERNEL_VOLUME); }
  const int accumulate = (flags & PCB_CONV_ACCUMULATE) ? 1 : 0;
  const int nsplit = conv_splits(K, n_out, Cin, Cout);
  if (nsplit > 1) PCB_ARG(ws && ws_bytes >= (size_t)nsplit * n_out * Cout * sizeof(float));
  if (bn) PCB_ARG(!bias && !accumulate && bn->n0 >= 1 && bn->n0 <= n_out && bn_done);
  if (int e = launch_conv_tcgen05(nullptr, 0, Xhi, Xlo, lds, w_tiles, tbl, tbl_stride, km, K, n_out, Cin, Cout, nullptr, nullptr, bias, Y, ldy,
                                  nsplit > 1 ? (float*)ws : nullptr, nsplit, pick_tile(Cout), accumulate, st,
                                  (flags & PCB_PLANES_A_FP16) ? 1 : 0, (flags & PCB_PLANES_B_FP16) ? 1 : 0)) return e;
  if (nsplit > 1) {
    if (bn) {
      *bn_done = 1;
      return bn_reduce_stats_launch((const float*)ws, nsplit, Y, ldy, n_out, bn->n0, Cout, bn->eps, bn->momentum, bn->mean, bn->invstd,
                                    bn->running_mean, bn->running_var, bn->ws, bn->ws_bytes, st);
    }
    int64_t n4 = n_out * (Cout / 4);
    launch_kernel(conv_split_reduce_kernel, (unsigned)((n4 + 255) / 256), 256, 0, st, (const float*)ws, nsplit, n_out, Cout, bias, Y, ldy, accumulate);
    return check_launch("conv_split_reduce_kernel");
  }
  return PCB_OK;
}
}  // namespace pcb

extern "C" int pcb_conv_forward_split(const uint16_t* Xhi, const uint16_t* Xlo, int lds, const int32_t* tbl, int64_t tbl_stride,
                                      const int32_t* kmap, int K, int64_t n_out, int Cin, int Cout, const void* w_tiles,
                                      const float* bias, float* Y, int ldy, void* ws, size_t ws_bytes,
                                      int flags, void* stream) {
  return pcb::conv_forward_split_impl(Xhi, Xlo, lds, tbl, tbl_stride, kmap, K, n_out, Cin, Cout, w_tiles, bias, Y, ldy, ws, ws_bytes, flags,
                                      (cudaStream_t)stream, nullptr, nullptr);
}

namespace {
int wgrad_split_splits(int K, int64_t n_out, int Ca, int Cb) {
  int tn = pick_tile(Cb);
  int64_t base = (int64_t)((K + 3) / 4) * ((Ca + 127) / 128) * (Cb / tn);      // CTAs per split: offset groups x channel blocks
  static double wwaves = 0.0;
  if (wwaves == 0.0) { const char* e = getenv("PCB_WGRAD_SPLIT_WAVES"); wwaves = e ? atof(e) : 1.0; if (wwaves < 0.05) wwaves = 0.05; }      // measured on C1: 1 wave best (0.5 under-fills, 2-3 add reduce traffic)
  int64_t s = (int64_t)(wwaves * num_sms()) / base;       // one CTA per SM: whole waves, never a nearly-empty extra one
  int64_t max_s = (n_out + 63) / 64;          // small levels: rather many short CTAs than a few long serial ones
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 96) s = 96;
  return (int)s;
}
}  // namespace

extern "C" size_t pcb_conv_wgrad_split_ws_bytes(int K, int64_t n_out, int Ca, int Cb) {
  if (Ca % 32 || Cb % 32 || n_out <= 0) return 256;
  return (size_t)wgrad_split_splits(K, n_out, Ca, Cb) * K * Ca * Cb * sizeof(float) + 256;
}

extern "C" int pcb_conv_wgrad_split(const uint16_t* Ahi, const uint16_t* Alo, int lda, const uint16_t* Bhi, const uint16_t* Blo, int ldb,
                                    const int32_t* tbl, int64_t tbl_stride, int K, int64_t n_out, int Ca, int Cb, float* dW,
                                    int transpose_out, void* ws, size_t ws_bytes, int flags, void* stream) {
  PCB_ARG(K >= 1 && K <= PCB_MAX_KERNEL_VOLUME && n_out >= 0 && Ca % 32 == 0 && Cb % 32 == 0 && Ca >= 32 && Cb >= 32 && dW);
  PCB_ARG(lda >= Ca && ldb >= Cb && lda % 8 == 0 && ldb % 8 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nW = (int64_t)K * Ca * Cb;
  if (n_out == 0) {
    if (!(flags & PCB_CONV_ACCUMULATE)) PCB_CUDA(cudaMemsetAsync(dW, 0, nW * sizeof(float), st));
    return PCB_OK;
  }
  PCB_ARG(Ahi && Alo && Bhi && Blo && tbl && ws && tbl_stride >= n_out);
  // tcgen05.mma.kind::f16 takes ONE 16-bit format for both operands (fp16 x bf16 is an illegal instruction)
  PCB_ARG(((flags & PCB_PLANES_A_FP16) != 0) == ((flags & PCB_PLANES_B_FP16) != 0));
  ProfScope prof(st, 1);
  const int splits = wgrad_split_splits(K, n_out, Ca, Cb);
  PCB_ARG(ws_bytes >= (size_t)splits * nW * sizeof(float));
  int64_t rps = (n_out + splits - 1) / splits;
  rps = (rps + 15) / 16 * 16;
  if (int e = launch_wgrad_tcgen05(Ahi, Alo, lda, Bhi, Blo, ldb, tbl, tbl_stride, K, n_out, Ca, Cb, (int)rps, splits, (float*)ws,
                                   transpose_out, pick_tile(Cb), st, (flags & PCB_PLANES_A_FP16) ? 1 : 0, (flags & PCB_PLANES_B_FP16) ? 1 : 0)) return e;
  launch_kernel(wgrad_reduce_kernel, (unsigned)((nW + 255) / 256), 256, 0, st, (const float*)ws, splits, nW, dW,
                                                                    (flags & PCB_CONV_ACCUMULATE) ? 1 : 0);
  return check_launch("wgrad_reduce_kernel");
}
