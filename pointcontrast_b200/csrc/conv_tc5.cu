// Blackwell-native sparse convolution: the same output-stationary gather-GEMM as conv.cu, with the per-offset
// Cin x Cout contraction issued as tcgen05.mma (5th-gen tensor cores), accumulators in Tensor Memory.
//
//   CTA = 128 output rows x BN output channels, 256 threads.
//   producers (all 8 warps): gather the neighbour rows of the current kernel offset with 128-bit loads, split fp32 ->
//       bf16 hi/lo, store into shared memory in the canonical UMMA K-major / no-swizzle core-matrix layout
//       (8 rows x 16 B per core matrix); weights (pre-split bf16, K-major = [K][Cout][Cin]) arrive by cp.async;
//   one elected thread: 2 k16-steps x 3 products (lo.hi + hi.lo + hi.hi) tcgen05.mma.cta_group::1.kind::f16,
//       M = 128, N = BN, fp32 accumulate in TMEM; tcgen05.commit -> mbarrier releases the smem stage (3-stage ring);
//   epilogue: tcgen05.ld 32x32b -> registers -> 64-byte contiguous stores per thread.
//
// This first kernel (conv_tcgen05_kernel) takes fp32 rows and serves the modular per-operator surface; the training / inference
// executor runs conv_tcgen05_split_kernel below on operands that already are 16-bit hi/lo planes.  Both skip kernel offsets without a
// neighbour in the tile and have an offset-split mode (partial planes + fixed-order reduce) for levels with few rows.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc5_ptx.cuh"

using namespace pcb;

namespace pcb {

namespace tc5 {

__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 h1 = __floats2bfloat162_rn(v.z, v.w);
  float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
  __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
  hi.x = *reinterpret_cast<uint32_t*>(&h0); hi.y = *reinterpret_cast<uint32_t*>(&h1);
  lo.x = *reinterpret_cast<uint32_t*>(&l0); lo.y = *reinterpret_cast<uint32_t*>(&l1);
}

constexpr int BM = 128, BK = 32, NTHR = 256, NS = 3;
constexpr int A_SBO = 128;
// k8-chunk stride of the A tile: +32 bytes, so that the four 16-byte chunks (t & 3) x two rows (t >> 2) written by a quarter-warp of a
// 128-bit st.shared cover eight DIFFERENT 16-byte slots of a 128-byte bank line.  (Round 1 used +64: chunks 0/2 and 1/3 collided --
// ncu counted 23.1 M of 39.8 M shared wavefronts as bank conflicts on the block8 shape.)
constexpr int A_LBO = (BM / 8) * 128 + 32;
constexpr int A_PLANE = (BK / 8) * A_LBO;

struct Args {
  const float* X; int ldx;
  const __nv_bfloat16* Xhi; const __nv_bfloat16* Xlo; int lds;      // SPLIT kernels: the input as bf16 hi/lo planes
  const int32_t* tbl; int64_t tbl_stride;
  int kmap[PCB_MAX_KERNEL_VOLUME]; int K;
  int64_t n_out; int Cin; int Cout;
  const __nv_bfloat16* wk_hi; const __nv_bfloat16* wk_lo;      // K-major weights: [K][Cout][Cin]
  const unsigned char* wt;                                      // split kernel: weights pre-tiled as shared-memory images
  const float* bias;
  float* Y; int ldy;
  float* partial;
  int accumulate;       // Y += result (direct mode only; the split mode accumulates in the reduce kernel)
  // operand formats of the tcgen05.mma instruction descriptor (bits 7-9: A = gathered rows, bits 10-12: B = weights; 0 fp16, 1 bf16)
  // and the factor applied to the accumulators on the way out (2^-10 when the weight tiles hold fp16(W * 2^10))
  uint32_t fmt_bits; float out_scale;
  int debug;            // timing experiments only (PCB_TC5_DEBUG): 1 skip A copies, 2 skip B copies, 4 skip MMAs, 8 skip proxy fence
  int consumer_fence;   // generic -> async proxy fence issued by the MMA thread after the "full" wait instead of by every producer (see consumer_fence())
};

// Where the generic-proxy -> async-proxy fence of a stage sits.  `fence.proxy.async` compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and
// the membar waits for EVERY outstanding memory operation of the issuing thread -- for a producer that is the gather loads it has just
// put in flight for the next stages, so each stage costs one full L2 round trip no matter how far ahead the loads are issued (ncu, round
// 2: 25 % of all warp samples of the kernel sat on these two instructions with long-scoreboard stalls; an iteration took ~1900 cycles at
// any occupancy).  The MMA-issuing thread has no loads in flight: producers st.shared -> mbarrier.arrive (release.cta), the issuer
// mbarrier.try_wait (acquire.cta) -> fence.proxy.async -> tcgen05.mma.  Measured (profiles/r2_results.md): forward / data-gradient kernel
// 339 -> 318 us on the block8 shape, 12.66 -> 12.08 ms per step.  The weight-gradient kernel with 4 offsets per CTA (12 MMAs per step on
// the issuing thread) was 5 % slower with it; in its two-CTAs-per-SM form (2 offsets per CTA) it gains 9-12 % (458 -> 415 us on the block8
// shape, 435 -> 382 us at 128 channels).  PCB_TC5_FENCE=producer / PCB_WG_FENCE=producer restore the writer-side fence.
inline int consumer_fence() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PCB_TC5_FENCE"); v = (e && !strcmp(e, "producer")) ? 0 : 1; }
  return v;
}
inline int wgrad_consumer_fence() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PCB_WG_FENCE"); v = (e && !strcmp(e, "producer")) ? 0 : 1; }
  return v;
}

template <int BN>
struct Smem {
  static constexpr int B_SBO = 128;
  static constexpr int B_LBO = (BN / 8) * 128 + 16;
  static constexpr int B_PLANE = (BK / 8) * B_LBO;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int IDX_OFF = NS * STAGE;
  static constexpr int META_OFF = IDX_OFF + PCB_MAX_KERNEL_VOLUME * BM * 4;     // flags[32] klist[32] nk, tmem ptr
  static constexpr int BAR_OFF = META_OFF + 72 * 4;
  static constexpr int TOTAL = BAR_OFF + (NS + 1) * 8 + 16;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(NTHR, 2) conv_tcgen05_kernel(const Args p) {
  using S = Smem<BN>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  int* s_idx = reinterpret_cast<int*>(smem + S::IDX_OFF);
  int* s_flag = reinterpret_cast<int*>(smem + S::META_OFF);
  int* s_klist = s_flag + 32;
  int* s_nk = s_klist + 32;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_nk + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + S::BAR_OFF;          // NS "stage free" barriers + 1 "accumulator done"

  if (tid == 0) {
    for (int i = 0; i <= NS; ++i) mbar_init(bar_base + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {      // one warp allocates the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(S::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  for (int e = tid; e < p.K * BM; e += NTHR) {
    int k = e / BM, r = e - k * BM;
    int64_t row = row0 + r;
    int v = -1;
    if (row < p.n_out) v = p.tbl[(int64_t)p.kmap[k] * p.tbl_stride + row];
    s_idx[e] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *s_tmem;
  for (int k = warp; k < p.K; k += NTHR / 32) {
    unsigned any = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) any |= __ballot_sync(0xffffffffu, s_idx[k * BM + s * 32 + lane] >= 0);
    if (lane == 0) s_flag[k] = any ? 1 : 0;
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
  const int it0 = (int)((int64_t)T * blockIdx.z / gridDim.z);
  const int it1 = (int)((int64_t)T * (blockIdx.z + 1) / gridDim.z);

  const int a_chunk = tid & 7, a_row = tid >> 3;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  auto load_B = [&](int stage, int it) {
    const int k = s_klist[it / nkc], kc = it % nkc;
    constexpr int PER_PLANE = BN * (BK / 8);             // 16-byte chunks: BN rows x 4 k-chunks
    for (int c = tid; c < 2 * PER_PLANE; c += NTHR) {
      int plane = c / PER_PLANE, rem = c - plane * PER_PLANE;
      int n = rem >> 2, k8 = rem & 3;
      const __nv_bfloat16* src = (plane ? p.wk_lo : p.wk_hi) + ((int64_t)k * p.Cout + n0 + n) * p.Cin + kc * BK + k8 * 8;
      uint32_t dst = smem_base + stage * S::STAGE + 2 * A_PLANE + plane * S::B_PLANE + k8 * S::B_LBO + (n >> 3) * S::B_SBO +
                     (n & 7) * 16;
      cp_async16(dst, src);
    }
    cp_async_commit();
  };
  auto load_A = [&](int it, float4 (&v)[4]) {
    const int k = s_klist[it / nkc], kc = it % nkc;
    if (SPLIT) {       // pure asynchronous copy: 128 rows x 4 chunks x 2 planes, zero-filled where there is no neighbour
      const int stage = (it - it0) % NS;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = tid + i * NTHR;
        const int plane = c >> 9, rem = c & 511, r = rem >> 2, k8 = rem & 3;
        const int idx = s_idx[k * BM + r];
        const __nv_bfloat16* src = (plane ? p.Xlo : p.Xhi) + (int64_t)(idx >= 0 ? idx : 0) * p.lds + kc * BK + k8 * 8;
        const uint32_t dst = smem_base + stage * S::STAGE + plane * A_PLANE + k8 * A_LBO + (r >> 3) * A_SBO + (r & 7) * 16;
        cp_async16_zfill(dst, src, idx >= 0 ? 16u : 0u);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = s_idx[k * BM + a_row + 32 * i];
      if (idx >= 0) v[i] = __ldg(reinterpret_cast<const float4*>(p.X + (int64_t)idx * p.ldx + kc * BK) + a_chunk);
      else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_A = [&](int stage, const float4 (&v)[4]) {
    if (SPLIT) return;
    unsigned char* base = smem + stage * S::STAGE + (a_chunk >> 1) * A_LBO + (a_chunk & 1) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = a_row + 32 * i;
      uint2 hi, lo;
      split4(v[i], hi, lo);
      const int off = (r >> 3) * A_SBO + (r & 7) * 16;
      *reinterpret_cast<uint2*>(base + off) = hi;
      *reinterpret_cast<uint2*>(base + A_PLANE + off) = lo;
    }
  };

  if (it1 > it0) {
    float4 v[4];
    load_A(it0, v);          // SPLIT: cp.async into stage 0 (joins the commit group closed by load_B)
    load_B(0, it0);
    for (int it = it0; it < it1; ++it) {
      const int i = it - it0;
      const int s = i % NS;
      const bool more = it + 1 < it1;
      store_A(s, v);
      if (more) {
        const int s1 = (i + 1) % NS, u1 = (i + 1) / NS;
        if (u1 >= 1) mbar_wait(bar_base + 8 * s1, (u1 - 1) & 1);     // the MMAs that read stage s1 have retired
        load_A(it + 1, v);
        load_B(s1, it + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      fence_proxy_async();              // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a_hi = smem_base + s * S::STAGE, a_lo = a_hi + A_PLANE;
        const uint32_t b_hi = a_hi + 2 * A_PLANE, b_lo = b_hi + S::B_PLANE;
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint64_t dah = make_desc(a_hi + j * 2 * A_LBO, A_LBO, A_SBO), dal = make_desc(a_lo + j * 2 * A_LBO, A_LBO, A_SBO);
          const uint64_t dbh = make_desc(b_hi + j * 2 * S::B_LBO, S::B_LBO, S::B_SBO);
          const uint64_t dbl = make_desc(b_lo + j * 2 * S::B_LBO, S::B_LBO, S::B_SBO);
          tc_mma(tmem_acc, dal, dbh, IDESC, (i > 0 || j > 0) ? 1u : 0u);
          tc_mma(tmem_acc, dah, dbl, IDESC, 1u);
          tc_mma(tmem_acc, dah, dbh, IDESC, 1u);
        }
        tc_commit(bar_base + 8 * s);
        if (!more) tc_commit(bar_base + 8 * NS);
      }
    }
    mbar_wait(bar_base + 8 * NS, 0);
    tc_fence_after();
  }

  // ---- epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (= tile rows) and column half w/4
  float* outp = p.partial ? p.partial + (int64_t)blockIdx.z * p.n_out * p.Cout : p.Y;
  const int ldo = p.partial ? p.Cout : p.ldy;
  const float* bias = p.partial ? nullptr : p.bias;
  const int q = warp & 3, half = warp >> 2;
  const int64_t row = row0 + q * 32 + lane;
  constexpr int HALF = BN / 2;
#pragma unroll
  for (int c0 = 0; c0 < HALF; c0 += 16) {
    const int col = half * HALF + c0;
    uint32_t r[16];
    if (it1 > it0) {
      tc_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
      tc_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = 0u;
    }
    if (row < p.n_out) {
      float* dst = outp + row * ldo + n0 + col;
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        float4 o = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
        if (bias) { o.x += bias[n0 + col + e]; o.y += bias[n0 + col + e + 1]; o.z += bias[n0 + col + e + 2]; o.w += bias[n0 + col + e + 3]; }
        if (p.accumulate && !p.partial) {
          float4 old = *reinterpret_cast<const float4*>(dst + e);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(dst + e) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_acc), "r"(S::TMEM_COLS));
  }
}


// ------------------------------------------------------------------------------------------------ split-operand forward / data gradient
// Warp-specialised: warps 0-7 are producers (the gathered hi/lo row segments: 128-bit loads three stages ahead in registers -> 128-bit
// st.shared; thread 0 also launches the stage's weight tile as one TMA bulk copy), warp 8 issues the tcgen05.mma's.  Hand-off through
// mbarriers only:
//   full[s]  : one arrival per producer warp (release, after its stores of stage s) + the weight tile's transaction bytes
//   empty[s] : tcgen05.commit of the MMAs that read stage s
// The generic -> async proxy fence of a stage is issued by the MMA thread after its acquire (consumer_fence() below).  Two CTAs per SM
// with a 3-slot ring each, accumulators in TMEM, same epilogue as above.

template <int BN, int DNS>
struct DSmem {
  static constexpr int B_SBO = 128;
  static constexpr int B_LBO = (BN / 8) * 128 + 16;
  static constexpr int B_PLANE = (BK / 8) * B_LBO;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int IDX_OFF = DNS * STAGE;
  static constexpr int META_OFF = IDX_OFF + PCB_MAX_KERNEL_VOLUME * BM * 4;
  static constexpr int BAR_OFF = META_OFF + 72 * 4;
  static constexpr int TOTAL = BAR_OFF + (2 * DNS + 1) * 8 + 16;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
};

// <DNS, DPROD, CTAS>: <6, 512, 1> = one CTA per SM with a 6-slot ring; <3, 256, 2> = two CTAs per SM with 3 slots each --
// every hand-off (mbarrier wake-up ~260 cycles, smem store -> fence -> arrive, MMA issue) is a serial chain inside a CTA,
// so two co-resident CTAs hide each other's chains.
template <int BN, int DNS, int DPROD, int CTAS, int PF = 3>
__global__ void __launch_bounds__(DPROD + 32, CTAS) conv_tcgen05_split_kernel(const Args p) {
  using S = DSmem<BN, DNS>;
  constexpr int DTHR = DPROD + 32;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  int* s_idx = reinterpret_cast<int*>(smem + S::IDX_OFF);
  int* s_flag = reinterpret_cast<int*>(smem + S::META_OFF);
  int* s_klist = s_flag + 32;
  int* s_nk = s_klist + 32;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_nk + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full_bar = smem_base + S::BAR_OFF, empty_bar = full_bar + 8 * DNS, done_bar = empty_bar + 8 * DNS;

  if (tid == 0) {
    for (int i = 0; i < DNS; ++i) { mbar_init(full_bar + 8 * i, DPROD / 32 + 1); mbar_init(empty_bar + 8 * i, 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(S::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  pdl_wait(); pdl_trigger();        // barriers + TMEM are set up while the preceding kernel drains; no global memory touched before this
  {
    // this tile's slice of the neighbour table -> shared memory.  All loads of a thread are issued before the first store: the
    // table is read once per tile straight from DRAM / L2, and a load-store loop would pay that latency FILL times in a row.
    constexpr int FILL = (PCB_MAX_KERNEL_VOLUME * BM + DTHR - 1) / DTHR;
    int vals[FILL];
#pragma unroll
    for (int f = 0; f < FILL; ++f) {
      const int e = tid + f * DTHR;
      int v = -1;
      if (e < p.K * BM) {
        const int k = e / BM, r = e - k * BM;
        const int64_t row = row0 + r;
        if (row < p.n_out) v = __ldg(p.tbl + (int64_t)p.kmap[k] * p.tbl_stride + row);
      }
      vals[f] = v;
    }
#pragma unroll
    for (int f = 0; f < FILL; ++f) {
      const int e = tid + f * DTHR;
      if (e < p.K * BM) s_idx[e] = vals[f];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *s_tmem;
  for (int k = warp; k < p.K; k += DTHR / 32) {
    unsigned any = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) any |= __ballot_sync(0xffffffffu, s_idx[k * BM + s * 32 + lane] >= 0);
    if (lane == 0) s_flag[k] = any ? 1 : 0;
  }
  __syncthreads();
  if (warp == 0) {                  // offsets with at least one neighbour in this tile, in order (K <= 27 < 32: one ballot)
    static_assert(PCB_MAX_KERNEL_VOLUME <= 32, "one ballot per tile");
    const int f = lane < p.K ? s_flag[lane] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, f != 0);
    if (f) s_klist[__popc(m & ((1u << lane) - 1u))] = lane;
    if (lane == 0) *s_nk = __popc(m);
  }
  __syncthreads();
  const int nk = *s_nk;
  const int nkc = p.Cin / BK;
  const int T = nk * nkc;
  const int it0 = (int)((int64_t)T * blockIdx.z / gridDim.z);
  const int it1 = (int)((int64_t)T * (blockIdx.z + 1) / gridDim.z);
  const int n_it = (p.debug & 16) ? 0 : it1 - it0;
  const uint32_t IDESC = (1u << 4) | p.fmt_bits | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  if (warp < DPROD / 32) {
    // ===== producers.  Thread t always owns the same two 16-byte chunks of the A tile (row t>>2, chunk t&3, hi + lo
    // plane): 128-bit global loads into registers PF stages ahead, 128-bit shared stores when the slot is free.
    // The weight tile of a stage is ONE TMA bulk copy (the weights are pre-tiled as shared-memory images), issued by
    // thread 0 and completed on the same "full" barrier through its transaction count.
    constexpr int RP = BM * 4 / DPROD;                  // (row, chunk) pairs per thread: rows ar + j * (DPROD / 4)
    constexpr uint32_t BLOB = 2 * S::B_PLANE;
    const int ar = tid >> 2, ak8 = tid & 3;
    uint32_t a_dst[RP];
#pragma unroll
    for (int j = 0; j < RP; ++j) { const int r = ar + j * (DPROD / 4); a_dst[j] = ak8 * A_LBO + (r >> 3) * A_SBO + (r & 7) * 16; }
    const int nblk = p.Cout / BN;
    // load cursor (runs PF stages ahead of the store cursor)
    int l_kq = it0 / nkc, l_kc = it0 - l_kq * nkc, loaded = 0;
    const __nv_bfloat16* l_hi[RP]; const __nv_bfloat16* l_lo[RP]; bool l_on[RP];
    auto set_k = [&]() {
      const int kb = s_klist[l_kq] * BM;
#pragma unroll
      for (int j = 0; j < RP; ++j) {
        const int idx = s_idx[kb + ar + j * (DPROD / 4)];
        l_on[j] = idx >= 0;
        const int64_t off = (int64_t)(l_on[j] ? idx : 0) * p.lds + ak8 * 8;
        l_hi[j] = p.Xhi + off; l_lo[j] = p.Xlo + off;
      }
    };
#pragma unroll
    for (int j = 0; j < RP; ++j) { l_hi[j] = p.Xhi; l_lo[j] = p.Xlo; l_on[j] = false; }
    if (n_it > 0) set_k();
    struct Regs { uint4 h[RP]; uint4 l[RP]; };
    auto load = [&](Regs& v) {
      if (loaded < n_it) {
#pragma unroll
        for (int j = 0; j < RP; ++j) {
          if (l_on[j] && !(p.debug & 1)) {
            v.h[j] = __ldg(reinterpret_cast<const uint4*>(l_hi[j] + l_kc * BK));
            v.l[j] = __ldg(reinterpret_cast<const uint4*>(l_lo[j] + l_kc * BK));
          } else {
            v.h[j] = make_uint4(0, 0, 0, 0); v.l[j] = make_uint4(0, 0, 0, 0);
          }
        }
        ++loaded;
        if (++l_kc == nkc) { l_kc = 0; ++l_kq; if (loaded < n_it) set_k(); }
      }
    };
    // store cursor
    int is = 0, iround = 0, s_kq = l_kq, s_kc = l_kc;
    auto store = [&](const Regs& v) {
      if (iround >= 1) {                 // slot reuse: one poller per warp waits for the MMAs that read it
        if (lane == 0) mbar_wait(empty_bar + 8 * is, (iround - 1) & 1);
        __syncwarp();
      }
      const uint32_t sb = smem_base + is * S::STAGE;
      if (tid == 0) {
        const int k = s_klist[s_kq];
        mbar_arrive_expect_tx(full_bar + 8 * is, (p.debug & 2) ? 0u : BLOB);
        if (!(p.debug & 2))
          tma_bulk_load(sb + 2 * A_PLANE, p.wt + ((int64_t)(k * nkc + s_kc) * nblk + blockIdx.y) * BLOB, BLOB, full_bar + 8 * is);
      }
#pragma unroll
      for (int j = 0; j < RP; ++j) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(sb + a_dst[j]), "r"(v.h[j].x), "r"(v.h[j].y), "r"(v.h[j].z), "r"(v.h[j].w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(sb + A_PLANE + a_dst[j]), "r"(v.l[j].x), "r"(v.l[j].y), "r"(v.l[j].z), "r"(v.l[j].w) : "memory");
      }
      if (!(p.debug & 8) && !p.consumer_fence) fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + 8 * is);
      if (++is == DNS) { is = 0; ++iround; }
      if (++s_kc == nkc) { s_kc = 0; ++s_kq; }
    };
    static_assert(PF == 3, "three register stages");
    Regs v0, v1, v2;
    load(v0); load(v1); load(v2);
    for (int i = 0; i < n_it; i += PF) {
      store(v0); load(v0);
      if (i + 1 < n_it) { store(v1); load(v1); }
      if (i + 2 < n_it) { store(v2); load(v2); }
    }
  } else if (warp == DPROD / 32 && lane == 0) {
    // ===== MMA issuer =====
    int s = 0, par = 0;
    for (int i = 0; i < n_it; ++i) {
      mbar_wait(full_bar + 8 * s, par);
      if (p.consumer_fence) fence_proxy_async();
      tc_fence_after();
      const uint32_t a_hi = smem_base + s * S::STAGE, a_lo = a_hi + A_PLANE;
      const uint32_t b_hi = a_hi + 2 * A_PLANE, b_lo = b_hi + S::B_PLANE;
#pragma unroll
      for (int j = 0; j < BK / 16; ++j) {
        if (p.debug & 4) break;
        const uint64_t dah = make_desc(a_hi + j * 2 * A_LBO, A_LBO, A_SBO), dal = make_desc(a_lo + j * 2 * A_LBO, A_LBO, A_SBO);
        const uint64_t dbh = make_desc(b_hi + j * 2 * S::B_LBO, S::B_LBO, S::B_SBO);
        const uint64_t dbl = make_desc(b_lo + j * 2 * S::B_LBO, S::B_LBO, S::B_SBO);
        tc_mma(tmem_acc, dal, dbh, IDESC, (i > 0 || j > 0) ? 1u : 0u);
        tc_mma(tmem_acc, dah, dbl, IDESC, 1u);
        tc_mma(tmem_acc, dah, dbh, IDESC, 1u);
      }
      tc_commit(empty_bar + 8 * s);
      if (i == n_it - 1) tc_commit(done_bar);
      if (++s == DNS) { s = 0; par ^= 1; }
    }
  }
  if (n_it > 0) {
    mbar_wait(done_bar, 0);
    tc_fence_after();
  }
  if (warp < 8) {
    float* outp = p.partial ? p.partial + (int64_t)blockIdx.z * p.n_out * p.Cout : p.Y;
    const int ldo = p.partial ? p.Cout : p.ldy;
    const float* bias = p.partial ? nullptr : p.bias;
    const int q = warp & 3, half = warp >> 2;
    const int64_t row = row0 + q * 32 + lane;
    constexpr int HALF = BN / 2;
#pragma unroll
    for (int c0 = 0; c0 < HALF; c0 += 16) {
      const int col = half * HALF + c0;
      uint32_t r[16];
      if (n_it > 0) {
        tc_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
        tc_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0u;
      }
      if (p.out_scale != 1.0f) {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * p.out_scale);
      }
      if (row < p.n_out) {
        float* dst = outp + row * ldo + n0 + col;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float4 o = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
          if (bias) { o.x += bias[n0 + col + e]; o.y += bias[n0 + col + e + 1]; o.z += bias[n0 + col + e + 2]; o.w += bias[n0 + col + e + 3]; }
          if (p.accumulate && !p.partial) {
            float4 old = *reinterpret_cast<const float4*>(dst + e);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *reinterpret_cast<float4*>(dst + e) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_acc), "r"(S::TMEM_COLS));
}

template <int BN, int DNS, int DPROD, int CTAS, int PF = 3>
int launch_split_cfg(const Args& a, int nsplit, cudaStream_t st) {
  using S = DSmem<BN, DNS>;
  static bool attr_set[64] = {};          // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    PCB_CUDA(cudaFuncSetAttribute(conv_tcgen05_split_kernel<BN, DNS, DPROD, CTAS, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set[dev_] = true;
  }
  dim3 grid((unsigned)((a.n_out + BM - 1) / BM), a.Cout / BN, nsplit);
  launch_kernel(conv_tcgen05_split_kernel<BN, DNS, DPROD, CTAS, PF>, grid, DPROD + 32, S::TOTAL, st, a);
  return check_launch("conv_tcgen05_split_kernel");
}

// One configuration is built: 3-slot ring, 8 producer warps, two CTAs per SM, gathers three stages ahead.  Measured and dropped in round 2
// (profiles/r2_results.md): one CTA per SM with a 6-slot ring and 16 producer warps, three CTAs per SM with 2-slot rings, gathers four
// stages ahead, 256-row tiles sharing each weight tile, a dedicated weight-loader warp, cp.async (LDGSTS) producers.
template <int BN>
int launch_split(const Args& a, int nsplit, cudaStream_t st) {
  return launch_split_cfg<BN, 3, 256, 2>(a, nsplit, st);
}

template <int BN, bool SPLIT>
int launch2(const Args& a, int nsplit, cudaStream_t st) {
  using S = Smem<BN>;
  static bool attr_set[64] = {};          // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    PCB_CUDA(cudaFuncSetAttribute(conv_tcgen05_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set[dev_] = true;
  }
  dim3 grid((unsigned)((a.n_out + BM - 1) / BM), a.Cout / BN, nsplit);
  conv_tcgen05_kernel<BN, SPLIT><<<grid, NTHR, S::TOTAL, st>>>(a);
  return check_launch(SPLIT ? "conv_tcgen05_kernel<split>" : "conv_tcgen05_kernel");
}

template <int BN>
int launch(const Args& a, int nsplit, cudaStream_t st) {
  return a.Xhi ? launch_split<BN>(a, nsplit, st) : launch2<BN, false>(a, nsplit, st);
}

}  // namespace tc5

// Called by pcb_conv_forward (conv.cu).  wk_*: K-major split weights [K][Cout][Cin] for this call's roles.
int launch_conv_tcgen05(const float* X, int ldx, const uint16_t* Xhi, const uint16_t* Xlo, int lds, const void* wt, const int32_t* tbl,
                        int64_t tbl_stride, const int* kmap, int K, int64_t n_out,
                        int Cin, int Cout, const uint16_t* wk_hi, const uint16_t* wk_lo, const float* bias, float* Y, int ldy,
                        float* partial, int nsplit, int bn, int accumulate, cudaStream_t st, int x_fp16, int w_fp16) {
  tc5::Args a;
  a.accumulate = accumulate;
  a.fmt_bits = ((x_fp16 ? 0u : 1u) << 7) | ((w_fp16 ? 0u : 1u) << 10);
  a.out_scale = w_fp16 ? 1.0f / 1024.0f : 1.0f;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("PCB_TC5_DEBUG"); dbg = e ? atoi(e) : 0; }
  a.debug = dbg;
  a.consumer_fence = tc5::consumer_fence();
  a.Xhi = (const __nv_bfloat16*)Xhi; a.Xlo = (const __nv_bfloat16*)Xlo; a.lds = lds;
  a.wt = (const unsigned char*)wt;
  a.X = X; a.ldx = ldx; a.tbl = tbl; a.tbl_stride = tbl_stride; a.K = K; a.n_out = n_out; a.Cin = Cin; a.Cout = Cout;
  for (int k = 0; k < K; ++k) a.kmap[k] = kmap[k];
  a.wk_hi = (const __nv_bfloat16*)wk_hi; a.wk_lo = (const __nv_bfloat16*)wk_lo; a.bias = bias; a.Y = Y; a.ldy = ldy;
  a.partial = partial;
  switch (bn) {
    case 128: return tc5::launch<128>(a, nsplit, st);
    case 96: return tc5::launch<96>(a, nsplit, st);
    case 64: return tc5::launch<64>(a, nsplit, st);
    default: return tc5::launch<32>(a, nsplit, st);
  }
}

// Offsets per weight-gradient CTA: 2 (two CTAs per SM; default) or 4 (PCB_WG_GROUP=4: one CTA per SM, the row-aligned tile staged once
// per 4 offsets).  Measured (profiles/r2_results.md): 464 -> 415 us on the block8 shape, 478 -> 382 us at 128 channels, 8.05 -> 7.0 ms per
// step.  conv.cu sizes the row splits and the partial buffer with it.
int wgrad_group() {
  static int v = 0;
  if (!v) { const char* e = getenv("PCB_WG_GROUP"); v = (e && atoi(e) == 4) ? 4 : 2; }
  return v;
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[k] (Ca x Cb) = sum_j A[tbl[k][j], :]^T . B[j, :] on split (bf16 hi/lo) operands.
// UMMA view: D_k[M = Ca-block (padded to 128)][N = Cb-block] += A_k[M x 16 rows] . B[16 rows x N]; both operands MN-major
// (a matrix row is contiguous along channels): core matrix = 8 rows (K) x 16 B (8 channels), channel-chunk stride SBO = 144 B,
// 8-row-group stride LBO.
// A CTA owns a GROUP of GK kernel offsets (2 by default, 4 optional) and a range of table rows: the row-aligned operand B is staged ONCE
// per 16-row step and shared by the GK gathered operands A_k, each accumulating into its own TMEM accumulator (GK x TN columns).
// Producer warps (thread = one row, one 16-byte channel chunk, two offsets, both planes; table entries prefetched three steps ahead of
// the 128-bit data loads, which are two steps ahead of the 128-bit shared stores) + 1 MMA-issuer warp; GK = 2: 8 producer warps, two
// CTAs per SM; GK = 4: 16 producer warps, one CTA per SM.
// grid: x = groups * mblocks * nblocks, y = row splits; partial tiles are reduced by wgrad_reduce_kernel (conv.cu).
namespace wg {

constexpr int WM = 128, WK = 16;
// <GK, WPROD, CTAS>: <4, 512, 1> = one CTA per SM, 4 offsets share every staged row-aligned tile (all 512 TMEM columns);
// <2, 256, 2> = two CTAs per SM with 2 offsets each: the per-step chain (table -> row loads -> st.shared -> fence -> arrive -> MMA ->
// commit) is serial inside a CTA, and two co-resident CTAs hide each other's (the same trade the forward kernel makes).

struct Args {
  const __nv_bfloat16* Ahi; const __nv_bfloat16* Alo; int lda;      // gathered operand (elements)
  const __nv_bfloat16* Bhi; const __nv_bfloat16* Blo; int ldb;      // row-aligned operand
  const int32_t* tbl; int64_t tbl_stride;
  int K; int64_t n_out; int Ca; int Cb; int rows_per_split;
  float* partial; int transpose_out;
  int ns;               // ring depth (host-chosen to fill shared memory)
  uint32_t fmt_bits;    // instruction-descriptor formats: bits 7-9 gathered operand, bits 10-12 row-aligned operand (0 fp16, 1 bf16)
  int consumer_fence;   // tc5::wgrad_consumer_fence()
};

// channel-chunk (core-matrix) stride 144 B, not 128: a producer warp writes the 16 chunks of ONE row, and a 128-byte stride
// would put all 16 stores on the same shared-memory banks (a 16..32-way conflict on every 128-bit store)
constexpr int SBO = 144;
__host__ __device__ inline int a_lbo(int mrows) { return (mrows / 8) * SBO + 16; }
__host__ __device__ inline int b_lbo(int tn) { return (tn / 8) * SBO + 16; }
__host__ __device__ inline int stage_bytes(int mrows, int tn, int gk) { return 4 * b_lbo(tn) + gk * 4 * a_lbo(mrows); }

template <int TN, int GK, int WPROD, int CTAS>
__global__ void __launch_bounds__(WPROD + 32, CTAS) wgrad_tcgen05_kernel(const Args p) {
  using namespace tc5;
  static_assert((GK == 4 && WPROD == 512) || (GK == 2 && WPROD == 256), "thread = (row, 16-byte chunk) of two offsets");
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mblocks = (p.Ca + WM - 1) / WM, nblocks = p.Cb / TN;
  int bx = blockIdx.x;
  const int nb = bx % nblocks; bx /= nblocks;
  const int mb = bx % mblocks; bx /= mblocks;
  const int k0 = bx * GK;
  const int nk = min(GK, p.K - k0);
  const int m0 = mb * WM, n0 = nb * TN;
  const int mrows = min(WM, p.Ca - m0);                 // valid M rows of this block (multiple of 32)
  const int ach = mrows / 8;
  constexpr int BCH = TN / 8;
  const int A_LBO = a_lbo(mrows), B_LBO = b_lbo(TN);
  const int STAGE = stage_bytes(mrows, TN, GK);
  const int NS = p.ns;
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_split;
  const int64_t r_end = min(p.n_out, r_begin + p.rows_per_split);
  const int nsteps = r_end > r_begin ? (int)((r_end - r_begin + WK - 1) / WK) : 0;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full_bar = smem_base + NS * STAGE + 4096, empty_bar = full_bar + 8 * NS, done_bar = empty_bar + 8 * NS;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + NS * STAGE + 4096 + (2 * NS + 1) * 8);
  constexpr int ACC_STRIDE = TN <= 32 ? 32 : (TN <= 64 ? 64 : 128);
  constexpr int TMEM_COLS = GK * ACC_STRIDE;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full_bar + 8 * i, WPROD / 32); mbar_init(empty_bar + 8 * i, 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  pdl_wait(); pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *s_tmem;
  // MN-major A and B (bits 15, 16), bf16 inputs, fp32 accumulate, M = 128, N = TN.  The A descriptor spans 16 channel
  // chunks although only `ach` are staged: rows >= mrows of D are garbage and never read back.
  const uint32_t IDESC = (1u << 4) | p.fmt_bits | (1u << 15) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);

  if (warp < WPROD / 32) {
    constexpr int PF = 2;
    // row of the step, offset pair (2kh, 2kh+1), 16-byte channel chunk
    const int r = GK == 4 ? tid >> 5 : tid >> 4, kh = GK == 4 ? (tid >> 4) & 1 : 0, ch = tid & 15;
    const bool a_on = ch < ach, b_on = (ch < BCH) && kh == 0;
    const uint32_t a_dst = kh * 8 * A_LBO + (r >> 3) * A_LBO + ch * SBO + (r & 7) * 16;   // A block of offset 2kh, hi plane
    const uint32_t b_dst = (r >> 3) * B_LBO + ch * SBO + (r & 7) * 16;
    const int64_t a_col = m0 + ch * 8, b_col = n0 + ch * 8;
    const int32_t* trow = p.tbl + (int64_t)(k0 + 2 * kh) * p.tbl_stride;
    const int32_t* trow_other = p.tbl + (int64_t)(k0 + 2 * (1 - kh)) * p.tbl_stride;
    const bool g0 = 2 * kh < nk, g1 = 2 * kh + 1 < nk, o0 = GK == 4 && 2 * (1 - kh) < nk, o1 = GK == 4 && 2 * (1 - kh) + 1 < nk;
    // Table entries are fetched TF steps ahead of the data loads that depend on them (a register ring, shifted once per step),
    // the data loads PF steps ahead of the shared-memory stores: neither dependent global-load latency (table -> row segment ->
    // st.shared, ~900 cycles each under load) sits on the per-step critical path.  (ncu, round 2: with the table only ONE step ahead
    // the producers spent 60 % of their samples in long-scoreboard stalls on exactly these two hops and a step took ~1800 cycles.)
    constexpr int TF = 3;
    int tq[TF][4] = {};        // [steps loaded .. loaded + TF - 1][own offset pair (2) + (B loaders) the other pair (2)]
    int loaded = 0;
    auto fetch_into = [&](int (&d)[4], int step) {          // independent loads only: nothing here waits on a previous load
      const int64_t row = r_begin + (int64_t)step * WK + r;
      const bool live = step < nsteps && row < r_end;
      d[0] = (live && g0) ? __ldg(trow + row) : -1;
      d[1] = (live && g1) ? __ldg(trow + p.tbl_stride + row) : -1;
      d[2] = (live && b_on && o0) ? __ldg(trow_other + row) : -1;
      d[3] = (live && b_on && o1) ? __ldg(trow_other + p.tbl_stride + row) : -1;
    };
    struct Regs { uint4 h0, l0, h1, l1, hb, lb; };
    const uint4 Z = make_uint4(0, 0, 0, 0);
    auto ldrow = [&](const __nv_bfloat16* base, int64_t off, bool on) {
      return on ? __ldg(reinterpret_cast<const uint4*>(base + off)) : Z;
    };
    auto load = [&](Regs& v) {
      if (loaded < nsteps) {
        const int64_t row = r_begin + (int64_t)loaded * WK + r;
        const int c0 = tq[0][0], c1 = tq[0][1];
        const bool any = (c0 >= 0) | (c1 >= 0) | (tq[0][2] >= 0) | (tq[0][3] >= 0);      // the B row is needed if ANY of the four offsets has a neighbour
#pragma unroll
        for (int f = 0; f + 1 < TF; ++f) {
#pragma unroll
          for (int e = 0; e < 4; ++e) tq[f][e] = tq[f + 1][e];
        }
        fetch_into(tq[TF - 1], loaded + TF);
        const int64_t q0 = (int64_t)c0 * p.lda + a_col, q1 = (int64_t)c1 * p.lda + a_col;
        v.h0 = ldrow(p.Ahi, q0, a_on && c0 >= 0); v.l0 = ldrow(p.Alo, q0, a_on && c0 >= 0);
        v.h1 = ldrow(p.Ahi, q1, a_on && c1 >= 0); v.l1 = ldrow(p.Alo, q1, a_on && c1 >= 0);
        v.hb = ldrow(p.Bhi, row * p.ldb + b_col, b_on && any); v.lb = ldrow(p.Blo, row * p.ldb + b_col, b_on && any);
        ++loaded;
      }
    };
    int is = 0, iround = 0;
    auto sts = [&](uint32_t addr, const uint4& x) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
    };
    auto store = [&](const Regs& v) {
      if (iround >= 1) {
        if (lane == 0) mbar_wait(empty_bar + 8 * is, (iround - 1) & 1);
        __syncwarp();
      }
      const uint32_t sb = smem_base + is * STAGE;
      if (b_on) { sts(sb + b_dst, v.hb); sts(sb + 2 * B_LBO + b_dst, v.lb); }
      if (a_on) {
        const uint32_t ab = sb + 4 * B_LBO + a_dst;
        sts(ab, v.h0); sts(ab + 2 * A_LBO, v.l0);
        sts(ab + 4 * A_LBO, v.h1); sts(ab + 6 * A_LBO, v.l1);
      }
      if (!p.consumer_fence) fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + 8 * is);
      if (++is == NS) { is = 0; ++iround; }
    };
    Regs v0, v1;
#pragma unroll
    for (int f = 0; f < TF; ++f) fetch_into(tq[f], f);
    load(v0); load(v1);
    for (int i = 0; i < nsteps; i += PF) {
      store(v0); load(v0);
      if (i + 1 < nsteps) { store(v1); load(v1); }
    }
  } else if (lane == 0) {
    int s = 0, par = 0;
    for (int i = 0; i < nsteps; ++i) {
      mbar_wait(full_bar + 8 * s, par);
      if (p.consumer_fence) fence_proxy_async();
      tc_fence_after();
      const uint32_t sb = smem_base + s * STAGE;
      const uint64_t dbh = make_desc(sb, B_LBO, SBO), dbl = make_desc(sb + 2 * B_LBO, B_LBO, SBO);
      for (int g = 0; g < nk; ++g) {
        const uint32_t ab = sb + 4 * B_LBO + g * 4 * A_LBO;
        const uint64_t dah = make_desc(ab, A_LBO, SBO), dal = make_desc(ab + 2 * A_LBO, A_LBO, SBO);
        const uint32_t acc = tmem_acc + g * ACC_STRIDE;
        tc_mma(acc, dal, dbh, IDESC, i > 0 ? 1u : 0u);
        tc_mma(acc, dah, dbl, IDESC, 1u);
        tc_mma(acc, dah, dbh, IDESC, 1u);
      }
      tc_commit(empty_bar + 8 * s);
      if (i == nsteps - 1) tc_commit(done_bar);
      if (++s == NS) { s = 0; par ^= 1; }
    }
  }
  if (nsteps > 0) {
    mbar_wait(done_bar, 0);
    tc_fence_after();
  }

  // ---- epilogue: accumulator g, row m = TMEM lane, column n -> partial tile of offset k0 + g
  if (warp < 8) {
    const int q = warp & 3, half = warp >> 2;
    const int m = q * 32 + lane;
    constexpr int HALF = TN / 2;
    for (int g = 0; g < nk; ++g) {
      float* out = p.partial + ((int64_t)blockIdx.y * p.K + k0 + g) * (int64_t)p.Ca * p.Cb;
#pragma unroll 1
      for (int c0 = 0; c0 < HALF; c0 += 16) {
        const int col = half * HALF + c0;
        uint32_t r[16];
        if (nsteps > 0) {
          tc_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * ACC_STRIDE + col), r);
          tc_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = 0u;
        }
        if (m < mrows) {
          if (!p.transpose_out) {
            float* dst = out + (int64_t)(m0 + m) * p.Cb + n0 + col;
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]),
                                                                __uint_as_float(r[e + 3]));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) out[(int64_t)(n0 + col + e) * p.Ca + m0 + m] = __uint_as_float(r[e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_acc), "r"(TMEM_COLS));
}

template <int TN, int GK, int WPROD, int CTAS>
int launch_cfg(Args a, int splits, cudaStream_t st) {
  static bool attr_set[64] = {};          // per device: the opt-in is a per-device function attribute
  constexpr int MAX_SMEM = CTAS == 1 ? 220 * 1024 : 112 * 1024;
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    PCB_CUDA(cudaFuncSetAttribute(wgrad_tcgen05_kernel<TN, GK, WPROD, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_set[dev_] = true;
  }
  const int mrows_max = a.Ca < WM ? a.Ca : WM;
  const int stage = stage_bytes(mrows_max, TN, GK);
  int ns = (MAX_SMEM - 6144) / stage;
  if (ns > 8) ns = 8;
  if (ns < 2) { set_error("wgrad_tcgen05: stage does not fit"); return PCB_ERR_ARG; }
  a.ns = ns;
  const int groups = (a.K + GK - 1) / GK;
  dim3 grid((unsigned)(groups * ((a.Ca + WM - 1) / WM) * (a.Cb / TN)), splits);
  // + 4 KB: the M = 128 descriptor of a 96-channel block reads (and ignores) a few hundred bytes past the last staged chunk
  launch_kernel(wgrad_tcgen05_kernel<TN, GK, WPROD, CTAS>, grid, WPROD + 32, (size_t)(ns * stage + (2 * ns + 1) * 8 + 64 + 4096), st, a);
  return check_launch("wgrad_tcgen05_kernel");
}

template <int TN>
int launch(Args a, int splits, cudaStream_t st) {
  return wgrad_group() == 2 ? launch_cfg<TN, 2, 256, 2>(a, splits, st) : launch_cfg<TN, 4, 512, 1>(a, splits, st);
}

}  // namespace wg

int launch_wgrad_tcgen05(const uint16_t* Ahi, const uint16_t* Alo, int lda, const uint16_t* Bhi, const uint16_t* Blo, int ldb,
                         const int32_t* tbl, int64_t tbl_stride, int K, int64_t n_out, int Ca, int Cb, int rows_per_split, int splits,
                         float* partial, int transpose_out, int tn, cudaStream_t st, int a_fp16, int b_fp16) {
  wg::Args a;
  a.fmt_bits = ((a_fp16 ? 0u : 1u) << 7) | ((b_fp16 ? 0u : 1u) << 10);
  a.Ahi = (const __nv_bfloat16*)Ahi; a.Alo = (const __nv_bfloat16*)Alo; a.lda = lda;
  a.Bhi = (const __nv_bfloat16*)Bhi; a.Blo = (const __nv_bfloat16*)Blo; a.ldb = ldb;
  a.tbl = tbl; a.tbl_stride = tbl_stride; a.K = K; a.n_out = n_out; a.Ca = Ca; a.Cb = Cb; a.rows_per_split = rows_per_split;
  a.partial = partial; a.transpose_out = transpose_out;
  a.consumer_fence = tc5::wgrad_consumer_fence();
  switch (tn) {
    case 128: return wg::launch<128>(a, splits, st);
    case 96: return wg::launch<96>(a, splits, st);
    case 64: return wg::launch<64>(a, splits, st);
    default: return wg::launch<32>(a, splits, st);
  }
}

}  // namespace pcb
