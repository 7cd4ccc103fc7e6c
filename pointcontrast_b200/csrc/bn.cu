// BatchNorm over the rows of a sparse tensor's feature matrix (training-mode statistics), with the
// affine-normalise (+ residual, + ReLU) fused into one elementwise pass.  HBM-bound: every kernel moves
// 16-byte vectors with consecutive threads on consecutive channels.
// Replaces MinkowskiBatchNorm == torch.nn.BatchNorm1d on .F (reference call sites in include/pcb200.h).
#include "common.cuh"

using namespace pcb;

namespace {

constexpr int ROWS_PER_CHUNK = 128;

// fp32 x4 -> bf16 hi x4 + bf16 lo x4 (x ~= hi + lo to 2^-17): the operand format of the tensor-core conv kernels
__device__ __forceinline__ void store_split4(const float4& v, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 h1 = __floats2bfloat162_rn(v.z, v.w);
  float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
  __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
  uint2 H, L;
  H.x = *reinterpret_cast<uint32_t*>(&h0); H.y = *reinterpret_cast<uint32_t*>(&h1);
  L.x = *reinterpret_cast<uint32_t*>(&l0); L.y = *reinterpret_cast<uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(hi) = H;
  *reinterpret_cast<uint2*>(lo) = L;
}

__global__ void split_rows_kernel(const float* __restrict__ X, int ldx, int64_t n4, int cv, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int lds) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  int c4 = (int)(i % cv);
  const int64_t row = i / cv;
  float4 x = __ldg(reinterpret_cast<const float4*>(X + row * ldx) + c4);
  store_split4(x, hi + row * lds + c4 * 4, lo + row * lds + c4 * 4);
}

// partial[chunk][0][C] = sum(a), partial[chunk][1][C] = sum(a*b)    (b == a for the forward statistics)
// block: (C/4) channel-vectors x RP row lanes; grid: one CTA per chunk of rows.
// Strides: lda / ldb / ldm (floats).  Mask (backward only): a is zeroed where mask <= 0 (ReLU folded into the BN backward).
struct Finalize {            // what the LAST CTA of a column reduction does with the per-chunk partial sums
  unsigned int* counter;     // zero before the launch; reset by the last CTA
  int mode;                  // 1: forward statistics, 2: backward sums
  int64_t n; float eps, momentum;
  float* mean; float* invstd; float* running_mean; float* running_var;      // mode 1
  float* dgamma; float* dbeta; int accumulate; float* sums;                  // mode 2
};

__device__ __forceinline__ void warp_sum2(const float* __restrict__ partial, int chunks, int C, int c, double& s1, double& s2);

template <bool TWO_INPUTS>
__global__ void colsum_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                              const float* __restrict__ Mask, int ldm, int64_t n, int C,
                              const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ partial,
                              const Finalize fin) {
  extern __shared__ float sm[];      // [RP][2][C]
  const int cv = C / 4;
  const int rp = blockDim.x / cv;    // row lanes
  const int c4 = threadIdx.x % cv;
  const int rl = threadIdx.x / cv;
  const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_CHUNK;
  const int64_t r1 = min(n, r0 + ROWS_PER_CHUNK);
  float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
  float4 mu = make_float4(0, 0, 0, 0), is = make_float4(1, 1, 1, 1);
  if (TWO_INPUTS && rl < rp) { mu = reinterpret_cast<const float4*>(mean)[c4]; is = reinterpret_cast<const float4*>(invstd)[c4]; }
  if (rl < rp) {
    for (int64_t r = r0 + rl; r < r1; r += rp) {
      float4 a = __ldg(reinterpret_cast<const float4*>(A + r * lda) + c4);
      if (TWO_INPUTS) {
        if (Mask) {
          float4 m = __ldg(reinterpret_cast<const float4*>(Mask + r * ldm) + c4);
          a.x = m.x > 0.f ? a.x : 0.f; a.y = m.y > 0.f ? a.y : 0.f; a.z = m.z > 0.f ? a.z : 0.f; a.w = m.w > 0.f ? a.w : 0.f;
        }
        float4 x = __ldg(reinterpret_cast<const float4*>(Bm + r * ldb) + c4);
        // a = dY, second sum = dY * xhat
        s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
        s2.x += a.x * ((x.x - mu.x) * is.x); s2.y += a.y * ((x.y - mu.y) * is.y);
        s2.z += a.z * ((x.z - mu.z) * is.z); s2.w += a.w * ((x.w - mu.w) * is.w);
      } else {
        s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
        s2.x += a.x * a.x; s2.y += a.y * a.y; s2.z += a.z * a.z; s2.w += a.w * a.w;
      }
    }
    float* d = sm + (int64_t)rl * 2 * C;
    reinterpret_cast<float4*>(d)[c4] = s1;
    reinterpret_cast<float4*>(d + C)[c4] = s2;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * C; e += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < rp; ++l) s += sm[(int64_t)l * 2 * C + e];
    partial[(int64_t)blockIdx.x * 2 * C + e] = s;
  }
  // ---- the last CTA to finish reduces the partials (fixed order -> deterministic) and finalises: no second launch
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(fin.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int chunks = gridDim.x;
  const int nwarps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = w; c < C; c += nwarps) {
    double s1, s2;
    warp_sum2(partial, chunks, C, c, s1, s2);
    if (lane != 0) continue;
    if (fin.mode == 1) {
      double m = s1 / (double)fin.n;
      double var = s2 / (double)fin.n - m * m;
      if (var < 0.0) var = 0.0;
      fin.mean[c] = (float)m;
      fin.invstd[c] = (float)(1.0 / sqrt(var + (double)fin.eps));
      if (fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)m;
      if (fin.running_var) {
        double unb = fin.n > 1 ? var * ((double)fin.n / (double)(fin.n - 1)) : var;
        fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unb;
      }
    } else {
      fin.sums[c] = (float)s1;
      fin.sums[C + c] = (float)s2;
      if (fin.accumulate) { fin.dbeta[c] += (float)s1; fin.dgamma[c] += (float)s2; }
      else { fin.dbeta[c] = (float)s1; fin.dgamma[c] = (float)s2; }
    }
  }
  if (threadIdx.x == 0) *fin.counter = 0u;
}

// one warp per channel: lanes stride over the row chunks, fp64 shuffle reduction (fixed order: deterministic).
// The partials were written by other CTAs of the same launch: read them through L2 (volatile), not the read-only path.
__device__ __forceinline__ void warp_sum2(const float* __restrict__ partial, int chunks, int C, int c, double& s1, double& s2) {
  const int lane = threadIdx.x & 31;
  s1 = 0.0; s2 = 0.0;
  const volatile float* pv = partial;
  for (int k = lane; k < chunks; k += 32) { s1 += pv[(int64_t)k * 2 * C + c]; s2 += pv[(int64_t)k * 2 * C + C + c]; }
  for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
}

__global__ void bn_apply_kernel(const float* __restrict__ X, int ldx, int64_t n4, int cv, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ residual, int ldr, int relu, float* __restrict__ Y, int ldy,
                                __nv_bfloat16* __restrict__ Yhi, __nv_bfloat16* __restrict__ Ylo, int lds) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  int c4 = (int)(i % cv);
  const int64_t row = i / cv;
  float4 x = __ldg(reinterpret_cast<const float4*>(X + row * ldx) + c4);
  float4 mu = reinterpret_cast<const float4*>(mean)[c4], is = reinterpret_cast<const float4*>(invstd)[c4];
  float4 g = reinterpret_cast<const float4*>(gamma)[c4], b = reinterpret_cast<const float4*>(beta)[c4];
  float4 y;
  y.x = (x.x - mu.x) * is.x * g.x + b.x; y.y = (x.y - mu.y) * is.y * g.y + b.y;
  y.z = (x.z - mu.z) * is.z * g.z + b.z; y.w = (x.w - mu.w) * is.w * g.w + b.w;
  if (residual) {
    float4 r = __ldg(reinterpret_cast<const float4*>(residual + row * ldr) + c4);
    y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
  }
  if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
  *reinterpret_cast<float4*>(Y + row * ldy + c4 * 4) = y;
  if (Yhi) store_split4(y, Yhi + row * lds + c4 * 4, Ylo + row * lds + c4 * 4);
}

// gout_mode: 0 none, 1 write, 2 accumulate -- the (ReLU-masked) incoming gradient, i.e. the gradient of the residual input
__global__ void bn_bwd_apply_kernel(const float* dY, int lddy, const float* __restrict__ X, int ldx,
                                    const float* __restrict__ Mask, int ldm, int64_t n4, int cv, float inv_n,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, float* __restrict__ dX, int lddx, float* gout, int ldg,
                                    int gout_mode, __nv_bfloat16* __restrict__ dXhi, __nv_bfloat16* __restrict__ dXlo, int lds) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  int c4 = (int)(i % cv);
  const int64_t row = i / cv;
  float4 dy = *(reinterpret_cast<const float4*>(dY + row * lddy) + c4);
  if (Mask) {
    float4 m = __ldg(reinterpret_cast<const float4*>(Mask + row * ldm) + c4);
    dy.x = m.x > 0.f ? dy.x : 0.f; dy.y = m.y > 0.f ? dy.y : 0.f; dy.z = m.z > 0.f ? dy.z : 0.f; dy.w = m.w > 0.f ? dy.w : 0.f;
  }
  if (gout_mode) {
    float4* gp = reinterpret_cast<float4*>(gout + row * ldg) + c4;
    float4 gv = dy;
    if (gout_mode == 2) { float4 o = *gp; gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w; }
    *gp = gv;
  }
  float4 x = __ldg(reinterpret_cast<const float4*>(X + row * ldx) + c4);
  float4 mu = reinterpret_cast<const float4*>(mean)[c4], is = reinterpret_cast<const float4*>(invstd)[c4];
  float4 g = reinterpret_cast<const float4*>(gamma)[c4];
  float4 dg = reinterpret_cast<const float4*>(dgamma)[c4], db = reinterpret_cast<const float4*>(dbeta)[c4];
  float4 o;
  o.x = g.x * is.x * (dy.x - db.x * inv_n - (x.x - mu.x) * is.x * dg.x * inv_n);
  o.y = g.y * is.y * (dy.y - db.y * inv_n - (x.y - mu.y) * is.y * dg.y * inv_n);
  o.z = g.z * is.z * (dy.z - db.z * inv_n - (x.z - mu.z) * is.z * dg.z * inv_n);
  o.w = g.w * is.w * (dy.w - db.w * inv_n - (x.w - mu.w) * is.w * dg.w * inv_n);
  if (dX) *reinterpret_cast<float4*>(dX + row * lddx + c4 * 4) = o;
  if (dXhi) store_split4(o, dXhi + row * lds + c4 * 4, dXlo + row * lds + c4 * 4);
}

// 4 bytes of control state per device for the "last CTA finalises" pattern (zeroed once; each launch leaves it at zero).
// Launches that share it are ordered by the stream; concurrent use from several streams of one device is not supported.
inline unsigned int* counter_for(cudaStream_t) {
  static unsigned int* ctr[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ctr[dev]) {
    if (cudaMalloc(&ctr[dev], 256) != cudaSuccess) { set_error("counter allocation failed"); return nullptr; }
    cudaMemset(ctr[dev], 0, 256);
  }
  return ctr[dev];
}

inline int chunks_for(int64_t n) { return (int)((n + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK); }

inline int colsum_threads(int C) {       // (C/4) * row lanes, <= 256, at least one row lane
  int cv = C / 4;
  int rp = 256 / cv; if (rp < 1) rp = 1;
  return cv * rp;
}

}  // namespace

extern "C" size_t pcb_bn_ws_bytes(int64_t n, int C) {
  if (n < 1) n = 1;
  return (size_t)(chunks_for(n) + 1) * 2 * C * sizeof(float) + 256;
}

extern "C" int pcb_bn_stats2(const float* X, int ldx, int64_t n, int C, float eps, float momentum, float* mean, float* invstd,
                             float* running_mean, float* running_var, void* ws, size_t ws_bytes, void* stream) {
  PCB_ARG(X && mean && invstd && ws && n >= 1 && C >= 4 && C % 4 == 0 && C <= 1024 && ldx >= C && ldx % 4 == 0);
  PCB_ARG(ws_bytes >= pcb_bn_ws_bytes(n, C) - 256);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = chunks_for(n);
  const int thr = colsum_threads(C);
  const int rp = thr / (C / 4);
  Finalize fin{};
  fin.counter = counter_for(st); fin.mode = 1; fin.n = n; fin.eps = eps; fin.momentum = momentum;
  fin.mean = mean; fin.invstd = invstd; fin.running_mean = running_mean; fin.running_var = running_var;
  if (!fin.counter) return PCB_ERR_CUDA;
  colsum_kernel<false><<<chunks, thr, (size_t)rp * 2 * C * sizeof(float), st>>>(X, ldx, nullptr, 0, nullptr, 0, n, C, nullptr, nullptr,
                                                                                 (float*)ws, fin);
  return check_launch("colsum_kernel");
}

extern "C" int pcb_bn_stats(const float* X, int64_t n, int C, float eps, float momentum, float* mean, float* invstd,
                            float* running_mean, float* running_var, void* ws, size_t ws_bytes, void* stream) {
  return pcb_bn_stats2(X, C, n, C, eps, momentum, mean, invstd, running_mean, running_var, ws, ws_bytes, stream);
}

extern "C" int pcb_split_rows(const float* X, int ldx, int64_t n, int C, uint16_t* hi, uint16_t* lo, int lds, void* stream) {
  PCB_ARG(n >= 0 && C >= 4 && C % 4 == 0 && ldx >= C && ldx % 4 == 0 && lds >= C && lds % 4 == 0);
  if (n == 0) return PCB_OK;
  PCB_ARG(X && hi && lo);
  int64_t n4 = n * (C / 4);
  split_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, n4, C / 4, (__nv_bfloat16*)hi,
                                                                                   (__nv_bfloat16*)lo, lds);
  return check_launch("split_rows_kernel");
}

extern "C" int pcb_bn_apply2(const float* X, int ldx, int64_t n, int C, const float* mean, const float* invstd, const float* gamma,
                             const float* beta, const float* residual, int ldr, int relu, float* Y, int ldy, uint16_t* Yhi,
                             uint16_t* Ylo, int lds, void* stream) {
  PCB_ARG(n >= 0 && C >= 4 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldx >= C && ldy >= C);
  if (n == 0) return PCB_OK;
  PCB_ARG(X && Y && mean && invstd && gamma && beta && (!residual || (ldr >= C && ldr % 4 == 0)));
  PCB_ARG(!Yhi || (Ylo && lds >= C && lds % 4 == 0));
  int64_t n4 = n * (C / 4);
  bn_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, n4, C / 4, mean, invstd, gamma, beta, residual,
                                                                                 ldr, relu, Y, ldy, (__nv_bfloat16*)Yhi,
                                                                                 (__nv_bfloat16*)Ylo, lds);
  return check_launch("bn_apply_kernel");
}

extern "C" int pcb_bn_apply(const float* X, int64_t n, int C, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, const float* residual, int relu, float* Y, void* stream) {
  return pcb_bn_apply2(X, C, n, C, mean, invstd, gamma, beta, residual, C, relu, Y, C, nullptr, nullptr, 0, stream);
}

extern "C" int pcb_bn_backward2(const float* dY, int lddy, const float* X, int ldx, const float* relu_out, int ldm, int64_t n, int C,
                                const float* mean, const float* invstd, const float* gamma, float* dX, int lddx, float* dgamma,
                                float* dbeta, int accumulate_param_grads, float* gout, int ldg, int gout_mode, uint16_t* dXhi,
                                uint16_t* dXlo, int lds, void* ws, size_t ws_bytes, void* stream) {
  PCB_ARG(dY && X && mean && invstd && gamma && (dX || dXhi) && dgamma && dbeta && ws && n >= 1 && C >= 4 && C % 4 == 0 && C <= 1024);
  PCB_ARG(lddy >= C && ldx >= C && lddy % 4 == 0 && ldx % 4 == 0 && (!dX || (lddx >= C && lddx % 4 == 0)));
  PCB_ARG(!dXhi || (dXlo && lds >= C && lds % 4 == 0));
  PCB_ARG(!relu_out || (ldm >= C && ldm % 4 == 0));
  PCB_ARG(gout_mode == 0 || (gout && ldg >= C && ldg % 4 == 0));
  PCB_ARG(ws_bytes >= pcb_bn_ws_bytes(n, C) - 256);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = chunks_for(n);
  const int thr = colsum_threads(C);
  const int rp = thr / (C / 4);
  float* partial = (float*)ws;
  float* sums = partial + (size_t)chunks * 2 * C;        // [2][C]: dbeta, dgamma of THIS call (the apply pass needs them)
  Finalize fin{};
  fin.counter = counter_for(st); fin.mode = 2; fin.n = n; fin.dgamma = dgamma; fin.dbeta = dbeta; fin.accumulate = accumulate_param_grads;
  fin.sums = sums;
  if (!fin.counter) return PCB_ERR_CUDA;
  colsum_kernel<true><<<chunks, thr, (size_t)rp * 2 * C * sizeof(float), st>>>(dY, lddy, X, ldx, relu_out, ldm, n, C, mean, invstd, partial,
                                                                               fin);
  if (int e = check_launch("colsum_kernel<bwd>")) return e;
  int64_t n4 = n * (C / 4);
  bn_bwd_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(dY, lddy, X, ldx, relu_out, ldm, n4, C / 4, 1.0f / (float)n, mean,
                                                                    invstd, gamma, sums + C, sums, dX, lddx, gout, ldg, gout_mode,
                                                                    (__nv_bfloat16*)dXhi, (__nv_bfloat16*)dXlo, lds);
  return check_launch("bn_bwd_apply_kernel");
}

extern "C" int pcb_bn_backward(const float* dY, const float* X, int64_t n, int C, const float* mean, const float* invstd,
                               const float* gamma, float* dX, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                               void* stream) {
  return pcb_bn_backward2(dY, C, X, C, nullptr, 0, n, C, mean, invstd, gamma, dX, C, dgamma, dbeta, 0, nullptr, 0, 0, nullptr, nullptr,
                          0, ws, ws_bytes, stream);
}
