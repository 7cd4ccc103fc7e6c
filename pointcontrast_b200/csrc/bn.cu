// BatchNorm over the rows of a sparse tensor's feature matrix (training-mode statistics), with the
// affine-normalise (+ residual, + ReLU) fused into one elementwise pass.  HBM-bound: every kernel moves
// 16-byte vectors with consecutive threads on consecutive channels.
// Replaces MinkowskiBatchNorm == torch.nn.BatchNorm1d on .F (reference call sites in include/pcb200.h).
#include "common.cuh"

using namespace pcb;

namespace {

constexpr int ROWS_PER_CHUNK = 512;

// partial[chunk][0][C] = sum(a), partial[chunk][1][C] = sum(a*b)    (b == a for the forward statistics)
// block: (C/4) channel-vectors x RP row lanes; grid: one CTA per chunk of rows.
template <bool TWO_INPUTS>
__global__ void colsum_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int64_t n, int C,
                              const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ partial) {
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
      float4 a = __ldg(reinterpret_cast<const float4*>(A + r * C) + c4);
      if (TWO_INPUTS) {
        float4 x = __ldg(reinterpret_cast<const float4*>(Bm + r * C) + c4);
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
}

__global__ void bn_finalize_kernel(const float* __restrict__ partial, int chunks, int64_t n, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* running_mean, float* running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < chunks; ++k) { s1 += partial[(int64_t)k * 2 * C + c]; s2 += partial[(int64_t)k * 2 * C + C + c]; }
  double m = s1 / (double)n;
  double var = s2 / (double)n - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var) {
    double unb = n > 1 ? var * ((double)n / (double)(n - 1)) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void bn_apply_kernel(const float* __restrict__ X, int64_t n4, int cv, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ residual, int relu, float* __restrict__ Y) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  int c4 = (int)(i % cv);
  float4 x = __ldg(reinterpret_cast<const float4*>(X) + i);
  float4 mu = reinterpret_cast<const float4*>(mean)[c4], is = reinterpret_cast<const float4*>(invstd)[c4];
  float4 g = reinterpret_cast<const float4*>(gamma)[c4], b = reinterpret_cast<const float4*>(beta)[c4];
  float4 y;
  y.x = (x.x - mu.x) * is.x * g.x + b.x; y.y = (x.y - mu.y) * is.y * g.y + b.y;
  y.z = (x.z - mu.z) * is.z * g.z + b.z; y.w = (x.w - mu.w) * is.w * g.w + b.w;
  if (residual) {
    float4 r = __ldg(reinterpret_cast<const float4*>(residual) + i);
    y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
  }
  if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
  reinterpret_cast<float4*>(Y)[i] = y;
}

// dgamma = sum(dY*xhat), dbeta = sum(dY); also leaves them in ws for the apply pass
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int chunks, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < chunks; ++k) { s1 += partial[(int64_t)k * 2 * C + c]; s2 += partial[(int64_t)k * 2 * C + C + c]; }
  dbeta[c] = (float)s1;
  dgamma[c] = (float)s2;
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dY, const float* __restrict__ X, int64_t n4, int cv, float inv_n,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, float* __restrict__ dX) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  int c4 = (int)(i % cv);
  float4 dy = __ldg(reinterpret_cast<const float4*>(dY) + i);
  float4 x = __ldg(reinterpret_cast<const float4*>(X) + i);
  float4 mu = reinterpret_cast<const float4*>(mean)[c4], is = reinterpret_cast<const float4*>(invstd)[c4];
  float4 g = reinterpret_cast<const float4*>(gamma)[c4];
  float4 dg = reinterpret_cast<const float4*>(dgamma)[c4], db = reinterpret_cast<const float4*>(dbeta)[c4];
  float4 o;
  o.x = g.x * is.x * (dy.x - db.x * inv_n - (x.x - mu.x) * is.x * dg.x * inv_n);
  o.y = g.y * is.y * (dy.y - db.y * inv_n - (x.y - mu.y) * is.y * dg.y * inv_n);
  o.z = g.z * is.z * (dy.z - db.z * inv_n - (x.z - mu.z) * is.z * dg.z * inv_n);
  o.w = g.w * is.w * (dy.w - db.w * inv_n - (x.w - mu.w) * is.w * dg.w * inv_n);
  reinterpret_cast<float4*>(dX)[i] = o;
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
  return (size_t)chunks_for(n) * 2 * C * sizeof(float) + 256;
}

extern "C" int pcb_bn_stats(const float* X, int64_t n, int C, float eps, float momentum, float* mean, float* invstd,
                            float* running_mean, float* running_var, void* ws, size_t ws_bytes, void* stream) {
  PCB_ARG(X && mean && invstd && ws && n >= 1 && C >= 4 && C % 4 == 0 && C <= 1024);
  PCB_ARG(ws_bytes >= pcb_bn_ws_bytes(n, C) - 256);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = chunks_for(n);
  const int thr = colsum_threads(C);
  const int rp = thr / (C / 4);
  colsum_kernel<false><<<chunks, thr, (size_t)rp * 2 * C * sizeof(float), st>>>(X, nullptr, n, C, nullptr, nullptr, (float*)ws);
  if (int e = check_launch("colsum_kernel")) return e;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>((const float*)ws, chunks, n, C, eps, momentum, mean, invstd, running_mean,
                                                      running_var);
  return check_launch("bn_finalize_kernel");
}

extern "C" int pcb_bn_apply(const float* X, int64_t n, int C, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, const float* residual, int relu, float* Y, void* stream) {
  PCB_ARG(n >= 0 && C >= 4 && C % 4 == 0);
  if (n == 0) return PCB_OK;
  PCB_ARG(X && Y && mean && invstd && gamma && beta);
  int64_t n4 = n * (C / 4);
  bn_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, n4, C / 4, mean, invstd, gamma, beta, residual,
                                                                                 relu, Y);
  return check_launch("bn_apply_kernel");
}

extern "C" int pcb_bn_backward(const float* dY, const float* X, int64_t n, int C, const float* mean, const float* invstd,
                               const float* gamma, float* dX, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                               void* stream) {
  PCB_ARG(dY && X && mean && invstd && gamma && dX && dgamma && dbeta && ws && n >= 1 && C >= 4 && C % 4 == 0 && C <= 1024);
  PCB_ARG(ws_bytes >= pcb_bn_ws_bytes(n, C) - 256);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = chunks_for(n);
  const int thr = colsum_threads(C);
  const int rp = thr / (C / 4);
  colsum_kernel<true><<<chunks, thr, (size_t)rp * 2 * C * sizeof(float), st>>>(dY, X, n, C, mean, invstd, (float*)ws);
  if (int e = check_launch("colsum_kernel<bwd>")) return e;
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>((const float*)ws, chunks, C, dgamma, dbeta);
  if (int e = check_launch("bn_bwd_finalize_kernel")) return e;
  int64_t n4 = n * (C / 4);
  bn_bwd_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(dY, X, n4, C / 4, 1.0f / (float)n, mean, invstd, gamma, dgamma,
                                                                    dbeta, dX);
  return check_launch("bn_bwd_apply_kernel");
}
