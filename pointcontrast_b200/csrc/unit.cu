// Fused "units" of the Res16UNet graph: convolution -> BatchNorm -> (+ residual) -> (ReLU) issued from one C call, and the
// reverse sweep of the same unit (include/pcb200.h: pcb_unit).  Host-side sequencing only -- the kernels live in
// conv_tc5.cu / conv.cu / bn.cu; what this file adds over calling them one by one from the host language:
//   * the BatchNorm statistics come from the convolution's own epilogue (or from its offset-split reduction pass), so the
//     separate column-sum pass over z and its launch disappear;
//   * one boundary crossing per unit instead of three to five.
// Replaces the per-module call sequence of `model/modules/resnet_block.py:44-60` / `model/res16unet.py:206-268`.
#include "common.cuh"

namespace pcb {
struct BnFuse { int64_t n0; float eps, momentum; float* mean; float* invstd; float* running_mean; float* running_var; void* ws; size_t ws_bytes; };
int conv_forward_split_impl(const uint16_t* Xhi, const uint16_t* Xlo, int lds, const int32_t* tbl, int64_t tbl_stride, const int32_t* kmap,
                            int K, int64_t n_out, int Cin, int Cout, const void* w_tiles, const float* bias, float* Y, int ldy, void* ws,
                            size_t ws_bytes, int flags, cudaStream_t st, const BnFuse* bn, int* bn_done);
int bn_eval_stats_launch(const float* running_mean, const float* running_var, int C, float eps, float* mean, float* invstd, cudaStream_t st);
int bn_backward_impl(const float* dY, int lddy, const float* X, int ldx, const float* relu_out, int ldm, const uint16_t* relu_hi, int ldmh,
                     int64_t n, int64_t n0, int C, const float* mean, const float* invstd, const float* gamma, float* dX, int lddx,
                     float* dgamma, float* dbeta, int accumulate_param_grads, float* gout, int ldg, int gout_mode, uint16_t* dXhi,
                     uint16_t* dXlo, int lds, void* ws, size_t ws_bytes, cudaStream_t st);
}

using namespace pcb;

// ------------------------------------------------------------------------------------------------ per-launch timing
#include <stdlib.h>
#include <vector>
namespace {
struct ProfRec { cudaEvent_t e0, e1; int kind; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t g_prof_open = nullptr;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
namespace pcb {
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("PCB_PDL"); on = (e && atoi(e) == 0) ? 0 : 1; }      // on by default; PCB_PDL=0 disables
  return on == 1;
}
void prof_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  g_prof_open = prof_event();
  cudaEventRecord(g_prof_open, st);
}
void prof_end(cudaStream_t st, int kind) {
  if (!g_prof_on || !g_prof_open) return;
  cudaEvent_t e1 = prof_event();
  cudaEventRecord(e1, st);
  g_prof.push_back({g_prof_open, e1, kind});
  g_prof_open = nullptr;
}
}  // namespace pcb

extern "C" int pcb_profile_enable(int on) {
  g_prof_on = on != 0;
  return PCB_OK;
}

extern "C" int pcb_profile_read(float* ms, int32_t* kinds, int max_records, int* count) {
  PCB_ARG(count && (max_records == 0 || (ms && kinds)));
  int n = 0;
  for (auto& r : g_prof) {
    PCB_CUDA(cudaEventSynchronize(r.e1));
    if (n < max_records) {
      float t = 0.f;
      PCB_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
      ms[n] = t; kinds[n] = r.kind;
    }
    ++n;
    g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1);
  }
  g_prof.clear();
  *count = n;
  return PCB_OK;
}

namespace {
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
inline bool tensor_core_shape(int Cin, int Cout) { return Cin % 32 == 0 && Cout % 32 == 0; }

// scratch layout: [convolution scratch (largest of forward / data gradient / weight gradient) | BatchNorm partial sums]
size_t conv_part_bytes(int K, int64_t n_in, int64_t n_out, int Cin, int Cout) {
  size_t a = pcb_conv_forward_ws_bytes(K, n_out, Cin, Cout), b = pcb_conv_forward_ws_bytes(K, n_in, Cout, Cin);
  size_t c = tensor_core_shape(Cin, Cout) ? pcb_conv_wgrad_split_ws_bytes(K, n_out > n_in ? n_out : n_in, Cin, Cout)
                                          : pcb_conv_wgrad_ws_bytes(K, n_out > n_in ? n_out : n_in, Cin, Cout);
  size_t d = tensor_core_shape(Cin, Cout) ? pcb_conv_wgrad_split_ws_bytes(K, n_out > n_in ? n_out : n_in, Cout, Cin) : 0;
  size_t m = a > b ? a : b;
  if (c > m) m = c;
  if (d > m) m = d;
  return up256(m);
}
}  // namespace

extern "C" size_t pcb_unit_ws_bytes(int K, int64_t n_in, int64_t n_out, int Cin, int Cout) {
  return conv_part_bytes(K, n_in, n_out, Cin, Cout) + up256(pcb_bn_ws_bytes(n_out, Cout));
}

extern "C" int pcb_unit_forward(const pcb_unit* u, void* stream) {
  PCB_ARG(u && u->K >= 1 && u->K <= PCB_MAX_KERNEL_VOLUME && u->n_out >= 1 && u->n_in >= 1 && u->n0 >= 1 && u->n0 <= u->n_out);
  PCB_ARG(u->fwd_tbl && u->z_p && u->out_hi && u->out_lo && u->mean && u->invstd && u->gamma && u->beta && u->ws);
  PCB_ARG(!(u->flags & PCB_UNIT_FP16_FORWARD) || (u->flags & PCB_UNIT_EVAL) || (u->out_bhi && u->out_blo));
  PCB_ARG(u->ws_bytes >= pcb_unit_ws_bytes(u->K, u->n_in, u->n_out, u->Cin, u->Cout));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t conv_bytes = conv_part_bytes(u->K, u->n_in, u->n_out, u->Cin, u->Cout);
  unsigned char* bn_ws = (unsigned char*)u->ws + conv_bytes;
  const size_t bn_bytes = u->ws_bytes - conv_bytes;
  int have_stats = 0;
  const bool f16 = (u->flags & PCB_UNIT_FP16_FORWARD) != 0;      // activations (x, out planes) and forward weight tiles are fp16 hi/lo
  if (tensor_core_shape(u->Cin, u->Cout)) {
    PCB_ARG(u->x_hi && u->x_lo && u->wt_fwd);
    BnFuse bn{u->n0, u->eps, u->momentum, u->mean, u->invstd, u->running_mean, u->running_var, bn_ws, bn_bytes};
    const bool fuse = !(u->flags & (PCB_UNIT_SEPARATE_STATS | PCB_UNIT_EVAL));
    if (int e = conv_forward_split_impl(u->x_hi, u->x_lo, u->x_lds, u->fwd_tbl, u->fwd_stride, u->fwd_kmap, u->K, u->n_out, u->Cin, u->Cout,
                                        u->wt_fwd, nullptr, u->z_p, u->z_ld, u->ws, conv_bytes, f16 ? (PCB_PLANES_A_FP16 | PCB_PLANES_B_FP16) : 0,
                                        st, fuse ? &bn : nullptr, &have_stats)) return e;
  } else {
    PCB_ARG(u->x_p && u->W);
    if (int e = pcb_conv_forward(u->x_p, u->x_ld, u->fwd_tbl, u->fwd_stride, u->fwd_kmap, u->K, u->n_out, u->Cin, u->Cout,
                                 nullptr, nullptr, u->W, nullptr, u->z_p, u->z_ld, nullptr, 0, 0, stream)) return e;
  }
  ProfScope prof(st, 2);                     // BatchNorm forward: statistics (unless fused into the split reduction) + normalise / residual / ReLU / planes
  if (u->flags & PCB_UNIT_EVAL) {            // eval-mode BatchNorm (`downstream/semseg/lib/test.py:95-117`): normalise with the running statistics
    PCB_ARG(u->n0 == u->n_out && u->running_mean && u->running_var);
    if (int e = bn_eval_stats_launch(u->running_mean, u->running_var, u->Cout, u->eps, u->mean, u->invstd, st)) return e;
  } else if (!have_stats) {
    if (int e = pcb_bn_stats_seg(u->z_p, u->z_ld, u->n_out, u->n0, u->Cout, u->eps, u->momentum, u->mean, u->invstd, u->running_mean,
                                 u->running_var, bn_ws, bn_bytes, stream)) return e;
  }
  return pcb_bn_apply_seg(u->z_p, u->z_ld, u->n_out, u->n0, u->Cout, u->mean, u->invstd, u->gamma, u->beta, u->res_p, u->res_ld,
                          (u->relu ? PCB_BN_RELU : 0) | (f16 ? PCB_PLANES_A_FP16 : 0), u->out_p, u->out_ld, u->out_hi, u->out_lo, u->out_lds,
                          f16 ? u->out_bhi : nullptr, f16 ? u->out_blo : nullptr, stream);
}

extern "C" int pcb_unit_backward(const pcb_unit* u, void* stream) {
  PCB_ARG(u && u->K >= 1 && u->K <= PCB_MAX_KERNEL_VOLUME && u->n_out >= 1 && u->n_in >= 1 && u->n0 >= 1 && u->n0 <= u->n_out);
  PCB_ARG(u->g_p && u->z_p && u->mean && u->invstd && u->gamma && u->dgamma && u->dbeta && u->dW && u->wg_tbl && u->ws);
  PCB_ARG(u->ws_bytes >= pcb_unit_ws_bytes(u->K, u->n_in, u->n_out, u->Cin, u->Cout));
  const bool tc = tensor_core_shape(u->Cin, u->Cout);
  const bool f16 = (u->flags & PCB_UNIT_FP16_FORWARD) != 0;
  PCB_ARG(tc ? (u->dz_hi && u->dz_lo && u->x_hi && u->x_lo) : (u->dz_p && u->x_p));
  PCB_ARG(tc || u->gin_mode == 0);               // only the 3-channel stem is not tensor-core shaped: its input wants no gradient
  cudaStream_t st = (cudaStream_t)stream;
  const size_t conv_bytes = conv_part_bytes(u->K, u->n_in, u->n_out, u->Cin, u->Cout);
  unsigned char* bn_ws = (unsigned char*)u->ws + conv_bytes;
  // 1. g * (out > 0) -> BatchNorm backward -> dz (split planes), residual-gradient fan-out, dgamma / dbeta accumulated
  prof_begin(st);
  if (int e = bn_backward_impl(u->g_p, u->g_ld, u->z_p, u->z_ld, nullptr, 0, u->relu ? u->out_hi : nullptr, u->out_lds, u->n_out, u->n0, u->Cout,
                               u->mean, u->invstd, u->gamma, u->dz_p, u->dz_ld, u->dgamma, u->dbeta, 1, u->gres_p, u->gres_ld, u->gres_mode,
                               u->dz_hi, u->dz_lo, u->dz_ld, bn_ws, u->ws_bytes - conv_bytes, st)) return e;
  prof_end(st, 3);
  // 2. weight gradient, accumulated into dW (the flat parameter-gradient buffer)
  if (tc) {
    // the activation operand as bf16 hi/lo planes (its fp16 planes serve the forward pass only): both MMA operands share one format
    const uint16_t* xh = f16 ? u->x_bhi : u->x_hi;
    const uint16_t* xl = f16 ? u->x_blo : u->x_lo;
    PCB_ARG(xh && xl);
    const uint16_t *Ahi, *Alo, *Bhi, *Blo; int lda, ldb, Ca, Cb, tr; int64_t rows;
    if (u->wg_gather_x) { Ahi = xh; Alo = xl; lda = u->x_lds; Bhi = u->dz_hi; Blo = u->dz_lo; ldb = u->dz_ld; Ca = u->Cin; Cb = u->Cout; tr = 0; rows = u->n_out; }
    else { Ahi = u->dz_hi; Alo = u->dz_lo; lda = u->dz_ld; Bhi = xh; Blo = xl; ldb = u->x_lds; Ca = u->Cout; Cb = u->Cin; tr = 1; rows = u->n_in; }
    if (int e = pcb_conv_wgrad_split(Ahi, Alo, lda, Bhi, Blo, ldb, u->wg_tbl, u->wg_stride, u->K, rows, Ca, Cb, u->dW, tr, u->ws, conv_bytes,
                                     PCB_CONV_ACCUMULATE, stream)) return e;
  } else {
    PCB_ARG(u->wg_gather_x);
    if (int e = pcb_conv_wgrad(u->x_p, u->x_ld, u->dz_p, u->dz_ld, u->wg_tbl, u->wg_stride, u->K, u->n_out, u->Cin, u->Cout, u->dW, 0, u->ws,
                               conv_bytes, PCB_CONV_ACCUMULATE, stream)) return e;
  }
  // 3. data gradient: the forward kernel on the data-gradient weight tiles and the opposite-offset table
  if (u->gin_mode) {
    PCB_ARG(u->gin_p && u->dg_tbl && u->wt_dg);
    if (int e = pcb_conv_forward_split(u->dz_hi, u->dz_lo, u->dz_ld, u->dg_tbl, u->dg_stride, u->dg_kmap, u->K, u->n_in, u->Cout, u->Cin, u->wt_dg,
                                       nullptr, u->gin_p, u->gin_ld, u->ws, conv_bytes, u->gin_mode == 2 ? PCB_CONV_ACCUMULATE : 0, stream)) return e;
  }
  return PCB_OK;
}
