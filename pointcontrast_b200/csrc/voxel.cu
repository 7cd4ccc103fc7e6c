// GPU voxelisation and correspondence search -- the per-sample work of the reference's data loader
// (`pretrain/pointcontrast/lib/ddp_data_loaders.py:196-265`), which today runs on CPU workers:
//   * `ME.utils.sparse_quantize(xyz / voxel_size, return_index=True)` (`:228-241`; semseg: `lib/voxelizer.py:113-148`):
//     one point per occupied voxel                                                          -> pcb_voxelize
//   * `get_matching_indices` (`:36-49`): an open3d KD-tree radius search PER POINT, radius 1.5 voxels   -> pcb_radius_pairs
// Both are integer / hashing work on the same primitives as the coordinate manager (radix sort + head flags + scan, the
// open-addressing hash table of common.cuh); results are exact (tests/test_gpu_voxel.py: vs numpy / scipy cKDTree).
#include <cub/cub.cuh>
#include "common.cuh"

using namespace pcb;

namespace {

constexpr int VB = 1 << 20;      // voxel / cell index bias: |index| < 2^20 per axis, 21 bits each

__device__ __forceinline__ bool cell_of(float x, float y, float z, float inv_unused, float size, int& cx, int& cy, int& cz) {
  // floor(v / size) in fp32, exactly what numpy does on float32 input (IEEE division, then floor)
  const float fx = floorf(x / size), fy = floorf(y / size), fz = floorf(z / size);
  cx = (int)fx; cy = (int)fy; cz = (int)fz;
  return fabsf(fx) < (float)VB && fabsf(fy) < (float)VB && fabsf(fz) < (float)VB;
}
__device__ __forceinline__ uint64_t cell_key(int cx, int cy, int cz) {
  return ((uint64_t)(cx + VB) << 42) | ((uint64_t)(cy + VB) << 21) | (uint64_t)(cz + VB);
}

__global__ void point_key_kernel(const float* __restrict__ xyz, int64_t n, float size, uint64_t* __restrict__ keys, int32_t* __restrict__ idx,
                                 int32_t* status) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx, cy, cz;
  if (!cell_of(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f, size, cx, cy, cz)) { atomicOr(status, PCB_ERR_RANGE); cx = cy = cz = 0; }
  keys[i] = cell_key(cx, cy, cz);
  idx[i] = (int32_t)i;
}

__global__ void head_kernel(const uint64_t* __restrict__ sk, int64_t n, int32_t* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || sk[i] != sk[i - 1]) ? 1 : 0;
}

// voxelisation output: one row per run of equal keys; the radix sort is stable, so the first element of a run is the point with the
// smallest original index (np.unique(..., return_index=True))
__global__ void voxel_write_kernel(const uint64_t* __restrict__ sk, const int32_t* __restrict__ sidx, const int32_t* __restrict__ rank,
                                   int64_t n, int32_t* __restrict__ coords, int32_t* __restrict__ sel, int64_t* m_out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || sk[i] != sk[i - 1]) {
    const int r = rank[i] - 1;
    const uint64_t k = sk[i];
    coords[3 * r] = (int)(k >> 42) - VB; coords[3 * r + 1] = (int)((k >> 21) & 0x1FFFFF) - VB; coords[3 * r + 2] = (int)(k & 0x1FFFFF) - VB;
    sel[r] = sidx[i];
  }
  if (i == n - 1) *m_out = rank[i];
}

// radius search: runs of the cell-sorted target points -> run start/end, hash (cell key -> run)
__global__ void run_bounds_kernel(const uint64_t* __restrict__ sk, const int32_t* __restrict__ rank, int64_t n, uint64_t* __restrict__ run_key,
                                  int32_t* __restrict__ run_start, int32_t* __restrict__ run_end, int64_t* n_runs) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rank[i] - 1;
  if (i == 0 || sk[i] != sk[i - 1]) { run_key[r] = sk[i]; run_start[r] = (int32_t)i; }
  if (i == n - 1 || sk[i] != sk[i + 1]) run_end[r] = (int32_t)i + 1;
  if (i == n - 1) *n_runs = rank[i];
}

__global__ void run_insert_kernel(const uint64_t* __restrict__ run_key, const int64_t* __restrict__ n_runs, uint64_t* tk, int32_t* tv, uint64_t mask) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= *n_runs) return;
  const uint64_t key = run_key[i];
  uint64_t slot = mix64(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(tk + slot), (unsigned long long)KEY_EMPTY, (unsigned long long)key);
    if (prev == KEY_EMPTY) { tv[slot] = (int32_t)i; return; }
    slot = (slot + 1) & mask;
  }
}

// FILL = false: cnt[i] = number of targets within the radius of source i;  FILL = true: writes them at pairs[off[i] ...], ascending j
template <bool FILL>
__global__ void radius_kernel(const float* __restrict__ src, int64_t ns, const float* __restrict__ dst, float radius, const uint64_t* __restrict__ tk,
                              const int32_t* __restrict__ tv, uint64_t mask, const int32_t* __restrict__ run_start,
                              const int32_t* __restrict__ run_end, const int32_t* __restrict__ sidx, int32_t* __restrict__ cnt,
                              const int64_t* __restrict__ off, int32_t* __restrict__ pairs, int64_t cap) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ns) return;
  const float px = src[3 * i], py = src[3 * i + 1], pz = src[3 * i + 2];
  int cx, cy, cz;
  int found = 0;
  const int64_t base = FILL ? off[i] : 0;
  if (cell_of(px, py, pz, 0.f, radius, cx, cy, cz)) {
    const float r2 = __fmul_rn(radius, radius);
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          const int qx = cx + dx, qy = cy + dy, qz = cz + dz;
          if (abs(qx) >= VB || abs(qy) >= VB || abs(qz) >= VB) continue;
          const int run = hash_lookup(tk, tv, mask, cell_key(qx, qy, qz));
          if (run < 0) continue;
          for (int s = run_start[run]; s < run_end[run]; ++s) {
            const int j = sidx[s];
            const float ex = dst[3 * j] - px, ey = dst[3 * j + 1] - py, ez = dst[3 * j + 2] - pz;
            if (__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)) < r2) {      // no FMA contraction: reproducible on the host
              if (FILL && base + found < cap) { pairs[2 * (base + found)] = (int32_t)i; pairs[2 * (base + found) + 1] = j; }
              ++found;
            }
          }
        }
  }
  if (!FILL) { cnt[i] = found; return; }
  // ascending j within the row (a handful of entries): insertion sort in place
  const int64_t m = min((int64_t)found, cap - base > 0 ? cap - base : 0);
  for (int64_t a = 1; a < m; ++a) {
    const int32_t v = pairs[2 * (base + a) + 1];
    int64_t b = a - 1;
    while (b >= 0 && pairs[2 * (base + b) + 1] > v) { pairs[2 * (base + b + 1) + 1] = pairs[2 * (base + b) + 1]; --b; }
    pairs[2 * (base + b + 1) + 1] = v;
  }
}

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }
inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sort_scan_bytes(int64_t n) {
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t*)nullptr, (uint64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (int32_t*)nullptr, (int64_t*)nullptr, (int)n);
  size_t m = a > b ? a : b;
  return m > c ? m : c;
}

struct SortWs { uint64_t* k; uint64_t* sk; int32_t* idx; int32_t* sidx; int32_t* flag; int32_t* rank; int64_t* count; int32_t* status; void* cub; size_t cub_bytes; char* end; };

SortWs carve(void* ws, int64_t n) {
  char* p = (char*)ws;
  SortWs w;
  w.k = (uint64_t*)p; p += align_up(n * 8);
  w.sk = (uint64_t*)p; p += align_up(n * 8);
  w.idx = (int32_t*)p; p += align_up(n * 4);
  w.sidx = (int32_t*)p; p += align_up(n * 4);
  w.flag = (int32_t*)p; p += align_up(n * 4);
  w.rank = (int32_t*)p; p += align_up(n * 4);
  w.count = (int64_t*)p; p += 256;
  w.status = (int32_t*)p; p += 256;
  w.cub = p; w.cub_bytes = sort_scan_bytes(n); p += align_up(w.cub_bytes);
  w.end = p;
  return w;
}
size_t carve_bytes(int64_t n) { return 2 * align_up(n * 8) + 4 * align_up(n * 4) + 512 + align_up(sort_scan_bytes(n)) + 256; }

// keys of the points' cells -> stable sort -> head flags -> inclusive scan (rank)
int sort_cells(const float* xyz, int64_t n, float size, SortWs& w, cudaStream_t st) {
  PCB_CUDA(cudaMemsetAsync(w.status, 0, sizeof(int32_t), st));
  point_key_kernel<<<blocks_for(n, 256), 256, 0, st>>>(xyz, n, size, w.k, w.idx, w.status);
  if (int e = check_launch("point_key_kernel")) return e;
  size_t cb = w.cub_bytes;
  PCB_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, cb, w.k, w.sk, w.idx, w.sidx, (int)n, 0, 63, st));
  head_kernel<<<blocks_for(n, 256), 256, 0, st>>>(w.sk, n, w.flag);
  if (int e = check_launch("head_kernel")) return e;
  cb = w.cub_bytes;
  PCB_CUDA(cub::DeviceScan::InclusiveSum(w.cub, cb, w.flag, w.rank, (int)n, st));
  g_launches.fetch_add(10);
  return PCB_OK;
}

}  // namespace

extern "C" size_t pcb_voxelize_ws_bytes(int64_t n) { return carve_bytes(n < 1 ? 1 : n); }

extern "C" int pcb_voxelize(const float* xyz, int64_t n, float voxel_size, int32_t* out_coords, int32_t* sel, int64_t* m_out, void* ws,
                            size_t ws_bytes, void* stream) {
  PCB_ARG(n >= 0 && n < (1ll << 31) && voxel_size > 0.f && m_out);
  *m_out = 0;
  if (n == 0) return PCB_OK;
  PCB_ARG(xyz && out_coords && sel && ws && ws_bytes >= pcb_voxelize_ws_bytes(n));
  cudaStream_t st = (cudaStream_t)stream;
  SortWs w = carve(ws, n);
  if (int e = sort_cells(xyz, n, voxel_size, w, st)) return e;
  voxel_write_kernel<<<blocks_for(n, 256), 256, 0, st>>>(w.sk, w.sidx, w.rank, n, out_coords, sel, w.count);
  if (int e = check_launch("voxel_write_kernel")) return e;
  int32_t status = 0;
  PCB_CUDA(cudaMemcpyAsync(m_out, w.count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PCB_CUDA(cudaMemcpyAsync(&status, w.status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  PCB_CUDA(cudaStreamSynchronize(st));
  if (status) { set_error("pcb_voxelize: a point lies outside +-2^20 voxels"); return PCB_ERR_RANGE; }
  return PCB_OK;
}

extern "C" size_t pcb_radius_pairs_ws_bytes(int64_t ns, int64_t nd) {
  if (nd < 1) nd = 1;
  if (ns < 1) ns = 1;
  int64_t cap = 16;
  while (cap < 2 * nd) cap <<= 1;
  return carve_bytes(nd) + align_up(nd * 8) + 2 * align_up(nd * 4) + align_up(cap * 8) + align_up(cap * 4) + align_up(ns * 4) +
         align_up(ns * 8) + align_up(sort_scan_bytes(ns)) + 1024;
}

// pairs == NULL / cap == 0: only counts (*n_pairs = total).  Otherwise writes min(total, cap) pairs (i ascending, j ascending within i).
extern "C" int pcb_radius_pairs(const float* src, int64_t ns, const float* dst, int64_t nd, float radius, int32_t* pairs, int64_t cap,
                                int64_t* n_pairs, void* ws, size_t ws_bytes, void* stream) {
  PCB_ARG(ns >= 0 && nd >= 0 && ns < (1ll << 31) && nd < (1ll << 31) && radius > 0.f && n_pairs);
  *n_pairs = 0;
  if (ns == 0 || nd == 0) return PCB_OK;
  PCB_ARG(src && dst && ws && ws_bytes >= pcb_radius_pairs_ws_bytes(ns, nd));
  cudaStream_t st = (cudaStream_t)stream;
  SortWs w = carve(ws, nd);
  if (int e = sort_cells(dst, nd, radius, w, st)) return e;
  char* p = w.end;
  uint64_t* run_key = (uint64_t*)p; p += align_up(nd * 8);
  int32_t* run_start = (int32_t*)p; p += align_up(nd * 4);
  int32_t* run_end = (int32_t*)p; p += align_up(nd * 4);
  int64_t tcap = 16;
  while (tcap < 2 * nd) tcap <<= 1;
  uint64_t* tk = (uint64_t*)p; p += align_up(tcap * 8);
  int32_t* tv = (int32_t*)p; p += align_up(tcap * 4);
  int32_t* cnt = (int32_t*)p; p += align_up(ns * 4);
  int64_t* off = (int64_t*)p; p += align_up(ns * 8);
  void* cub2 = p; size_t cub2_bytes = sort_scan_bytes(ns);
  run_bounds_kernel<<<blocks_for(nd, 256), 256, 0, st>>>(w.sk, w.rank, nd, run_key, run_start, run_end, w.count);
  if (int e = check_launch("run_bounds_kernel")) return e;
  PCB_CUDA(cudaMemsetAsync(tk, 0xFF, (size_t)tcap * 8, st));
  run_insert_kernel<<<blocks_for(nd, 256), 256, 0, st>>>(run_key, w.count, tk, tv, (uint64_t)tcap - 1);
  if (int e = check_launch("run_insert_kernel")) return e;
  radius_kernel<false><<<blocks_for(ns, 128), 128, 0, st>>>(src, ns, dst, radius, tk, tv, (uint64_t)tcap - 1, run_start, run_end, w.sidx, cnt,
                                                            nullptr, nullptr, 0);
  if (int e = check_launch("radius_kernel<count>")) return e;
  PCB_CUDA(cub::DeviceScan::ExclusiveSum(cub2, cub2_bytes, cnt, off, (int)ns, st));
  g_launches.fetch_add(2);
  int64_t last_off = 0; int32_t last_cnt = 0;
  PCB_CUDA(cudaMemcpyAsync(&last_off, off + ns - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PCB_CUDA(cudaMemcpyAsync(&last_cnt, cnt + ns - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  PCB_CUDA(cudaStreamSynchronize(st));
  *n_pairs = last_off + last_cnt;
  if (!pairs || cap <= 0) return PCB_OK;
  radius_kernel<true><<<blocks_for(ns, 128), 128, 0, st>>>(src, ns, dst, radius, tk, tv, (uint64_t)tcap - 1, run_start, run_end, w.sidx, cnt, off,
                                                           pairs, cap);
  return check_launch("radius_kernel<fill>");
}
