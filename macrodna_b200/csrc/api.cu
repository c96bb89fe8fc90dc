// C ABI of the hot path (include/macrodna_b200.h): context, workspace, step loop (K4), fused driver.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstring>
#include <new>
#include <thread>

#include "mcd_internal.cuh"

// ------------------------------------------------------------------------------------------------
// context plumbing
// ------------------------------------------------------------------------------------------------
int mcd_fail(mcd_context* h, int status, const char* what, cudaError_t e) {
  if (h != nullptr) {
    char buf[512];
    if (e != cudaSuccess)
      snprintf(buf, sizeof buf, "%s: %s (%s)", mcd_strerror(status), what, cudaGetErrorString(e));
    else
      snprintf(buf, sizeof buf, "%s: %s", mcd_strerror(status), what);
    h->err = buf;
  }
  return status;
}

int mcd_ws(mcd_context* h, int slot, size_t bytes, void** out) {
  mcd_buffer& b = h->ws[slot];
  if (bytes == 0) bytes = 256;
  if (b.bytes < bytes) {
    if (b.ptr != nullptr) {
      cudaStreamSynchronize(h->stream);
      cudaFree(b.ptr);
      b.ptr = nullptr;
      b.bytes = 0;
    }
    cudaError_t e = cudaMalloc(&b.ptr, bytes);
    if (e != cudaSuccess) {
      b.ptr = nullptr;
      return mcd_fail(h, MCD_ERR_NOMEM, "cudaMalloc workspace", e);
    }
    b.bytes = bytes;
  }
  *out = b.ptr;
  return MCD_OK;
}

extern "C" {

int mcd_abi_version(void) { return MCD_ABI_VERSION; }

const char* mcd_strerror(int status) {
  switch (status) {
    case MCD_OK: return "ok";
    case MCD_ERR_INVALID: return "invalid argument";
    case MCD_ERR_CUDA: return "CUDA error";
    case MCD_ERR_NOMEM: return "out of memory";
    case MCD_ERR_NONFINITE: return "non-finite value in input";
    case MCD_ERR_UNSUPPORTED: return "unsupported";
    case MCD_ERR_NOT_CONVERGED: return "assignment solver did not converge";
    default: return "unknown status";
  }
}

int mcd_create(mcd_handle* out, int device) {
  if (out == nullptr) return MCD_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MCD_ERR_CUDA;
  if (device < 0 || device >= ndev) return MCD_ERR_INVALID;
  mcd_context* h = new (std::nothrow) mcd_context();
  if (h == nullptr) return MCD_ERR_NOMEM;
  h->device = device;
  cudaError_t e = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess && prop.major < 10) {
    delete h;
    return MCD_ERR_UNSUPPORTED;  // sm_100a-only build
  }
  if (e == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_flags, 64);
  if (e == cudaSuccess) e = cudaMemsetAsync(h->d_flags, 0, 64, h->stream);
  if (e != cudaSuccess) {
    delete h;
    return MCD_ERR_CUDA;
  }
  *out = h;
  return MCD_OK;
}

int mcd_destroy(mcd_handle h) {
  if (h == nullptr) return MCD_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (mcd_context* w : h->workers) mcd_destroy(w);
  h->workers.clear();
  for (auto& b : h->ws)
    if (b.ptr) cudaFree(b.ptr);
  for (auto ev : h->ev) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) {
    if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
    if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
  }
  if (h->stager) mcd_stager_destroy(h->stager);
  if (h->d_flags) cudaFree(h->d_flags);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return MCD_OK;
}

const char* mcd_last_error(mcd_handle h) { return h ? h->err.c_str() : "null handle"; }
int mcd_device_sm_count(mcd_handle h) { return h ? h->sm_count : 0; }
void* mcd_stream(mcd_handle h) { return h ? (void*)h->stream : nullptr; }

int mcd_synchronize(mcd_handle h) {
  if (!h) return MCD_ERR_INVALID;
  MCD_CUDA(h, cudaSetDevice(h->device));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  return MCD_OK;
}

int64_t mcd_padded_k(int64_t G) { return (G + 15) / 16 * 16; }
int64_t mcd_padded_k_split(int64_t G) { return (G + 63) / 64 * 64; }
int64_t mcd_num_steps(int64_t M, int64_t N) { return N > 0 ? (M + N - 1) / N : 0; }

int mcd_standardize(mcd_handle h, const double* X, int64_t ncells, int64_t G, int64_t ldx, double* centred,
                    double* norms) {
  if (!h) return MCD_ERR_INVALID;
  if (!X || !centred || !norms || ncells < 0 || G < 1 || ldx < G || G > 0x7fffffff)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_standardize arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  return mcd_launch_standardize(h, X, ncells, G, ldx, centred, mcd_padded_k(G), nullptr, nullptr, 0, norms);
}

int mcd_standardize_split(mcd_handle h, const double* X, int64_t ncells, int64_t G, int64_t ldx, uint16_t* slices,
                           double* norms) {
  if (!h) return MCD_ERR_INVALID;
  if (!X || !slices || !norms || ncells < 0 || G < 1 || ldx < G || G > 0x7fffffff)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_standardize_split arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  const int64_t ldk16 = mcd_padded_k_split(G);
  return mcd_launch_standardize(h, X, ncells, G, ldx, nullptr, 0, slices, slices + ncells * ldk16, ldk16, norms);
}

int mcd_ozaki_default_slices(void) { return 6; }

int mcd_ozaki_slices_for(int64_t M, int64_t N, int64_t G) {
  // 8 slices (36 products, error at the level of an FP64 GEMM's own rounding) while the contraction is cheap
  // anyway; 6 slices (21 products, ~1e-12 absolute) once it is the large-instance bottleneck
  return (double)M * (double)N * (double)G <= 2.0e11 ? 8 : 6;
}

int mcd_ozaki_slices(mcd_handle h, int64_t M, int64_t N, int64_t G) {
  if (h != nullptr && h->opt.ozaki_slices >= 2)
    return h->opt.ozaki_slices > MCD_OZAKI_MAX_SLICES ? MCD_OZAKI_MAX_SLICES : h->opt.ozaki_slices;
  return mcd_ozaki_slices_for(M, N, G);
}

namespace {
struct OptEntry {
  const char* name;
  int is_double;
  size_t off;
};
#define MCD_OPT_I(n, f) {n, 0, offsetof(mcd_options, f)}
#define MCD_OPT_D(n, f) {n, 1, offsetof(mcd_options, f)}
const OptEntry kOptions[] = {
    MCD_OPT_I("certify", certify),
    MCD_OPT_I("debug", debug),
    MCD_OPT_I("corr_only", corr_only),
    MCD_OPT_I("deterministic", deterministic),
    MCD_OPT_I("ozaki.slices", ozaki_slices),
    MCD_OPT_I("ozaki.align", ozaki_align),
    MCD_OPT_I("ozaki.plan", ozaki_plan),
    MCD_OPT_I("k1.generic", k1_generic),
    MCD_OPT_I("k1.no_stream", k1_no_stream),
    MCD_OPT_D("lap.theta", lap_theta),
    MCD_OPT_D("lap.eps_min", lap_eps_min),
    MCD_OPT_D("lap.eps0", lap_eps0),
    MCD_OPT_I("lap.scaling", lap_scaling),
    MCD_OPT_D("lap.max_rounds", lap_max_rounds),
    MCD_OPT_D("lap.tail_budget", lap_tail_budget),
    MCD_OPT_I("lap.blocks_per_sm", lap_blocks_per_sm),
    MCD_OPT_I("lap.grid_blocks", lap_grid_blocks),
    MCD_OPT_I("lap.list_max_m", lap_list_max_m),
    MCD_OPT_I("lap.lists", lap_lists),
    MCD_OPT_I("lap.list_min_nu", lap_list_min_nu),
    MCD_OPT_I("lap.tail_cluster", lap_tail_cluster),
    MCD_OPT_I("lap.tail_mh", lap_tail_mh),
    MCD_OPT_I("lap.tail_sym", lap_tail_sym),
    MCD_OPT_I("lap.async", lap_async),
    MCD_OPT_I("lap.async_nu", lap_async_nu),
    MCD_OPT_I("lap.async_threads", lap_async_threads),
    MCD_OPT_I("lap.async_blocks_per_sm", lap_async_blocks_per_sm),
    MCD_OPT_I("lap.async_stop", lap_async_stop),
    MCD_OPT_I("lap.prefetch_rows", lap_prefetch_rows),
    MCD_OPT_I("lap.tail_nu", lap_tail_nu),
    MCD_OPT_I("lap.mh_nu", lap_mh_nu),
    MCD_OPT_I("lap.scale_cut", lap_scale_cut),
    MCD_OPT_I("lap.scale_full_phases", lap_scale_full_phases),
    MCD_OPT_D("lap.scale_tail_rounds", lap_scale_tail_rounds),
    MCD_OPT_I("lap.aug_nu", lap_aug_nu),
    MCD_OPT_I("lap.aug_nu_square", lap_aug_nu_square),
    MCD_OPT_I("lap.rank_select", lap_rank_select),
    MCD_OPT_I("lap.min_chunk", lap_min_chunk),
    MCD_OPT_I("lap.chunk_waves", lap_chunk_waves),
};
const OptEntry* find_option(const char* name) {
  if (name == nullptr) return nullptr;
  for (const OptEntry& e : kOptions)
    if (strcmp(e.name, name) == 0) return &e;
  return nullptr;
}
}  // namespace

int mcd_set_option(mcd_handle h, const char* name, double value) {
  if (!h) return MCD_ERR_INVALID;
  const OptEntry* e = find_option(name);
  if (e == nullptr || !(value == value)) return mcd_fail(h, MCD_ERR_INVALID, "mcd_set_option: unknown option or NaN value");
  char* base = reinterpret_cast<char*>(&h->opt);
  if (e->is_double)
    *reinterpret_cast<double*>(base + e->off) = value;
  else
    *reinterpret_cast<int*>(base + e->off) = (int)value;
  return MCD_OK;
}

int mcd_get_option(mcd_handle h, const char* name, double* value) {
  if (!h || !value) return MCD_ERR_INVALID;
  const OptEntry* e = find_option(name);
  if (e == nullptr) return mcd_fail(h, MCD_ERR_INVALID, "mcd_get_option: unknown option");
  const char* base = reinterpret_cast<const char*>(&h->opt);
  *value = e->is_double ? *reinterpret_cast<const double*>(base + e->off)
                        : (double)*reinterpret_cast<const int*>(base + e->off);
  return MCD_OK;
}

int mcd_standardize_ozaki(mcd_handle h, const double* X, int64_t ncells, int64_t G, int64_t ldx, int8_t* digits,
                          int nsl, double* scale, double* norms) {
  if (!h) return MCD_ERR_INVALID;
  if (!X || !digits || !scale || !norms || ncells < 0 || G < 1 || ldx < G || G > 0x7fffffff || nsl < 2 ||
      nsl > MCD_OZAKI_MAX_SLICES)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_standardize_ozaki arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  mcd_ozaki_out oz;
  oz.digits = digits;
  oz.ldk8 = mcd_padded_k_split(G);
  oz.slice_stride = ncells * oz.ldk8;
  oz.nsl = nsl;
  oz.scale = scale;
  return mcd_launch_standardize(h, X, ncells, G, ldx, nullptr, 0, nullptr, nullptr, 0, norms, nullptr, &oz);
}

int mcd_corr_ozaki(mcd_handle h, const int8_t* A8, int64_t M, const int8_t* B8, int64_t N, int64_t G, int64_t ldk8,
                   int nsl, const double* sA, const double* sB, const double* nA, const double* nB, double* C,
                   int64_t ldc, double* Ct, int64_t ldct) {
  if (!h) return MCD_ERR_INVALID;
  if (!A8 || !B8 || !sA || !sB || !nA || !nB || (!C && !Ct) || M < 0 || N < 0 || G < 1 || ldk8 < G ||
      (ldk8 % 64) != 0 || (C && ldc < N) || (Ct && ldct < M))
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_corr_ozaki arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  return mcd_launch_corr_ozaki(h, A8, M * ldk8, M, B8, N * ldk8, N, ldk8, nsl, sA, sB, nA, nB, C, ldc, Ct, ldct);
}

int mcd_check_finite(mcd_handle h) {
  if (!h) return MCD_ERR_INVALID;
  MCD_CUDA(h, cudaSetDevice(h->device));
  int flag = 0;
  MCD_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  if (flag) return mcd_fail(h, MCD_ERR_NONFINITE, "NaN or Inf in the expression / copy-number matrix");
  return MCD_OK;
}

int mcd_corr_fp64(mcd_handle h, const double* A, int64_t M, const double* B, int64_t N, int64_t G, int64_t ldk,
                  const double* nA, const double* nB, double* C, int64_t ldc, double* Ct, int64_t ldct) {
  if (!h) return MCD_ERR_INVALID;
  if (!A || !B || !nA || !nB || (!C && !Ct) || M < 0 || N < 0 || G < 1 || ldk < G || (ldk % 16) != 0 ||
      (C && ldc < N) || (Ct && ldct < M))
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_corr_fp64 arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  return mcd_launch_corr_fp64(h, A, M, B, N, ldk, nA, nB, C, ldc, Ct, ldct);
}

int mcd_corr_split(mcd_handle h, const uint16_t* A3, int64_t M, const uint16_t* B3, int64_t N, int64_t G,
                    int64_t ldk16, const double* nA, const double* nB, double* C, int64_t ldc, double* Ct,
                    int64_t ldct) {
  if (!h) return MCD_ERR_INVALID;
  if (!A3 || !B3 || !nA || !nB || (!C && !Ct) || M < 0 || N < 0 || G < 1 || ldk16 < G || (ldk16 % 64) != 0 ||
      (C && ldc < N) || (Ct && ldct < M))
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_corr_split arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  return mcd_launch_corr_split(h, A3, A3 + M * ldk16, M, B3, B3 + N * ldk16, N, ldk16, nA, nB, C, ldc, Ct, ldct);
}

static int lap_max_impl(mcd_handle h, const double* W, int64_t n, int64_t m, int64_t ldw, int32_t* col4row,
                        double* objective, double* prices, double* cert_out) {
  if (!h) return MCD_ERR_INVALID;
  if (!W || !col4row || n < 0 || m < n || ldw < m) return mcd_fail(h, MCD_ERR_INVALID, "mcd_lap_max arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  void* work = nullptr;
  int st = mcd_ws(h, WS_LAP, mcd_lap_workspace_bytes(n, m) + 512, &work);
  if (st) return st;
  // counters and certificate live in the first 512 bytes of the slot
  mcd_lap_counters* cnt = static_cast<mcd_lap_counters*>(work);
  mcd_lap_cert* cert = reinterpret_cast<mcd_lap_cert*>(static_cast<char*>(work) + 256);
  MCD_CUDA(h, cudaMemsetAsync(work, 0, 512, h->stream));
  int st2 = mcd_launch_lap(h, W, n, m, ldw, col4row, objective, static_cast<char*>(work) + 512, cnt, true, cert, prices);
  if (st2) return st2;
  mcd_lap_counters hc;
  mcd_lap_cert hcert;
  MCD_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof hc, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(&hcert, cert, sizeof hcert, cudaMemcpyDeviceToHost, h->stream));
  if ((st = mcd_check_finite(h))) return st;  // synchronises
  if (cert_out) {
    cert_out[0] = hcert.rel_gap;
    cert_out[1] = hcert.gap;
    cert_out[2] = hcert.max_violation;
    cert_out[3] = (double)hcert.n_bad;
  }
  if (n > 0 && (hc.status & 1)) return mcd_fail(h, MCD_ERR_NOT_CONVERGED, "a person was left unassigned");
  if (n > 0 && (hc.status & 2)) return mcd_fail(h, MCD_ERR_NOT_CONVERGED, "the dual certificate of the assignment failed");
  return MCD_OK;
}

int mcd_lap_certify(mcd_handle h, const double* W, int64_t n, int64_t m, int64_t ldw, const int32_t* col4row,
                    const double* prices, double* cert_out) {
  if (!h) return MCD_ERR_INVALID;
  if (!W || !col4row || !prices || !cert_out || n < 1 || m < n || ldw < m || m > 0x3fffffff)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_lap_certify arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  void* work = nullptr;
  int st = mcd_ws(h, WS_LAP, 512 + (size_t)m * 4 + (size_t)n * 8 + 1024, &work);
  if (st) return st;
  mcd_lap_counters* cnt = static_cast<mcd_lap_counters*>(work);
  mcd_lap_cert* cert = reinterpret_cast<mcd_lap_cert*>(static_cast<char*>(work) + 256);
  MCD_CUDA(h, cudaMemsetAsync(work, 0, 512, h->stream));
  if ((st = mcd_launch_lap_certify(h, W, n, m, ldw, col4row, prices, static_cast<char*>(work) + 512, cert, cnt)))
    return st;
  mcd_lap_cert hcert;
  MCD_CUDA(h, cudaMemcpyAsync(&hcert, cert, sizeof hcert, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  cert_out[0] = hcert.rel_gap;
  cert_out[1] = hcert.gap;
  cert_out[2] = hcert.max_violation;
  cert_out[3] = (double)hcert.n_bad;
  return MCD_OK;
}

int mcd_lap_max(mcd_handle h, const double* W, int64_t n, int64_t m, int64_t ldw, int32_t* col4row,
                double* objective) {
  return lap_max_impl(h, W, n, m, ldw, col4row, objective, nullptr, nullptr);
}

int mcd_lap_max_certified(mcd_handle h, const double* W, int64_t n, int64_t m, int64_t ldw, int32_t* col4row,
                          double* objective, double* prices, double* cert) {
  return lap_max_impl(h, W, n, m, ldw, col4row, objective, prices, cert);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// K4: step-loop bookkeeping kernels (reference macrodna.py:110-145, O(M) instead of dense M x N)
// ------------------------------------------------------------------------------------------------
namespace {

__global__ void iota_flags_kernel(int* act, int* flag, int* assign, int* step, int M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) {
    act[i] = i;
    flag[i] = 1;
    assign[i] = -1;
    step[i] = 0;
  }
}

// W[j, k] = Ct[j, act[k]]  (persons = DNA rows of Ct, objects = still-unassigned RNA cells)
__global__ void gather_cols_kernel(const double* __restrict__ Ct, int64_t ldct, const int* __restrict__ act, int R,
                                   double* __restrict__ W, int64_t ldw, int64_t nrows) {
  for (int64_t j = blockIdx.y; j < nrows; j += gridDim.y) {  // gridDim.y is capped at 65535
    const double* src = Ct + j * ldct;
    double* dst = W + j * ldw;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < R; k += gridDim.x * blockDim.x) dst[k] = src[act[k]];
  }
}

// W[k, :] = C[act[k], :]  (persons = still-unassigned RNA cells, objects = DNA cells)
__global__ void gather_rows_kernel(const double* __restrict__ C, int64_t ldc, const int* __restrict__ act, int N,
                                   double* __restrict__ W, int64_t ldw, int64_t nrows) {
  for (int64_t k = blockIdx.y; k < nrows; k += gridDim.y) {
    const double* src = C + (int64_t)act[k] * ldc;
    double* dst = W + k * ldw;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) dst[j] = src[j];
  }
}

// persons were DNA cells: col4row[j] = index into act
__global__ void record_dna_major_kernel(const int* __restrict__ col4row, int N, const int* __restrict__ act,
                                        int* assign, int* step, int* flag, int tag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) {
    const int k = col4row[j];
    if (k >= 0) {
      const int g = act[k];
      assign[g] = j;
      step[g] = tag;
      flag[g] = 0;
    }
  }
}
// persons were RNA cells: col4row[k] = DNA column
__global__ void record_rna_major_kernel(const int* __restrict__ col4row, int R, const int* __restrict__ act,
                                        int* assign, int* step, int* flag, int tag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < R) {
    const int j = col4row[k];
    if (j >= 0) {
      const int g = act[k];
      assign[g] = j;
      step[g] = tag;
      flag[g] = 0;
    }
  }
}

// Ordered stream compaction of the still-unassigned RNA rows (ascending, macrodna.py:141-145). One CTA.
__global__ void __launch_bounds__(1024) compact_active_kernel(const int* __restrict__ flag, int M, int* act_out) {
  __shared__ int sums[1024];
  const int t = threadIdx.x;
  const int per = (M + 1023) / 1024;
  const int lo = min(M, t * per), hi = min(M, lo + per);
  int c = 0;
  for (int i = lo; i < hi; ++i) c += flag[i];
  sums[t] = c;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    const int v = t >= o ? sums[t - o] : 0;
    __syncthreads();
    sums[t] += v;
    __syncthreads();
  }
  int pos = sums[t] - c;
  for (int i = lo; i < hi; ++i)
    if (flag[i]) act_out[pos++] = i;
}

// dst[c, r] = src[r, c] through a padded 32x32 shared tile (coalesced on both sides)
__global__ void transpose_f64_kernel(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t lds,
                                     double* __restrict__ dst, int64_t ldd) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[c * ldd + r] = tile[threadIdx.x][i];
  }
}

__global__ void gather_match_kernel(const double* __restrict__ C, int64_t ldc, const int* __restrict__ assign, int64_t M,
                                    double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) {
    const int j = assign[i];
    out[i] = j >= 0 ? C[i * ldc + j] : 0.0;
  }
}

// sub[i, j] = C[rows[i], cols[j]]  (rows / cols may repeat: resampled replicates carry duplicate DNA cells)
__global__ void gather_sub_kernel(const double* __restrict__ C, int64_t ldc, const int* __restrict__ rows,
                                  const int* __restrict__ cols, int64_t m, int64_t n, double* __restrict__ sub,
                                  int64_t lds) {
  for (int64_t i = blockIdx.y; i < m; i += gridDim.y) {
    const double* src = C + (int64_t)(rows ? rows[i] : i) * ldc;
    double* dst = sub + i * lds;
    for (int64_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
      dst[j] = src[cols ? cols[j] : j];
  }
}

__global__ void gather_rows_out_kernel(const double* __restrict__ C, int64_t ldc, const int* __restrict__ rows,
                                       int64_t nrows, int64_t n, double* __restrict__ out) {
  for (int64_t i = blockIdx.y; i < nrows; i += gridDim.y) {
    const double* src = C + (int64_t)rows[i] * ldc;
    for (int64_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
      out[i * n + j] = src[j];
  }
}

__global__ void gather_pairs_kernel(const double* __restrict__ C, int64_t ldc, const int* __restrict__ rows,
                                    const int* __restrict__ cols, int64_t n, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = C[(int64_t)rows[k] * ldc + cols[k]];
}

// cls[j] = first position of cols that holds the same cell as position j (copies of a cell share an id < n); returns
// the number of extra copies.  O(n log n).
int64_t duplicate_classes(const int32_t* cols, int64_t n, std::vector<int>& cls) {
  std::vector<int64_t> order((size_t)n);
  for (int64_t j = 0; j < n; ++j) order[(size_t)j] = j;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cols[a] < cols[b]; });
  cls.assign((size_t)n, 0);
  int64_t extra = 0;
  for (int64_t q = 0; q < n; ++q) {
    const int64_t j = order[(size_t)q];
    if (q > 0 && cols[j] == cols[order[(size_t)q - 1]]) {
      cls[(size_t)j] = cls[(size_t)order[(size_t)q - 1]];
      ++extra;
    } else {
      cls[(size_t)j] = (int)j;
    }
  }
  return extra;
}

// rows of a gather go to gridDim.y, which the hardware caps at 65535: the kernels loop over the rest
inline unsigned grid_rows(int64_t rows) { return (unsigned)(rows < 65535 ? (rows < 1 ? 1 : rows) : 65535); }

cudaEvent_t get_event(mcd_context* h, size_t idx) {
  while (h->ev.size() <= idx) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev.push_back(e);
  }
  return h->ev[idx];
}

struct StepPlan {
  int64_t nsteps;
  size_t w_bytes;
  size_t lap_bytes;
};

StepPlan plan_steps(int64_t M, int64_t N) {
  StepPlan p;
  p.nsteps = mcd_num_steps(M, N);
  p.w_bytes = 0;
  p.lap_bytes = 0;
  int64_t R = M;
  for (int64_t s = 0; s < p.nsteps; ++s, R -= N) {
    const int64_t n = R > N ? N : R;
    const int64_t m = R > N ? R : N;
    const size_t lb = mcd_lap_workspace_bytes(n, m);
    if (lb > p.lap_bytes) p.lap_bytes = lb;
    if (s > 0) {
      const size_t wb = (size_t)n * (size_t)((m + 1) & ~1LL) * 8;
      if (wb > p.w_bytes) p.w_bytes = wb;
    }
  }
  return p;
}

// Device-side outputs of one step loop, carved out of one block: assign[M], step[M], obj[nsteps], one counter
// block and one dual certificate per step.
struct StepOut {
  int* assign;
  int* step;
  double* obj;
  mcd_lap_counters* cnt;
  mcd_lap_cert* cert;
};
size_t step_out_bytes(int64_t M, int64_t nsteps) {
  const size_t mi = ((size_t)M * 4 + 255) / 256 * 256;
  const size_t ob = ((size_t)nsteps * 8 + 255) / 256 * 256;
  return 2 * mi + ob + (sizeof(mcd_lap_counters) + sizeof(mcd_lap_cert)) * (size_t)nsteps;
}
StepOut carve_step_out(void* base, int64_t M, int64_t nsteps) {
  const size_t mi = ((size_t)M * 4 + 255) / 256 * 256;
  const size_t ob = ((size_t)nsteps * 8 + 255) / 256 * 256;
  char* mb = static_cast<char*>(base);
  StepOut o;
  o.assign = reinterpret_cast<int*>(mb);
  o.step = reinterpret_cast<int*>(mb + mi);
  o.obj = reinterpret_cast<double*>(mb + 2 * mi);
  o.cnt = reinterpret_cast<mcd_lap_counters*>(mb + 2 * mi + ob);
  o.cert = reinterpret_cast<mcd_lap_cert*>(mb + 2 * mi + ob + sizeof(mcd_lap_counters) * (size_t)nsteps);
  return o;
}

// Host side of the per-step records: fills the solver part of `stats`, prints the debug lines, returns the OR of
// the steps' status bits (1 = a cell left unassigned, 2 = dual certificate failed).
int fold_step_records(mcd_context* h, int64_t M, int64_t N, int64_t nsteps, const mcd_lap_counters* hc,
                      const mcd_lap_cert* hcert, mcd_stats* stats, size_t ev_lap, bool have_events) {
  int bad = 0;
  if (h->opt.debug) {
    int64_t R = M;
    for (int64_t s = 0; s < nsteps; ++s, R -= N) {
      const int64_t mm = R > N ? R : N;
      fprintf(stderr, "[lap step %lld] n=%lld m=%lld rounds=%lld bids=%lld sweeps=%lld aug=%lld/%lld cert_gap=%.3e cyc=",
              (long long)s, (long long)(R > N ? N : R), (long long)mm, hc[s].rounds, hc[s].bids, hc[s].bytes / (mm * 8),
              hc[s].aug_rows, hc[s].aug_steps, hcert[s].rel_gap);
      for (int q = 0; q < 8; ++q) fprintf(stderr, "%lld ", hc[s].t_phase[q]);
      fprintf(stderr, " async_us=");
      for (int q = 0; q < 8; ++q) fprintf(stderr, "%lld ", hc[s].a_ts[q] / 1000);
      fprintf(stderr, " async_kbids=");
      for (int q = 0; q < 8; ++q) fprintf(stderr, "%lld ", hc[s].a_hops[q] / 1000);
      fprintf(stderr, "\n");
    }
  }
  for (int64_t s = 0; s < nsteps; ++s) {
    bad |= hc[s].status;
    if (!stats) continue;
    stats->lap_rounds += hc[s].rounds;
    stats->lap_bids += hc[s].bids;
    stats->lap_bytes += hc[s].bytes;
    stats->lap_aug_rows += hc[s].aug_rows;
    stats->lap_aug_steps += hc[s].aug_steps;
    for (int q = 0; q < 8; ++q) stats->lap_cycles[q] += hc[s].t_phase[q];
    if (h->opt.certify) {
      if (hcert[s].rel_gap > stats->cert_rel_gap) stats->cert_rel_gap = hcert[s].rel_gap;
      if (hcert[s].max_violation > stats->cert_max_violation) stats->cert_max_violation = hcert[s].max_violation;
      stats->cert_bad += hcert[s].n_bad;
      stats->cert_steps += 1;
    }
    if (s < MCD_MAX_STEP_STATS) {
      if (have_events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, get_event(h, ev_lap + s), get_event(h, ev_lap + s + 1));
        stats->step_ms[s] = ms;
      }
      stats->step_rounds[s] = hc[s].rounds;
      stats->step_bids[s] = hc[s].bids;
      stats->step_cert_gap[s] = h->opt.certify ? hcert[s].rel_gap : -1.0;
    }
  }
  return bad;
}

int step_status(mcd_context* h, int bad) {
  if (bad & 1) return mcd_fail(h, MCD_ERR_NOT_CONVERGED, "a step left an RNA/DNA cell unassigned");
  if (bad & 2) return mcd_fail(h, MCD_ERR_NOT_CONVERGED, "the dual certificate of a step's assignment failed");
  return MCD_OK;
}

// Enqueue the whole step loop on h->stream (no host synchronisation: every step's shape is known
// up front because each non-final step matches exactly N RNA cells).
// Device outputs: d_assign[M], d_step[M], d_obj[nsteps], d_counters[nsteps].
// dna_class: optional DEVICE int [N], equal ids = DNA cells that are copies of one another (identical columns of C):
// in the steps where the DNA cells are the solver's persons they bid as a class (see mcd_launch_lap).
int enqueue_step_loop(mcd_context* h, const double* C, int64_t ldc, const double* Ct, int64_t ldct, int64_t M,
                      int64_t N, const StepOut& out, size_t ev_base, bool check_finite, bool record_events = true,
                      const int* dna_class = nullptr) {
  int* d_assign = out.assign;
  int* d_step = out.step;
  double* d_obj = out.obj;
  mcd_lap_counters* d_counters = out.cnt;
  const StepPlan plan = plan_steps(M, N);
  void* wbuf = nullptr;
  void* lapbuf = nullptr;
  void* stepbuf = nullptr;
  int st;
  if ((st = mcd_ws(h, WS_W, plan.w_bytes, &wbuf))) return st;
  if ((st = mcd_ws(h, WS_LAP, plan.lap_bytes, &lapbuf))) return st;
  const size_t mi = ((size_t)M * 4 + 255) / 256 * 256;
  const size_t ci = ((size_t)(M > N ? M : N) * 4 + 255) / 256 * 256;
  if ((st = mcd_ws(h, WS_STEP, 3 * mi + ci, &stepbuf))) return st;
  char* sb = static_cast<char*>(stepbuf);
  int* act[2] = {reinterpret_cast<int*>(sb), reinterpret_cast<int*>(sb + mi)};
  int* flag = reinterpret_cast<int*>(sb + 2 * mi);
  int* col4row = reinterpret_cast<int*>(sb + 3 * mi);
  double* W = static_cast<double*>(wbuf);

  MCD_CUDA(h, cudaMemsetAsync(d_counters, 0, (sizeof(mcd_lap_counters) + sizeof(mcd_lap_cert)) * plan.nsteps, h->stream));
  iota_flags_kernel<<<(unsigned)((M + 255) / 256), 256, 0, h->stream>>>(act[0], flag, d_assign, d_step, (int)M);
  MCD_LAUNCH_CHECK(h, "iota_flags_kernel");

  int64_t R = M;
  int cur = 0;
  for (int64_t s = 0; s < plan.nsteps; ++s) {
    if (record_events && s < MCD_MAX_STEP_STATS) MCD_CUDA(h, cudaEventRecord(get_event(h, ev_base + s), h->stream));
    if (R > N) {
      // all N DNA cells take one RNA cell each (macrodna.py:29,53 with n_min = N)
      const double* Wp;
      int64_t ldw;
      if (s == 0) {
        Wp = Ct;
        ldw = ldct;
      } else {
        ldw = (R + 1) & ~1LL;
        dim3 grid((unsigned)((R + 1023) / 1024 < 64 ? (R + 1023) / 1024 : 64), grid_rows(N));
        gather_cols_kernel<<<grid, 256, 0, h->stream>>>(Ct, ldct, act[cur], (int)R, W, ldw, N);
        MCD_LAUNCH_CHECK(h, "gather_cols_kernel");
        Wp = W;
      }
      if ((st = mcd_launch_lap(h, Wp, N, R, ldw, col4row, d_obj + s, lapbuf, d_counters + s, check_finite && s == 0,
                               out.cert + s, nullptr, dna_class)))
        return st;
      record_dna_major_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(col4row, (int)N, act[cur], d_assign,
                                                                                 d_step, flag, (int)(s + 1));
      MCD_LAUNCH_CHECK(h, "record_dna_major_kernel");
      if (s + 1 < plan.nsteps) {
        compact_active_kernel<<<1, 1024, 0, h->stream>>>(flag, (int)M, act[cur ^ 1]);
        MCD_LAUNCH_CHECK(h, "compact_active_kernel");
        cur ^= 1;
      }
    } else {
      // last step: every remaining RNA cell takes a DNA cell (n_min = |R|)
      const double* Wp;
      int64_t ldw;
      if (s == 0) {
        Wp = C;
        ldw = ldc;
      } else {
        ldw = (N + 1) & ~1LL;
        dim3 grid((unsigned)((N + 1023) / 1024 < 64 ? (N + 1023) / 1024 : 64), grid_rows(R));
        gather_rows_kernel<<<grid, 256, 0, h->stream>>>(C, ldc, act[cur], (int)N, W, ldw, R);
        MCD_LAUNCH_CHECK(h, "gather_rows_kernel");
        Wp = W;
      }
      if ((st = mcd_launch_lap(h, Wp, R, N, ldw, col4row, d_obj + s, lapbuf, d_counters + s, check_finite && s == 0,
                               out.cert + s)))
        return st;
      record_rna_major_kernel<<<(unsigned)((R + 255) / 256), 256, 0, h->stream>>>(col4row, (int)R, act[cur], d_assign,
                                                                                 d_step, flag, (int)(s + 1));
      MCD_LAUNCH_CHECK(h, "record_rna_major_kernel");
    }
    R -= N;
  }
  if (record_events)
    MCD_CUDA(h, cudaEventRecord(
        get_event(h, ev_base + (plan.nsteps < MCD_MAX_STEP_STATS ? plan.nsteps : MCD_MAX_STEP_STATS)), h->stream));
  return MCD_OK;
}

}  // namespace

extern "C" {

int mcd_transpose_f64(mcd_handle h, const double* src, int64_t rows, int64_t cols, int64_t lds, double* dst,
                      int64_t ldd) {
  if (!h) return MCD_ERR_INVALID;
  if (!src || !dst || rows < 0 || cols < 0 || lds < cols || ldd < rows)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_transpose_f64 arguments");
  if (rows == 0 || cols == 0) return MCD_OK;
  MCD_CUDA(h, cudaSetDevice(h->device));
  const int64_t gy = (rows + 31) / 32;
  if (gy > 65535) {
    // split the row range so gridDim.y stays legal
    for (int64_t r0 = 0; r0 < rows; r0 += 65535LL * 32) {
      const int64_t nr = rows - r0 < 65535LL * 32 ? rows - r0 : 65535LL * 32;
      dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((nr + 31) / 32));
      transpose_f64_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(src + r0 * lds, nr, cols, lds, dst + r0, ldd);
      MCD_LAUNCH_CHECK(h, "transpose_f64_kernel");
    }
    return MCD_OK;
  }
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)gy);
  transpose_f64_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(src, rows, cols, lds, dst, ldd);
  MCD_LAUNCH_CHECK(h, "transpose_f64_kernel");
  return MCD_OK;
}

int mcd_last_match_values(mcd_handle h, double* out, int64_t M, int out_space) {
  if (!h) return MCD_ERR_INVALID;
  if (!out || M < 1 || M != h->last_M || h->last_assign == nullptr)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_last_match_values: no matching mcd_cell2cell result is resident");
  MCD_CUDA(h, cudaSetDevice(h->device));
  const double* C = static_cast<const double*>(h->ws[WS_C].ptr);
  double* tmp = out;
  void* scratch = nullptr;
  if (out_space != MCD_MEM_DEVICE) {
    int st = mcd_ws(h, WS_W, (size_t)M * 8, &scratch);  // the step-loop block buffer is free between calls
    if (st) return st;
    tmp = static_cast<double*>(scratch);
  }
  gather_match_kernel<<<(unsigned)((M + 255) / 256), 256, 0, h->stream>>>(C, h->last_ldc, h->last_assign, M, tmp);
  MCD_LAUNCH_CHECK(h, "gather_match_kernel");
  if (out_space != MCD_MEM_DEVICE)
    MCD_CUDA(h, cudaMemcpyAsync(out, tmp, (size_t)M * 8, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  return MCD_OK;
}

int mcd_lap_steps(mcd_handle h, const double* C, int64_t ldc, const double* Ct, int64_t ldct, int64_t M, int64_t N,
                  int32_t* assign, int32_t* step, double* step_obj, int out_space, mcd_stats* stats) {
  if (!h) return MCD_ERR_INVALID;
  if (!C || !Ct || !assign || !step || M < 1 || N < 1 || ldc < N || ldct < M || M > 0x3fffffff || N > 0x3fffffff)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_lap_steps arguments");
  MCD_CUDA(h, cudaSetDevice(h->device));
  h->last_M = 0;  // the resident cell2cell result (if any) is about to be overwritten
  h->last_assign = nullptr;
  const int64_t nsteps = mcd_num_steps(M, N);
  void* misc = nullptr;
  int st = mcd_ws(h, WS_MISC, step_out_bytes(M, nsteps), &misc);
  if (st) return st;
  const StepOut out = carve_step_out(misc, M, nsteps);
  const int64_t launches0 = h->launches;
  const size_t EV_LAP = 8;
  MCD_CUDA(h, cudaEventRecord(get_event(h, 6), h->stream));
  if ((st = enqueue_step_loop(h, C, ldc, Ct, ldct, M, N, out, EV_LAP, true))) return st;
  MCD_CUDA(h, cudaEventRecord(get_event(h, 7), h->stream));

  const cudaMemcpyKind kind = out_space == MCD_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  MCD_CUDA(h, cudaMemcpyAsync(assign, out.assign, (size_t)M * 4, kind, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(step, out.step, (size_t)M * 4, kind, h->stream));
  if (step_obj) MCD_CUDA(h, cudaMemcpyAsync(step_obj, out.obj, (size_t)nsteps * 8, kind, h->stream));
  std::vector<mcd_lap_counters> hc((size_t)nsteps);
  std::vector<mcd_lap_cert> hcert((size_t)nsteps);
  int flag = 0;
  MCD_CUDA(h, cudaMemcpyAsync(hc.data(), out.cnt, sizeof(mcd_lap_counters) * nsteps, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(hcert.data(), out.cert, sizeof(mcd_lap_cert) * nsteps, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  if (flag) return mcd_fail(h, MCD_ERR_NONFINITE, "NaN or Inf in the correlation matrix");
  if (stats) memset(stats, 0, sizeof *stats);
  const int bad = fold_step_records(h, M, N, nsteps, hc.data(), hcert.data(), stats, EV_LAP, true);
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, get_event(h, 6), get_event(h, 7));
    stats->ms_lap = ms;
    stats->ms_total = ms;
    stats->n_steps = nsteps;
    stats->kernel_launches = h->launches - launches0;
  }
  return step_status(h, bad);
}

int mcd_cell2cell(mcd_handle h, const double* rna, int64_t ld_rna, const double* dna, int64_t ld_dna, int64_t M,
                  int64_t N, int64_t G, int in_space, int precision, int32_t* assign, int32_t* step, double* step_obj,
                  double* corr_out, int out_space, mcd_stats* stats) {
  return mcd_cell2cell_gather(h, rna, ld_rna, nullptr, dna, ld_dna, nullptr, M, N, G, in_space, precision, assign, step,
                              step_obj, corr_out, out_space, stats);
}

int mcd_cell2cell_gather(mcd_handle h, const double* rna, int64_t ld_rna, const int32_t* rna_gene_idx,
                         const double* dna, int64_t ld_dna, const int32_t* dna_gene_idx, int64_t M, int64_t N,
                         int64_t G, int in_space, int precision, int32_t* assign, int32_t* step, double* step_obj,
                         double* corr_out, int out_space, mcd_stats* stats) {
  if (!h) return MCD_ERR_INVALID;
  if (!rna || !dna || !assign || !step || M < 1 || N < 1 || G < 1 || (!rna_gene_idx && ld_rna < G) ||
      (!dna_gene_idx && ld_dna < G) || ld_rna < 1 || ld_dna < 1 || M > 0x3fffffff || N > 0x3fffffff || G > 0x7fffffff)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_cell2cell arguments");
  for (int64_t g = 0; g < G; ++g) {
    if ((rna_gene_idx && (rna_gene_idx[g] < 0 || rna_gene_idx[g] >= ld_rna)) ||
        (dna_gene_idx && (dna_gene_idx[g] < 0 || dna_gene_idx[g] >= ld_dna)))
      return mcd_fail(h, MCD_ERR_INVALID, "mcd_cell2cell_gather: gene index out of range");
  }
  if (precision != MCD_PREC_FP64 && precision != MCD_PREC_SPLIT_FP16 && precision != MCD_PREC_OZAKI_INT8)
    return mcd_fail(h, MCD_ERR_INVALID, "unknown precision");
  const int nsl = mcd_ozaki_slices(h, M, N, G);
  // exact int32 accumulation bounds the gene count of the integer path; longer rows use the FP64 pipe
  if (precision == MCD_PREC_OZAKI_INT8 && (double)nsl * 4096.0 * (double)mcd_padded_k_split(G) >= 2147483648.0)
    precision = MCD_PREC_FP64;
  MCD_CUDA(h, cudaSetDevice(h->device));
  h->last_M = 0;  // whatever was resident is about to be overwritten (also on an early error return)
  h->last_assign = nullptr;
  const int64_t launches0 = h->launches;
  enum { EV_T0 = 0, EV_H2D, EV_STD, EV_CORR, EV_LAPEND, EV_D2H, EV_LAP = 8 };
  const size_t EV_DNA_K1 = 90;  // 2 events around the DNA operand's K1 launch
  int st;
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_T0), h->stream));

  // ---- outputs of K1/K2
  const int64_t ldc = (N + 1) & ~1LL, ldct = (M + 1) & ~1LL;
  void *pC = nullptr, *pCt = nullptr, *pnA = nullptr, *pnB = nullptr;
  if ((st = mcd_ws(h, WS_C, (size_t)M * ldc * 8, &pC))) return st;
  if ((st = mcd_ws(h, WS_CT, (size_t)N * ldct * 8, &pCt))) return st;
  if ((st = mcd_ws(h, WS_NORM_A, (size_t)M * 8, &pnA))) return st;
  if ((st = mcd_ws(h, WS_NORM_B, (size_t)N * 8, &pnB))) return st;
  double* C = static_cast<double*>(pC);
  double* Ct = static_cast<double*>(pCt);
  double* nA = static_cast<double*>(pnA);
  double* nB = static_cast<double*>(pnB);
  const int64_t ldk = precision == MCD_PREC_FP64 ? mcd_padded_k(G) : mcd_padded_k_split(G);
  void *pa = nullptr, *pb = nullptr;
  double *sA = nullptr, *sB = nullptr;
  if (precision == MCD_PREC_FP64) {
    if ((st = mcd_ws(h, WS_RNA_C, (size_t)M * ldk * 8, &pa))) return st;
    if ((st = mcd_ws(h, WS_DNA_C, (size_t)N * ldk * 8, &pb))) return st;
  } else if (precision == MCD_PREC_OZAKI_INT8) {
    void* ps = nullptr;
    if ((st = mcd_ws(h, WS_SLICES_A, (size_t)nsl * M * ldk, &pa))) return st;
    if ((st = mcd_ws(h, WS_SLICES_B, (size_t)nsl * N * ldk, &pb))) return st;
    if ((st = mcd_ws(h, WS_SCALE, (size_t)(M + N) * 8, &ps))) return st;
    sA = static_cast<double*>(ps);
    sB = sA + M;
  } else {
    if ((st = mcd_ws(h, WS_SLICES_A, (size_t)2 * M * ldk * 2, &pa))) return st;
    if ((st = mcd_ws(h, WS_SLICES_B, (size_t)2 * N * ldk * 2, &pb))) return st;
  }
  uint16_t* a_hi = static_cast<uint16_t*>(pa);
  uint16_t* a_lo = a_hi + M * ldk;
  uint16_t* b_hi = static_cast<uint16_t*>(pb);
  uint16_t* b_lo = b_hi + N * ldk;

  // ---- stage inputs.  Host inputs: the DNA operand first, then the RNA rows in chunks on the copy stream,
  //      so K1 + K2 of chunk c overlap the H2D copy of chunk c+1 (chunks are tile aligned: same results).
  const double* d_rna = rna;
  const double* d_dna = dna;
  int64_t ldr = ld_rna, ldd = ld_dna;
  int nchunk = 1;
  bool pageable_in = false;
  // gene gather indices (device copies); with a gather the staged rows keep all ld columns of the host block
  const int* d_ridx = nullptr;
  const int* d_didx = nullptr;
  if (rna_gene_idx || dna_gene_idx) {
    void* pg = nullptr;
    if ((st = mcd_ws(h, WS_GIDX, (size_t)2 * G * 4, &pg))) return st;
    int* gi = static_cast<int*>(pg);
    if (rna_gene_idx) {
      MCD_CUDA(h, cudaMemcpyAsync(gi, rna_gene_idx, (size_t)G * 4, cudaMemcpyHostToDevice, h->stream));
      d_ridx = gi;
    }
    if (dna_gene_idx) {
      MCD_CUDA(h, cudaMemcpyAsync(gi + G, dna_gene_idx, (size_t)G * 4, cudaMemcpyHostToDevice, h->stream));
      d_didx = gi + G;
    }
  }
  const int64_t wr = rna_gene_idx ? ld_rna : G;  // columns staged per RNA row
  const int64_t wd = dna_gene_idx ? ld_dna : G;
  if (in_space == MCD_MEM_HOST) {
    void *pr = nullptr, *pd = nullptr;
    if ((st = mcd_ws(h, WS_RNA_IN, (size_t)M * wr * 8, &pr))) return st;
    if ((st = mcd_ws(h, WS_DNA_IN, (size_t)N * wd * 8, &pd))) return st;
    pageable_in = mcd_is_pageable(rna) || mcd_is_pageable(dna);
    if (pageable_in) {
      if ((st = mcd_staged_h2d(h, static_cast<double*>(pd), wd, dna, ld_dna, wd, N, h->stream))) return st;
    } else {
      MCD_CUDA(h, cudaMemcpy2DAsync(pd, (size_t)wd * 8, dna, (size_t)ld_dna * 8, (size_t)wd * 8, (size_t)N,
                                    cudaMemcpyHostToDevice, h->stream));
    }
    d_rna = static_cast<const double*>(pr);
    d_dna = static_cast<const double*>(pd);
    ldr = wr;
    ldd = wd;
    const double bytes = (double)M * wr * 8;
    nchunk = (int)(bytes / (768.0 * 1024 * 1024)) + 1;
    if (nchunk > 16) nchunk = 16;
  }
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_H2D), h->stream));  // DNA operand resident
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_DNA_K1), h->stream));
  // DNA operand: K1 once
  mcd_ozaki_out ozb, oza;
  ozb.digits = static_cast<int8_t*>(pb);
  ozb.ldk8 = ldk;
  ozb.slice_stride = N * ldk;
  ozb.nsl = nsl;
  ozb.scale = sB;
  oza = ozb;
  oza.slice_stride = M * ldk;
  if (precision == MCD_PREC_FP64)
    st = mcd_launch_standardize(h, d_dna, N, G, ldd, (double*)pb, ldk, nullptr, nullptr, 0, nB, d_didx);
  else if (precision == MCD_PREC_OZAKI_INT8)
    st = mcd_launch_standardize(h, d_dna, N, G, ldd, nullptr, 0, nullptr, nullptr, 0, nB, d_didx, &ozb);
  else
    st = mcd_launch_standardize(h, d_dna, N, G, ldd, nullptr, 0, b_hi, b_lo, ldk, nB, d_didx);
  if (st) return st;
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_DNA_K1 + 1), h->stream));
  int64_t rows_per = ((M + nchunk - 1) / nchunk + 127) / 128 * 128;
  if (rows_per < 128) rows_per = 128;
  const size_t EV_CHUNK = 100;  // 4 events per chunk: copy done, K1 start, K1 end, K2 end
  int nchunk_used = 0;
  if (in_space == MCD_MEM_HOST) MCD_CUDA(h, cudaStreamWaitEvent(h->copy_stream, get_event(h, EV_T0), 0));
  for (int64_t r0 = 0; r0 < M; r0 += rows_per, ++nchunk_used) {
    const int64_t mr = (M - r0 < rows_per) ? (M - r0) : rows_per;
    const int c = nchunk_used;
    if (in_space == MCD_MEM_HOST) {
      double* dst = const_cast<double*>(d_rna) + r0 * wr;
      if (pageable_in) {
        if ((st = mcd_staged_h2d(h, dst, wr, rna + r0 * ld_rna, ld_rna, wr, mr, h->copy_stream))) return st;
      } else {
        MCD_CUDA(h, cudaMemcpy2DAsync(dst, (size_t)wr * 8, rna + r0 * ld_rna, (size_t)ld_rna * 8, (size_t)wr * 8,
                                      (size_t)mr, cudaMemcpyHostToDevice, h->copy_stream));
      }
      MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CHUNK + 4 * c), h->copy_stream));
      MCD_CUDA(h, cudaStreamWaitEvent(h->stream, get_event(h, EV_CHUNK + 4 * c), 0));
    } else {
      MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CHUNK + 4 * c), h->stream));
    }
    MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CHUNK + 4 * c + 1), h->stream));
    const double* xr = d_rna + r0 * ldr;
    if (precision == MCD_PREC_FP64) {
      double* ac = (double*)pa + r0 * ldk;
      if ((st = mcd_launch_standardize(h, xr, mr, G, ldr, ac, ldk, nullptr, nullptr, 0, nA + r0, d_ridx))) return st;
      MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CHUNK + 4 * c + 2), h->stream));
      if ((st = mcd_launch_corr_fp64(h, ac, mr, (double*)pb, N, ldk, nA + r0, nB, C + r0 * ldc, ldc, Ct + r0, ldct)))
        return st;
    } else if (precision == MCD_PREC_OZAKI_INT8) {
      oza.digits = static_cast<int8_t*>(pa) + r0 * ldk;
      oza.scale = sA + r0;
      if ((st = mcd_launch_standardize(h, xr, mr, G, ldr, nullptr, 0, nullptr, nullptr, 0, nA + r0, d_ridx, &oza)))
        return st;
      MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CHUNK + 4 * c + 2), h->stream));
      if ((st = mcd_launch_corr_ozaki(h, oza.digits, oza.slice_stride, mr, ozb.digits, ozb.slice_stride, N, ldk, nsl,
                                      sA + r0, sB, nA + r0, nB, C + r0 * ldc, ldc, Ct + r0, ldct)))
        return st;
    } else {
      if ((st = mcd_launch_standardize(h, xr, mr, G, ldr, nullptr, 0, a_hi + r0 * ldk, a_lo + r0 * ldk, ldk, nA + r0,
                                       d_ridx)))
        return st;
      MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CHUNK + 4 * c + 2), h->stream));
      if ((st = mcd_launch_corr_split(h, a_hi + r0 * ldk, a_lo + r0 * ldk, mr, b_hi, b_lo, N, ldk, nA + r0, nB,
                                      C + r0 * ldc, ldc, Ct + r0, ldct)))
        return st;
    }
    MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CHUNK + 4 * c + 3), h->stream));
  }
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_STD), h->stream));
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_CORR), h->stream));

  // ---- K3 + K4
  const int64_t nsteps = mcd_num_steps(M, N);
  void* misc = nullptr;
  if ((st = mcd_ws(h, WS_MISC, step_out_bytes(M, nsteps), &misc))) return st;
  const StepOut out = carve_step_out(misc, M, nsteps);
  if (h->opt.corr_only) {
    // the caller only wants the correlation matrix resident (it will solve views of it): no step loop
    MCD_CUDA(h, cudaMemsetAsync(misc, 0, step_out_bytes(M, nsteps), h->stream));
    MCD_CUDA(h, cudaMemsetAsync(out.assign, 0xFF, (size_t)M * 4, h->stream));
    for (size_t e = 0; e <= (size_t)(nsteps < MCD_MAX_STEP_STATS ? nsteps : MCD_MAX_STEP_STATS); ++e)
      MCD_CUDA(h, cudaEventRecord(get_event(h, EV_LAP + e), h->stream));
  } else if ((st = enqueue_step_loop(h, C, ldc, Ct, ldct, M, N, out, EV_LAP, false))) {
    return st;
  }
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_LAPEND), h->stream));

  // ---- outputs
  const cudaMemcpyKind kind = out_space == MCD_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  MCD_CUDA(h, cudaMemcpyAsync(assign, out.assign, (size_t)M * 4, kind, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(step, out.step, (size_t)M * 4, kind, h->stream));
  if (step_obj) MCD_CUDA(h, cudaMemcpyAsync(step_obj, out.obj, (size_t)nsteps * 8, kind, h->stream));
  if (corr_out)
    MCD_CUDA(h, cudaMemcpy2DAsync(corr_out, (size_t)N * 8, C, (size_t)ldc * 8, (size_t)N * 8, (size_t)M, kind,
                                  h->stream));
  std::vector<mcd_lap_counters> hc((size_t)nsteps);
  std::vector<mcd_lap_cert> hcert((size_t)nsteps);
  int flag = 0;
  MCD_CUDA(h, cudaMemcpyAsync(hc.data(), out.cnt, sizeof(mcd_lap_counters) * nsteps, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(hcert.data(), out.cert, sizeof(mcd_lap_cert) * nsteps, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemsetAsync(h->d_flags, 0, sizeof(int), h->stream));
  MCD_CUDA(h, cudaEventRecord(get_event(h, EV_D2H), h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));

  if (stats) memset(stats, 0, sizeof *stats);
  const int bad = h->opt.corr_only ? 0 : fold_step_records(h, M, N, nsteps, hc.data(), hcert.data(), stats, EV_LAP, true);
  if (stats) {
    auto el = [&](int a, int b) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, get_event(h, a), get_event(h, b));
      return (double)ms;
    };
    // K1 / K2 run per RNA chunk (interleaved with the H2D copies of the next chunk): sum the chunk spans, plus the
    // DNA operand's K1 launch.  ms_h2d is the EXPOSED copy time: everything of [T0, last K2] that is neither K1 nor K2.
    double k1 = el((int)EV_DNA_K1, (int)EV_DNA_K1 + 1), k2 = 0.0;
    for (int c = 0; c < nchunk_used; ++c) {
      k1 += el((int)EV_CHUNK + 4 * c + 1, (int)EV_CHUNK + 4 * c + 2);
      k2 += el((int)EV_CHUNK + 4 * c + 2, (int)EV_CHUNK + 4 * c + 3);
    }
    const double span = el(EV_T0, EV_CORR);
    stats->ms_standardize = k1;
    stats->ms_corr = k2;
    stats->ms_h2d = span - k1 - k2 > 0.0 ? span - k1 - k2 : 0.0;
    stats->ms_lap = el(EV_CORR, EV_LAPEND);
    stats->ms_d2h = el(EV_LAPEND, EV_D2H);
    stats->ms_total = el(EV_T0, EV_D2H);
    stats->n_steps = nsteps;
    stats->kernel_launches = h->launches - launches0;
  }
  if (flag) return mcd_fail(h, MCD_ERR_NONFINITE, "NaN or Inf in the expression / copy-number matrix");
  if ((st = step_status(h, bad))) return st;
  h->last_M = M;
  h->last_N = N;
  h->last_ldc = ldc;
  h->last_assign = out.assign;
  return MCD_OK;
}

// ------------------------------------------------------------------------------------------------
// Several GPUs, one caller.  SURVEY.md section 8e: K1 + K2 shard by RNA rows with no exchange inside the contraction;
// the correlation shards travel once to the first device (peer copies over NVLink), which runs the step loop.
// ------------------------------------------------------------------------------------------------
namespace {
struct ShardOperands {
  void* pa = nullptr;
  void* pb = nullptr;
  double* sA = nullptr;
  double* sB = nullptr;
  double* nA = nullptr;
  double* nB = nullptr;
};

// K1 of both operands + K2 of `rows` RNA rows on handle h; C_loc [rows, ldc] (no transpose).  All on h->stream.
int shard_standardize_correlate(mcd_context* h, const double* d_rna, int64_t ldr, const int* d_ridx, const double* d_dna,
                                int64_t ldd, const int* d_didx, int64_t rows, int64_t N, int64_t G, int precision, int nsl,
                                double* C_loc, int64_t ldc) {
  int st;
  ShardOperands o;
  void *pnA = nullptr, *pnB = nullptr;
  const int64_t ldk = precision == MCD_PREC_FP64 ? mcd_padded_k(G) : mcd_padded_k_split(G);
  if ((st = mcd_ws(h, WS_NORM_A, (size_t)(rows > 0 ? rows : 1) * 8, &pnA))) return st;
  if ((st = mcd_ws(h, WS_NORM_B, (size_t)N * 8, &pnB))) return st;
  o.nA = static_cast<double*>(pnA);
  o.nB = static_cast<double*>(pnB);
  const int64_t mr = rows > 0 ? rows : 1;
  if (precision == MCD_PREC_FP64) {
    if ((st = mcd_ws(h, WS_RNA_C, (size_t)mr * ldk * 8, &o.pa))) return st;
    if ((st = mcd_ws(h, WS_DNA_C, (size_t)N * ldk * 8, &o.pb))) return st;
    if ((st = mcd_launch_standardize(h, d_dna, N, G, ldd, (double*)o.pb, ldk, nullptr, nullptr, 0, o.nB, d_didx))) return st;
    if (rows > 0) {
      if ((st = mcd_launch_standardize(h, d_rna, rows, G, ldr, (double*)o.pa, ldk, nullptr, nullptr, 0, o.nA, d_ridx))) return st;
      st = mcd_launch_corr_fp64(h, (double*)o.pa, rows, (double*)o.pb, N, ldk, o.nA, o.nB, C_loc, ldc, nullptr, 0);
    }
  } else if (precision == MCD_PREC_OZAKI_INT8) {
    void* ps = nullptr;
    if ((st = mcd_ws(h, WS_SLICES_A, (size_t)nsl * mr * ldk, &o.pa))) return st;
    if ((st = mcd_ws(h, WS_SLICES_B, (size_t)nsl * N * ldk, &o.pb))) return st;
    if ((st = mcd_ws(h, WS_SCALE, (size_t)(mr + N) * 8, &ps))) return st;
    o.sA = static_cast<double*>(ps);
    o.sB = o.sA + mr;
    mcd_ozaki_out oza, ozb;
    ozb.digits = static_cast<int8_t*>(o.pb);
    ozb.ldk8 = ldk;
    ozb.slice_stride = N * ldk;
    ozb.nsl = nsl;
    ozb.scale = o.sB;
    oza = ozb;
    oza.digits = static_cast<int8_t*>(o.pa);
    oza.slice_stride = mr * ldk;
    oza.scale = o.sA;
    if ((st = mcd_launch_standardize(h, d_dna, N, G, ldd, nullptr, 0, nullptr, nullptr, 0, o.nB, d_didx, &ozb))) return st;
    if (rows > 0) {
      if ((st = mcd_launch_standardize(h, d_rna, rows, G, ldr, nullptr, 0, nullptr, nullptr, 0, o.nA, d_ridx, &oza))) return st;
      st = mcd_launch_corr_ozaki(h, oza.digits, oza.slice_stride, rows, ozb.digits, ozb.slice_stride, N, ldk, nsl, o.sA, o.sB,
                                 o.nA, o.nB, C_loc, ldc, nullptr, 0);
    }
  } else {
    if ((st = mcd_ws(h, WS_SLICES_A, (size_t)2 * mr * ldk * 2, &o.pa))) return st;
    if ((st = mcd_ws(h, WS_SLICES_B, (size_t)2 * N * ldk * 2, &o.pb))) return st;
    uint16_t* a_hi = static_cast<uint16_t*>(o.pa);
    uint16_t* b_hi = static_cast<uint16_t*>(o.pb);
    if ((st = mcd_launch_standardize(h, d_dna, N, G, ldd, nullptr, 0, b_hi, b_hi + N * ldk, ldk, o.nB, d_didx))) return st;
    if (rows > 0) {
      if ((st = mcd_launch_standardize(h, d_rna, rows, G, ldr, nullptr, 0, a_hi, a_hi + mr * ldk, ldk, o.nA, d_ridx))) return st;
      st = mcd_launch_corr_split(h, a_hi, a_hi + mr * ldk, rows, b_hi, b_hi + N * ldk, N, ldk, o.nA, o.nB, C_loc, ldc, nullptr, 0);
    }
  }
  return st;
}
}  // namespace

int mcd_cell2cell_multi(mcd_handle* hs, int ndev, const double* rna, int64_t ld_rna, const int32_t* rna_gene_idx,
                        const double* dna, int64_t ld_dna, const int32_t* dna_gene_idx, int64_t M, int64_t N, int64_t G,
                        int precision, int32_t* assign, int32_t* step, double* step_obj, mcd_stats* stats) {
  if (!hs || ndev < 1 || !hs[0]) return MCD_ERR_INVALID;
  mcd_context* h0 = hs[0];
  if (ndev == 1)
    return mcd_cell2cell_gather(h0, rna, ld_rna, rna_gene_idx, dna, ld_dna, dna_gene_idx, M, N, G, MCD_MEM_HOST, precision,
                                assign, step, step_obj, nullptr, MCD_MEM_HOST, stats);
  for (int d = 0; d < ndev; ++d)
    if (!hs[d]) return mcd_fail(h0, MCD_ERR_INVALID, "mcd_cell2cell_multi: NULL handle");
  if (!rna || !dna || !assign || !step || M < 1 || N < 1 || G < 1 || (!rna_gene_idx && ld_rna < G) ||
      (!dna_gene_idx && ld_dna < G) || M > 0x3fffffff || N > 0x3fffffff || G > 0x7fffffff)
    return mcd_fail(h0, MCD_ERR_INVALID, "mcd_cell2cell_multi arguments");
  if (precision != MCD_PREC_FP64 && precision != MCD_PREC_SPLIT_FP16 && precision != MCD_PREC_OZAKI_INT8)
    return mcd_fail(h0, MCD_ERR_INVALID, "unknown precision");
  for (int64_t g = 0; g < G; ++g)
    if ((rna_gene_idx && (rna_gene_idx[g] < 0 || rna_gene_idx[g] >= ld_rna)) ||
        (dna_gene_idx && (dna_gene_idx[g] < 0 || dna_gene_idx[g] >= ld_dna)))
      return mcd_fail(h0, MCD_ERR_INVALID, "mcd_cell2cell_multi: gene index out of range");
  const int nsl = mcd_ozaki_slices(h0, M, N, G);
  if (precision == MCD_PREC_OZAKI_INT8 && (double)nsl * 4096.0 * (double)mcd_padded_k_split(G) >= 2147483648.0)
    precision = MCD_PREC_FP64;
  h0->last_M = 0;
  h0->last_assign = nullptr;
  const int64_t ldc = (N + 1) & ~1LL, ldct = (M + 1) & ~1LL;
  const int64_t per = (M + ndev - 1) / ndev;
  const int64_t wr = rna_gene_idx ? ld_rna : G, wd = dna_gene_idx ? ld_dna : G;
  int st;
  // the whole matrix lives on the first device
  MCD_CUDA(h0, cudaSetDevice(h0->device));
  void *pC = nullptr, *pCt = nullptr;
  if ((st = mcd_ws(h0, WS_C, (size_t)M * ldc * 8, &pC))) return st;
  if ((st = mcd_ws(h0, WS_CT, (size_t)N * ldct * 8, &pCt))) return st;
  double* C = static_cast<double*>(pC);
  double* Ct = static_cast<double*>(pCt);
  const int64_t launches0 = h0->launches;
  MCD_CUDA(h0, cudaEventRecord(get_event(h0, 0), h0->stream));
  for (int d = 0; d < ndev; ++d) {
    mcd_context* h = hs[d];
    const int64_t lo = d * per < M ? d * per : M, hi = lo + per < M ? lo + per : M;
    const int64_t rows = hi - lo;
    MCD_CUDA(h0, cudaSetDevice(h->device));
    if (d > 0) {
      int can = 0;
      if (h->device != h0->device && cudaDeviceCanAccessPeer(&can, h->device, h0->device) == cudaSuccess && can) {
        cudaError_t pe = cudaDeviceEnablePeerAccess(h0->device, 0);
        if (pe != cudaSuccess) cudaGetLastError();  // already enabled
      }
    }
    void *pr = nullptr, *pd = nullptr, *pg = nullptr, *pcl = nullptr;
    if ((st = mcd_ws(h, WS_RNA_IN, (size_t)(rows > 0 ? rows : 1) * wr * 8, &pr))) return mcd_fail(h0, st, h->err.c_str());
    if ((st = mcd_ws(h, WS_DNA_IN, (size_t)N * wd * 8, &pd))) return mcd_fail(h0, st, h->err.c_str());
    const int* d_ridx = nullptr;
    const int* d_didx = nullptr;
    if (rna_gene_idx || dna_gene_idx) {
      if ((st = mcd_ws(h, WS_GIDX, (size_t)2 * G * 4, &pg))) return mcd_fail(h0, st, h->err.c_str());
      int* gi = static_cast<int*>(pg);
      if (rna_gene_idx) {
        MCD_CUDA(h0, cudaMemcpyAsync(gi, rna_gene_idx, (size_t)G * 4, cudaMemcpyHostToDevice, h->stream));
        d_ridx = gi;
      }
      if (dna_gene_idx) {
        MCD_CUDA(h0, cudaMemcpyAsync(gi + G, dna_gene_idx, (size_t)G * 4, cudaMemcpyHostToDevice, h->stream));
        d_didx = gi + G;
      }
    }
    // every device copies its RNA rows and the (small) DNA operand over its own PCIe link
    MCD_CUDA(h0, cudaMemcpy2DAsync(pd, (size_t)wd * 8, dna, (size_t)ld_dna * 8, (size_t)wd * 8, (size_t)N,
                                   cudaMemcpyHostToDevice, h->stream));
    if (rows > 0)
      MCD_CUDA(h0, cudaMemcpy2DAsync(pr, (size_t)wr * 8, rna + lo * ld_rna, (size_t)ld_rna * 8, (size_t)wr * 8, (size_t)rows,
                                     cudaMemcpyHostToDevice, h->stream));
    double* C_loc = C + lo * ldc;  // device 0 writes its shard in place
    if (d > 0) {
      if ((st = mcd_ws(h, WS_C, (size_t)(rows > 0 ? rows : 1) * ldc * 8, &pcl))) return mcd_fail(h0, st, h->err.c_str());
      C_loc = static_cast<double*>(pcl);
    } else {
      MCD_CUDA(h0, cudaStreamWaitEvent(h->stream, get_event(h0, 0), 0));
    }
    if ((st = shard_standardize_correlate(h, static_cast<const double*>(pr), wr, d_ridx, static_cast<const double*>(pd), wd,
                                          d_didx, rows, N, G, precision, nsl, C_loc, ldc)))
      return mcd_fail(h0, st, h->err.c_str());
    if (d > 0 && rows > 0)
      MCD_CUDA(h0, cudaMemcpyPeerAsync(C + lo * ldc, h0->device, C_loc, h->device, (size_t)rows * ldc * 8, h->stream));
    if (d > 0) MCD_CUDA(h0, cudaEventRecord(get_event(h, 0), h->stream));
  }
  MCD_CUDA(h0, cudaSetDevice(h0->device));
  for (int d = 1; d < ndev; ++d) MCD_CUDA(h0, cudaStreamWaitEvent(h0->stream, get_event(hs[d], 0), 0));
  MCD_CUDA(h0, cudaEventRecord(get_event(h0, 3), h0->stream));
  if ((st = mcd_transpose_f64(h0, C, M, N, ldc, Ct, ldct))) return st;
  const int64_t nsteps = mcd_num_steps(M, N);
  void* misc = nullptr;
  if ((st = mcd_ws(h0, WS_MISC, step_out_bytes(M, nsteps), &misc))) return st;
  const StepOut out = carve_step_out(misc, M, nsteps);
  const size_t EV_LAP = 8;
  // check_finite: the non-finite flag K1 raises lives on the device that saw the bad value; the first solve
  // re-scans the assembled matrix so that hs[0]'s solver kernels no-op on NaN / Inf instead of iterating on them
  if ((st = enqueue_step_loop(h0, C, ldc, Ct, ldct, M, N, out, EV_LAP, true))) return st;
  MCD_CUDA(h0, cudaEventRecord(get_event(h0, 4), h0->stream));
  MCD_CUDA(h0, cudaMemcpyAsync(assign, out.assign, (size_t)M * 4, cudaMemcpyDeviceToHost, h0->stream));
  MCD_CUDA(h0, cudaMemcpyAsync(step, out.step, (size_t)M * 4, cudaMemcpyDeviceToHost, h0->stream));
  if (step_obj) MCD_CUDA(h0, cudaMemcpyAsync(step_obj, out.obj, (size_t)nsteps * 8, cudaMemcpyDeviceToHost, h0->stream));
  std::vector<mcd_lap_counters> hc((size_t)nsteps);
  std::vector<mcd_lap_cert> hcert((size_t)nsteps);
  MCD_CUDA(h0, cudaMemcpyAsync(hc.data(), out.cnt, sizeof(mcd_lap_counters) * nsteps, cudaMemcpyDeviceToHost, h0->stream));
  MCD_CUDA(h0, cudaMemcpyAsync(hcert.data(), out.cert, sizeof(mcd_lap_cert) * nsteps, cudaMemcpyDeviceToHost, h0->stream));
  MCD_CUDA(h0, cudaEventRecord(get_event(h0, 5), h0->stream));
  MCD_CUDA(h0, cudaStreamSynchronize(h0->stream));
  int flag = 0;
  for (int d = 0; d < ndev; ++d) {  // non-finite input is flagged by K1 on the device that saw it
    int f = 0;
    MCD_CUDA(h0, cudaSetDevice(hs[d]->device));
    MCD_CUDA(h0, cudaMemcpy(&f, hs[d]->d_flags, sizeof(int), cudaMemcpyDeviceToHost));
    MCD_CUDA(h0, cudaMemset(hs[d]->d_flags, 0, sizeof(int)));
    flag |= f;
  }
  MCD_CUDA(h0, cudaSetDevice(h0->device));
  if (stats) memset(stats, 0, sizeof *stats);
  const int bad = fold_step_records(h0, M, N, nsteps, hc.data(), hcert.data(), stats, EV_LAP, true);
  if (stats) {
    auto el = [&](int a, int b) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, get_event(h0, a), get_event(h0, b));
      return (double)ms;
    };
    stats->ms_corr = el(0, 3);  // H2D + K1 + K2 + shard transfer, all devices (overlapped)
    stats->ms_lap = el(3, 4);
    stats->ms_d2h = el(4, 5);
    stats->ms_total = el(0, 5);
    stats->n_steps = nsteps;
    int64_t launches = h0->launches - launches0;
    stats->kernel_launches = launches;
  }
  if (flag) return mcd_fail(h0, MCD_ERR_NONFINITE, "NaN or Inf in the expression / copy-number matrix");
  if ((st = step_status(h0, bad))) return st;
  h0->last_M = M;
  h0->last_N = N;
  h0->last_ldc = ldc;
  h0->last_assign = out.assign;
  return MCD_OK;
}

int mcd_corr_rows(mcd_handle h, const int32_t* rows, int64_t nrows, double* out, int out_space) {
  if (!h) return MCD_ERR_INVALID;
  if (!rows || !out || nrows < 1 || h->last_M < 1 || h->last_assign == nullptr)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_corr_rows: no mcd_cell2cell result is resident");
  for (int64_t i = 0; i < nrows; ++i)
    if (rows[i] < 0 || rows[i] >= h->last_M) return mcd_fail(h, MCD_ERR_INVALID, "mcd_corr_rows: row out of range");
  MCD_CUDA(h, cudaSetDevice(h->device));
  const int64_t N = h->last_N;
  void* pi = nullptr;
  void* po = nullptr;
  int st;
  if ((st = mcd_ws(h, WS_SUB_IDX, (size_t)nrows * 4, &pi))) return st;
  MCD_CUDA(h, cudaMemcpyAsync(pi, rows, (size_t)nrows * 4, cudaMemcpyHostToDevice, h->stream));
  double* dst = out;
  if (out_space != MCD_MEM_DEVICE) {
    if ((st = mcd_ws(h, WS_SUB_C, (size_t)nrows * N * 8, &po))) return st;
    dst = static_cast<double*>(po);
  }
  dim3 grid((unsigned)((N + 1023) / 1024 < 64 ? (N + 1023) / 1024 : 64), grid_rows(nrows));
  gather_rows_out_kernel<<<grid, 256, 0, h->stream>>>(static_cast<const double*>(h->ws[WS_C].ptr), h->last_ldc,
                                                      static_cast<const int*>(pi), nrows, N, dst);
  MCD_LAUNCH_CHECK(h, "gather_rows_out_kernel");
  if (out_space != MCD_MEM_DEVICE)
    MCD_CUDA(h, cudaMemcpyAsync(out, dst, (size_t)nrows * N * 8, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  return MCD_OK;
}

int mcd_corr_pairs(mcd_handle h, const int32_t* rows, const int32_t* cols, int64_t n, double* out) {
  if (!h) return MCD_ERR_INVALID;
  if (!rows || !cols || !out || n < 1 || h->last_M < 1 || h->last_assign == nullptr)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_corr_pairs: no mcd_cell2cell result is resident");
  for (int64_t k = 0; k < n; ++k)
    if (rows[k] < 0 || rows[k] >= h->last_M || cols[k] < 0 || cols[k] >= h->last_N)
      return mcd_fail(h, MCD_ERR_INVALID, "mcd_corr_pairs: index out of range");
  MCD_CUDA(h, cudaSetDevice(h->device));
  void* pi = nullptr;
  void* po = nullptr;
  int st;
  if ((st = mcd_ws(h, WS_SUB_IDX, (size_t)n * 8, &pi))) return st;
  if ((st = mcd_ws(h, WS_SUB_MISC, (size_t)n * 8, &po))) return st;
  int* d_r = static_cast<int*>(pi);
  int* d_c = d_r + n;
  MCD_CUDA(h, cudaMemcpyAsync(d_r, rows, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(d_c, cols, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  gather_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(static_cast<const double*>(h->ws[WS_C].ptr),
                                                                        h->last_ldc, d_r, d_c, n,
                                                                        static_cast<double*>(po));
  MCD_LAUNCH_CHECK(h, "gather_pairs_kernel");
  MCD_CUDA(h, cudaMemcpyAsync(out, po, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  return MCD_OK;
}

static int subinstance_steps_impl(mcd_handle h, const int32_t* rna_rows, int64_t m_sub, const int32_t* dna_cols,
                                  int64_t n_sub, int32_t* assign, int32_t* step, double* step_obj, int out_space,
                                  mcd_stats* stats, bool use_classes, int* bad_out);

int mcd_subinstance_steps(mcd_handle h, const int32_t* rna_rows, int64_t m_sub, const int32_t* dna_cols, int64_t n_sub,
                          int32_t* assign, int32_t* step, double* step_obj, int out_space, mcd_stats* stats) {
  int bad = 0;
  int st = subinstance_steps_impl(h, rna_rows, m_sub, dna_cols, n_sub, assign, step, step_obj, out_space, stats, true, &bad);
  if (st == MCD_ERR_NOT_CONVERGED && (bad & 4)) {
    // The class treatment of duplicated cells leaves one window open (two copies settling at different levels in the
    // very round a third party buys the dearer copy's object); the certificate catches it.  Such a replicate is
    // solved again with every copy as a person of its own: exact ties then stall the auction and the
    // augmenting-path kernel finishes them -- slower, never wrong.
    st = subinstance_steps_impl(h, rna_rows, m_sub, dna_cols, n_sub, assign, step, step_obj, out_space, stats, false, &bad);
  }
  return st;
}

static int subinstance_steps_impl(mcd_handle h, const int32_t* rna_rows, int64_t m_sub, const int32_t* dna_cols,
                                  int64_t n_sub, int32_t* assign, int32_t* step, double* step_obj, int out_space,
                                  mcd_stats* stats, bool use_classes, int* bad_out) {
  if (!h) return MCD_ERR_INVALID;
  if (!assign || !step || m_sub < 1 || n_sub < 1 || h->last_M < 1 || h->last_assign == nullptr)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_steps: no mcd_cell2cell result is resident");
  if ((!rna_rows && m_sub != h->last_M) || (!dna_cols && n_sub != h->last_N))
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_steps: NULL index means all rows / columns");
  for (int64_t i = 0; rna_rows && i < m_sub; ++i)
    if (rna_rows[i] < 0 || rna_rows[i] >= h->last_M)
      return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_steps: RNA row out of range");
  for (int64_t j = 0; dna_cols && j < n_sub; ++j)
    if (dna_cols[j] < 0 || dna_cols[j] >= h->last_N)
      return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_steps: DNA column out of range");
  MCD_CUDA(h, cudaSetDevice(h->device));
  int st;
  const int64_t lds = (n_sub + 1) & ~1LL, ldst = (m_sub + 1) & ~1LL;
  void *pc = nullptr, *pct = nullptr, *pi = nullptr, *misc = nullptr;
  if ((st = mcd_ws(h, WS_SUB_C, (size_t)m_sub * lds * 8, &pc))) return st;
  if ((st = mcd_ws(h, WS_SUB_CT, (size_t)n_sub * ldst * 8, &pct))) return st;
  if ((st = mcd_ws(h, WS_SUB_IDX, (size_t)(m_sub + 2 * n_sub) * 4, &pi))) return st;
  int* d_rows = rna_rows ? static_cast<int*>(pi) : nullptr;
  int* d_cols = dna_cols ? static_cast<int*>(pi) + m_sub : nullptr;
  int* d_cls = nullptr;
  if (rna_rows) MCD_CUDA(h, cudaMemcpyAsync(d_rows, rna_rows, (size_t)m_sub * 4, cudaMemcpyHostToDevice, h->stream));
  if (dna_cols) MCD_CUDA(h, cudaMemcpyAsync(d_cols, dna_cols, (size_t)n_sub * 4, cudaMemcpyHostToDevice, h->stream));
  // copies of a DNA cell (resampling with replacement): identical columns, i.e. exact ties by construction.  They
  // are told to the solver as classes of similar persons.
  std::vector<int> cls;
  if (use_classes && dna_cols && duplicate_classes(dna_cols, n_sub, cls) > 0) {
    d_cls = static_cast<int*>(pi) + m_sub + n_sub;
    MCD_CUDA(h, cudaMemcpyAsync(d_cls, cls.data(), (size_t)n_sub * 4, cudaMemcpyHostToDevice, h->stream));
  }
  const int64_t launches0 = h->launches;
  const size_t EV_LAP = 8;
  MCD_CUDA(h, cudaEventRecord(get_event(h, 6), h->stream));
  double* subC = static_cast<double*>(pc);
  double* subCt = static_cast<double*>(pct);
  {
    dim3 grid((unsigned)((n_sub + 1023) / 1024 < 64 ? (n_sub + 1023) / 1024 : 64), grid_rows(m_sub));
    gather_sub_kernel<<<grid, 256, 0, h->stream>>>(static_cast<const double*>(h->ws[WS_C].ptr), h->last_ldc, d_rows,
                                                   d_cols, m_sub, n_sub, subC, lds);
    MCD_LAUNCH_CHECK(h, "gather_sub_kernel");
  }
  if ((st = mcd_transpose_f64(h, subC, m_sub, n_sub, lds, subCt, ldst))) return st;
  const int64_t nsteps = mcd_num_steps(m_sub, n_sub);
  if ((st = mcd_ws(h, WS_SUB_MISC, step_out_bytes(m_sub, nsteps), &misc))) return st;
  const StepOut out = carve_step_out(misc, m_sub, nsteps);
  if ((st = enqueue_step_loop(h, subC, lds, subCt, ldst, m_sub, n_sub, out, EV_LAP, false, true, d_cls))) return st;
  MCD_CUDA(h, cudaEventRecord(get_event(h, 7), h->stream));
  const cudaMemcpyKind kind = out_space == MCD_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  MCD_CUDA(h, cudaMemcpyAsync(assign, out.assign, (size_t)m_sub * 4, kind, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(step, out.step, (size_t)m_sub * 4, kind, h->stream));
  if (step_obj) MCD_CUDA(h, cudaMemcpyAsync(step_obj, out.obj, (size_t)nsteps * 8, kind, h->stream));
  std::vector<mcd_lap_counters> hc((size_t)nsteps);
  std::vector<mcd_lap_cert> hcert((size_t)nsteps);
  MCD_CUDA(h, cudaMemcpyAsync(hc.data(), out.cnt, sizeof(mcd_lap_counters) * nsteps, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaMemcpyAsync(hcert.data(), out.cert, sizeof(mcd_lap_cert) * nsteps, cudaMemcpyDeviceToHost, h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  if (stats) memset(stats, 0, sizeof *stats);
  const int bad = fold_step_records(h, m_sub, n_sub, nsteps, hc.data(), hcert.data(), stats, EV_LAP, true);
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, get_event(h, 6), get_event(h, 7));
    stats->ms_lap = ms;
    stats->ms_total = ms;
    stats->n_steps = nsteps;
    stats->kernel_launches = h->launches - launches0;
  }
  if (bad_out) *bad_out = bad | ((bad & 2) && d_cls != nullptr ? 4 : 0);  // bit 2: certificate failed WITH classes in use
  return step_status(h, bad);
}

// worker k of a handle (created on first use)
static int sweep_worker(mcd_context* h, size_t k, mcd_context** out) {
  while (h->workers.size() <= k) {
    mcd_context* w = new (std::nothrow) mcd_context();
    if (w == nullptr) return mcd_fail(h, MCD_ERR_NOMEM, "worker context");
    w->device = h->device;
    w->sm_count = h->sm_count;
    cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_flags, 64);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_flags, 0, 64, w->stream);
    if (e != cudaSuccess) {
      mcd_destroy(w);
      return mcd_fail(h, MCD_ERR_CUDA, "worker context", e);
    }
    h->workers.push_back(w);
  }
  *out = h->workers[k];
  return MCD_OK;
}

int mcd_subinstance_sweep(mcd_handle h, int64_t nrep, const int32_t* rna_rows, int64_t m_sub, const int32_t* dna_cols,
                          int64_t n_sub, int32_t* assign, int32_t* step, double* step_obj, double* cert_gap,
                          int concurrency, mcd_stats* stats) {
  if (!h) return MCD_ERR_INVALID;
  if (!assign || !step || !dna_cols || nrep < 1 || m_sub < 1 || n_sub < 1 || h->last_M < 1 || h->last_assign == nullptr)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_sweep: no mcd_cell2cell result is resident / bad arguments");
  if (!rna_rows && m_sub != h->last_M)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_sweep: NULL rna_rows means all rows");
  for (int64_t i = 0; rna_rows && i < m_sub; ++i)
    if (rna_rows[i] < 0 || rna_rows[i] >= h->last_M)
      return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_sweep: RNA row out of range");
  for (int64_t j = 0; j < nrep * n_sub; ++j)
    if (dna_cols[j] < 0 || dna_cols[j] >= h->last_N)
      return mcd_fail(h, MCD_ERR_INVALID, "mcd_subinstance_sweep: DNA column out of range");
  MCD_CUDA(h, cudaSetDevice(h->device));
  int K = concurrency < 1 ? 8 : concurrency;
  if (K > 32) K = 32;
  if ((int64_t)K > nrep) K = (int)nrep;
  const int64_t nsteps = mcd_num_steps(m_sub, n_sub);
  const size_t rep_bytes = (step_out_bytes(m_sub, nsteps) + 255) / 256 * 256;
  int64_t batch = (int64_t)((size_t)1 << 30) / (int64_t)rep_bytes;  // <= 1 GiB of per-replicate outputs at a time
  if (batch < K) batch = K;
  if (batch > nrep) batch = nrep;
  int st;
  void *p_idx = nullptr, *p_out = nullptr;
  if ((st = mcd_ws(h, WS_SWEEP_IDX, (size_t)(m_sub + 2 * batch * n_sub) * 4, &p_idx))) return st;
  if ((st = mcd_ws(h, WS_SWEEP_OUT, (size_t)batch * rep_bytes, &p_out))) return st;
  int* d_rows = rna_rows ? static_cast<int*>(p_idx) : nullptr;
  int* d_cols = static_cast<int*>(p_idx) + m_sub;
  int* d_clss = d_cols + batch * n_sub;
  std::vector<int> clss((size_t)(batch * n_sub)), cls1;
  std::vector<int64_t> n_extra((size_t)batch);
  if (rna_rows) MCD_CUDA(h, cudaMemcpyAsync(d_rows, rna_rows, (size_t)m_sub * 4, cudaMemcpyHostToDevice, h->stream));
  const int64_t lds = (n_sub + 1) & ~1LL, ldst = (m_sub + 1) & ~1LL;
  const double* C = static_cast<const double*>(h->ws[WS_C].ptr);
  // every worker solves on a slice of the chip: the cooperative grid of the wide rounds is capped so that K of them
  // are co-resident (the narrow rounds run on one 8- or 16-CTA cluster per solve anyway)
  std::vector<mcd_context*> ws((size_t)K);
  const int total_blocks = h->sm_count * (h->opt.lap_blocks_per_sm < 1 ? 1 : h->opt.lap_blocks_per_sm);
  for (int k = 0; k < K; ++k) {
    if ((st = sweep_worker(h, (size_t)k, &ws[k]))) return st;
    ws[k]->opt = h->opt;
    ws[k]->launches = 0;
    if (h->opt.lap_grid_blocks <= 0) {
      int gb = total_blocks / K;
      ws[k]->opt.lap_grid_blocks = gb < 16 ? 16 : gb;
    }
  }
  const int64_t launches0 = h->launches;
  // (events 6 / 7 and 8 .. 8 + steps belong to mcd_subinstance_steps, which the class-free re-solve of a replicate runs
  //  on this handle: the sweep times itself with events of its own)
  MCD_CUDA(h, cudaEventRecord(get_event(h, 96), h->stream));
  std::vector<mcd_lap_counters> hc((size_t)(nsteps * batch));
  std::vector<mcd_lap_cert> hcert((size_t)(nsteps * batch));
  if (stats) memset(stats, 0, sizeof *stats);
  int bad = 0;
  double host_enqueue_ms = 0.0;  // wall time of the enqueuing phase of every batch: reported as ms_h2d
  for (int64_t r0 = 0; r0 < nrep; r0 += batch) {
    const int64_t nb = nrep - r0 < batch ? nrep - r0 : batch;
    MCD_CUDA(h, cudaMemcpyAsync(d_cols, dna_cols + r0 * n_sub, (size_t)nb * n_sub * 4, cudaMemcpyHostToDevice, h->stream));
    for (int64_t b = 0; b < nb; ++b) {
      n_extra[(size_t)b] = duplicate_classes(dna_cols + (r0 + b) * n_sub, n_sub, cls1);
      std::copy(cls1.begin(), cls1.end(), clss.begin() + b * n_sub);
    }
    MCD_CUDA(h, cudaMemcpyAsync(d_clss, clss.data(), (size_t)nb * n_sub * 4, cudaMemcpyHostToDevice, h->stream));
    MCD_CUDA(h, cudaEventRecord(get_event(h, 5), h->stream));
    for (int k = 0; k < K; ++k) MCD_CUDA(h, cudaStreamWaitEvent(ws[k]->stream, get_event(h, 5), 0));
    // Replicates differ a lot in cost (a resampled replicate with a starved clone needs 10x the rounds of a balanced
    // one), so they are dealt dynamically from a shared counter.  Every worker has its own HOST thread: a cooperative /
    // cluster launch may block its caller until the stream's earlier work has drained (measured: with one enqueuing
    // thread 75 % of the sweep's wall time was spent inside launch calls and the workers starved), and the runtime is
    // thread-safe.  A thread keeps at most two replicates queued on its stream.
    const auto host_t0 = std::chrono::steady_clock::now();
    std::atomic<int64_t> next_rep(0);
    std::vector<int> wstatus((size_t)K, 0);
    auto worker_main = [&](int k) {
      mcd_context* w = ws[(size_t)k];
      if (cudaSetDevice(h->device) != cudaSuccess) {
        wstatus[(size_t)k] = MCD_ERR_CUDA;
        return;
      }
      int issued = 0;
      for (;;) {
        const int64_t b = next_rep.fetch_add(1);
        if (b >= nb) break;
        if (issued >= 2) cudaEventSynchronize(get_event(w, 1 + (size_t)(issued & 1)));  // replicate issued - 2 is done
        int stw;
        void *pc = nullptr, *pct = nullptr;
        if ((stw = mcd_ws(w, WS_SUB_C, (size_t)m_sub * lds * 8, &pc)) ||
            (stw = mcd_ws(w, WS_SUB_CT, (size_t)n_sub * ldst * 8, &pct))) {
          wstatus[(size_t)k] = stw;
          return;
        }
        double* subC = static_cast<double*>(pc);
        double* subCt = static_cast<double*>(pct);
        dim3 grid((unsigned)((n_sub + 1023) / 1024 < 64 ? (n_sub + 1023) / 1024 : 64), grid_rows(m_sub));
        gather_sub_kernel<<<grid, 256, 0, w->stream>>>(C, h->last_ldc, d_rows, d_cols + b * n_sub, m_sub, n_sub, subC, lds);
        w->launches++;
        if (cudaGetLastError() != cudaSuccess) {
          wstatus[(size_t)k] = mcd_fail(w, MCD_ERR_CUDA, "gather_sub_kernel");
          return;
        }
        const StepOut out = carve_step_out(static_cast<char*>(p_out) + (size_t)b * rep_bytes, m_sub, nsteps);
        if ((stw = mcd_transpose_f64(w, subC, m_sub, n_sub, lds, subCt, ldst)) ||
            (stw = enqueue_step_loop(w, subC, lds, subCt, ldst, m_sub, n_sub, out, 0, false, false,
                                     n_extra[(size_t)b] > 0 ? d_clss + b * n_sub : nullptr))) {
          wstatus[(size_t)k] = stw;
          return;
        }
        cudaEventRecord(get_event(w, 1 + (size_t)(issued & 1)), w->stream);
        issued++;
      }
    };
    for (int k = 0; k < K; ++k) {  // events are created on the calling thread (get_event grows a vector)
      get_event(ws[(size_t)k], 0);
      get_event(ws[(size_t)k], 1);
      get_event(ws[(size_t)k], 2);
    }
    {
      std::vector<std::thread> threads;
      for (int k = 1; k < K; ++k) threads.emplace_back(worker_main, k);
      worker_main(0);
      for (auto& t : threads) t.join();
    }
    for (int k = 0; k < K; ++k)
      if (wstatus[(size_t)k]) return mcd_fail(h, wstatus[(size_t)k], ws[(size_t)k]->err.c_str());
    host_enqueue_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    for (int k = 0; k < K; ++k) {
      MCD_CUDA(h, cudaEventRecord(get_event(ws[k], 0), ws[k]->stream));
      MCD_CUDA(h, cudaStreamWaitEvent(h->stream, get_event(ws[k], 0), 0));
    }
    // results of the batch: strided device block -> dense host arrays
    const StepOut o0 = carve_step_out(p_out, m_sub, nsteps);
    MCD_CUDA(h, cudaMemcpy2DAsync(assign + r0 * m_sub, (size_t)m_sub * 4, o0.assign, rep_bytes, (size_t)m_sub * 4, (size_t)nb,
                                  cudaMemcpyDeviceToHost, h->stream));
    MCD_CUDA(h, cudaMemcpy2DAsync(step + r0 * m_sub, (size_t)m_sub * 4, o0.step, rep_bytes, (size_t)m_sub * 4, (size_t)nb,
                                  cudaMemcpyDeviceToHost, h->stream));
    if (step_obj)
      MCD_CUDA(h, cudaMemcpy2DAsync(step_obj + r0 * nsteps, (size_t)nsteps * 8, o0.obj, rep_bytes, (size_t)nsteps * 8,
                                    (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
    MCD_CUDA(h, cudaMemcpy2DAsync(hc.data(), sizeof(mcd_lap_counters) * nsteps, o0.cnt, rep_bytes,
                                  sizeof(mcd_lap_counters) * nsteps, (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
    MCD_CUDA(h, cudaMemcpy2DAsync(hcert.data(), sizeof(mcd_lap_cert) * nsteps, o0.cert, rep_bytes,
                                  sizeof(mcd_lap_cert) * nsteps, (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
    MCD_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int64_t b = 0; b < nb; ++b) {
      const mcd_lap_cert* rc = hcert.data() + b * nsteps;
      int bad_b = fold_step_records(h, m_sub, n_sub, nsteps, hc.data() + b * nsteps, rc, stats, 0, false);
      double g = 0.0;
      for (int64_t s2 = 0; s2 < nsteps; ++s2) g = rc[s2].rel_gap > g ? rc[s2].rel_gap : g;
      if ((bad_b & 2) && n_extra[(size_t)b] > 0) {
        // certificate failed with the class treatment of duplicated cells: solve this replicate again without it
        // (see mcd_subinstance_steps)
        mcd_stats st2;
        const int r = (int)(r0 + b);
        const int rc2 = subinstance_steps_impl(h, rna_rows, m_sub, dna_cols + (int64_t)r * n_sub, n_sub,
                                               assign + (int64_t)r * m_sub, step + (int64_t)r * m_sub,
                                               step_obj ? step_obj + (int64_t)r * nsteps : nullptr, MCD_MEM_HOST, &st2, false,
                                               &bad_b);
        if (rc2 != MCD_OK && rc2 != MCD_ERR_NOT_CONVERGED) return rc2;
        g = st2.cert_rel_gap;
        if (stats) stats->sweep_fallbacks += 1;
      }
      bad |= bad_b;
      if (cert_gap) cert_gap[r0 + b] = h->opt.certify ? g : -1.0;
    }
  }
  MCD_CUDA(h, cudaEventRecord(get_event(h, 97), h->stream));
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, get_event(h, 96), get_event(h, 97));
    stats->ms_lap = ms;
    stats->ms_total = ms;
    stats->ms_h2d = host_enqueue_ms;
    stats->n_steps = nsteps;
    int64_t launches = h->launches - launches0;
    for (int k = 0; k < K; ++k) launches += ws[k]->launches;
    stats->kernel_launches = launches;
  }
  return step_status(h, bad);
}

int mcd_null_assignments(mcd_handle h, int64_t trials, uint64_t seed, double* sums, double* medians, int out_space) {
  if (!h) return MCD_ERR_INVALID;
  if (!sums || trials < 1 || h->last_M < 1 || h->last_assign == nullptr)
    return mcd_fail(h, MCD_ERR_INVALID, "mcd_null_assignments: no mcd_cell2cell result is resident");
  MCD_CUDA(h, cudaSetDevice(h->device));
  double* d_s = sums;
  double* d_m = medians;
  int st;
  if (out_space != MCD_MEM_DEVICE) {
    void* p = nullptr;
    if ((st = mcd_ws(h, WS_SUB_MISC, (size_t)trials * 16, &p))) return st;
    d_s = static_cast<double*>(p);
    d_m = medians ? d_s + trials : nullptr;
  }
  if ((st = mcd_launch_null_assignments(h, static_cast<const double*>(h->ws[WS_C].ptr), h->last_ldc, h->last_M,
                                        h->last_N, trials, seed, d_s, d_m)))
    return st;
  if (out_space != MCD_MEM_DEVICE) {
    MCD_CUDA(h, cudaMemcpyAsync(sums, d_s, (size_t)trials * 8, cudaMemcpyDeviceToHost, h->stream));
    if (medians) MCD_CUDA(h, cudaMemcpyAsync(medians, d_m, (size_t)trials * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  MCD_CUDA(h, cudaStreamSynchronize(h->stream));
  return MCD_OK;
}

}  // extern "C"
