// Random-assignment null test on the resident correlation matrix.
//
// Replaces `random_test.assign` of the reference
// (Resampling_stability_analyses/BE_data_analyses/random_assignment_test.py:233-258, 1e8 trials per biopsy):
// per trial, every step pairs a uniformly random N-subset of the still-unassigned RNA cells with the DNA cells
// (a random bijection), the last step maps the remaining r <= N RNA cells injectively onto a random r-subset of
// the DNA cells; the statistic is the sum (random_assignment_test.py) or the median
// (random_assignment_test_median.py) of the matched correlations.
//
// Here one CTA draws one trial: a uniformly random permutation of the RNA cells is the ascending order of M
// independent 51-bit hash keys (bitonic sort in shared memory); position p of the permutation belongs to step
// p / N and, in the full steps, takes DNA cell p % N; the last partial step takes the first r entries of a second
// random permutation (of the DNA cells).  That is the same distribution as the reference's
// `np.random.choice(..., replace=False)` chain -- NOT the same random stream (NumPy's Mersenne Twister is not
// reproduced), so results are compared as distributions.  The matched values are gathered from C (L2 resident),
// summed in FP64 and, when asked for, sorted once more for the median.
#include "mcd_internal.cuh"

namespace {

constexpr int NT = 256;
constexpr int NULL_MAX = 4096;  // cells per side that fit the shared-memory sort

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ascending bitonic sort of n (power of two) 64-bit words in shared memory
__device__ void bitonic_sort_u64(unsigned long long* a, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int idx = threadIdx.x; idx < n; idx += NT) {
        const int ixj = idx ^ j;
        if (ixj > idx) {
          const unsigned long long x = a[idx], y = a[ixj];
          const bool up = (idx & k) == 0;
          if (up ? x > y : x < y) {
            a[idx] = y;
            a[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

// order-preserving map double -> uint64 (for sorting the matched values)
__device__ __forceinline__ unsigned long long f64_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(NT) null_assign_kernel(const double* __restrict__ C, int64_t ldc, int M, int N, int Mp,
                                                          int Np, long long trials, unsigned long long seed,
                                                          double* __restrict__ sums, double* __restrict__ medians) {
  extern __shared__ unsigned long long nsm[];
  unsigned long long* rk = nsm;        // [Mp] RNA keys -> permutation -> matched values
  unsigned long long* dk = nsm + Mp;   // [Np] DNA keys (last partial step)
  __shared__ double red[NT / 32];
  const int q = M / N, r = M - q * N;
  for (long long trial = blockIdx.x; trial < trials; trial += gridDim.x) {
    const unsigned long long base = mix64(seed ^ mix64((unsigned long long)trial));
    for (int i = threadIdx.x; i < Mp; i += NT)
      rk[i] = i < M ? ((mix64(base + 2ull * (unsigned long long)i) >> 13) << 13) | (unsigned long long)i : ~0ull;
    if (r > 0)
      for (int j = threadIdx.x; j < Np; j += NT)
        dk[j] = j < N ? ((mix64(base + 2ull * (unsigned long long)j + 1ull) >> 13) << 13) | (unsigned long long)j : ~0ull;
    __syncthreads();
    bitonic_sort_u64(rk, Mp);
    if (r > 0) bitonic_sort_u64(dk, Np);
    double acc = 0.0;
    for (int p = threadIdx.x; p < M; p += NT) {
      const int row = (int)(rk[p] & 8191ull);
      const int col = p < q * N ? p % N : (int)(dk[p - q * N] & 8191ull);
      const double v = C[(int64_t)row * ldc + col];
      acc += v;
      if (medians != nullptr) rk[p] = f64_key(v);  // each thread overwrites only the slots it just read
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < NT / 32; ++w) s += red[w];
      sums[trial] = s;
    }
    if (medians != nullptr) {
      for (int i = M + threadIdx.x; i < Mp; i += NT) rk[i] = ~0ull;
      __syncthreads();
      bitonic_sort_u64(rk, Mp);
      if (threadIdx.x == 0) medians[trial] = 0.5 * (key_f64(rk[(M - 1) / 2]) + key_f64(rk[M / 2]));  // numpy.median
    }
    __syncthreads();
  }
}

int pow2_at_least(int64_t n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace

int mcd_launch_null_assignments(mcd_context* h, const double* C, int64_t ldc, int64_t M, int64_t N, int64_t trials,
                                uint64_t seed, double* d_sums, double* d_medians) {
  if (M > NULL_MAX || N > NULL_MAX)
    return mcd_fail(h, MCD_ERR_UNSUPPORTED, "null test: more than 4096 cells on a side");
  const int Mp = pow2_at_least(M), Np = pow2_at_least(N);
  const size_t smem = (size_t)(Mp + Np) * 8;
  MCD_CUDA(h, cudaFuncSetAttribute(null_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  MCD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, null_assign_kernel, NT, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)h->sm_count * per_sm;
  if (grid > trials) grid = trials;
  null_assign_kernel<<<(unsigned)grid, NT, smem, h->stream>>>(C, ldc, (int)M, (int)N, Mp, Np, (long long)trials,
                                                              (unsigned long long)seed, d_sums, d_medians);
  MCD_LAUNCH_CHECK(h, "null_assign_kernel");
  return MCD_OK;
}
