// K2 (parity mode) -- FP64 correlation contraction on the FP64 tensor pipe (DMMA).
//
// Replaces the RNA x DNA double loop of the reference (src/MaCroDNA/macrodna.py:103-107,
// formula :24-25):   C[i,j] = dot(a_i, b_j) / (1e-10 + |a_i| |b_j|)
// with a_i, b_j the centred rows K1 produced.  Both operands are K-major (cells x genes
// row-major), i.e. a "TN" GEMM  C = A * B^T  over the gene axis.
//
// Tiling: 128x128 output tile per CTA, BK = 16 genes (one 128-byte line per row) per stage,
// 4-stage cp.async pipeline (128 KB smem), 8 warps of 64x32, mma.sync.m8n8k4.f64 (SASS
// DMMA.8x8x4 on sm_100a).  Shared tiles use a 16-byte-chunk XOR swizzle (chunk ^= (row&3)<<1)
// so both the cp.async fill and the 64-bit fragment reads are bank-conflict free.  The
// epilogue applies the reference's epsilon'd denominator and writes C and/or C^T (the
// assignment solver scans DNA-major rows in the |R| > N steps).
#include "mcd_internal.cuh"

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;  // doubles
constexpr int STAGES = 4;
constexpr int THREADS = 256;
constexpr int GROUP_M = 16;
constexpr int TILE_BYTES = BM * BK * 8;  // 16 KB
constexpr int SMEM_BYTES = STAGES * 2 * TILE_BYTES;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
  return v;
}

// Fill one stage: rows [row0, row0+128) x genes [k0, k0+16) of a K-major matrix, zero beyond nrows.
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const double* __restrict__ P, int64_t nrows, int64_t ld,
                                          int64_t row0, int64_t k0, int tid) {
#pragma unroll
  for (int it = 0; it < (BM * 8) / THREADS; ++it) {
    const int idx = tid + it * THREADS;
    const int r = idx >> 3;
    const int ch = idx & 7;
    const int64_t row = row0 + r;
    const bool ok = row < nrows;
    const double* src = P + (ok ? row : 0) * ld + k0 + ch * 2;
    const uint32_t dst = smem_tile + r * 128 + ((ch ^ ((r & 3) << 1)) << 4);
    cp_async16(dst, src, ok ? 16 : 0);
  }
}

__global__ void __launch_bounds__(THREADS, 1)
corr_fp64_kernel(const double* __restrict__ A, int64_t M, const double* __restrict__ B, int64_t N, int64_t ldk,
                 const double* __restrict__ nA, const double* __restrict__ nB, double* __restrict__ C, int64_t ldc,
                 double* __restrict__ Ct, int64_t ldct, int tiles_m, int tiles_n) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp >> 2;  // 0..1 -> 64 rows
  const int wn = warp & 3;   // 0..3 -> 32 cols
  const int r = lane >> 2;
  const int kq = lane & 3;

  // grouped rasterisation: GROUP_M row-tiles share each DNA panel while it is L2-hot
  const int per_group = GROUP_M * tiles_n;
  const int bid = blockIdx.x;
  const int group = bid / per_group;
  const int first_m = group * GROUP_M;
  const int gsize = min(tiles_m - first_m, GROUP_M);
  const int tm = first_m + (bid % per_group) % gsize;
  const int tn = (bid % per_group) / gsize;
  const int64_t row0 = (int64_t)tm * BM;
  const int64_t col0 = (int64_t)tn * BN;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (int)(ldk / BK);
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      load_tile(smem_base + s * 2 * TILE_BYTES, A, M, ldk, row0, (int64_t)s * BK, tid);
      load_tile(smem_base + s * 2 * TILE_BYTES + TILE_BYTES, B, N, ldk, col0, (int64_t)s * BK, tid);
    }
    cp_async_commit();
  }

  // per-thread fragment offsets inside a tile (bytes), without the k-step part
  const int sw = (r & 3) << 1;
  uint32_t a_off[8], b_off[4];
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) a_off[mi] = (wm * 64 + mi * 8 + r) * 128 + (kq & 1) * 8;
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) b_off[ni] = TILE_BYTES + (wn * 32 + ni * 8 + r) * 128 + (kq & 1) * 8;

  for (int kb = 0; kb < nk; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kb + STAGES - 1;
      if (nxt < nk) {
        const uint32_t st = smem_base + (nxt % STAGES) * 2 * TILE_BYTES;
        load_tile(st, A, M, ldk, row0, (int64_t)nxt * BK, tid);
        load_tile(st + TILE_BYTES, B, N, ldk, col0, (int64_t)nxt * BK, tid);
      }
      cp_async_commit();
    }
    const uint32_t st = smem_base + (kb % STAGES) * 2 * TILE_BYTES;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const uint32_t chunk = (uint32_t)(((ks * 2 + (kq >> 1)) ^ sw) << 4);
      double a[8], b[4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) a[mi] = lds64(st + a_off[mi] + chunk);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = lds64(st + b_off[ni] + chunk);
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();

  // epilogue: reference denominator (macrodna.py:25), C and/or C^T
  const bool vec_ok = (ldc & 1) == 0;
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) {
    const int64_t j = col0 + wn * 32 + ni * 8 + 2 * kq;
    const double nb0 = j < N ? nB[j] : 0.0;
    const double nb1 = j + 1 < N ? nB[j + 1] : 0.0;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
      const int64_t i = row0 + wm * 64 + mi * 8 + r;
      if (i >= M || j >= N) continue;
      const double na = nA[i];
      const double c0 = acc[mi][ni][0] / (1e-10 + na * nb0);
      const double c1 = acc[mi][ni][1] / (1e-10 + na * nb1);
      if (C != nullptr) {
        double* p = C + i * ldc + j;
        if (j + 1 < N && vec_ok) {
          *reinterpret_cast<double2*>(p) = make_double2(c0, c1);
        } else {
          p[0] = c0;
          if (j + 1 < N) p[1] = c1;
        }
      }
      if (Ct != nullptr) {
        Ct[j * ldct + i] = c0;
        if (j + 1 < N) Ct[(j + 1) * ldct + i] = c1;
      }
    }
  }
}

}  // namespace

int mcd_launch_corr_fp64(mcd_context* h, const double* A, int64_t M, const double* B, int64_t N, int64_t ldk,
                         const double* nA, const double* nB, double* C, int64_t ldc, double* Ct, int64_t ldct) {
  if (M == 0 || N == 0) return MCD_OK;
  static bool attr_set = false;  // per process; harmless if repeated
  if (!attr_set) {
    MCD_CUDA(h, cudaFuncSetAttribute(corr_fp64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int64_t tiles_m = (M + BM - 1) / BM;
  const int64_t tiles_n = (N + BN - 1) / BN;
  if (tiles_m * tiles_n > 0x7fffffffLL) return mcd_fail(h, MCD_ERR_UNSUPPORTED, "corr_fp64: too many tiles");
  corr_fp64_kernel<<<(unsigned)(tiles_m * tiles_n), THREADS, SMEM_BYTES, h->stream>>>(
      A, M, B, N, ldk, nA, nB, C, ldc, Ct, ldct, (int)tiles_m, (int)tiles_n);
  MCD_LAUNCH_CHECK(h, "corr_fp64_kernel");
  return MCD_OK;
}
