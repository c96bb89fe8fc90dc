// K1 -- per-cell standardisation (HBM-bound streaming kernel).
//
// Replaces the per-operand half of cosine_similarity_np (reference
// src/MaCroDNA/macrodna.py:24-25): for every cell row x[0..G) compute
//   mean = sum(x)/G,  c = x - mean,  norm = sqrt(sum(c^2))
// once, instead of twice per (RNA, DNA) pair as the reference does.
//
// Layout: X is cells x genes row-major (macrodna.py:93-94 hand-off), so one cell is
// one contiguous vector.  A group of T threads owns one row and keeps it in
// registers between the mean pass and the norm pass, so HBM sees exactly one read
// (8 B/element) and one write (8 B/element FP64 centred rows, or 4 B/element for the
// two fp16 slices of the tcgen05 path).  Loads/stores are 128-bit and coalesced; reductions are warp
// shuffles plus one shared-memory hop.
#include "mcd_internal.cuh"

#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the T threads that own one row.  `red` has BLOCK/32 doubles.
template <int T, int BLOCK>
__device__ __forceinline__ double group_sum(double v, double* red) {
  v = warp_sum(v);
  if (T == 32) return v;
  constexpr int WPG = T / 32;  // warps per row group
  const int warp = threadIdx.x >> 5;
  __syncthreads();  // protect `red` reuse between the two reductions
  if ((threadIdx.x & 31) == 0) red[warp] = v;
  __syncthreads();
  const int g0 = (warp / WPG) * WPG;
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < WPG; ++w) s += red[g0 + w];
  return s;
}

// Max over the T threads that own one row (same protocol as group_sum).
template <int T, int BLOCK>
__device__ __forceinline__ double group_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (T == 32) return v;
  constexpr int WPG = T / 32;
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[warp] = v;
  __syncthreads();
  const int g0 = (warp / WPG) * WPG;
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < WPG; ++w) s = fmax(s, red[g0 + w]);
  return s;
}

// Integer-slice operand of the int8 tensor-core path (corr_ozaki.cu).  y (|y| < 0.5) is rounded to the
// fixed-point integer q = rint(y * 2^(7*nsl-1)) and written as nsl balanced radix-128 digits d_t in [-64, 63],
//   q = sum_t d_t * 128^(nsl-1-t)      (t = 0 is the most significant slice).
// The products of two such digit vectors are exact in the tensor core's int32 accumulators.
// ys = y * 2^(7 nsl - 1), already scaled by the caller (one multiply by a per-row power of two).
__device__ __forceinline__ void ozaki_digits(double ys, int nsl, int d[MCD_OZAKI_MAX_SLICES]) {
  const long long q = __double2ll_rn(ys);  // |q| <= 2^(7 nsl - 2)
  // Adding 64 to every 7-bit field turns the balanced digits into plain bit fields: q + B = sum (d_t + 64) 128^..
  const unsigned long long B = (0x0102040810204081ull << 6) & ((1ull << (7 * nsl)) - 1ull);
  const unsigned long long qb = (unsigned long long)q + B;
#pragma unroll
  for (int t = 0; t < MCD_OZAKI_MAX_SLICES; ++t)
    d[t] = (t < nsl) ? (int)((qb >> (7 * (nsl - 1 - t))) & 127ull) - 64 : 0;
}

// Split-precision operand for the tcgen05 path: y (|y| <= 1, a unit-norm centred value) is scaled by
// 2^8 and written as fp16 hi + fp16 lo (22 significant bits; the scale keeps `lo` out of the fp16
// subnormal range for every |y| > 5e-4).  hi*hi + hi*lo + lo*hi reproduces the product to ~2^-22.
__device__ __forceinline__ void split_fp16x2(double y, uint16_t& s0, uint16_t& s1) {
  const double ys = y * 256.0;
  const __half h0 = __double2half(ys);
  const double r = ys - (double)__half2float(h0);
  const __half h1 = __double2half(r);
  s0 = __half_as_ushort(h0);
  s1 = __half_as_ushort(h1);
}

// T threads per row, NV double2 per thread (row capacity 2*T*NV elements), VEC = 128-bit loads legal.
// gidx != nullptr: gene gather -- element g of the standardised row is X[row, gidx[g]] (the host only computes
// the index arrays of the gene intersection, macrodna.py:89-91; the data never gets re-indexed on the host).
// The row is then read with scattered 8-byte loads, but it is L1/L2 resident while its CTA works on it.
template <int T, int NV, bool VEC>
__global__ void __launch_bounds__((T > 512 ? T : 512), 1)
standardize_rows(const double* __restrict__ X, const int* __restrict__ gidx, int64_t ncells, int G, int64_t ldx,
                 double* __restrict__ Y,
                 int64_t ldk, uint16_t* __restrict__ S, uint16_t* __restrict__ S_lo, int64_t ldk16,
                 double* __restrict__ norms, int* __restrict__ flags, const mcd_ozaki_out oz) {
  constexpr int BLOCK = (T > 512 ? T : 512);
  constexpr int ROWS = BLOCK / T;
  __shared__ double red[BLOCK / 32];
  const int t = threadIdx.x % T;
  const int64_t row = (int64_t)blockIdx.x * ROWS + threadIdx.x / T;
  const bool live = row < ncells;
  const double* x = X + (live ? row : 0) * ldx;

  double2 v[NV];
  double sum = 0.0;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int e = 2 * (t + k * T);
    double2 d = make_double2(0.0, 0.0);
    if (live) {
      if (gidx != nullptr) {
        if (e < G) d.x = __ldg(x + __ldg(gidx + e));
        if (e + 1 < G) d.y = __ldg(x + __ldg(gidx + e + 1));
      } else if (e + 1 < G) {
        if (VEC) {
          d = __ldcs(reinterpret_cast<const double2*>(x + e));
        } else {
          d.x = __ldcs(x + e);
          d.y = __ldcs(x + e + 1);
        }
      } else if (e < G) {
        d.x = __ldcs(x + e);
      }
    }
    v[k] = d;
    sum += d.x + d.y;
  }
  sum = group_sum<T, BLOCK>(sum, red);
  const double mean = sum / (double)G;

  double ss = 0.0;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int e = 2 * (t + k * T);
    double cx = (e < G) ? v[k].x - mean : 0.0;
    double cy = (e + 1 < G) ? v[k].y - mean : 0.0;
    v[k] = make_double2(cx, cy);
    ss += cx * cx + cy * cy;
  }
  ss = group_sum<T, BLOCK>(ss, red);
  const double nrm = sqrt(ss);
  double amax = 0.0;
  if (oz.digits != nullptr) {  // uniform over the block
#pragma unroll
    for (int k = 0; k < NV; ++k) amax = fmax(amax, fmax(fabs(v[k].x), fabs(v[k].y)));
    amax = group_max<T, BLOCK>(amax, red);
  }
  if (!live) return;
  if (t == 0) {
    norms[row] = nrm;
    if (!isfinite(sum) || !isfinite(ss)) atomicOr(flags, 1);
  }
  if (oz.digits != nullptr) {
    // unit row u = c / nrm scaled by 2^e so that max|u| * 2^e lies in [0.25, 0.5)
    const double inv = (nrm > 0.0 && isfinite(nrm)) ? 1.0 / nrm : 0.0;
    const double umax = amax * inv;
    int ex = 0;
    if (umax > 0.0 && isfinite(umax)) (void)frexp(umax, &ex);  // umax = f * 2^ex, f in [0.5, 1)
    const int e = -ex - 1;
    if (t == 0) oz.scale[row] = scalbn(1.0, -e);
    // one multiplier per row: 1/nrm times the exact power of two 2^(e + 7 nsl - 1)  (bit-identical to scaling
    // the rounded quotient afterwards: a power-of-two factor commutes with rounding)
    const double mul = inv * scalbn(1.0, e + 7 * oz.nsl - 1);
    int8_t* o = oz.digits + row * oz.ldk8;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int g = 2 * (t + k * T);
      if (g < G) {
        int d0[MCD_OZAKI_MAX_SLICES], d1[MCD_OZAKI_MAX_SLICES];
        ozaki_digits(v[k].x * mul, oz.nsl, d0);
        ozaki_digits((g + 1 < G) ? v[k].y * mul : 0.0, oz.nsl, d1);
#pragma unroll
        for (int sl = 0; sl < MCD_OZAKI_MAX_SLICES; ++sl)
          if (sl < oz.nsl)  // g is even and ldk8 a multiple of 64: aligned 2-byte store
            *reinterpret_cast<uint16_t*>(o + sl * oz.slice_stride + g) =
                (uint16_t)((uint32_t)(d0[sl] & 0xff) | ((uint32_t)(d1[sl] & 0xff) << 8));
      }
    }
    const int64_t g2 = (G + 1) & ~1;
    for (int sl = 0; sl < oz.nsl; ++sl)
      for (int64_t c = g2 + 2 * t; c < oz.ldk8; c += 2 * T)
        *reinterpret_cast<uint16_t*>(o + sl * oz.slice_stride + c) = 0;
  }

  if (Y != nullptr) {
    double* y = Y + row * ldk;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int e = 2 * (t + k * T);
      if (e + 1 < G) {
        *reinterpret_cast<double2*>(y + e) = v[k];  // ldk is a multiple of 16 -> aligned
      } else if (e < G) {
        y[e] = v[k].x;
      }
    }
    for (int64_t c = G + t; c < ldk; c += T) y[c] = 0.0;
  }
  if (S != nullptr) {
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    uint16_t* s0 = S + row * ldk16;
    uint16_t* s1 = S_lo + row * ldk16;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int e = 2 * (t + k * T);
      if (e < G) {
        uint16_t a0, a1, b0 = 0, b1 = 0;
        split_fp16x2(v[k].x * inv, a0, a1);
        if (e + 1 < G) split_fp16x2(v[k].y * inv, b0, b1);
        // e is even and ldk16 is a multiple of 64: 4-byte aligned pair store (pad column is zero)
        *reinterpret_cast<uint32_t*>(s0 + e) = (uint32_t)a0 | ((uint32_t)b0 << 16);
        *reinterpret_cast<uint32_t*>(s1 + e) = (uint32_t)a1 | ((uint32_t)b1 << 16);
      }
    }
    const int64_t g2 = (G + 1) & ~1;
    for (int64_t c = g2 + 2 * t; c < ldk16; c += 2 * T) {
      *reinterpret_cast<uint32_t*>(s0 + c) = 0u;
      *reinterpret_cast<uint32_t*>(s1 + c) = 0u;
    }
  }
}

// Rows longer than the register-resident capacity: one 512-thread block per row, three sweeps
// (the row stays L2-resident between sweeps; declared as a 3-read variant in DESIGN.md).
__global__ void __launch_bounds__(512)
standardize_rows_long(const double* __restrict__ X, const int* __restrict__ gidx, int64_t ncells, int64_t G,
                      int64_t ldx, double* __restrict__ Y,
                      int64_t ldk, uint16_t* __restrict__ S, uint16_t* __restrict__ S_lo, int64_t ldk16,
                      double* __restrict__ norms, int* __restrict__ flags, const mcd_ozaki_out oz) {
  __shared__ double red[16];
  const int64_t row = blockIdx.x;
  const double* x = X + row * ldx;
  auto at = [&](int64_t e) { return gidx != nullptr ? x[gidx[e]] : x[e]; };
  double sum = 0.0;
  for (int64_t e = threadIdx.x; e < G; e += 512) sum += at(e);
  sum = group_sum<512, 512>(sum, red);
  const double mean = sum / (double)G;
  double ss = 0.0;
  for (int64_t e = threadIdx.x; e < G; e += 512) {
    const double c = at(e) - mean;
    ss += c * c;
  }
  ss = group_sum<512, 512>(ss, red);
  const double nrm = sqrt(ss);
  if (threadIdx.x == 0) {
    norms[row] = nrm;
    if (!isfinite(sum) || !isfinite(ss)) atomicOr(flags, 1);
  }
  const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
  if (Y != nullptr) {
    double* y = Y + row * ldk;
    for (int64_t e = threadIdx.x; e < ldk; e += 512) y[e] = e < G ? at(e) - mean : 0.0;
  }
  if (oz.digits != nullptr) {
    double amax = 0.0;
    for (int64_t e = threadIdx.x; e < G; e += 512) amax = fmax(amax, fabs(at(e) - mean));
    amax = group_max<512, 512>(amax, red);
    const double inv1 = (nrm > 0.0 && isfinite(nrm)) ? 1.0 / nrm : 0.0;
    const double umax = amax * inv1;
    int ex = 0;
    if (umax > 0.0 && isfinite(umax)) (void)frexp(umax, &ex);
    const int sh = -ex - 1;
    if (threadIdx.x == 0) oz.scale[row] = scalbn(1.0, -sh);
    const double mul = inv1 * scalbn(1.0, sh + 7 * oz.nsl - 1);
    int8_t* o = oz.digits + row * oz.ldk8;
    for (int64_t e = threadIdx.x; e < oz.ldk8; e += 512) {
      int d[MCD_OZAKI_MAX_SLICES];
      ozaki_digits(e < G ? (at(e) - mean) * mul : 0.0, oz.nsl, d);
#pragma unroll
      for (int sl = 0; sl < MCD_OZAKI_MAX_SLICES; ++sl)
        if (sl < oz.nsl) o[sl * oz.slice_stride + e] = (int8_t)d[sl];
    }
  }
  if (S != nullptr) {
    uint16_t* s0 = S + row * ldk16;
    uint16_t* s1 = S_lo + row * ldk16;
    for (int64_t e = threadIdx.x; e < ldk16; e += 512) {
      uint16_t a0 = 0, a1 = 0;
      if (e < G) split_fp16x2((at(e) - mean) * inv, a0, a1);
      s0[e] = a0;
      s1[e] = a1;
    }
  }
}

template <int T, int NV>
int launch_t(mcd_context* h, bool vec, const double* X, const int* gidx, int64_t ncells, int G, int64_t ldx, double* Y,
             int64_t ldk,
             uint16_t* S, uint16_t* S_lo, int64_t ldk16, double* norms, const mcd_ozaki_out& oz) {
  constexpr int BLOCK = (T > 512 ? T : 512);
  constexpr int ROWS = BLOCK / T;
  const int64_t grid = (ncells + ROWS - 1) / ROWS;
  if (vec)
    standardize_rows<T, NV, true><<<(unsigned)grid, BLOCK, 0, h->stream>>>(X, gidx, ncells, G, ldx, Y, ldk, S, S_lo,
                                                                          ldk16, norms, h->d_flags, oz);
  else
    standardize_rows<T, NV, false><<<(unsigned)grid, BLOCK, 0, h->stream>>>(X, gidx, ncells, G, ldx, Y, ldk, S, S_lo,
                                                                           ldk16, norms, h->d_flags, oz);
  MCD_LAUNCH_CHECK(h, "standardize_rows");
  return MCD_OK;
}

}  // namespace

int mcd_launch_standardize(mcd_context* h, const double* X, int64_t ncells, int64_t G, int64_t ldx, double* centred,
                           int64_t ldk, uint16_t* slices, uint16_t* slices_lo, int64_t ldk16, double* norms,
                           const int* gidx, const mcd_ozaki_out* ozaki) {
  if (ncells == 0) return MCD_OK;
  mcd_ozaki_out oz{};
  if (ozaki != nullptr) oz = *ozaki;
  const bool vec = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((ldx & 1) == 0);
  const int g = (int)G;
  if (G <= 256) return launch_t<32, 4>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  if (G <= 1024) return launch_t<128, 4>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  if (G <= 4096) return launch_t<512, 4>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  if (G <= 8192) return launch_t<512, 8>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  if (G <= 12288) return launch_t<512, 12>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  if (G <= 16384) return launch_t<512, 16>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  if (G <= 20480) return launch_t<512, 20>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  if (G <= 24576) return launch_t<512, 24>(h, vec, X, gidx, ncells, g, ldx, centred, ldk, slices, slices_lo, ldk16, norms, oz);
  standardize_rows_long<<<(unsigned)ncells, 512, 0, h->stream>>>(X, gidx, ncells, G, ldx, centred, ldk, slices,
                                                                slices_lo, ldk16, norms, h->d_flags, oz);
  MCD_LAUNCH_CHECK(h, "standardize_rows_long");
  return MCD_OK;
}
