// Host-side check of K1's digit fast path (standardize.cu): the fma-with-magic-constant rounding, the field
// extraction and the PRMT / SWAR packing must give the same int8 digits as the plain reference ozaki_digits() applied
// to the exactly rounded fixed-point integer.  The device helpers are pasted in by tests/test_host_logic.py
// (DIGIT_SNIPPET) and run with host stand-ins for the three intrinsics they use.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define MCD_OZAKI_MAX_SLICES 8
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
  uint8_t src[8];
  std::memcpy(src, &a, 4);
  std::memcpy(src + 4, &b, 4);
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const int n = (s >> (4 * i)) & 0xf;
    uint8_t v = src[n & 7];
    if (n & 8) v = (v & 0x80) ? 0xff : 0;
    r |= (uint32_t)v << (8 * i);
  }
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31));
}
static inline long long __double_as_longlong(double d) {
  long long r;
  std::memcpy(&r, &d, 8);
  return r;
}
static inline long long __double2ll_rn(double d) { return llrint(d); }
#include DIGIT_SNIPPET

template <int NSL>
static long run(long iters) {
  const unsigned long long B = (0x0102040810204081ull << 6) & ((1ull << (7 * NSL)) - 1ull);
  const double magic = 6755399441055744.0 + (double)B;
  long bad = 0;
  for (long it = 0; it < iters; ++it) {
    const double mul = ldexp(1.0 + drand48(), (rand() % 5) + 7 * NSL - 3);
    double c[4];
    uint32_t lo[4], hi[4];
    int dref[4][8];
    for (int i = 0; i < 4; ++i) {
      c[i] = (drand48() - 0.5) * 0.49999 * ldexp(1.0, -(rand() % 5)) / (mul / ldexp(1.0, 7 * NSL - 1));
      if (it % 97 == 0 && i == 0) c[i] = 0.0;
      if (it % 89 == 0 && i == 1) c[i] = std::floor(c[i] * mul) / mul;  // exact integers and
      if (it % 83 == 0 && i == 2) c[i] = (std::floor(c[i] * mul) + 0.5) / mul;  // exact ties
      if (std::fabs(c[i] * mul) >= ldexp(1.0, 7 * NSL - 2)) c[i] = 0.0;
      ozaki_bits<NSL>(c[i], mul, magic, lo[i], hi[i]);
      double ys = c[i] * mul;  // reference: the fixed-point integer = the EXACT product, rounded once (NSL <= 7)
      if (NSL <= 7) {
        const double p = ys, err = std::fma(c[i], mul, -p), fl = std::floor(p);
        if (p - fl == 0.5)
          ys = err > 0 ? fl + 1 : err < 0 ? fl : std::rint(p);
        else
          ys = std::rint(p);
      }
      ozaki_digits(ys, NSL, dref[i]);
    }
    int8_t out[8 * 16];
    std::memset(out, 0x55, sizeof out);
    ozaki_store_slices<NSL, 0>(lo, hi, out, 16);
    for (int sl = 0; sl < NSL; ++sl)
      for (int i = 0; i < 4; ++i) bad += out[sl * 16 + i] != (int8_t)dref[i][sl];
  }
  return bad;
}
int main() {
  const long bad = run<5>(300000) + run<6>(1000000) + run<7>(300000) + run<8>(1000000);
  std::printf("digit bits: %ld mismatches\n", bad);
  return bad != 0;
}
