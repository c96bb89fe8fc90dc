// K2 (throughput mode) -- tcgen05 / TMEM split-precision correlation contraction for sm_100a.
//
// Replaces the RNA x DNA double loop of the reference (src/MaCroDNA/macrodna.py:103-107) with a
// TMA-fed 5th-generation tensor-core GEMM.  Operands are the split-precision rows K1 emits:
//   256 * unit(x - mean) = hi + lo   (two fp16 slices, 22 significant bits)
// and each output needs three tensor-core products  hi.hi + hi.lo + lo.hi  accumulated in FP32
// in TMEM (the dropped lo.lo term is < 2^-22).  The epilogue rescales by 2^-16 and by
// nn / (1e-10 + nn), nn = |r_i| |d_j|, which is exactly the reference's epsilon'd denominator
// applied to unit vectors, and writes C and/or C^T in FP64 for the assignment solver.
//
// Accuracy: the tensor core's FP32 accumulation TRUNCATES (measured here: a bias of ~2^-24.5 of the
// running sum per tcgen05.mma, i.e. ~1e-4 absolute at 20k genes if one accumulator ran over all of K).
// So the contraction is chunked: every CHUNK_KB k-blocks (128 genes, 24 MMAs) the MMA warp moves on to
// the next of four 128-column TMEM accumulators, and the epilogue warps drain the finished one into
// FP64 registers (exact summation of the chunk partials).  That holds the error at the 1e-7 level
// (north-star gate 1e-6) independent of K.
//
// Structure (one CTA per SM, persistent over 128 x 128 output tiles; CTAs run as CLUSTER PAIRS that work on
// two horizontally adjacent tiles, i.e. share the RNA operand):
//   warp 0      TMA producer: cp.async.bulk.tensor (128B swizzle) into a 3-stage ring (64 KB/stage).  Each
//               CTA fetches its own B_hi/B_lo tiles [128 x 64] but only HALF (64 rows) of A_hi/A_lo, with
//               .multicast::cluster to both CTAs of the pair: L2->SM traffic per k-block drops from 64 to
//               48 KB per CTA, which is what bounds this kernel (1 KB of operand per 128x128x3 MMA-k).
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M128 N128 K16),
//               12 per k-block; tcgen05.commit frees the smem stage / publishes the chunk accumulator
//   warps 2-9   continuous epilogue: tcgen05.ld 32 lanes x 32 columns, FP64 accumulation of the chunk,
//               and at the end of the tile the reference's scaling + stores of C / C^T
#include <cuda.h>

#include "mcd_internal.cuh"

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 64;  // fp16 elements = 128 bytes = one swizzle-128B row
constexpr int STAGES = 3;
constexpr int UMMA_K = 16;
constexpr int A_SLICE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_SLICE_BYTES = BN * BK * 2;  // 16 KB
constexpr int STAGE_BYTES = 2 * A_SLICE_BYTES + 2 * B_SLICE_BYTES;  // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 320;  // 10 warps: TMA, MMA, 8 epilogue
constexpr int NUM_ACC = 4;        // TMEM accumulator ring
constexpr int TMEM_COLS = NUM_ACC * BN;  // 512 columns
#ifndef MCD_CHUNK_KB
#define MCD_CHUNK_KB 2
#endif
constexpr int CHUNK_KB = MCD_CHUNK_KB;  // k-blocks per accumulator fill before promotion to FP64
constexpr int EPI_WARPS = 8;
constexpr int EPI_COLS = BN / 2;  // columns per epilogue thread (two warps share a TMEM lane quadrant)

// instruction descriptor, kind::f16: D=F32 (bits 4-5 = 1), A=B=F16 (0), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO = 64 x 16 B),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;            // version
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// commit that arrives on the barrier at the same smem offset in every CTA of `mask` (both CTAs of the pair)
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y), "h"(mask)
      : "memory");
}
// The producer and MMA warps run their loops WARP-UNIFORMLY (all 32 lanes: stages, phases and descriptors live in
// uniform registers) and only the instruction that must be issued once is predicated on elect.sync.  Inside an
// `if (lane == 0)` region the compiler cannot prove the descriptors uniform and wraps every UTCHMMA / UTMALDG in an
// ELECT / R2UR.BROADCAST waterfall loop (see corr_ozaki.cu, where this cost 16 % of the kernel).
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
      "}\n" ::"r"(bar),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      "}\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y,
                                                     uint16_t mask) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4}], [%2], %5;\n"
      "}\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                               uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc_elect(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();  // .aligned: the whole warp must be converged here
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// Pair-tile rasterisation.  The 74 pairs that run concurrently should touch few distinct operand panels so
// that HBM sees each panel about once per wave (the operands do not fit L2: ncu showed 305 GB of DRAM reads
// for 4.8 GB of operands with a row-major tile order).  Bands of BAND_M tile-rows, column-major inside a band:
// one wave = ~12 RNA panels x ~12 DNA panels.
constexpr int BAND_M = 12;
__device__ __forceinline__ void decode_tile(int t, int tiles_m, int tiles_np, int& tm, int& tnp) {
  const int per_band = BAND_M * tiles_np;
  const int band = t / per_band;
  const int rem = t - band * per_band;
  const int hb = min(BAND_M, tiles_m - band * BAND_M);
  tnp = rem / hb;
  tm = band * BAND_M + (rem - tnp * hb);
}

struct TcParams {
  int64_t M, N;
  int num_kb;
  int tiles_m, tiles_n;
  const double* nA;
  const double* nB;
  double* C;
  int64_t ldc;
  double* Ct;
  int64_t ldct;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_split_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024 B alignment
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barriers (8 B each): full[STAGES], empty[STAGES], tmem_full[NUM_ACC], tmem_empty[NUM_ACC]; then the TMEM base slot
  const uint32_t full_bar = bar_base;
  const uint32_t empty_bar = bar_base + 8 * STAGES;
  const uint32_t tfull_bar = bar_base + 16 * STAGES;
  const uint32_t tempty_bar = tfull_bar + 8 * NUM_ACC;
  const uint32_t tmem_slot = tempty_bar + 8 * NUM_ACC;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_np = (p.tiles_n + 1) >> 1;  // pair-columns; the odd one out is a phantom tile (all columns masked)
  const int num_tiles = p.tiles_m * tiles_np;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 2);  // both CTAs of the pair must have consumed the stage (A is multicast)
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
  }
  if (warp == 1) {  // one warp allocates all 512 TMEM columns (1 CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loops, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_half = cta_rank * (A_SLICE_BYTES / 2);  // this CTA fetches rows [64*rank, 64*rank+64) of A
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int tm, tnp;
        decode_tile(tile, p.tiles_m, tiles_np, tm, tnp);
        const int tn = 2 * tnp + (int)cta_rank;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          mbar_expect_tx_elect(fb, STAGE_BYTES);  // 2 x 2 multicast A halves (own + peer's) + own B tiles
          const uint32_t st = smem_base + stage * STAGE_BYTES;
          tma_load_2d_mc_elect(st + a_half, &map_a_hi, fb, kb * BK, tm * BM + (int)cta_rank * (BM / 2), (uint16_t)3);
          tma_load_2d_mc_elect(st + A_SLICE_BYTES + a_half, &map_a_lo, fb, kb * BK, tm * BM + (int)cta_rank * (BM / 2),
                               (uint16_t)3);
          tma_load_2d_elect(st + 2 * A_SLICE_BYTES, &map_b_hi, fb, kb * BK, tn * BN);
          tma_load_2d_elect(st + 2 * A_SLICE_BYTES + B_SLICE_BYTES, &map_b_lo, fb, kb * BK, tn * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loops, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;  // chunk counter across tiles -> accumulator ring slot and parity
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        for (int kb0 = 0; kb0 < p.num_kb; kb0 += CHUNK_KB, ++g) {
          const uint32_t acc = g % NUM_ACC;
          const uint32_t acc_phase = (g / NUM_ACC) & 1;
          mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);  // epilogue drained this accumulator
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + acc * BN;
          const int kb1 = min(p.num_kb, kb0 + CHUNK_KB);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar + 8 * stage, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = smem_base + stage * STAGE_BYTES;
            const uint64_t da_hi = make_smem_desc(st);
            const uint64_t da_lo = make_smem_desc(st + A_SLICE_BYTES);
            const uint64_t db_hi = make_smem_desc(st + 2 * A_SLICE_BYTES);
            const uint64_t db_lo = make_smem_desc(st + 2 * A_SLICE_BYTES + B_SLICE_BYTES);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);  // +32 B per k-step inside the swizzle row
              umma_f16_elect(tmem_d, da_hi + adv, db_hi + adv, IDESC, (kb != kb0 || ks != 0) ? 1u : 0u);
              umma_f16_elect(tmem_d, da_hi + adv, db_lo + adv, IDESC, 1u);
              umma_f16_elect(tmem_d, da_lo + adv, db_hi + adv, IDESC, 1u);
            }
            umma_commit_mc_elect(empty_bar + 8 * stage, (uint16_t)3);  // stage free (in both CTAs) once these MMAs retire
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_elect(tfull_bar + 8 * acc);  // chunk accumulator complete
        }
      }
    }
  } else {
    // ===================== continuous epilogue (warps 2..9) =====================
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // which 64-column half of the tile
    uint32_t g = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int tm, tnp;
      decode_tile(tile, p.tiles_m, tiles_np, tm, tnp);
      const int tn = 2 * tnp + (int)cta_rank;
      double tot[EPI_COLS];
#pragma unroll
      for (int q = 0; q < EPI_COLS; ++q) tot[q] = 0.0;
      for (int kb0 = 0; kb0 < p.num_kb; kb0 += CHUNK_KB, ++g) {
        const uint32_t acc = g % NUM_ACC;
        const uint32_t acc_phase = (g / NUM_ACC) & 1;
        mbar_wait(tfull_bar + 8 * acc, acc_phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + half * EPI_COLS;
#pragma unroll
        for (int c = 0; c < EPI_COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int q = 0; q < 32; ++q) tot[c * 32 + q] += (double)__uint_as_float(r[q]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      }
      // tile done: reference scaling and stores
      const int64_t row = (int64_t)tm * BM + quad * 32 + lane;
      if (row < p.M) {
        const double na = p.nA[row];
        const int64_t col0 = (int64_t)tn * BN + half * EPI_COLS;
#pragma unroll
        for (int q = 0; q < EPI_COLS; ++q) {
          const int64_t col = col0 + q;
          if (col < p.N) {
            const double nn = na * __ldg(p.nB + col);
            // unit-vector dot (scaled by 2^16) * nn/(1e-10+nn) == dot(xc, yc)/(1e-10 + |xc||yc|)  (macrodna.py:25)
            const double v = tot[q] * (1.0 / 65536.0) * (nn / (1e-10 + nn));
            if (p.C) p.C[row * p.ldc + col] = v;
            if (p.Ct) p.Ct[col * p.ldct + row] = v;
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D map over one fp16 slice [rows, ldk] (K-major): box = 64 elements (128 B) x box_rows, 128 B swizzle.
bool make_map(CUtensorMap* map, const uint16_t* base, int64_t rows, int64_t ldk, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)ldk, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldk * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int mcd_launch_corr_split(mcd_context* h, const uint16_t* A2, const uint16_t* A_lo, int64_t M, const uint16_t* B2,
                          const uint16_t* B_lo, int64_t N, int64_t ldk16, const double* nA, const double* nB, double* C,
                          int64_t ldc, double* Ct, int64_t ldct) {
  if (M == 0 || N == 0) return MCD_OK;
  if ((reinterpret_cast<uintptr_t>(A2) & 15) || (reinterpret_cast<uintptr_t>(B2) & 15) ||
      (reinterpret_cast<uintptr_t>(A_lo) & 15) || (reinterpret_cast<uintptr_t>(B_lo) & 15) || (ldk16 % BK) != 0)
    return mcd_fail(h, MCD_ERR_INVALID, "corr_split: operands must be 16-byte aligned with ldk a multiple of 64");
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  if (!make_map(&ma_hi, A2, M, ldk16, BM / 2) || !make_map(&ma_lo, A_lo, M, ldk16, BM / 2) ||
      !make_map(&mb_hi, B2, N, ldk16, BN) || !make_map(&mb_lo, B_lo, N, ldk16, BN))
    return mcd_fail(h, MCD_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  TcParams p;
  p.M = M;
  p.N = N;
  p.num_kb = (int)(ldk16 / BK);
  p.tiles_m = (int)((M + BM - 1) / BM);
  p.tiles_n = (int)((N + BN - 1) / BN);
  p.nA = nA;
  p.nB = nB;
  p.C = C;
  p.ldc = ldc;
  p.Ct = Ct;
  p.ldct = ldct;
  MCD_CUDA(h, cudaFuncSetAttribute(corr_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int64_t pair_tiles = (int64_t)p.tiles_m * ((p.tiles_n + 1) / 2);
  const int64_t max_pairs = h->sm_count / 2;
  const int grid = 2 * (int)(pair_tiles < max_pairs ? pair_tiles : max_pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MCD_CUDA(h, cudaLaunchKernelEx(&cfg, corr_split_kernel, ma_hi, ma_lo, mb_hi, mb_lo, p));
  h->launches++;
  return MCD_OK;
}
