// K3 -- exact rectangular assignment on the GPU.
//
// Replaces one `ilp` call of the reference (src/MaCroDNA/macrodna.py:27-84): maximise
// sum_i W[i, col(i)] over injective maps of n "persons" (rows) into m >= n "objects"
// (columns).  The Gurobi model there is a totally unimodular assignment polytope, so the
// exact LAP optimum is the ILP optimum.
//
// Algorithm (all FP64, exact complementary slackness -- no epsilon in the final answer):
//   phase A  Jacobi forward auction in a persistent cooperative kernel.  Every unassigned
//            person scans its cost row against the current prices (best, second best),
//            bids  price += best - second  (the "naive" eps = 0 increment, which keeps
//            exact CS: the bidder is indifferent between its object and its runner-up),
//            one winner per object (64-bit atomicMax on (increment, person)), evicted owners
//            rejoin the bidder list.  Long rows are split into chunks over several CTAs
//            and merged by the last CTA to finish, so the many rounds with few bidders
//            are latency- not single-SM-bandwidth-bound.
//            For n == m (square, every object must be sold) the naive auction degenerates
//            into long price wars, so it is preceded by eps-scaling phases (eps = range/4,
//            /16, ... down to 1e-6*range); their only product is the price vector the final
//            eps = 0 phase starts from -- valid for square problems because every object ends
//            up assigned, so no "unassigned objects are cheapest" condition is needed.
//            For n < m all prices start at 0 and an object once assigned stays assigned, so
//            unassigned objects keep price 0 = the minimum: the rectangular optimality
//            condition holds by construction.
//   phase B  whatever the auction leaves (exact ties give zero increments -> no progress) is
//            finished by shortest-augmenting-path (Jonker-Volgenant/Dijkstra) steps that
//            start from the auction's dual-feasible prices/profits and tight partial matching.
//
// Bound: HBM/L2 bandwidth in the wide rounds (each bid reads one cost row: 8 B/object),
// launch/sync latency in the narrow ones.
#include <cstdlib>

#include "mcd_internal.cuh"

namespace {

// Grid-wide barrier for the persistent auction kernel (launched with cudaLaunchCooperativeKernel so
// all CTAs are co-resident).  One monotonic counter; each CTA's thread 0 arrives with a RELEASE
// reduction (MEMBAR.ALL.GPU + REDG) and spins on a relaxed gpu-scope load.  There is deliberately
// no acquire fence: on sm_100a every gpu-scope acquire (ld.acquire, fence.acq_rel, __threadfence,
// cooperative_groups grid.sync) ends in CCTL.IVALL, a whole-L1 invalidate that ncu shows costing
// ~4x the actual wait in this latency-bound round loop.  Instead every load of state that other
// CTAs mutate goes through ldm() = ld.global.cg (L2, the coherence point), so there is no stale L1
// line to drop.  The cost matrix W is immutable during the kernel and keeps the cached path.
struct GridBarrier {
  unsigned int* counter;
  unsigned int target;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned int seen;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      } while ((int)(seen - target) < 0);
    }
    __syncthreads();
  }
};

// load of inter-CTA mutable state: L2-coherent, never served from a stale L1 line
template <typename T>
__device__ __forceinline__ T ldm(const T* p) {
  return __ldcg(p);
}

constexpr int LAP_THREADS = 256;
constexpr int JV_THREADS = 1024;
constexpr double NEG_INF = -1.0e300;  // finite sentinel: inputs are finite correlations

struct LapCtrl {
  int cnt[2];       // bidder-list lengths (double buffered)
  int progress[2];  // successful bids per round (double buffered by round parity)
  int cur;          // which list is current
  int stalled;      // phase A ended with bidders left
  double wmin, wmax;
  unsigned int barrier;  // GridBarrier counter
  int pad[1];
};

struct LapState {
  const double* W;
  int n, m;
  int64_t ldw;
  int max_chunks;  // upper bound on chunks per row (partial-result slots = grid size)
  int vec;         // 128-bit loads legal
  double* price;     // [m]
  int* owner;        // [m]
  unsigned long long* key;  // [m]
  double* profit;    // [n]
  int* col4row;      // [n]  (caller's output)
  int* bj;           // [n] bid object per list slot
  double* gam;       // [n] bid increment per list slot
  double* bval;      // [n] bidder's value (W - price) of the object it bids on, per list slot
  int* un[2];        // [n] bidder lists
  int* done;         // [n] chunks finished per list slot
  double* pv1;       // [grid slots] partial best
  double* pv2;       // [grid slots] partial second
  int* pj1;          // [grid slots]
  int* pj2;          // [grid slots]
  // augmentation scratch
  double* sp;        // [m] shortest path cost
  int* pred;         // [m]
  int* sc_col;       // [n + 1] scanned columns in scan order
  double* sc_val;    // [n + 1] their path cost when scanned
  LapCtrl* ctrl;
  mcd_lap_counters* counters;
  double theta;        // eps-scaling factor (square case)
  double eps_min_rel;  // last scaling eps relative to the cost range
  long long max_rounds;
  int square_scaling;
};

struct Top2 {
  double v1, v2;
  int j1, j2;
};

__device__ __forceinline__ bool better(double va, int ja, double vb, int jb) {
  return va > vb || (va == vb && ja < jb);
}
__device__ __forceinline__ void top2_push(Top2& t, double v, int j) {
  if (better(v, j, t.v1, t.j1)) {
    t.v2 = t.v1;
    t.j2 = t.j1;
    t.v1 = v;
    t.j1 = j;
  } else if (better(v, j, t.v2, t.j2)) {
    t.v2 = v;
    t.j2 = j;
  }
}
// in-thread variant: candidates arrive in increasing j, so strict '>' keeps the smallest index on ties
__device__ __forceinline__ void top2_push_seq(Top2& t, double v, int j) {
  if (v > t.v1) {
    t.v2 = t.v1;
    t.j2 = t.j1;
    t.v1 = v;
    t.j1 = j;
  } else if (v > t.v2) {
    t.v2 = v;
    t.j2 = j;
  }
}
__device__ __forceinline__ void top2_merge(Top2& a, const Top2& b) {
  top2_push(a, b.v1, b.j1);
  top2_push(a, b.v2, b.j2);
}
__device__ __forceinline__ Top2 top2_shfl(const Top2& t, int o) {
  Top2 r;
  r.v1 = __shfl_xor_sync(0xffffffffu, t.v1, o);
  r.v2 = __shfl_xor_sync(0xffffffffu, t.v2, o);
  r.j1 = __shfl_xor_sync(0xffffffffu, t.j1, o);
  r.j2 = __shfl_xor_sync(0xffffffffu, t.j2, o);
  return r;
}

__device__ __forceinline__ unsigned long long pack_bid(double gamma, int person) {
  // increments are >= 0: the float32 bit pattern is monotone; person id breaks ties deterministically
  const float g = (float)gamma;
  return ((unsigned long long)__float_as_uint(g) << 32) | (unsigned)(person + 1);
}

// Record the bid of list slot k (person i) once its whole row has been scanned.
__device__ __forceinline__ void finalize_bid(const LapState& s, int k, int i, Top2 t, double eps) {
  int j = t.j1;
  if (eps == 0.0 && t.j2 >= 0 && t.v1 == t.v2 && ldm(&s.owner[j]) >= 0 && ldm(&s.owner[t.j2]) < 0) j = t.j2;  // exact tie
  const double gamma = (t.j2 >= 0 ? (t.v1 - t.v2) : 0.0) + eps;
  s.bj[k] = j;
  s.gam[k] = gamma;
  s.bval[k] = (j == t.j1) ? t.v1 : t.v2;
  atomicMax(&s.key[j], pack_bid(gamma, i));
}

// Phase A.  One cooperative launch runs every round of every eps phase.
__global__ void __launch_bounds__(LAP_THREADS) lap_auction_kernel(LapState s) {
  GridBarrier grid{&s.ctrl->barrier, 0u};
  __shared__ Top2 wred[LAP_THREADS / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int gtid = blockIdx.x * blockDim.x + tid;
  const int gthreads = gridDim.x * blockDim.x;
  LapCtrl* ctrl = s.ctrl;

  for (int j = gtid; j < s.m; j += gthreads) s.price[j] = 0.0;

  const double range = ctrl->wmax - ctrl->wmin;
  double eps = 0.0;
  const bool scaling = s.square_scaling && range > 0.0;
  if (scaling) eps = range / s.theta;
  long long rounds = 0, bids = 0, bytes = 0;
  long long tph[4] = {0, 0, 0, 0};
  bool guard_hit = false;

  for (;;) {  // eps phases
    // (re)start with everybody unassigned; prices are kept
    for (int j = gtid; j < s.m; j += gthreads) {
      s.owner[j] = -1;
      s.key[j] = 0ull;
    }
    for (int i = gtid; i < s.n; i += gthreads) {
      s.col4row[i] = -1;
      s.un[0][i] = i;
      s.done[i] = 0;
    }
    if (gtid == 0) {
      ctrl->cnt[0] = s.n;
      ctrl->cnt[1] = 0;
      ctrl->progress[0] = 0;
      ctrl->progress[1] = 0;
    }
    grid.sync();
    int cur = 0;
    int parity = 0;
    int nu = s.n;
    bool stalled = false;

    while (nu > 0) {
      if (rounds >= s.max_rounds) {
        guard_hit = true;
        break;
      }
      // ---- bidding: (list slot, chunk) work items over the whole grid
      const long long t0 = clock64();
      const int* un = s.un[cur];
      // few bidders: split every row over ~grid/nu CTAs so the round costs one memory latency, not a row sweep
      int nch = 1;
      if (nu < (int)gridDim.x) nch = min(s.max_chunks, (int)gridDim.x / nu);
      const int chunk = (((s.m + nch - 1) / nch) + 1) & ~1;
      const long long items = (long long)nu * nch;
      for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int k = (int)(item / nch);
        const int c = (int)(item - (long long)k * nch);
        const int i = ldm(&un[k]);
        const double* w = s.W + (int64_t)i * s.ldw;
        const int j0 = min(s.m, c * chunk);
        const int j1 = min(s.m, j0 + chunk);
        Top2 t{NEG_INF, NEG_INF, -1, -1};
        if (s.vec) {
          constexpr int S = 2 * LAP_THREADS;
          int j = j0 + 2 * tid;
          // 4 independent 128-bit load pairs in flight per thread (latency-bound when few rows are active)
          for (; j + 3 * S + 1 < j1; j += 4 * S) {
            const double2 w0 = __ldg(reinterpret_cast<const double2*>(w + j));
            const double2 w1 = __ldg(reinterpret_cast<const double2*>(w + j + S));
            const double2 w2 = __ldg(reinterpret_cast<const double2*>(w + j + 2 * S));
            const double2 w3 = __ldg(reinterpret_cast<const double2*>(w + j + 3 * S));
            const double2 p0 = ldm(reinterpret_cast<const double2*>(s.price + j));
            const double2 p1 = ldm(reinterpret_cast<const double2*>(s.price + j + S));
            const double2 p2 = ldm(reinterpret_cast<const double2*>(s.price + j + 2 * S));
            const double2 p3 = ldm(reinterpret_cast<const double2*>(s.price + j + 3 * S));
            top2_push_seq(t, w0.x - p0.x, j);
            top2_push_seq(t, w0.y - p0.y, j + 1);
            top2_push_seq(t, w1.x - p1.x, j + S);
            top2_push_seq(t, w1.y - p1.y, j + S + 1);
            top2_push_seq(t, w2.x - p2.x, j + 2 * S);
            top2_push_seq(t, w2.y - p2.y, j + 2 * S + 1);
            top2_push_seq(t, w3.x - p3.x, j + 3 * S);
            top2_push_seq(t, w3.y - p3.y, j + 3 * S + 1);
          }
          for (; j < j1; j += S) {
            if (j + 1 < j1) {
              const double2 wv = __ldg(reinterpret_cast<const double2*>(w + j));
              const double2 pv = ldm(reinterpret_cast<const double2*>(s.price + j));
              top2_push_seq(t, wv.x - pv.x, j);
              top2_push_seq(t, wv.y - pv.y, j + 1);
            } else {
              top2_push_seq(t, __ldg(w + j) - ldm(&s.price[j]), j);
            }
          }
        } else {
          for (int j = j0 + tid; j < j1; j += LAP_THREADS) top2_push_seq(t, __ldg(w + j) - ldm(&s.price[j]), j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          Top2 r = top2_shfl(t, o);
          top2_merge(t, r);
        }
        if ((tid & 31) == 0) wred[tid >> 5] = t;
        __syncthreads();
        if (tid == 0) {
#pragma unroll
          for (int wi = 1; wi < LAP_THREADS / 32; ++wi) top2_merge(t, wred[wi]);
          if (nch == 1) {
            finalize_bid(s, k, i, t, eps);
          } else {
            const int64_t slot = (int64_t)k * nch + c;
            s.pv1[slot] = t.v1;
            s.pv2[slot] = t.v2;
            s.pj1[slot] = t.j1;
            s.pj2[slot] = t.j2;
            int prev;  // release: the partial above is visible before the count; no L1-invalidating fence
            asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&s.done[k]) : "memory");
            if (prev == nch - 1) {
              Top2 a{NEG_INF, NEG_INF, -1, -1};
              for (int cc = 0; cc < nch; ++cc) {
                const int64_t sl = (int64_t)k * nch + cc;
                Top2 b{__ldcg(&s.pv1[sl]), __ldcg(&s.pv2[sl]), __ldcg(&s.pj1[sl]), __ldcg(&s.pj2[sl])};
                top2_merge(a, b);
              }
              s.done[k] = 0;
              finalize_bid(s, k, i, a, eps);
            }
          }
        }
        __syncthreads();
      }
      const long long t1 = clock64();
      grid.sync();
      const long long t2 = clock64();
      // ---- resolution: one thread per bidder; the winner of each object applies its bid
      int* nxt = s.un[cur ^ 1];
      for (int k = gtid; k < nu; k += gthreads) {
        const int i = ldm(&un[k]);
        const int j = ldm(&s.bj[k]);
        const unsigned long long kj = ldm(&s.key[j]);
        bool requeue = true;
        if ((unsigned)(kj & 0xffffffffull) == (unsigned)(i + 1)) {
          const double p_old = ldm(&s.price[j]);
          const double p_new = p_old + ldm(&s.gam[k]);
          const int prev = ldm(&s.owner[j]);
          if (prev < 0 || p_new > p_old) {
            if (prev >= 0) {
              s.col4row[prev] = -1;
              nxt[atomicAdd(&ctrl->cnt[cur ^ 1], 1)] = prev;
            }
            s.owner[j] = i;
            s.col4row[i] = j;
            s.price[j] = p_new;
            // profit := value of the owned object at its new price.  bval = fl(W - p_old) from the scan;
            // (bval + p_old) - p_new reproduces W - p_new to rounding, keeping the matched edge tight
            // at the 1-ulp level without re-reading W.
            s.profit[i] = (ldm(&s.bval[k]) + p_old) - p_new;
            atomicAdd(&ctrl->progress[parity], 1);
            requeue = false;
          }
          s.key[j] = 0ull;
        }
        if (requeue) nxt[atomicAdd(&ctrl->cnt[cur ^ 1], 1)] = i;
      }
      rounds++;
      bids += nu;
      bytes += (long long)nu * s.m * 8;
      const long long t3 = clock64();
      grid.sync();
      const long long t4 = clock64();
      tph[0] += t1 - t0;
      tph[1] += t2 - t1;
      tph[2] += t3 - t2;
      tph[3] += t4 - t3;
      const int nu_next = ldm(&ctrl->cnt[cur ^ 1]);
      const int prog = ldm(&ctrl->progress[parity]);
      if (gtid == 0) {
        ctrl->cnt[cur] = 0;            // becomes the "next" list of the coming round
        ctrl->progress[parity ^ 1] = 0;
      }
      cur ^= 1;
      parity ^= 1;
      nu = nu_next;
      if (prog == 0 && nu > 0) {
        stalled = true;
        break;
      }
    }
    if (guard_hit || stalled || eps == 0.0) {
      if (gtid == 0) {
        ctrl->cur = cur;
        ctrl->stalled = nu > 0 ? 1 : 0;
      }
      break;
    }
    // next eps phase (square case only)
    eps /= s.theta;
    if (eps < s.eps_min_rel * range) eps = 0.0;
    grid.sync();  // everyone has read ctrl->cnt before the phase reset rewrites it
  }
  if (gtid == 0) {
    s.counters->rounds += rounds;
    s.counters->bids += bids;
    s.counters->bytes += bytes;
    for (int q = 0; q < 4; ++q) s.counters->t_phase[q] += tph[q];
  }
}

// Block-wide arg-min with the (value, prefer-unowned, index) order of the augmentation step.
struct MinItem {
  double v;
  int j;
  int free_;  // 1 if the column is unowned
};
__device__ __forceinline__ bool min_better(const MinItem& a, const MinItem& b) {
  if (a.v != b.v) return a.v < b.v;
  if (a.free_ != b.free_) return a.free_ > b.free_;
  return a.j < b.j;
}

// Phase B: shortest augmenting paths for the persons phase A left unassigned.  One CTA; the
// per-step work is one cost row (m objects) spread over 1024 threads.  min-form duals:
// u_i = -profit_i, v_j = -price_j, cost' = -W.
__global__ void __launch_bounds__(JV_THREADS) lap_augment_kernel(LapState s) {
  LapCtrl* ctrl = s.ctrl;
  if (!ctrl->stalled) return;
  __shared__ MinItem red[JV_THREADS / 32];
  __shared__ MinItem best;
  const int tid = threadIdx.x;
  const int cur_list = ctrl->cur;
  const int nfree = ctrl->cnt[cur_list];
  const int* freelist = s.un[cur_list];
  long long steps = 0, bytes = 0;

  for (int f = 0; f < nfree; ++f) {
    const int cur_row = freelist[f];
    for (int j = tid; j < s.m; j += JV_THREADS) s.sp[j] = 1.0e300;
    __syncthreads();
    double minval = 0.0;
    int i = cur_row;
    double ui = 0.0;  // dual of the row being scanned (min-form); the free row starts at 0
    int nsc = 0;
    int sink = -1;
    while (sink < 0) {
      const double* w = s.W + (int64_t)i * s.ldw;
      MinItem loc{1.0e300, 0x7fffffff, 0};
      for (int j = tid; j < s.m; j += JV_THREADS) {
        double spj = s.sp[j];
        if (spj == NEG_INF) continue;  // already scanned
        const double r = minval - w[j] - ui + s.price[j];
        if (r < spj) {
          spj = r;
          s.sp[j] = r;
          s.pred[j] = i;
        }
        MinItem it{spj, j, s.owner[j] < 0 ? 1 : 0};
        if (min_better(it, loc)) loc = it;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        MinItem ot{__shfl_xor_sync(0xffffffffu, loc.v, o), __shfl_xor_sync(0xffffffffu, loc.j, o),
                   __shfl_xor_sync(0xffffffffu, loc.free_, o)};
        if (min_better(ot, loc)) loc = ot;
      }
      if ((tid & 31) == 0) red[tid >> 5] = loc;
      __syncthreads();
      if (tid < 32) {
        MinItem x = red[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          MinItem ot{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.j, o),
                     __shfl_xor_sync(0xffffffffu, x.free_, o)};
          if (min_better(ot, x)) x = ot;
        }
        if (tid == 0) {
          best = x;
          s.sc_col[nsc] = x.j;
          s.sc_val[nsc] = x.v;
          s.sp[x.j] = NEG_INF;
        }
      }
      __syncthreads();
      const MinItem b = best;
      minval = b.v;
      nsc++;
      steps++;
      bytes += (long long)s.m * 8;
      if (b.free_) {
        sink = b.j;
      } else {
        i = s.owner[b.j];
        ui = -s.profit[i];
      }
      __syncthreads();
    }
    // dual update (every scanned column but the sink is owned; its owner is the row it led to)
    for (int q = tid; q < nsc; q += JV_THREADS) {
      const int j = s.sc_col[q];
      const double d = minval - s.sc_val[q];
      s.price[j] += d;  // v_j -= d
      const int r = s.owner[j];
      if (r >= 0) s.profit[r] -= d;  // u_r += d
    }
    __syncthreads();
    // augment along the predecessor chain (sequential, short)
    if (tid == 0) {
      s.profit[cur_row] = -minval;  // u_cur = 0 + minval
      int j = sink;
      for (;;) {
        const int r = s.pred[j];
        s.owner[j] = r;
        const int jn = s.col4row[r];
        s.col4row[r] = j;
        if (r == cur_row) break;
        j = jn;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    s.counters->aug_rows += nfree;
    s.counters->aug_steps += steps;
    s.counters->bytes += bytes;
    ctrl->stalled = 0;
  }
}

__global__ void __launch_bounds__(1024) lap_minmax_kernel(const double* __restrict__ W, int n, int m, int64_t ldw,
                                                          LapCtrl* ctrl) {
  __shared__ double smin[32], smax[32];
  double lo = 1.0e300, hi = -1.0e300;
  const int64_t total = (int64_t)n * m;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const double v = W[(e / m) * ldw + (e % m)];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = lo;
    smax[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = smin[threadIdx.x];
    hi = smax[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (threadIdx.x == 0) {
      // doubles of one sign order like their bit patterns; use CAS loops for generality
      unsigned long long* pmin = reinterpret_cast<unsigned long long*>(&ctrl->wmin);
      unsigned long long old = *pmin;
      while (__longlong_as_double((long long)old) > lo) {
        const unsigned long long seen = atomicCAS(pmin, old, (unsigned long long)__double_as_longlong(lo));
        if (seen == old) break;
        old = seen;
      }
      unsigned long long* pmax = reinterpret_cast<unsigned long long*>(&ctrl->wmax);
      old = *pmax;
      while (__longlong_as_double((long long)old) < hi) {
        const unsigned long long seen = atomicCAS(pmax, old, (unsigned long long)__double_as_longlong(hi));
        if (seen == old) break;
        old = seen;
      }
    }
  }
}

__global__ void lap_ctrl_init_kernel(LapCtrl* ctrl, mcd_lap_counters* counters, int zero_counters) {
  ctrl->cnt[0] = ctrl->cnt[1] = 0;
  ctrl->progress[0] = ctrl->progress[1] = 0;
  ctrl->cur = 0;
  ctrl->stalled = 0;
  ctrl->wmin = 1.0e300;
  ctrl->wmax = -1.0e300;
  ctrl->barrier = 0u;
  if (zero_counters) {
    counters->rounds = counters->bids = counters->bytes = counters->aug_rows = counters->aug_steps = 0;
    counters->status = 0;
    for (int q = 0; q < 4; ++q) counters->t_phase[q] = 0;
  }
}

// objective = sum_i W[i, col4row[i]], fixed summation order (deterministic bits)
__global__ void __launch_bounds__(1024) lap_objective_kernel(const double* __restrict__ W, int n, int64_t ldw,
                                                             const int* __restrict__ col4row, double* out,
                                                             mcd_lap_counters* counters) {
  __shared__ double red[32];
  double acc = 0.0;
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const int j = col4row[i];
    if (j < 0)
      bad = 1;
    else
      acc += W[(int64_t)i * ldw + j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  bad = __syncthreads_or(bad);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      if (out != nullptr) *out = acc;
      if (bad) counters->status = 1;
    }
  }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int MAX_GRID_SLOTS = 4096;  // >= cooperative grid size (sm_count * blocks/SM)

int pick_max_chunks(int64_t m) {
  const char* e = getenv("MCD_LAP_MIN_CHUNK");
  int min_chunk = e ? atoi(e) : 4096;  // objects per CTA below which splitting stops paying
  if (min_chunk < 2) min_chunk = 2;
  int64_t mc = m / min_chunk;
  if (mc < 1) mc = 1;
  if (mc > 256) mc = 256;
  return (int)mc;
}

}  // namespace

size_t mcd_lap_workspace_bytes(int64_t n, int64_t m) {
  const int64_t nch = 1;
  (void)nch;
  size_t b = 0;
  b += align_up(sizeof(LapCtrl), 256);
  b += align_up(m * 8, 256);          // price
  b += align_up(m * 4, 256);          // owner
  b += align_up(m * 8, 256);          // key
  b += align_up(n * 8, 256);          // profit
  b += align_up(n * 4, 256);          // bj
  b += align_up(n * 8, 256);          // gam
  b += align_up(n * 8, 256);          // bval
  b += 2 * align_up(n * 4, 256);      // un lists
  b += align_up(n * 4, 256);          // done
  b += 2 * align_up(MAX_GRID_SLOTS * 8, 256);  // pv1 pv2
  b += 2 * align_up(MAX_GRID_SLOTS * 4, 256);  // pj1 pj2
  b += align_up(m * 8, 256);          // sp
  b += align_up(m * 4, 256);          // pred
  b += align_up((n + 1) * 4, 256);    // sc_col
  b += align_up((n + 1) * 8, 256);    // sc_val
  return b;
}

int mcd_launch_lap(mcd_context* h, const double* W, int64_t n, int64_t m, int64_t ldw, int32_t* col4row,
                   double* objective, void* work, mcd_lap_counters* d_counters) {
  if (n <= 0) return MCD_OK;
  if (n > m) return mcd_fail(h, MCD_ERR_INVALID, "lap: rows must be the smaller side");
  if (m > 0x3fffffff) return mcd_fail(h, MCD_ERR_UNSUPPORTED, "lap: too many objects");
  LapState s;
  s.W = W;
  s.n = (int)n;
  s.m = (int)m;
  s.ldw = ldw;
  s.max_chunks = pick_max_chunks(m);
  s.vec = ((reinterpret_cast<uintptr_t>(W) & 15) == 0) && ((ldw & 1) == 0);
  char* p = static_cast<char*>(work);
  auto take = [&](size_t bytes) {
    char* r = p;
    p += align_up(bytes, 256);
    return r;
  };
  s.ctrl = reinterpret_cast<LapCtrl*>(take(sizeof(LapCtrl)));
  s.price = reinterpret_cast<double*>(take(m * 8));
  s.owner = reinterpret_cast<int*>(take(m * 4));
  s.key = reinterpret_cast<unsigned long long*>(take(m * 8));
  s.profit = reinterpret_cast<double*>(take(n * 8));
  s.bj = reinterpret_cast<int*>(take(n * 4));
  s.gam = reinterpret_cast<double*>(take(n * 8));
  s.bval = reinterpret_cast<double*>(take(n * 8));
  s.un[0] = reinterpret_cast<int*>(take(n * 4));
  s.un[1] = reinterpret_cast<int*>(take(n * 4));
  s.done = reinterpret_cast<int*>(take(n * 4));
  s.pv1 = reinterpret_cast<double*>(take(MAX_GRID_SLOTS * 8));
  s.pv2 = reinterpret_cast<double*>(take(MAX_GRID_SLOTS * 8));
  s.pj1 = reinterpret_cast<int*>(take(MAX_GRID_SLOTS * 4));
  s.pj2 = reinterpret_cast<int*>(take(MAX_GRID_SLOTS * 4));
  s.sp = reinterpret_cast<double*>(take(m * 8));
  s.pred = reinterpret_cast<int*>(take(m * 4));
  s.sc_col = reinterpret_cast<int*>(take((n + 1) * 4));
  s.sc_val = reinterpret_cast<double*>(take((n + 1) * 8));
  s.col4row = col4row;
  s.counters = d_counters;
  const char* e;
  s.theta = (e = getenv("MCD_LAP_THETA")) ? atof(e) : 4.0;
  if (!(s.theta > 1.0)) s.theta = 4.0;
  s.eps_min_rel = (e = getenv("MCD_LAP_EPS_MIN")) ? atof(e) : 1e-6;
  s.square_scaling = (n == m && n > 1) ? 1 : 0;
  if ((e = getenv("MCD_LAP_NO_SCALING")) && atoi(e)) s.square_scaling = 0;
  s.max_rounds = 200000 + 64 * (long long)n;
  if ((e = getenv("MCD_LAP_MAX_ROUNDS"))) s.max_rounds = atoll(e);

  lap_ctrl_init_kernel<<<1, 1, 0, h->stream>>>(s.ctrl, d_counters, 0);
  MCD_LAUNCH_CHECK(h, "lap_ctrl_init_kernel");
  if (s.square_scaling) {
    lap_minmax_kernel<<<h->sm_count * 2, 1024, 0, h->stream>>>(W, s.n, s.m, ldw, s.ctrl);
    MCD_LAUNCH_CHECK(h, "lap_minmax_kernel");
  }
  int per_sm = 0;
  MCD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lap_auction_kernel, LAP_THREADS, 0));
  if (per_sm < 1) return mcd_fail(h, MCD_ERR_CUDA, "lap_auction_kernel cannot be resident");
  int want = (e = getenv("MCD_LAP_BLOCKS_PER_SM")) ? atoi(e) : 4;
  if (want < 1) want = 1;
  int blocks = h->sm_count * (per_sm < want ? per_sm : want);
  if (blocks > MAX_GRID_SLOTS) blocks = MAX_GRID_SLOTS;
  void* args[] = {&s};
  MCD_CUDA(h, cudaLaunchCooperativeKernel((const void*)lap_auction_kernel, dim3(blocks), dim3(LAP_THREADS), args, 0,
                                          h->stream));
  h->launches++;
  lap_augment_kernel<<<1, JV_THREADS, 0, h->stream>>>(s);
  MCD_LAUNCH_CHECK(h, "lap_augment_kernel");
  lap_objective_kernel<<<1, 1024, 0, h->stream>>>(W, s.n, ldw, col4row, objective, d_counters);
  MCD_LAUNCH_CHECK(h, "lap_objective_kernel");
  return MCD_OK;
}
