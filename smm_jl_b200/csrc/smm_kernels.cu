// smm_kernels.cu -- sm_100a kernels of the BGP hot path.
//
// Three ways to run an iteration, same device functions, bit-identical results:
//
//  (A) multi-launch (exchange_mode 0):
//      bgp_eval_kernel      grid (n_split, L) x 128 threads: proposal -> simulate -> [last CTA of the chain]
//                           moments, distance, accept/reject, trace.
//      [ncclAllGather of the last-accepted records when world > 1]
//      bgp_exchange_kernel  exchangeMoves! on the gathered records.
//
//  (B) persistent with grid barriers (exchange_mode 1): bgp_persistent_kernel<false>, ONE cooperative launch for up
//      to kPairChunk iterations, one kPersistThreads-thread CTA per SM, two grid barriers per iteration:
//        [exchange of iteration i-1 (replicated per CTA), proposals of iteration i for the chains the
//         CTA owns, several chains at a time in warp groups]                               -- barrier --
//        [every CTA takes an equal share of the flattened (chain, draw) space; inside the CTA every warp walks a
//         fixed share and then pulls shrinking grabs of draws from a shared-memory counter, so all warps finish
//         together whatever the warp scheduler favours; the CTA that completes a chain finishes it (moments,
//         distance, accept/reject, trace) in one warp while the other warps already simulate the next chain;
//         with world > 1 the finished record is stored straight into every peer GPU's gather buffer
//         over NVLink]                                                                      -- barrier,
//         folded with a cross-GPU flag exchange: the all-gather costs no launch and no extra barrier --
//
//  (C) persistent, barrier-free (exchange_mode 2; exchange_mode 3 in the -DSMM_LL_TU build): bgp_persistent_kernel<true>.  The warp that finishes a chain
//      publishes a per-chain completion tag; every CTA waits for the tags of iteration i-1, replays the exchange,
//      computes the proposals of the chains IT simulates, simulates, finishes -- no grid barrier anywhere.
//
// Reference lines: proposal AlgoBGP.jl:424-471 (mysample :400-410, mapto_01/ab mprob.jl:246-272);
// objfunc_norm ObjExamples.jl:59-116; doAcceptReject! AlgoBGP.jl:324-392; set_eval! :220-245;
// set_acceptRate! :253-257; exchangeMoves! :647-691; swap_ev_ij! :734-749; pair sample :653-656.
//
// Draws never touch memory: a thread owns one simulated dimension k, generates its normals in registers
// (Philox4x32-10 + the 256-layer ziggurat of include/smm_stream.h, rare branch deferred to warp-sized batches) and
// adds them to ORDER-INVARIANT accumulators (see simulate_*), so the result does not depend on how the draw space is
// cut up.  Proposals and the panel simulator use the Box-Muller transform of the same header.
#include "smm_device.cuh"

// This file is compiled TWICE (smm_jl_b200/build.py): as it stands, and with -DSMM_LL_TU, which adds exchange_mode 3
// (flag-in-data hand-over, see ll_store) to the barrier-free persistent kernel and puts everything into smm::ll, from
// which smm_api.cu takes only the persistent launcher.  Two builds instead of one more template parameter because the
// persistent kernel sits at its register budget: with the mode-3 code in the same instantiation family, ptxas moved
// spills into the simulate loop of the OTHER modes (measured: 56.6 -> 58.8-61.4 us per C2 iteration in exchange_mode 2
// from code that never executes there).
namespace smm {
#ifdef SMM_LL_TU
namespace ll {
#endif

// ------------------------------------------------------------------------------------------------
// thread groups: a CTA or a warp-aligned part of one, synchronised with a named barrier
// ------------------------------------------------------------------------------------------------
struct Grp {
  int tid;  // index of this thread in the group
  int n;    // threads in the group (multiple of 32)
  int bar;  // named barrier id (0 = the CTA-wide barrier, i.e. __syncthreads when n == blockDim.x)
};
__device__ __forceinline__ void gsync(const Grp &g) {
  if (g.n == 32)
    __syncwarp();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(g.bar), "r"(g.n) : "memory");
}

// scratch of one proposal (shared memory)
struct PropScratch {
  double *pp;              // [P]  out: proposed parameter vector
  double *mu01;            // [P]
  double *cand;            // [attempts per round][P]
  unsigned char *okf;      // [attempts per round][P]
  int *first;              // [P]  per batch: first in-support attempt of this round
  int *resolved;           // [P]  per batch: attempts used (0 = unresolved)
  const smm_logent *logtab;
};

// scratch of one finalisation (shared memory)
struct FinScratch {
  double *pp;    // [P]   parameters of the evaluation
  double *tot;   // [2D]  totals
  double *mom;   // [M]   simulated moments
  double *value; // [2]   value, prob
  int *flags;    // [2]   accepted, status
};

// acquire/release accesses (PTX memory model) -- lighter than __threadfence(), which is a sequentially
// consistent fence plus an L1 invalidate; every cross-CTA read in this file bypasses L1 (__ldcg) anyway
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_acq_rel_gpu(unsigned *p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned *p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;  // 4 s: a stuck peer becomes an error, not a hang

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");  // "memory": stays on its side of barriers
  return t;
}
#define PHASE_STAMP(slot, i)                                                                  \
  do {                                                                                        \
    if (st.phase_ts && threadIdx.x == 0) st.phase_ts[(size_t)(slot)*4 + (i)] = gtimer();      \
  } while (0)

#ifdef SMM_LL_TU
// ---- flag-in-data words (exchange_mode 3; this translation unit is compiled a second time with -DSMM_LL_TU) ----------
// A double travels as two 8-byte words {high half | tag}, {low half | tag}; an 8-byte store is single-copy atomic, so a
// reader that sees the tag sees the payload -- no fence and no separate flag round trip between the producer's store and
// the consumer's load (the LL protocol of collective libraries).  tag = iteration (never 0; the table starts zeroed).
__device__ __forceinline__ void ll_store(unsigned long long *p, double v, uint32_t tag) {
  const unsigned long long a = ((unsigned long long)(uint32_t)__double2hiint(v) << 32) | tag;
  const unsigned long long b = ((unsigned long long)(uint32_t)__double2loint(v) << 32) | tag;
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// spin until both words carry `tag`; false = gave up (another CTA aborted, or the timeout: sticky error flag)
__device__ __forceinline__ bool ll_load(const DevState &st, const unsigned long long *p, uint32_t tag, double &v) {
  unsigned long long a, b, t0 = 0ull;
  for (unsigned spins = 0;;) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    if ((uint32_t)a == tag && (uint32_t)b == tag) break;
    __nanosleep(40);  // back off: thousands of threads polling L2 in a tight loop slow down the chains still finishing
    if ((++spins & 255u) == 0) {
      if (t0 == 0ull) t0 = gtimer();
      const bool aborted = (*(volatile unsigned *)&st.bar->gen & 0x80000000u) != 0u;
      if (aborted || gtimer() - t0 > kSpinTimeoutNs) {
        if (!aborted) {
          atomicOr(st.err, kErrTimeout);
          atomicOr(&st.bar->gen, 0x80000000u);
        }
        v = 0.0;
        return false;
      }
    }
  }
  v = __hiloint2double((int)(a >> 32), (int)(b >> 32));
  return true;
}
#endif

__device__ __forceinline__ void load_logtab(smm_logent *dst) {
  const smm_logent *src = smm_logtab();
  for (int i = threadIdx.x; i < (1 << SMM_LOG_BITS); i += blockDim.x) dst[i] = src[i];
}

// Philox4x32-10 with the round keys of seed_sim taken from the kernel-parameter constant bank
__device__ __forceinline__ smm_u32x4 philox_sim(const DevProblem &pb, uint32_t c0, uint32_t c1, uint32_t c2,
                                                uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)SMM_PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)SMM_PHILOX_M1 * c2;
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ pb.rk_sim0[r];
    c1 = (uint32_t)p1;
    c2 = (uint32_t)(p0 >> 32) ^ c3 ^ pb.rk_sim1[r];
    c3 = (uint32_t)p0;
  }
  smm_u32x4 out;
  out.x = c0;
  out.y = c1;
  out.z = c2;
  out.w = c3;
  return out;
}

// ------------------------------------------------------------------------------------------------
// proposal(c) -- AlgoBGP.jl:424-471.  Attempts are counter-indexed, so a round evaluates g.n/kp attempts
// at once and the lowest in-support attempt wins: identical to the reference's sequential rejection loop.
// Result in ps.pp.
// ------------------------------------------------------------------------------------------------
// centre_sh (exchange_mode 3 only): shared [P + 1] = the centre's parameters and the chain's sigma, already fetched.
__device__ void group_proposal(const DevProblem &pb, const DevState &st, const Grp &g, const PropScratch &ps, int c,
                               int gc, int iter, bool count, const double *centre = nullptr
#ifdef SMM_LL_TU
                               ,
                               const double *centre_sh = nullptr
#endif
) {
  const int P = pb.P, tid = g.tid, nthr = g.n;
  if (iter == 1) {
    for (int k = tid; k < P; k += nthr) ps.pp[k] = pb.init[k];
    gsync(g);
    return;
  }
  const int R = rec_len(pb.P, pb.M);
  const double *la = centre ? centre : st.la_cur + (size_t)c * R;  // record of the last accepted evaluation
#ifdef SMM_LL_TU
  const double sigma = centre_sh ? centre_sh[P] : __ldcg(st.sigma + c);
#else
  const double sigma = __ldcg(st.sigma + c);
#endif
  const int kp = (P + 1) >> 1;
  const int A = nthr / kp;  // attempts per round
  const int bs = pb.batch_size, nb = P / bs;
  for (int k = tid; k < P; k += nthr) {
#ifdef SMM_LL_TU
    ps.mu01[k] = __ddiv_rn(__dsub_rn(centre_sh ? centre_sh[k] : __ldcg(la + 3 + k), pb.lb[k]), __dsub_rn(pb.ub[k], pb.lb[k]));
#else
    ps.mu01[k] = __ddiv_rn(__dsub_rn(__ldcg(la + 3 + k), pb.lb[k]), __dsub_rn(pb.ub[k], pb.lb[k]));
#endif
    ps.pp[k] = 0.0;  // pp = zero(mu01) (:445)
  }
  for (int b = tid; b < nb; b += nthr) ps.resolved[b] = 0;
  gsync(g);
  int unresolved = nb;
  for (int base = 0; base < pb.smpl_iters && unresolved > 0; base += A) {
    for (int b = tid; b < nb; b += nthr) ps.first[b] = 0x7fffffff;
    if (tid < A * kp) {
      const int a_loc = tid / kp, kq = tid - a_loc * kp;
      const int a = base + a_loc;
      if (a < pb.smpl_iters) {
        double z0, z1;
        smm_normal_pair_tab(smm_prop_block(pb.seed_algo, (uint32_t)gc, (uint32_t)iter, (uint32_t)a, (uint32_t)kq),
                            ps.logtab, &z0, &z1);
        const int k0 = 2 * kq, k1 = k0 + 1;
        const double x0 = __dadd_rn(ps.mu01[k0], __dmul_rn(sigma, z0));
        ps.cand[a_loc * P + k0] = x0;
        ps.okf[a_loc * P + k0] = (x0 >= 0.0) && (x0 <= 1.0);
        if (k1 < P) {
          const double x1 = __dadd_rn(ps.mu01[k1], __dmul_rn(sigma, z1));
          ps.cand[a_loc * P + k1] = x1;
          ps.okf[a_loc * P + k1] = (x1 >= 0.0) && (x1 <= 1.0);
        }
      }
    }
    gsync(g);
    for (int t = tid; t < A * nb; t += nthr) {
      const int a_loc = t / nb, b = t - a_loc * nb;
      if (ps.resolved[b] == 0 && base + a_loc < pb.smpl_iters) {
        bool ok = true;
        for (int k = b * bs; k < (b + 1) * bs; ++k) ok = ok && ps.okf[a_loc * P + k];
        if (ok) atomicMin(&ps.first[b], a_loc);
      }
    }
    gsync(g);
    for (int k = tid; k < P; k += nthr) {
      const int b = k / bs;
      if (ps.resolved[b] == 0 && ps.first[b] != 0x7fffffff) ps.pp[k] = ps.cand[ps.first[b] * P + k];
    }
    int still = 0;
    for (int b = 0; b < nb; ++b)  // every thread computes the same count (nb <= 64)
      if (ps.resolved[b] == 0 && ps.first[b] == 0x7fffffff) ++still;
    unresolved = still;
    gsync(g);
    for (int b = tid; b < nb; b += nthr)
      if (ps.resolved[b] == 0 && ps.first[b] != 0x7fffffff) ps.resolved[b] = base + ps.first[b] + 1;
    gsync(g);
  }
  if (unresolved > 0 && nb == 1) {
    // single batch: `error("no draw in support ...")` (:409) aborts the run -> sticky error flag;
    // (several batches: the exception is logged and swallowed, pp[i] stays 0, :447-451)
    if (tid == 0) atomicOr(st.err, kErrExhausted);
    for (int k = tid; k < P; k += nthr) ps.pp[k] = ps.mu01[k];
  }
  if (count && tid == 0) {
    unsigned long long att = 0;
    for (int b = 0; b < nb; ++b) att += ps.resolved[b] ? ps.resolved[b] : pb.smpl_iters;
    atomicAdd(&st.counters[2], att);
  }
  gsync(g);
  for (int k = tid; k < P; k += nthr)
    ps.pp[k] = __dadd_rn(__dmul_rn(ps.pp[k], __dsub_rn(pb.ub[k], pb.lb[k])), pb.lb[k]);
  gsync(g);
}

// ------------------------------------------------------------------------------------------------
// Simulation of the MvNormal objectives (ObjExamples.jl:76-79): X[k,s] = p_k + Z[k,s], reduced on the
// fly to sum_s X and sum_s X^2 per row k, for Philox blocks j in [j0, j1) of an evaluation.
//
// ORDER-INVARIANT ACCUMULATION.  Every term is rounded once to a fixed-point grid (x -> x + M with
// M = 1.5 * 2^(52-F): the low mantissa bits of the sum are round(x * 2^F)) and the 64-bit patterns are
// added as integers, so the totals do not depend on how draws are split over threads, warps, CTAs, launches
// or GPUs: chains with equal parameters get bit-equal values (the exchange step compares values; ties must
// stay ties as in the sequential reference), and 1-GPU and N-GPU runs agree to the bit.  Grid error per
// term <= 2^-(F+1), F chosen by the host from S and the parameter box (2^-43 / 2^-37 for the C2 shapes).
// ------------------------------------------------------------------------------------------------
struct Acc {
  unsigned long long sum, sq;  // sums of raw bit patterns; the n * bits(M) offset is removed at the end
};

// ------------------------------------------------------------------------------------------------
// Zsim of the MvNormal objectives is a ziggurat (include/smm_stream.h): ONE Philox block gives THREE normals, and
// 99.2 % of the draws take the FAST path (one 8-byte table entry from shared memory, a shift, a compare, two fp64
// operations).  The rest -- wedge test with exp, tail with two logs, retries -- would cost every warp a divergent detour
// on half of its steps, so it is DEFERRED: the fast path always accumulates its candidates; a lane whose block holds a
// rejected candidate pushes the block's index onto its warp's queue (ballot + popc, two stores), and when 32 entries
// have gathered the whole warp re-derives those blocks, one per lane, resolves them with smm_zig_normal_tab and adds
// [pattern(true block sums) - pattern(candidate block sums)] to the accumulators with 64-bit integer atomics.
// Integer accumulation is exact and order-free, so the totals are those of the sequential definition.
//
// BLOCK-WISE ACCUMULATION.  The three values X = p + Z of a block are summed in double, in stream order,
//   s = fl(fl(X0 + X1) + X2),  q = fl(fma(X2, X2, fl(fma(X1, X1, fl(X0 X0)))))      (missing draws of the last block: +0)
// and s, q are what is rounded to the fixed-point grid -- a block always lives in one lane, so this is as independent of
// the work split as a per-draw rounding, at a third of the integer adds.
// ------------------------------------------------------------------------------------------------
constexpr int kZigQCap = 64;                    // entries per warp: < 32 before a step, <= 63 after its push
constexpr int kZigQWords = 2 * kZigQCap;        // {j, kpack} pairs (8-byte aligned)
constexpr int kZigSigned = 2 * SMM_ZIG_LAYERS;  // shared-memory table: one 8-byte entry per (sign, layer)

// Everything the hot loop needs lives in 32-bit registers (shared-window addresses, not generic pointers).
struct ZigCtx {
  uint32_t ztab;  // shared address of the signed layer table: entry s = {+-W'[i] 2^-32 | KH[i]}, on an 8 KB boundary
  uint32_t hi52;  // 0x43300000 (high word of 2^52) held in a register, so that building 2^52 + u costs no MOV per draw
  uint32_t q;     // shared address of this warp's queue
  uint32_t lt;    // %lanemask_lt
  uint32_t pvec;  // shared address: [D] parameters of the evaluation being simulated (f64)
  uint32_t fix;   // shared address: [2D] u64 where corrections are added
  uint32_t c2, c3;  // counter words 2, 3 of the blocks being simulated (what the drain needs to re-derive a block)
  uint32_t segc2;   // persistent kernel: shared address of c2 per segment (queue entries carry their segment); 0 = use c2
  int D;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// The signed copy of the layer table (index = sign << 9 | layer; the sign bit is set in the entries of the upper half, so
// the sign costs nothing) must start on an 8 KB boundary of the shared window, because zig_fast_dev forms entry
// addresses with OR.  A declared __align__(8192) is not enough: static shared memory starts 1 KB into the window on
// sm_100, and the compiler aligns relative to that.  So the kernels reserve 16 KB and use the aligned half.
constexpr int kZigBufEntries = 2 * kZigSigned;
static_assert(kZigSigned * 8 == 8192, "address masks below assume an 8 KB table");
__device__ __forceinline__ uint32_t zig_table_addr(const unsigned long long *buf) { return (smem_addr(buf) + 8191u) & ~8191u; }
__device__ __forceinline__ void load_zigtab(unsigned long long *buf) {
  unsigned long long *dst = buf + (zig_table_addr(buf) - smem_addr(buf)) / 8u;
  const smm_zigent *src = smm_zigtab();
  for (int i = threadIdx.x; i < kZigSigned; i += blockDim.x)
    dst[i] = src[i & (SMM_ZIG_LAYERS - 1)] | (i >= SMM_ZIG_LAYERS ? 0x8000000000000000ull : 0ull);
}
// `opaque` is any kernel argument known to be non-negative: OR-ing its sign bit into the table address keeps ptxas from
// folding it back into an immediate (a LOP3 takes one immediate; with the base in a register the mask-and-merge of the
// entry address is a single instruction)
__device__ __forceinline__ ZigCtx zig_ctx(const unsigned long long *zbuf, uint32_t *q_warp, int opaque) {
  ZigCtx cx;
  const uint32_t zero = (uint32_t)opaque >> 31;
  cx.ztab = zig_table_addr(zbuf) | zero;
  cx.hi52 = 0x43300000u | zero;
  cx.q = smem_addr(q_warp);
  asm volatile("" : "+r"(cx.q));  // opaque: keep the address in a register instead of recomputing it from %tid per push
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(cx.lt));
  cx.pvec = 0;
  cx.fix = 0;
  cx.c2 = 0;
  cx.c3 = 0;
  cx.segc2 = 0;
  cx.D = 0;
  return cx;
}

// candidate of the fast path (same value as smm_zig_fast) and whether it is final.  `sel8` carries the draw's select
// field (sign, layer) in bits 3..12, anything elsewhere:  LOP3, LDS.64, SHF, ISETP, DADD, DMUL
__device__ __forceinline__ double zig_fast_dev(uint32_t u, uint32_t sel8, const ZigCtx &cx, bool &ok) {
  uint32_t addr, e0, e1;
  asm("lop3.b32 %0, %1, 0x1FF8, %2, 0xEA;" : "=r"(addr) : "r"(sel8), "r"(cx.ztab));
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e0), "=r"(e1) : "r"(addr));
  ok = u < (e0 << 20);
  const double t = __dsub_rn(__hiloint2double((int)cx.hi52, (int)u), 4503599627370496.0);  // (double)u, exactly
  return __dmul_rn(t, __hiloint2double((int)e1, (int)e0));
}

// block sums of three values (see BLOCK-WISE ACCUMULATION above)
__device__ __forceinline__ void block_sums(double x0, double x1, double x2, double &s, double &q) {
  s = __dadd_rn(__dadd_rn(x0, x1), x2);
  q = __fma_rn(x2, x2, __fma_rn(x1, x1, __dmul_rn(x0, x0)));
}

// the warp resolves `cnt` (<= 32) queued blocks starting at entry `first`, one block per lane: the block is re-derived,
// its FIRST rejected candidate goes through smm_zig_slow in a single convergent call (one in 5000 blocks holds a second
// one: a divergent tail), and the difference of the block's fixed-point patterns is added to the shared accumulators
__device__ __noinline__ void zig_drain(const DevProblem &pb, const ZigCtx cx, int first, int cnt) {
  const int lane = threadIdx.x & 31;
  if (lane < cnt) {
    uint32_t j, kp;
    const uint32_t qa = cx.q + 8u * (uint32_t)(first + lane);
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(j), "=r"(kp) : "r"(qa) : "memory");
    // entry tag: row k (8 bits) | active draws (2 bits) | segment of the CTA's share (persistent kernel; 0 elsewhere)
    const uint32_t kk = kp & 0xFFu, nact = (kp >> 8) & 3u, seg = kp >> 10;
    uint32_t c2 = cx.c2;
    if (cx.segc2) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c2) : "r"(cx.segc2 + 4u * seg) : "memory");
    const smm_u32x4 r = smm_philox4x32_10(j, kk, c2, cx.c3, (uint32_t)pb.seed_sim, (uint32_t)(pb.seed_sim >> 32));
    const uint32_t D = (uint32_t)cx.D;
    double p;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(p) : "r"(cx.pvec + 8u * (seg * D + kk)) : "memory");
    bool ok0, ok1, ok2;
    const double zf0 = zig_fast_dev(r.x, r.w, cx, ok0);
    const double zf1 = zig_fast_dev(r.y, r.w >> 10, cx, ok1);
    const double zf2 = zig_fast_dev(r.z, __funnelshift_l(r.w, r.w, 12), cx, ok2);
    ok1 = ok1 || nact < 2;
    ok2 = ok2 || nact < 3;
    // the unsigned half of the shared table is the stream's table: the rare branch reads it through a generic pointer
    const smm_zigent *tab = (const smm_zigent *)__cvta_shared_to_generic((size_t)(cx.ztab & ~7u));
    const int t1 = !ok0 ? 0 : (!ok1 ? 1 : 2);
    const double z1 = smm_zig_slow(t1 == 0 ? r.x : (t1 == 1 ? r.y : r.z), smm_zig_select(r.w, t1), tab, smm_logtab());
    double zs0 = t1 == 0 ? z1 : zf0, zs1 = t1 == 1 ? z1 : zf1, zs2 = t1 == 2 ? z1 : zf2;
    if (t1 == 0 && !ok1) zs1 = smm_zig_slow(r.y, smm_zig_select(r.w, 1), tab, smm_logtab());
    if (t1 < 2 && !ok2) zs2 = smm_zig_slow(r.z, smm_zig_select(r.w, 2), tab, smm_logtab());
    const double xf0 = __dadd_rn(p, zf0), xs0 = __dadd_rn(p, zs0);
    const double xf1 = nact < 2 ? 0.0 : __dadd_rn(p, zf1), xs1 = nact < 2 ? 0.0 : __dadd_rn(p, zs1);
    const double xf2 = nact < 3 ? 0.0 : __dadd_rn(p, zf2), xs2 = nact < 3 ? 0.0 : __dadd_rn(p, zs2);
    double sf, qf, ss, qs;
    block_sums(xf0, xf1, xf2, sf, qf);
    block_sums(xs0, xs1, xs2, ss, qs);
    const unsigned long long dsum = (unsigned long long)__double_as_longlong(__dadd_rn(ss, pb.magic_sum)) -
                                    (unsigned long long)__double_as_longlong(__dadd_rn(sf, pb.magic_sum));
    const unsigned long long dsq = (unsigned long long)__double_as_longlong(__dadd_rn(qs, pb.magic_sq)) -
                                   (unsigned long long)__double_as_longlong(__dadd_rn(qf, pb.magic_sq));
    asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(cx.fix + 8u * (seg * 2u * D + kk)), "l"(dsum) : "memory");
    asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(cx.fix + 8u * (seg * 2u * D + D + kk)), "l"(dsq) : "memory");
  }
  __syncwarp();
}

// between steps: bring the queue back below 32 entries (full warps of work only)
__device__ __forceinline__ void zig_relieve(const DevProblem &pb, const ZigCtx &cx, int &qn) {
  while (qn >= 32) {
    __syncwarp();
    zig_drain(pb, cx, qn - 32, 32);
    qn -= 32;
  }
}
__device__ __forceinline__ void zig_flush(const DevProblem &pb, const ZigCtx &cx, int &qn) {
  zig_relieve(pb, cx, qn);
  if (qn > 0) {
    __syncwarp();
    zig_drain(pb, cx, 0, qn);
    qn = 0;
  }
}

// Philox4x32-10 of the simulator stream with the 20 round keys of seed_sim held in registers
struct SimKeys {
  uint32_t a[10], b[10];
};
__device__ __forceinline__ smm_u32x4 philox_keys(const SimKeys &ks, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)SMM_PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)SMM_PHILOX_M1 * c2;
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ ks.a[r];
    c1 = (uint32_t)p1;
    c2 = (uint32_t)(p0 >> 32) ^ c3 ^ ks.b[r];
    c3 = (uint32_t)p0;
  }
  smm_u32x4 out;
  out.x = c0;
  out.y = c1;
  out.z = c2;
  out.w = c3;
  return out;
}
__device__ __forceinline__ SimKeys sim_keys(const DevProblem &pb) {
  SimKeys ks;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    ks.a[r] = pb.rk_sim0[r];
    ks.b[r] = pb.rk_sim1[r];
  }
  return ks;
}
// the same keys parked in shared memory ([20] u32, 16-byte aligned) for the out-of-line hot loop
__device__ __forceinline__ void store_keys(const DevProblem &pb, uint32_t *dst) {
  if (threadIdx.x < 10) dst[threadIdx.x] = pb.rk_sim0[threadIdx.x];
  else if (threadIdx.x < 20) dst[threadIdx.x] = pb.rk_sim1[threadIdx.x - 10];
}
__device__ __forceinline__ SimKeys load_keys(uint32_t keys_s) {
  uint32_t w[20];
#pragma unroll
  for (int i = 0; i < 5; ++i)
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[4 * i]), "=r"(w[4 * i + 1]), "=r"(w[4 * i + 2]), "=r"(w[4 * i + 3])
                 : "r"(keys_s + 16u * i));
  SimKeys ks;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    ks.a[r] = w[r];
    ks.b[r] = w[10 + r];
  }
  return ks;
}

// One Philox block of row k: X = p + Z for its three draws, block-summed and added to the accumulators.  Called by all
// 32 lanes of a warp; nact = how many of this lane's draws count (kMasked = false: all three always do; 0 = the lane
// sits this step out).  Needs qn < 32 on entry; the caller resolves the queue between steps (kept out of here so that
// hot loops contain no call).  kpack = k | nact << 8 is what the queue stores next to j.
// kVariant != 0: ablations for the throughput micro-benchmark only (1 = deferred queue off, 2 = also no table lookup,
// 3 = Philox alone)
template <bool kMasked, int kVariant = 0>
__device__ __forceinline__ void add_block(const SimKeys &ks, double magic_sum, double magic_sq, const ZigCtx &cx, int &qn,
                                          Acc &a, double p, uint32_t j, uint32_t k, uint32_t kpack, int nact) {
  const smm_u32x4 r = philox_keys(ks, j, k, cx.c2, cx.c3);
  if (kVariant == 3) {
    a.sum += r.x ^ r.y;
    a.sq += r.z ^ r.w;
    return;
  }
  if (kVariant == 2) {
    const double w = __hiloint2double(0x3DF00000 | (int)(cx.ztab >> 28), 0);
    const double x0 = __dadd_rn(p, __dmul_rn(__dsub_rn(__hiloint2double(0x43300000, (int)r.x), 4503599627370496.0), w));
    const double x1 = __dadd_rn(p, __dmul_rn(__dsub_rn(__hiloint2double(0x43300000, (int)r.y), 4503599627370496.0), w));
    const double x2 = __dadd_rn(p, __dmul_rn(__dsub_rn(__hiloint2double(0x43300000, (int)r.z), 4503599627370496.0), w));
    double s, q;
    block_sums(x0, x1, x2, s, q);
    a.sum += (unsigned long long)__double_as_longlong(__dadd_rn(s, magic_sum)) + r.w;
    a.sq += (unsigned long long)__double_as_longlong(__dadd_rn(q, magic_sq));
    return;
  }
  bool ok0, ok1, ok2;
  double x0 = __dadd_rn(p, zig_fast_dev(r.x, r.w, cx, ok0));
  double x1 = __dadd_rn(p, zig_fast_dev(r.y, r.w >> 10, cx, ok1));
  double x2 = __dadd_rn(p, zig_fast_dev(r.z, __funnelshift_l(r.w, r.w, 12), cx, ok2));
  bool slow = !(ok0 && ok1 && ok2);
  if (kMasked) {
    if (nact < 3) {
      x2 = 0.0;
      ok2 = true;
    }
    if (nact < 2) {
      x1 = 0.0;
      ok1 = true;
    }
    slow = nact > 0 && !(ok0 && ok1 && ok2);
  }
  double s, q;
  block_sums(x0, x1, x2, s, q);
  const unsigned long long sb = (unsigned long long)__double_as_longlong(__dadd_rn(s, magic_sum));
  const unsigned long long qb = (unsigned long long)__double_as_longlong(__dadd_rn(q, magic_sq));
  if (!kMasked || nact > 0) {
    a.sum += sb;
    a.sq += qb;
  }
  if (kVariant == 1) {
    a.sum += slow;
    return;
  }
  if (__any_sync(0xffffffffu, slow)) {  // half of the steps: some lane's block holds a candidate that left the fast path
    const unsigned m = __ballot_sync(0xffffffffu, slow);
    if (slow) {
      const uint32_t qa = cx.q + 8u * (uint32_t)(qn + __popc(m & cx.lt));
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(qa), "r"(j), "r"(kpack) : "memory");
    }
    qn += __popc(m);
  }
}

// THE HOT LOOP, out of line so that it is register-allocated on its own (the persistent kernel around it carries far
// too much state): up to n_steps warp steps of row k, blocks j, j + dj, ...; it stops early when 32 deferred blocks have
// gathered (the caller drains them and comes back).  The round keys travel through shared memory into 20 registers.
struct SimRet {
  unsigned long long sum, sq;
  int done, qn;
};
template <int kVariant = 0>
__device__ __noinline__ SimRet sim_steps_full(uint32_t keys_s, uint32_t ztab, uint32_t q_s, uint32_t c2, uint32_t c3, double p,
                                              double magic_sum, double magic_sq, uint32_t j, uint32_t dj, int n_steps,
                                              uint32_t kpack3, int qn, unsigned long long sum, unsigned long long sq) {
  const SimKeys ks = load_keys(keys_s);
  ZigCtx cx;
  cx.ztab = ztab;
  cx.hi52 = 0x43300000u | (ztab >> 31);  // opaque (shared addresses are small): stays in a register
  cx.q = q_s;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(cx.lt));
  cx.c2 = c2;
  cx.c3 = c3;
  cx.pvec = cx.fix = 0;
  cx.D = 0;
  Acc a{sum, sq};
  const uint32_t k = kpack3 & 0xFFu;
  int q = 0;
#pragma unroll 1
  for (; q < n_steps && qn < 32; ++q, j += dj) add_block<false, kVariant>(ks, magic_sum, magic_sq, cx, qn, a, p, j, k, kpack3, 3);
  return SimRet{a.sum, a.sq, q, qn};
}
// the same for steps in which lanes or draws are masked out: lane active iff lane_on and j < jlimit, with nact_on draws
__device__ __noinline__ SimRet sim_steps_masked(uint32_t keys_s, uint32_t ztab, uint32_t q_s, uint32_t c2, uint32_t c3, double p,
                                                double magic_sum, double magic_sq, uint32_t j, uint32_t dj, int n_steps,
                                                uint32_t k, int qn, unsigned long long sum, unsigned long long sq,
                                                bool lane_on, uint32_t jlimit, int nact_on) {
  const SimKeys ks = load_keys(keys_s);
  ZigCtx cx;
  cx.ztab = ztab;
  cx.hi52 = 0x43300000u | (ztab >> 31);
  cx.q = q_s;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(cx.lt));
  cx.c2 = c2;
  cx.c3 = c3;
  cx.pvec = cx.fix = 0;
  cx.D = 0;
  Acc a{sum, sq};
  int q = 0;
#pragma unroll 1
  for (; q < n_steps && qn < 32; ++q, j += dj) {
    const int nact = (lane_on && j < jlimit) ? nact_on : 0;
    add_block<true>(ks, magic_sum, magic_sq, cx, qn, a, p, j, k & 0xFFu, k | ((uint32_t)nact << 8), nact);  // k: row | segment << 10
  }
  return SimRet{a.sum, a.sq, q, qn};
}

// Static mapping, any D <= g.n: thread t owns row k = t % D and blocks j0 + t/D, +lanes, ...
// zt: the 16 KB buffer holding the signed layer table (shared); red: shared [2 * g.n] u64; zfix: shared [2 * D] u64 (zeroed
// here); zq: shared [g.n / 32][kZigQWords]; pp: shared.
// Writes the group's partial sums [2][D] (u64 patterns) to `part`.
__device__ void simulate_static(const DevProblem &pb, const Grp &g, const unsigned long long *zt, const double *pp, int j0,
                                int j1, uint32_t uid, uint32_t rep, unsigned long long *red, unsigned long long *zfix,
                                uint32_t *zq, double *part) {
  const int D = pb.P, S = pb.S, tid = g.tid;
  const int lanes = g.n / D;
  const int n_full = S / 3;      // blocks whose three normals are all used
  const int n_tail = S - 3 * n_full;  // draws of the last, partial block (0 = there is none)
  for (int e = tid; e < 2 * D; e += g.n) zfix[e] = 0ull;
  gsync(g);
  Acc a{0ull, 0ull};
  ZigCtx cx = zig_ctx(zt, zq + (size_t)(tid >> 5) * kZigQWords, S);
  cx.pvec = smem_addr(pp);
  cx.fix = smem_addr(zfix);
  cx.D = D;
  cx.c2 = pb.noseed ? uid : 0u;
  cx.c3 = (SMM_STREAM_SIM << 28) | (pb.noseed ? (rep & SMM_ITER_MASK) : 0u);
  int qn = 0;
  const SimKeys ks = sim_keys(pb);
  const bool on = tid < lanes * D;
  const int k = tid % D, ln = tid / D;
  const double p = pp[k];
  const int jend = j1 < n_full ? j1 : n_full;
  // every warp runs the same number of steps (the ballots inside add_block need all 32 lanes)
  const int n_steps = jend > j0 ? (jend - j0 + lanes - 1) / lanes : 0;
  for (int t = 0; t < n_steps; ++t) {
    const int j = j0 + ln + t * lanes;
    const int nact = (on && j < jend) ? 3 : 0;
    add_block<true>(ks, pb.magic_sum, pb.magic_sq, cx, qn, a, p, (uint32_t)j, (uint32_t)k, (uint32_t)k | ((uint32_t)nact << 8), nact);
    if (qn >= 32) zig_relieve(pb, cx, qn);
  }
  if (n_tail && j0 <= n_full && n_full < j1) {  // S not a multiple of 3: the last block contributes one or two draws
    const int nact = (on && ln == 0) ? n_tail : 0;
    add_block<true>(ks, pb.magic_sum, pb.magic_sq, cx, qn, a, p, (uint32_t)n_full, (uint32_t)k, (uint32_t)k | ((uint32_t)nact << 8), nact);
  }
  zig_flush(pb, cx, qn);
  red[2 * tid] = a.sum;
  red[2 * tid + 1] = a.sq;
  gsync(g);
  if (tid < 2 * D) {
    const int kk = tid % D, which = tid / D;
    unsigned long long acc = zfix[which * D + kk];
    for (int l = 0; l < lanes; ++l) acc += red[2 * (l * D + kk) + which];
    ((unsigned long long *)part)[which * D + kk] = acc;
  }
  gsync(g);
}

constexpr int kUnitSteps = 1;   // warp steps per work unit of the persistent kernel
constexpr int kMaxGrab = 8;     // units a warp takes from its CTA's queue at once (guided: fewer towards the end)
constexpr int kMinGrab = 2;     // ... and at least (a call of the out-of-line hot loop costs about half a step)
constexpr int kStaticNum = 7, kStaticDen = 8;  // share of a CTA's units that is split statically over its warps
constexpr int kTputSteps = 8;   // sim_throughput_kernel: steps per queue access

__device__ void group_distance(const DevProblem &pb, const Grp &g, const FinScratch &fs);

// Sum the n_seg partials of an evaluation (integer adds: exact), then moments + weighted distance.
// Result in fs.mom, fs.value[0], fs.flags[1] (status).
__device__ void group_finalize(const DevProblem &pb, const Grp &g, const FinScratch &fs, const double *part_base,
                               int n_seg, int part_len) {
  const int D = pb.P, tid = g.tid;
  if (pb.obj == SMM_OBJ_FAILS) {
    // the objective throws -> caught by evaluateObjective: status -2, value stays -1.0, no moments
    for (int k = tid; k < pb.M; k += g.n) fs.mom[k] = __longlong_as_double(0x7ff8000000000000ll);
    if (tid == 0) {
      fs.value[0] = -1.0;
      fs.flags[1] = -2;
    }
    gsync(g);
    return;
  }
  for (int e = tid; e < 2 * D; e += g.n) {
    unsigned long long acc = 0ull;
    const unsigned long long *pu = (const unsigned long long *)part_base;
    for (int s = 0; s < n_seg; ++s) acc += __ldcg(pu + (size_t)s * part_len + e);
    // remove one copy of bits(M) per block (mod 2^64: exact), then fixed point -> double
    const bool is_sq = e >= D;
    const unsigned long long mb = (unsigned long long)__double_as_longlong(is_sq ? pb.magic_sq : pb.magic_sum);
    const long long fixed = (long long)(acc - (unsigned long long)zig_blocks(pb.S) * mb);
    fs.tot[e] = __dmul_rn((double)fixed, is_sq ? pb.scale_sq : pb.scale_sum);
  }
  gsync(g);
  const double S = (double)pb.S;
  for (int k = tid; k < D; k += g.n) {
    const double mean = __ddiv_rn(fs.tot[k], S);
    fs.mom[k] = mean;
    if (pb.obj == SMM_OBJ_NORM_MV) {
      // sum (x - mean)^2 = sum x^2 - mean * sum x
      const double ss = __dsub_rn(fs.tot[D + k], __dmul_rn(mean, fs.tot[k]));
      fs.mom[D + k] = __ddiv_rn(ss, S - 1.0);
    }
  }
  gsync(g);
  group_distance(pb, g, fs);
}

// value = mean_k ((sim_k - data_k) / w_k)^2 (ObjExamples.jl:90-101) of the moments in fs.mom -> fs.value[0],
// fs.flags[1] = 1.  Divisions in parallel, summed in moment order.
__device__ void group_distance(const DevProblem &pb, const Grp &g, const FinScratch &fs) {
  const int tid = g.tid;
  if (tid < 32) {
    double acc = 0.0;
    for (int k0 = 0; k0 < pb.M; k0 += 32) {
      const int k = k0 + tid;
      double d2 = 0.0;
      if (k < pb.M) {
        const double d = __ddiv_rn(__dsub_rn(fs.mom[k], pb.data[k]), pb.w[k]);
        d2 = __dmul_rn(d, d);
      }
      const int n = pb.M - k0 < 32 ? pb.M - k0 : 32;
      for (int l = 0; l < n; ++l) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, d2, l));
    }
    if (tid == 0) {
      fs.value[0] = __ddiv_rn(acc, (double)pb.M);
      fs.flags[1] = 1;
    }
  }
  gsync(g);
}

__device__ __forceinline__ void slow_spin(double seconds) {
  // objfunc_norm_slow: sleep(0.1) (ObjExamples.jl:130)
  const unsigned long long t0 = gtimer();
  const unsigned long long ns = (unsigned long long)(seconds * 1e9);
  while (gtimer() - t0 < ns) __nanosleep(20000);
}

// ------------------------------------------------------------------------------------------------
// doAcceptReject! (AlgoBGP.jl:324-392) + set_eval! (:220-245) for chain c with the evaluation in
// fs.value / fs.flags[1] / fs.mom / fs.pp.  Writes the trace slot, the last-accepted record (la_cur),
// the published record (la_pub) and -- fused multi-GPU mode -- the record and its value into every
// rank's gather buffer (peer stores over NVLink).
// ------------------------------------------------------------------------------------------------
// flow (barrier-free persistent kernel): wait_applied != 0 -> first wait until the owner CTA has applied the exchange
// of iteration iter-1 to this chain (st.applied[c] >= iter-1).  Completion is published per CTA, not per chain: see
// publish_completions.
// The completion counter of a rank (exchange_mode 2): ONE monotone 64-bit count of finished chain evaluations, over all
// ranks, since the handle was created; every CTA adds the chains it finished to the counter of every rank.
__device__ __forceinline__ unsigned long long *done_counter(const DevProblem &pb, double *val_all) {
  return (unsigned long long *)(val_all + 2 * (size_t)pb.N);
}
// The chain state doAcceptReject!/set_eval! read, fetched ahead of time by a 32-lane group (persistent kernel: before
// the arrival atomic of the partial sums, so that the loads travel together with it instead of after it).
constexpr int kPreRec = (3 + SMM_MAX_PARAMS + SMM_MAX_MOMENTS + 31) / 32;
struct AcceptPre {
  double old_value, sigma0, curr_prev, best_prev, u_acc;  // lane 0
  int n_noex0, n_acc0, bestid_prev;                       // lane 0
  double la[kPreRec];                                     // record entries lane, lane + 32, ... of la_cur[c]
};
__device__ __forceinline__ void wait_exchange_applied(const DevState &st, const Grp &g, int c, int iter) {
  if (g.tid == 0) {
    const unsigned long long t0 = gtimer();
    unsigned spins = 0;
    while (ld_acquire_gpu(st.applied + c) < (unsigned)(iter - 1)) {
      if ((++spins & 255u) == 0 && ((ld_relaxed_gpu(&st.bar->gen) & 0x80000000u) || gtimer() - t0 > kSpinTimeoutNs)) {
        atomicOr(st.err, kErrTimeout);
        break;
      }
    }
  }
  gsync(g);
}
__device__ __forceinline__ void accept_prefetch(const DevProblem &pb, const DevState &st, int lane, int c, int gc, int iter,
                                                AcceptPre &pre) {
  const int R = rec_len(pb.P, pb.M), L = pb.L;
  const double *la = st.la_cur + (size_t)c * R;
  const size_t slot = (size_t)(iter - 1) * L + c;
#pragma unroll
  for (int q = 0; q < kPreRec; ++q) pre.la[q] = lane + 32 * q < R ? __ldcg(la + lane + 32 * q) : 0.0;
  if (lane == 0) {
    pre.old_value = pre.la[0];
    pre.n_noex0 = __ldcg(st.n_noex + c);
    pre.n_acc0 = __ldcg(st.n_acc + c);
    pre.sigma0 = __ldcg(st.sigma + c);
    const size_t prev_slot = iter > 1 ? slot - L : slot;
    pre.curr_prev = __ldcg(st.t_curr + prev_slot);
    pre.best_prev = __ldcg(st.t_best + prev_slot);
    pre.bestid_prev = __ldcg(st.t_bestid + prev_slot);
  }
}

__device__ void group_accept_store(const DevProblem &pb, const DevState &st, const Grp &g, const FinScratch &fs, int c,
                                   int gc, int iter, bool fused, bool flow = false, bool wait_applied = false,
                                   const AcceptPre *pre = nullptr) {
#ifdef SMM_LL_TU
  const bool ll = flow && st.ll != nullptr && pre != nullptr;  // exchange_mode 3: one warp, chain state prefetched
#endif
  const int tid = g.tid;
  const int P = pb.P, M = pb.M, R = rec_len(P, M), L = pb.L;
  if (flow && wait_applied && !pre) wait_exchange_applied(st, g, c, iter);
  double *la = st.la_cur + (size_t)c * R;
  double *pub = st.la_pub + (size_t)c * R;
  const size_t slot = (size_t)(iter - 1) * L + c;
  if (tid == 0) {
    // everything this thread will need from global memory, requested at once (one L2 round trip) -- or already here
    const double old_value = pre ? pre->old_value : __ldcg(la);
    const int n_noex0 = pre ? pre->n_noex0 : __ldcg(st.n_noex + c), n_acc0 = pre ? pre->n_acc0 : __ldcg(st.n_acc + c);
    const double sigma0 = pre ? pre->sigma0 : __ldcg(st.sigma + c);
    const size_t prev_slot = iter > 1 ? slot - L : slot;
    const double curr_prev = pre ? pre->curr_prev : __ldcg(st.t_curr + prev_slot);
    const double best_prev = pre ? pre->best_prev : __ldcg(st.t_best + prev_slot);
    const int bestid_prev = pre ? pre->bestid_prev : __ldcg(st.t_bestid + prev_slot);
    const double value = fs.value[0];
    double prob;
    int accepted, status = fs.flags[1];
    if (iter == 1) {
      prob = 1.0;
      accepted = 1;
      status = 1;
    } else {
      if (status < 0) {
        prob = 0.0;
        accepted = 0;
      } else {
        if (!(value >= 0.0)) atomicOr(st.err, kErrNegative);  // `error(...)` upstream (:341)
        const double e = exp(__dmul_rn(pb.acc_tuner[gc], __dsub_rn(old_value, value)));
        prob = isnan(e) ? e : (e < 1.0 ? e : 1.0);  // minimum([1.0, e]) propagates NaN
        if (!isfinite(prob)) {
          prob = 0.0;
          accepted = 0;
          status = -1;
        } else if (!isfinite(old_value)) {
          prob = 1.0;
          accepted = 1;
        } else {
          status = 1;
          accepted = prob > (pre ? pre->u_acc : smm_acc_uniform(pb.seed_algo, (uint32_t)gc, (uint32_t)iter));
        }
      }
    }
    // set_acceptRate! (:253-257): this iteration has exchanged == 0 at this point
    const int n_noex = n_noex0 + 1, n_acc = n_acc0 + accepted;
    st.n_noex[c] = n_noex;
    st.n_acc[c] = n_acc;
    const double rate = __ddiv_rn((double)n_acc, (double)n_noex);
    st.accept_rate[c] = rate;
#ifdef SMM_LL_TU
    double sigma_cur = sigma0;
    if (iter > 1 && iter % pb.sigma_update_steps == 0) {
      sigma_cur = rate > 0.234 ? __dmul_rn(sigma0, __dadd_rn(1.0, pb.sigma_adjust_by))
                               : __dmul_rn(sigma0, __dsub_rn(1.0, pb.sigma_adjust_by));
      st.sigma[c] = sigma_cur;
    }
    if (ll) ((double *)fs.flags)[1] = sigma_cur;  // (the spare double behind the two flags of the persistent kernel's scratch)
#else
    if (iter > 1 && iter % pb.sigma_update_steps == 0) {
      st.sigma[c] = rate > 0.234 ? __dmul_rn(sigma0, __dadd_rn(1.0, pb.sigma_adjust_by))
                                 : __dmul_rn(sigma0, __dsub_rn(1.0, pb.sigma_adjust_by));
    }
#endif
    // set_eval!
    double curr, best;
    int best_id;
    if (iter == 1) {
      curr = value;
      best = value;
      best_id = 1;
    } else {
      curr = accepted ? value : curr_prev;
      if (value < best_prev) {
        best = value;
        best_id = iter;
      } else {
        best = best_prev;
        best_id = bestid_prev;
      }
    }
    st.t_value[slot] = value;
    st.t_prob[slot] = prob;
    st.t_curr[slot] = curr;
    st.t_best[slot] = best;
    st.t_acc[slot] = (uint8_t)accepted;
    st.t_status[slot] = status;
    st.t_exch[slot] = 0;
    st.t_bestid[slot] = best_id;
    if (accepted) atomicAdd(&st.counters[0], 1ull);
    fs.flags[0] = accepted;
    fs.flags[1] = status;
    fs.value[1] = prob;
  }
  gsync(g);
  const int acc = fs.flags[0];
  // trace rows + last-accepted record (coalesced over threads)
  for (int k = tid; k < P; k += g.n) st.t_params[slot * P + k] = fs.pp[k];
  for (int k = tid; k < M; k += g.n) st.t_mom[slot * M + k] = fs.mom[k];
  const int par = fused ? (iter & 1) : 0;
#ifdef SMM_LL_TU
  if (ll) {
    // exchange_mode 3.  What the next iteration's critical path needs -- value (exchange), sigma and parameters
    // (proposal centre) -- goes to every rank as flag-in-data words: no system fence, no counter round trip.  The
    // chain's local state written above is ordered before them for this GPU's readers by a gpu-scope fence (the warp
    // barrier makes it cumulative over the lanes); the full records for the owners' swap_ev_ij! follow and are covered
    // by the completion counter, which nobody on the critical path waits for.
    auto published = [&](int k, double old) -> double {  // entry k of the chain's last accepted record
      if (!acc) return old;
      return k == 0 ? fs.value[0] : k == 1 ? fs.value[1] : k == 2 ? (double)fs.flags[1] : k < 3 + P ? fs.pp[k - 3] : fs.mom[k - 3 - P];
    };
#pragma unroll
    for (int j = 0; j < kPreRec; ++j) {  // entry tid + 32 j of la_cur[c] sits in pre->la[j] (static indices only)
      const int k = tid + 32 * j;
      if (k < R) {
        const double v = published(k, pre->la[j]);
        if (acc) la[k] = v;
        pub[k] = v;
      }
    }
    gsync(g);
    const int W = ll_words(P);
    const double sigma_cur = ((const double *)fs.flags)[1];
    for (int q0 = 0; q0 < P + 2; q0 += 32) {
      const int q = q0 + tid;                                                // double q of the LL record: value, sigma, params
      const int kq = q == 0 ? 0 : (q >= 2 && q < P + 2 ? 3 + (q - 2) : -1);  // its entry of the record
      double vold = 0.0;
#pragma unroll
      for (int j = 0; j < kPreRec; ++j) {
        if (32 * j < 3 + P) {  // (warp uniform: value and parameters sit in the first 3 + P entries of the record)
          const double t = __shfl_sync(0xffffffffu, pre->la[j], (kq < 0 ? 0 : kq) & 31);
          if (kq >= 0 && (kq >> 5) == j) vold = t;
        }
      }
      if (q < P + 2) {
        const double v = q == 1 ? sigma_cur : (acc ? (kq == 0 ? fs.value[0] : fs.pp[kq - 3]) : vold);
        fence_acq_rel_gpu();
        for (int r = 0; r < pb.world; ++r)
          ll_store(st.peer_ll[r] + ((size_t)par * pb.N + gc) * W + 2 * q, v, (uint32_t)iter);
      }
    }
#pragma unroll
    for (int j = 0; j < kPreRec; ++j) {  // the all-gather of the full records: one coalesced row per peer
      const int k = tid + 32 * j;
      if (k < R) {
        const double v = published(k, pre->la[j]);
        for (int r = 0; r < pb.world; ++r) st.peer_la_all[r][((size_t)par * pb.N + gc) * R + k] = v;
        if (k == 0)
          for (int r = 0; r < pb.world; ++r) st.peer_val_all[r][(size_t)par * pb.N + gc] = v;
      }
    }
    gsync(g);
    return;
  }
#endif
  for (int k = tid; k < R; k += g.n) {
    double v;
    if (acc) {
      v = k == 0 ? fs.value[0] : k == 1 ? fs.value[1] : k == 2 ? (double)fs.flags[1] : k < 3 + P ? fs.pp[k - 3] : fs.mom[k - 3 - P];
      la[k] = v;
    } else {
      v = pre ? pre->la[(k - tid) >> 5] : __ldcg(la + k);
    }
    pub[k] = v;
    if (fused) {
      // the all-gather: one coalesced row per peer, straight into its gather buffer
      for (int r = 0; r < pb.world; ++r) st.peer_la_all[r][((size_t)par * pb.N + gc) * R + k] = v;
      if (k == 0)
        for (int r = 0; r < pb.world; ++r) st.peer_val_all[r][(size_t)par * pb.N + gc] = v;
    } else if (k == 0) {
      st.val_all[gc] = v;  // compact copy of the values for the exchange step
    }
  }
  gsync(g);
}

// ------------------------------------------------------------------------------------------------
// exchangeMoves! (AlgoBGP.jl:647-691): the sequential pair loop, run level-parallel over the schedule of
// iteration `iter` (pairs inside a level share no chain).  val/own/exch live in shared memory.
// ------------------------------------------------------------------------------------------------
__device__ unsigned exchange_levels(const DevProblem &pb, const DevState &st, const Grp &g, int sched_idx, int n_s,
                                    double *val, unsigned short *own, unsigned short *exch) {
  const int *ij = st.sched_ij + (size_t)sched_idx * n_s * 2;
  const int *off = st.sched_off + (size_t)sched_idx * (n_s + 1);
  const int nlev = st.sched_nlev[sched_idx];
  unsigned n_swaps = 0;
  for (int l = 0; l < nlev; ++l) {
    const int lo = off[l], hi = off[l + 1];
    for (int t = lo + g.tid; t < hi; t += g.n) {
      const int i = ij[2 * t], j = ij[2 * t + 1];
      const double vi = val[i], vj = val[j];
      if (__dsub_rn(vi, vj) > pb.min_improve[i]) {  // dist_fun(evi.value, evj.value) > min_improve[i]
        val[i] = vj;
        val[j] = vi;
        const unsigned short oi = own[i];
        own[i] = own[j];
        own[j] = oi;
        exch[i] = (unsigned short)(j + 1);
        exch[j] = (unsigned short)(i + 1);
        ++n_swaps;
      }
    }
    gsync(g);
  }
  return n_swaps;
}

// swap_ev_ij! (:734-749) for one chain this rank owns: set_eval!(ci, ej) + set_exchanged!(ci, j).
// One warp; `src` is the record that ended up on this chain.
__device__ void exchange_apply_chain(const DevProblem &pb, const DevState &st, int iter, int c, int partner,
                                     const double *src, int lane) {
  const int P = pb.P, M = pb.M, R = rec_len(P, M), L = pb.L;
  double *la = st.la_cur + (size_t)c * R;
  const size_t slot = (size_t)(iter - 1) * L + c;
  for (int k = lane; k < R; k += 32) {
    const double v = __ldcg(src + k);
    la[k] = v;
    if (k >= 3 && k < 3 + P) st.t_params[slot * P + (k - 3)] = v;
    if (k >= 3 + P) st.t_mom[slot * M + (k - 3 - P)] = v;
  }
  if (lane == 0) {
    const double value = __ldcg(src), prob = __ldcg(src + 1), status = __ldcg(src + 2);
    const int n_noex0 = __ldcg(st.n_noex + c), n_acc0 = __ldcg(st.n_acc + c);
    const int acc0 = (int)__ldcg(st.t_acc + slot);
    const double bprev = __ldcg(st.t_best + slot - L);
    const int bidprev = __ldcg(st.t_bestid + slot - L);
    // this iteration no longer counts towards the acceptance rate (exchanged != 0)
    st.n_noex[c] = n_noex0 - 1;
    st.n_acc[c] = n_acc0 - acc0;
    st.t_value[slot] = value;
    st.t_prob[slot] = prob;
    st.t_status[slot] = (int)status;
    st.t_acc[slot] = 1;  // the swapped-in eval is an accepted one
    st.t_curr[slot] = value;
    if (value < bprev) {
      st.t_best[slot] = value;
      st.t_bestid[slot] = iter;
    } else {
      st.t_best[slot] = bprev;
      st.t_bestid[slot] = bidprev;
    }
    st.t_exch[slot] = partner;
  }
}

// ------------------------------------------------------------------------------------------------
// (A) multi-launch mode
// ------------------------------------------------------------------------------------------------
struct EvalSmem {
  double pp[SMM_MAX_PARAMS];
  double mu01[SMM_MAX_PARAMS];
  double cand[2 * kEvalThreads];
  unsigned long long red[2 * kEvalThreads];
  double tot[2 * SMM_MAX_PARAMS];
  double mom[SMM_MAX_MOMENTS];
  smm_logent logtab[1 << SMM_LOG_BITS];
  unsigned long long zfix[2 * SMM_MAX_PARAMS];
  uint32_t zq[(kEvalThreads / 32) * kZigQWords];
  unsigned char okf[2 * kEvalThreads];
  int first[SMM_MAX_PARAMS];
  int resolved[SMM_MAX_PARAMS];
  double value[2];
  int flags[2];
  int is_last;
};

__device__ __forceinline__ PropScratch prop_scratch(EvalSmem &sm) {
  return PropScratch{sm.pp, sm.mu01, sm.cand, sm.okf, sm.first, sm.resolved, sm.logtab};
}
__device__ __forceinline__ FinScratch fin_scratch(EvalSmem &sm) {
  return FinScratch{sm.pp, sm.tot, sm.mom, sm.value, sm.flags};
}

__global__ void __launch_bounds__(kEvalThreads) bgp_eval_kernel(DevProblem pb, DevState st, int iter, int n_split,
                                                                int part_len) {
  __shared__ EvalSmem sm;
  __shared__ unsigned long long s_zigtab[kZigBufEntries];  // 16 KB: the 8 KB-aligned half holds the table (see load_zigtab)
  const int c = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  const int gc = global_chain(pb, c);
  const Grp g{tid, (int)blockDim.x, 0};
  const size_t stamp = (size_t)c * n_split + split;
  PHASE_STAMP(stamp, 0);
  load_logtab(sm.logtab);
  load_zigtab(s_zigtab);
  __syncthreads();
  group_proposal(pb, st, g, prop_scratch(sm), c, gc, iter, split == 0);
  PHASE_STAMP(stamp, 1);
  double *part_base = st.partials + (size_t)c * n_split * part_len;
  if (pb.obj == SMM_OBJ_FAILS) {
    if (split != 0) return;
  } else {
    if (pb.obj == SMM_OBJ_NORM_SLOW) slow_spin(pb.slow_seconds);
    const int nb = zig_blocks(pb.S);
    const int j0 = (int)(((long long)nb * split) / n_split), j1 = (int)(((long long)nb * (split + 1)) / n_split);
    simulate_static(pb, g, s_zigtab, sm.pp, j0, j1, (uint32_t)gc, (uint32_t)iter, sm.red, sm.zfix, sm.zq,
                    part_base + (size_t)split * part_len);
    PHASE_STAMP(stamp, 2);
    if (n_split > 1) {
      if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(st.arrive + c, 1u);
        sm.is_last = (prev == (unsigned)(n_split - 1));
        if (sm.is_last) st.arrive[c] = 0u;  // re-arm for the next iteration
        __threadfence();
      }
      __syncthreads();
      if (!sm.is_last) return;
    }
  }
  group_finalize(pb, g, fin_scratch(sm), part_base, n_split, part_len);
  group_accept_store(pb, st, g, fin_scratch(sm), c, gc, iter, false);
  PHASE_STAMP(stamp, 3);
}

// dynamic smem: val[N] (double) own[N] exch[N] (u16)
__global__ void __launch_bounds__(kExchThreads) bgp_exchange_kernel(DevProblem pb, DevState st, int iter,
                                                                    int sched_idx, int n_s) {
  extern __shared__ double smem_d[];
  const int N = pb.N, tid = threadIdx.x, nthr = blockDim.x;
  const int R = rec_len(pb.P, pb.M), L = pb.L;
  const Grp g{tid, nthr, 0};
  double *val = smem_d;
  unsigned short *own = (unsigned short *)(val + N), *exch = own + N;
  for (int i = tid; i < N; i += nthr) {
    val[i] = st.la_all[(size_t)gather_slot(pb, i) * R];  // (with one rank la_all is la_pub and the slot is i)
    own[i] = (unsigned short)i;
    exch[i] = 0;
  }
  __syncthreads();
  const unsigned n_swaps = exchange_levels(pb, st, g, sched_idx, n_s, val, own, exch);
  if (pb.rank == 0 && n_swaps) atomicAdd(&st.counters[1], (unsigned long long)n_swaps);
  for (int c = tid / 32; c < L; c += nthr / 32) {  // one warp per chain
    const int gc = global_chain(pb, c);
    const int partner = exch[gc];
    if (partner == 0) continue;
    exchange_apply_chain(pb, st, iter, c, partner, st.la_all + (size_t)gather_slot(pb, own[gc]) * R, tid & 31);
  }
}

// ------------------------------------------------------------------------------------------------
// (B) persistent mode
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) { return *(const volatile unsigned *)p; }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  return *(const volatile unsigned long long *)p;
}


// Grid barrier (CTA 0 is the master).  Arrivals are release-adds (each CTA's writes become visible before its
// arrival); the master waits for all of them, fences once (acquire for its own reads + release for the store
// that follows) and publishes the new generation; the others spin on it with relaxed loads.  No fence on the
// waiting side: everything a CTA reads after a barrier that another CTA wrote is fetched with ld.global.cg
// (L2, never a stale L1 line) by instructions issued after the spin loop has seen the new generation.
// Measured 2.0 us per barrier on 148 x 1024 threads (tools/barrier_bench.py) against 3.1 us with fences on
// both sides.  The generation counter runs in the low 31 bits; bit 31 tells everybody to abort.
// With `cross`, the master also exchanges sequence flags with every peer GPU (system scope) before
// releasing, so the records stored into our gather buffer by the peers are complete.
__device__ bool grid_barrier(const DevProblem &pb, const DevState &st, unsigned &gen, bool cross,
                             unsigned long long &seq) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    bool ok = true;
    const unsigned G = gridDim.x;
    gen = (gen + 1u) & 0x7fffffffu;
    const unsigned target = gen;
    if (cross) {
      ++seq;
      __threadfence_system();  // this CTA's peer stores (all threads, ordered by the bar.sync above)
    }
    if (blockIdx.x == 0) {
      const unsigned long long t0 = gtimer();
      unsigned spins = 0;
      while (ld_relaxed_gpu(&st.bar->arrive) < G - 1) {
        if ((++spins & 4095u) == 0 && gtimer() - t0 > kSpinTimeoutNs) {
          ok = false;
          break;
        }
      }
      fence_acq_rel_gpu();
      st_relaxed_gpu(&st.bar->arrive, 0u);
      if (cross && ok) {
        for (int r = 0; r < pb.world; ++r) st_release_sys_u64(st.peer_flags[r] + pb.rank, seq);
        for (int r = 0; r < pb.world && ok; ++r) {
          spins = 0;
          while (ld_acquire_sys_u64(st.flags + r) < seq) {
            if ((++spins & 1023u) == 0 && gtimer() - t0 > kSpinTimeoutNs) {
              ok = false;
              break;
            }
          }
        }
        fence_acq_rel_gpu();
      }
      if (!ok) atomicOr(st.err, kErrTimeout);
      st_relaxed_gpu(&st.bar->gen, ok ? target : (target | 0x80000000u));
    } else {
      red_release_gpu(&st.bar->arrive, 1u);
      const unsigned long long t0 = gtimer();
      unsigned spins = 0, g;
      for (;;) {
        g = ld_relaxed_gpu(&st.bar->gen);
        if ((g & 0x80000000u) || (int)(((g & 0x7fffffffu) - target) << 1) >= 0) break;
        if ((++spins & 4095u) == 0 && gtimer() - t0 > 2 * kSpinTimeoutNs) {
          atomicOr(st.err, kErrTimeout);
          g = 0x80000000u;
          break;
        }
      }
      if (g & 0x80000000u) ok = false;
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// block that holds flattened index x when T items are split evenly over G blocks: [floor(b*T/G), floor((b+1)*T/G))
__device__ __forceinline__ long long block_of(long long x, long long T, long long G) {
  return ((x + 1) * G + T - 1) / T - 1;
}

constexpr int kPGroups = 8;                      // proposal groups per CTA (128 threads each at most)
constexpr int kExchWarps = 8;                    // warps walking the exchange levels when N > 512
constexpr int kExchBarrier = 10;                 // their named barrier (0 = CTA, 1..kPGroups = proposal groups)
constexpr int kPropCand = 2 * kPersistThreads;   // candidate slots shared by the groups
constexpr int kMaxCtaSeg = 32;                   // chains (segments) one CTA may touch per iteration

struct PersistSmem {
  smm_logent logtab[1 << SMM_LOG_BITS];
  __align__(16) uint32_t keys[20];  // Philox round keys of seed_sim for the out-of-line hot loop
  // segment geometry of this CTA's share (iteration invariant)
  int n_seg, total_units;
  int seg_c[kMaxCtaSeg], seg_j0[kMaxCtaSeg], seg_j1[kMaxCtaSeg], seg_unit0[kMaxCtaSeg + 1];
  int seg_slot[kMaxCtaSeg], seg_nseg[kMaxCtaSeg];   // partial slot of this CTA / CTAs sharing the chain
  uint32_t seg_c2[kMaxCtaSeg];                       // counter word 2 of the segment's blocks (chain id with noseed, else 0)
  // per-iteration work queue
  int next_unit;
  unsigned n_finished;  // chains this CTA finished in the current iteration
  int nlev;  // levels of the prefetched exchange schedule
  // proposals
  double g_pp[kPGroups][SMM_MAX_PARAMS];
  double g_mu01[kPGroups][SMM_MAX_PARAMS];
  int g_first[kPGroups][SMM_MAX_PARAMS];
  int g_resolved[kPGroups][SMM_MAX_PARAMS];
#ifdef SMM_LL_TU
  double g_cen[kPGroups][SMM_MAX_PARAMS + 1];  // exchange_mode 3: centre parameters + sigma of the group's proposal
  unsigned n_finished_prev;                    // exchange_mode 3: completions of the previous iteration, published late
#endif
};

// Level schedule of exchange `pit` -> shared memory (called at the start of phase A of iteration pit, far
// from the critical path).  dynamic smem: sij[n_s] u32 (i | j << 16) | soff[n_s + 1] i32
__device__ __forceinline__ void prefetch_schedule(const DevState &st, int pit, int sched_iter0, int n_s, unsigned *sij,
                                                  int *soff, int *nlev_out) {
  const int sidx = pit - sched_iter0;
  const int *ij = st.sched_ij + (size_t)sidx * n_s * 2;
  const int *off = st.sched_off + (size_t)sidx * (n_s + 1);
  for (int t = threadIdx.x; t < n_s; t += blockDim.x) sij[t] = (unsigned)ij[2 * t] | ((unsigned)ij[2 * t + 1] << 16);
  for (int t = threadIdx.x; t <= n_s; t += blockDim.x) soff[t] = off[t];
  if (threadIdx.x == 0) *nlev_out = st.sched_nlev[sidx];
}

// swap_ev_ij! for the chains this CTA owns, one warp per chain, from the replayed outcome in shared memory
__device__ void persistent_exchange_apply(const DevProblem &pb, const DevState &st, int pit, bool fused,
                                          const unsigned short *own, const unsigned short *exch, bool flow,
                                          bool one_warp = false) {
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
  const int N = pb.N, L = pb.L, R = rec_len(pb.P, pb.M);
  const double *la_all = st.la_all + (size_t)(fused ? (pit & 1) : 0) * N * R;
  const int warp = one_warp ? 0 : tid >> 5, nwarps = one_warp ? 1 : blockDim.x >> 5;  // one_warp: the caller's warp does them all
  for (int c = b + warp * G; c < L; c += nwarps * G) {  // one warp per owned chain
    const int gc = global_chain(pb, c);
    const int partner = exch[gc];
    if (partner != 0) {
      exchange_apply_chain(pb, st, pit, c, partner, la_all + (size_t)own[gc] * R, tid & 31);
      if (flow) {  // tell the CTA that will finish this chain's next evaluation that its state is up to date
        __syncwarp();
        if ((tid & 31) == 0) st_release_gpu(st.applied + c, (unsigned)pit);
      }
    }
  }
}

// exchange of iteration `pit` for the chains this CTA owns.  The pair loop is replicated in every owner CTA
// on the prefetched schedule: only the N values come from L2 here; one warp walks the levels (a __syncwarp
// per level), then one warp per owned chain applies the outcome.
// dynamic smem (persistent kernel): val[N] f64 | own[N] exch[N] u16 | sij | soff
__device__ void persistent_exchange(const DevProblem &pb, const DevState &st, int pit, bool fused, double *val,
                                    unsigned short *own, unsigned short *exch, const unsigned *sij, const int *soff,
                                    int nlev, bool flow = false, bool apply = true, const double *min_improve = nullptr) {
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
  const int N = pb.N, L = pb.L, R = rec_len(pb.P, pb.M);
  const int par = fused ? (pit & 1) : 0;
  const double *la_all = st.la_all + (size_t)par * N * R;
  const double *val_all = st.val_all + (size_t)par * N;
#ifdef SMM_LL_TU
  if (!(flow && st.ll))  // (exchange_mode 3 filled val / own / exch while it waited for the values)
#endif
  {
    for (int i = tid; i < N; i += blockDim.x) {
      val[i] = __ldcg(val_all + i);
      own[i] = (unsigned short)i;
      exch[i] = 0;
    }
    __syncthreads();
  }
  // The levels are walked by one warp (a __syncwarp per level) or, when the levels are wide (many chains over several
  // GPUs), by kExchWarps warps with a named barrier per level; the pairs of a level share no chain.
  const int nw = N > 512 ? kExchWarps : 1;
  if (tid < 32 * nw) {
    unsigned n_swaps = 0;
    for (int l = 0; l < nlev; ++l) {
      const int lo = soff[l], hi = soff[l + 1];
      for (int t = lo + tid; t < hi; t += 32 * nw) {
        const int i = (int)(sij[t] & 0xffffu), j = (int)(sij[t] >> 16);
        const double vi = val[i], vj = val[j];
        const double thr = min_improve ? min_improve[i] : pb.min_improve[i];  // shared copy: no global load per level
        if (__dsub_rn(vi, vj) > thr) {  // dist_fun(evi.value, evj.value) > min_improve[i]
          val[i] = vj;
          val[j] = vi;
          const unsigned short oi = own[i];
          own[i] = own[j];
          own[j] = oi;
          exch[i] = (unsigned short)(j + 1);
          exch[j] = (unsigned short)(i + 1);
          ++n_swaps;
        }
      }
      if (nw == 1)
        __syncwarp();
      else
        asm volatile("bar.sync %0, %1;" ::"n"(kExchBarrier), "r"(32 * nw) : "memory");
    }
    if (b == 0 && pb.rank == 0 && n_swaps) atomicAdd(&st.counters[1], (unsigned long long)n_swaps);
  }
  __syncthreads();
  if (!apply) return;
  persistent_exchange_apply(pb, st, pit, fused, own, exch, flow);
  __syncthreads();
}

// Barrier-free hand-over between iterations: every CTA waits until the rank's completion counter says that all N chains
// (of every rank) have finished the iteration -- their trace rows, records and state were stored before the finishing
// CTA's fence and its add to the counter.  One thread polls one word; the acquire fence plus the CTA barrier make those
// writes visible to the whole CTA.
__device__ bool wait_all_done(const DevProblem &pb, const DevState &st, unsigned long long target) {
  __shared__ int s_bad;
  if (threadIdx.x == 0) {
    const unsigned long long *ctr = done_counter(pb, st.val_all);
    const unsigned long long t0 = gtimer();
    unsigned spins = 0;
    int bad = 0;
    for (;;) {
      unsigned long long v;
      if (pb.world > 1)
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
      else
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if ((++spins & 63u) == 0) {
        if (ld_relaxed_gpu(&st.bar->gen) & 0x80000000u) bad = 1;  // another CTA gave up
        if (gtimer() - t0 > kSpinTimeoutNs) {
          atomicOr(st.err, kErrTimeout);
          atomicOr(&st.bar->gen, 0x80000000u);
          bad = 1;
        }
        if (bad) break;
      }
    }
    if (pb.world > 1)
      asm volatile("fence.acq_rel.sys;" ::: "memory");
    else
      fence_acq_rel_gpu();
    s_bad = bad;
  }
  __syncthreads();
  return s_bad == 0;
}

#ifdef SMM_LL_TU
// exchange_mode 3, before persistent_exchange_apply: a warp that owns a swapped chain is about to copy full records, which
// are covered by the completion counter (published a moment ago by every CTA of every rank): it waits for it here
__device__ __forceinline__ void owner_wait_records(const DevProblem &pb, const DevState &st, const unsigned short *exch,
                                                   unsigned long long target) {
  bool mine = false;  // does this CTA own a chain that was swapped?
  for (int c = blockIdx.x; c < pb.L; c += gridDim.x) mine = mine || exch[global_chain(pb, c)] != 0;
  if (!mine) return;
  if ((threadIdx.x & 31) == 0) {
    const unsigned long long *ctr = done_counter(pb, st.val_all);
    const unsigned long long t0 = gtimer();
    for (unsigned spins = 0;;) {
      unsigned long long v;
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if ((++spins & 63u) == 0 && ((ld_relaxed_gpu(&st.bar->gen) & 0x80000000u) || gtimer() - t0 > kSpinTimeoutNs)) {
        atomicOr(st.err, kErrTimeout);
        atomicOr(&st.bar->gen, 0x80000000u);
        break;
      }
    }
    if (pb.world > 1)
      asm volatile("fence.acq_rel.sys;" ::: "memory");
    else
      fence_acq_rel_gpu();
  }
  __syncwarp();
}
#endif

// The other half: after the CTA's finishing warps have stored everything (CTA barrier by the caller), one thread fences
// once -- system scope when the records went to peer GPUs -- and adds the number of chains this CTA finished to the
// completion counter of every rank (a remote atomic over NVLink for the peers).
// `stamp` (debug, with SMM_PHASE_TS): {before the fence, after it, after the adds} of this CTA.
__device__ __forceinline__ void publish_completions(const DevProblem &pb, const DevState &st, unsigned n_finished,
                                                    unsigned long long *stamp = nullptr) {
  if (n_finished == 0) return;
  if (pb.world > 1) {
    if (stamp) stamp[0] = gtimer();
    __threadfence_system();
    if (stamp) stamp[1] = gtimer();
    for (int r = 0; r < pb.world; ++r)
      asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(done_counter(pb, st.peer_val_all[r])),
                   "l"((unsigned long long)n_finished)
                   : "memory");
    if (stamp) stamp[2] = gtimer();
  } else {
    fence_acq_rel_gpu();
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(done_counter(pb, st.val_all)),
                 "l"((unsigned long long)n_finished)
                 : "memory");
  }
}

// One warp completes segment s of its CTA: publish the CTA's exact partial sums; if this was the last CTA of
// the chain, finish the chain (moments, distance, accept/reject, trace, record) -- all inside the warp, while
// the other 31 warps keep simulating.
__device__ bool warp_publish_segment(const DevProblem &pb, const DevState &st, PersistSmem &sm, int s, int it,
                                     bool fused, int part_len, int max_seg, const unsigned long long *acc,
                                     double *pp_seg, double *fscratch, int fs_len, bool flow = false,
                                     const unsigned short *exch = nullptr) {
  const int lane = threadIdx.x & 31, D = pb.P;
  const int c = sm.seg_c[s];
  double *part_base = st.partials + (size_t)c * max_seg * part_len;
  unsigned long long *part = (unsigned long long *)(part_base + (size_t)sm.seg_slot[s] * part_len);
  for (int e = lane; e < 2 * D; e += 32) part[e] = acc[(size_t)s * 2 * D + e];
  // The chain state the finishing warp will need is requested NOW: the release half of the arrival atomic waits for
  // these loads together with the partial stores, instead of a second and third L2 round trip after it.  (Whether
  // this CTA is the chain's last is not known yet; the few wasted loads of the others cost nothing.)
  const Grp gw{lane, 32, 0};
  const int gc = global_chain(pb, c);
  const unsigned long long t_pub = st.phase_ts ? gtimer() : 0ull;  // debug: when this warp started publishing
  AcceptPre pre;
  if (flow && exch && exch[gc] != 0) wait_exchange_applied(st, gw, c, it);
  accept_prefetch(pb, st, lane, c, gc, it, pre);
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    const unsigned prev = atom_acq_rel_gpu(st.arrive + c, 1u);  // releases the warp's partial, acquires the others'
    pre.u_acc = smm_acc_uniform(pb.seed_algo, (uint32_t)gc, (uint32_t)it);  // pure arithmetic while the atomic travels
    last = (prev == (unsigned)(sm.seg_nseg[s] - 1));
    if (last) st_relaxed_gpu(st.arrive + c, 0u);  // re-arm (ordered before the next iteration by the grid barrier)
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return false;
  double *f = fscratch + (size_t)s * fs_len;
  const FinScratch fs{pp_seg + (size_t)s * D, f, f + 2 * D, f + 2 * D + pb.M, (int *)(f + 2 * D + pb.M + 2)};
  if (st.phase_ts && lane == 0) st.phase_ts[(size_t)((blockIdx.x * 2 + (it & 1)) * 2 + 1) * 4 + 2] = t_pub;
  group_finalize(pb, gw, fs, part_base, sm.seg_nseg[s], part_len);
  group_accept_store(pb, st, gw, fs, c, gc, it, fused, flow, false, &pre);
  if (st.phase_ts && lane == 0) {
    st.phase_ts[(size_t)((blockIdx.x * 2 + (it & 1)) * 2 + 1) * 4 + 3] = gtimer();
    unsigned long long *row = st.phase_ts + ((size_t)gridDim.x * 4 + c) * 4;  // per-chain row of the debug buffer
    row[0] = t_pub;
    row[1] = gtimer();
    row[2] = blockIdx.x;
    row[3] = (unsigned long long)it;
  }
  return true;
}

// dynamic smem: val[N] (double) own[N] exch[N] (u16) | pp_seg[n][D] | acc[n][2D] (u64) | fscratch[n][2D+M+4] |
//               cand[kPropCand] (proposal candidates) | mi[N] | zq[32 warps][kZigQWords] (u32: deferred ziggurat draws) |
//               okf[kPropCand] (u8)
// kFlow = false (exchange_mode 1): two grid barriers per iteration, proposals by the CTA that owns the chain.
// kFlow = true  (exchange_mode 2): no grid barrier at all.  A CTA waits until every chain carries the completion tag of
//   the previous iteration, replays the exchange, computes the proposals of the chains IT simulates (from the gathered
//   records; a chain shared by several CTAs is proposed redundantly, attempts are counter-indexed), simulates, and the
//   warp finishing a chain publishes the chain's tag.  Owners apply the exchange to their chains' state off the critical
//   path and raise applied[c], which the finishing warp checks before it reads that state ~50 us later.
template <bool kFlow>
__global__ void __launch_bounds__(kPersistThreads, 1) bgp_persistent_kernel(DevProblem pb, DevState st, int iter0,
                                                                            int n_iters, int sched_iter0, int n_s,
                                                                            int part_len, int max_seg, int cta_seg,
                                                                            unsigned long long done_base) {
  __shared__ PersistSmem sm;
  __shared__ unsigned long long s_zigtab[kZigBufEntries];  // 16 KB: the 8 KB-aligned half holds the table (see load_zigtab)
  extern __shared__ double smem_d[];
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
  const int lane = tid & 31;
  const int N = pb.N, L = pb.L, P = pb.P, D = pb.P, S = pb.S;
  const bool fused = kFlow || pb.world > 1;  // records double-buffered by iteration parity in la_all / val_all
  double *val = smem_d;
  unsigned short *own = (unsigned short *)(val + N), *exch = own + N;
  const int fs_len = 2 * D + pb.M + 4;
  unsigned *sij = (unsigned *)(own + 2 * N);  // [n_s]
  int *soff = (int *)(sij + n_s);             // [n_s + 1]
  double *pp_seg = val + N + (N + 1) / 2 + (n_s + 1);  // u16 arrays = N/2 doubles, schedule = n_s + 1/2 doubles
  unsigned long long *acc = (unsigned long long *)(pp_seg + (size_t)cta_seg * D);
  double *fscratch = (double *)(acc + (size_t)cta_seg * 2 * D);
  double *cand = fscratch + (size_t)cta_seg * fs_len;                      // [kPropCand] proposal candidates
  double *mi = cand + kPropCand;                                           // [N] min_improve (read in the pair loop)
  uint32_t *zq = (uint32_t *)(mi + N) + (size_t)(tid >> 5) * kZigQWords;
  unsigned char *okf = (unsigned char *)((uint32_t *)(mi + N) + (kPersistThreads / 32) * kZigQWords);  // [kPropCand]
  for (int i = tid; i < N; i += kPersistThreads) mi[i] = pb.min_improve[i];
  load_logtab(sm.logtab);
  load_zigtab(s_zigtab);
  store_keys(pb, sm.keys);
  const uint32_t keys_s = smem_addr(sm.keys);
  unsigned gen = ld_volatile_u32(&st.bar->gen) & 0x7fffffffu;
  unsigned long long seq = ld_volatile_u64(st.sync_seq);
  const int nb = zig_blocks(S);          // Philox blocks per row of an evaluation (three draws each)
  const int n_full = S / 3;              // ... of which this many are complete
  const int n_tail = S - 3 * n_full;     // draws of the last, partial block (0 = none)
  const long long T = (long long)L * nb;
  const long long Gw = T < G ? T : G;  // CTAs that take a share of the draw space (all of them unless T is tiny)
  const long long lo = b < Gw ? (T * b) / Gw : 0, hi = b < Gw ? (T * (b + 1)) / Gw : 0;
  const int rows = 32 / D;                // Philox blocks one warp step covers (lanes >= rows*D idle if D does not divide 32)
  const int unit_j = rows * kUnitSteps;   // blocks per work unit
  if (tid == 0) {                         // this CTA's segments: one per chain its share touches
    int n = 0, u0 = 0;
    for (long long x = lo; x < hi && n < kMaxCtaSeg;) {
      const int c = (int)(x / nb);
      const long long cbase = (long long)c * nb;
      const long long xe = (cbase + nb < hi) ? cbase + nb : hi;
      const int b_first = (int)block_of(cbase, T, Gw), b_last = (int)block_of(cbase + nb - 1, T, Gw);
      sm.seg_c[n] = c;
      sm.seg_j0[n] = (int)(x - cbase);
      sm.seg_j1[n] = (int)(xe - cbase);
      sm.seg_unit0[n] = u0;
      sm.seg_slot[n] = b - b_first;
      sm.seg_nseg[n] = b_last - b_first + 1;
      sm.seg_c2[n] = pb.noseed ? (uint32_t)(global_chain(pb, c)) : 0u;
      u0 += (sm.seg_j1[n] - sm.seg_j0[n] + unit_j - 1) / unit_j;
      ++n;
      x = xe;
    }
    sm.seg_unit0[n] = u0;
    sm.n_seg = n;
    sm.total_units = u0;
  }
  __syncthreads();
  const int n_seg = sm.n_seg, total_units = sm.total_units;
  // work distribution inside the CTA: every warp first walks a fixed share (kStaticNum/kStaticDen of an equal split,
  // one queue access instead of many), the rest is handed out dynamically in shrinking grabs so that all 32 warps
  // finish within a step of each other whatever the warp scheduler favours
#ifdef SMM_LL_TU
  // exchange_mode 3: the last warp is the CTA's SERVICE warp -- at the start of phase A it publishes the previous
  // iteration's completions (a system fence: ~3.5 us) and applies the exchange to the chains the CTA owns (after waiting
  // for the completion counter: a few more) -- so it takes no fixed share and joins the dynamic hand-out when it is done;
  // the other warps never stall on either
  constexpr int kStaticWarps = kPersistThreads / 32 - 1;
#else
  constexpr int kStaticWarps = kPersistThreads / 32;
#endif
  const int static_units = (int)(((long long)total_units * kStaticNum / kStaticDen) / kStaticWarps);
  const int n_owned = b < L ? (L - b + G - 1) / G : 0;  // chains b, b + G, ...
  // proposal groups: as many threads per chain as the CTA can spare (more attempts per round)
  const int n_prop = kFlow ? n_seg : n_owned;  // chains this CTA computes proposals for
  int ngroups = 1;
  while (ngroups < n_prop && ngroups < kPGroups) ngroups <<= 1;
  const int gsize = kPersistThreads / ngroups;
  const int gi = tid / gsize;
  const Grp gprop{tid - gi * gsize, gsize, 1 + gi};
  const int cand_per_group = kPropCand / ngroups;
  const PropScratch ps{sm.g_pp[gi], sm.g_mu01[gi], cand + (size_t)gi * cand_per_group,
                       okf + (size_t)gi * cand_per_group, sm.g_first[gi], sm.g_resolved[gi], sm.logtab};
  const int k = lane % D, jo = lane / D;
  const bool lane_on = lane < rows * D;
  const bool all_on = rows * D == 32;

  for (int it = iter0; it < iter0 + n_iters; ++it) {
    bool have_ex = false;
    if (kFlow) {
      // ---- wait for iteration it-1 of every chain, replay its exchange, propose for the chains simulated here ----
#ifdef SMM_LL_TU
      // exchange_mode 3: the values themselves say when they have arrived (flag-in-data words, no counter round trip)
      const uint32_t tag = (uint32_t)(it - 1);
      const int W = ll_words(P);
      const unsigned long long *llp = st.ll + (size_t)((it - 1) & 1) * N * W;
      if (it > iter0) {
        int bad = 0;
        // (four warps poll, the other twenty wait at the barrier: every poller is L2 traffic that competes with the
        // warps still finishing their chains)
        for (int i = tid; i < N; i += 128) {
          if (tid >= 128) break;
          double v;
          if (!ll_load(st, llp + (size_t)i * W, tag, v)) bad = 1;
          val[i] = v;
          own[i] = (unsigned short)i;
          exch[i] = 0;
        }
        fence_acq_rel_gpu();  // pairs with the finishing warp's fence: the chains' local state is visible behind the values
        if (__syncthreads_or(bad)) return;
      }
#else
      if (it > iter0 && !wait_all_done(pb, st, done_base + (unsigned long long)N * (unsigned)(it - iter0))) return;
#endif
      PHASE_STAMP((b * 2 + (it & 1)) * 2, 3);
      have_ex = N > 1 && it - 1 >= 2 && it > iter0;
      if (have_ex && (n_owned > 0 || n_seg > 0))
        persistent_exchange(pb, st, it - 1, true, val, own, exch, sij, soff, sm.nlev, true, false, mi);  // replay only
      PHASE_STAMP((b * 2 + (it & 1)) * 2 + 1, 0);  // exchange done
      const double *la_prev = st.la_all + (size_t)((it - 1) & 1) * N * rec_len(P, pb.M);
      for (int r = 0; r * ngroups < n_seg; ++r) {
        const int sidx = r * ngroups + gi;
        if (sidx < n_seg) {
          const int c = sm.seg_c[sidx], gc = global_chain(pb, c);
          // centre: the record that sits on this chain after the exchange (its own last accepted one if not swapped)
          const double *centre = have_ex ? la_prev + (size_t)own[gc] * rec_len(P, pb.M) : nullptr;
#ifdef SMM_LL_TU
          const double *centre_sh = nullptr;
          if (it > iter0) {
            // the parameters that sit on this chain after the exchange come from the flag-in-data table, like the chain's
            // sigma (which stays with the chain): no read of a record that only the completion counter covers
            const int o = have_ex ? (int)own[gc] : gc;
            if (gprop.tid <= P) {
              const int q = gprop.tid < P ? 2 + gprop.tid : 1, rec = gprop.tid < P ? o : gc;
              double v;
              ll_load(st, llp + (size_t)rec * W + 2 * q, tag, v);
              sm.g_cen[gi][gprop.tid] = v;
            }
            gsync(gprop);
            centre_sh = sm.g_cen[gi];
          }
          group_proposal(pb, st, gprop, ps, c, gc, it, sm.seg_slot[sidx] == 0, centre, centre_sh);
#else
          group_proposal(pb, st, gprop, ps, c, gc, it, sm.seg_slot[sidx] == 0, centre);
#endif
          for (int q = gprop.tid; q < P; q += gsize) pp_seg[(size_t)sidx * D + q] = ps.pp[q];
        }
      }
      PHASE_STAMP((b * 2 + (it & 1)) * 2 + 1, 1);  // proposals done (group 0)
    }
    // ---- exchange of iteration it-1 (AlgoBGP.jl:637), then this iteration's proposals ----
    if (!kFlow && n_owned > 0) {
      if (N > 1 && it - 1 >= 2 && it > iter0)
        persistent_exchange(pb, st, it - 1, fused, val, own, exch, sij, soff, sm.nlev, false, true, mi);
      PHASE_STAMP((b * 2 + (it & 1)) * 2 + 1, 0);  // exchange done
      for (int r = 0; r * ngroups < n_owned; ++r) {
        const int o = r * ngroups + gi;  // index among the owned chains
        if (o < n_owned) {
          const int c = b + o * G;
          group_proposal(pb, st, gprop, ps, c, global_chain(pb, c), it, true);
          for (int q = gprop.tid; q < P; q += gsize) st.pp[(size_t)c * P + q] = ps.pp[q];
        }
      }
    }
    if (!kFlow) {
      PHASE_STAMP((b * 2 + (it & 1)) * 2 + 1, 1);  // proposals done (group 0)
      if (!grid_barrier(pb, st, gen, false, seq)) return;
    }

    // ---- phase A: this CTA's share of the flattened (chain, Philox block) space as a warp-level work queue ----
    PHASE_STAMP((b * 2 + (it & 1)) * 2, 0);
    if (!kFlow)
      for (int e = tid; e < n_seg * D; e += kPersistThreads) pp_seg[e] = __ldcg(st.pp + (size_t)sm.seg_c[e / D] * P + e % D);
    for (int e = tid; e < n_seg * 2 * D; e += kPersistThreads) acc[e] = 0ull;
    if (tid == 0) {
      sm.next_unit = static_units * kStaticWarps;
#ifdef SMM_LL_TU
      sm.n_finished_prev = sm.n_finished;
#endif
      sm.n_finished = 0u;
    }
    if ((n_owned > 0 || (kFlow && n_seg > 0)) && N > 1 && it >= 2) prefetch_schedule(st, it, sched_iter0, n_s, sij, soff, &sm.nlev);
    __syncthreads();
#ifdef SMM_LL_TU
    // exchange_mode 3, service warp: the previous iteration's completions are published HERE, off the critical path
    // (nobody on it reads the counter any more: only the swap_ev_ij! right below and the end of the launch do), then the
    // exchange is applied to the chains this CTA owns
    if (kFlow && (tid >> 5) == kStaticWarps) {
      if (lane == 0 && it > iter0)
        publish_completions(pb, st, sm.n_finished_prev,
                            st.phase_ts ? st.phase_ts + ((size_t)G * 4 + L + (size_t)b * 2 + ((it - 1) & 1)) * 4 : nullptr);
      __syncwarp();
      if (have_ex && n_owned > 0) {
        owner_wait_records(pb, st, exch, done_base + (unsigned long long)N * (unsigned)(it - iter0));
        persistent_exchange_apply(pb, st, it - 1, true, own, exch, true, true);
      }
    }
#endif
    if (pb.obj == SMM_OBJ_NORM_SLOW) slow_spin(pb.slow_seconds);
    {
      const uint32_t c3base = SMM_STREAM_SIM << 28;
      int cur = -1, j0 = 0, jfull = 0, j1 = 0;
      double p = 0.0;
      Acc a{0ull, 0ull};
      // barrier-free mode: the owner's part of the exchange (trace slot, state of the swapped chains) is off the
      // critical path -- its warps do it here, before their first units; finishers wait on applied[c] much later
#ifndef SMM_LL_TU
      if (kFlow && have_ex && n_owned > 0) persistent_exchange_apply(pb, st, it - 1, true, own, exch, true);
#endif
      ZigCtx cx = zig_ctx(s_zigtab, zq, S);
      cx.D = D;
      cx.c3 = c3base | (pb.noseed ? ((uint32_t)it & SMM_ITER_MASK) : 0u);
      cx.pvec = smem_addr(pp_seg);
      cx.fix = smem_addr(acc);
      cx.segc2 = smem_addr(sm.seg_c2);
      uint32_t ktag = (uint32_t)k;  // queue tag of this lane's blocks: row | segment << 10 (| active draws << 8)
      int qn = 0;                   // deferred ziggurat blocks of this warp; entries carry their segment, so the queue
                                    // lives across segment changes and is emptied once, when the warp is out of units
      // Leaving a segment: add this warp's exact sums to the CTA's.  (A warp may leave and re-enter a segment.)
      auto fold = [&](int seg) {
        for (int r = 1; r < rows; ++r) {
          const unsigned long long os = __shfl_down_sync(0xffffffffu, a.sum, r * D);
          const unsigned long long oq = __shfl_down_sync(0xffffffffu, a.sq, r * D);
          if (lane < D) {
            a.sum += os;
            a.sq += oq;
          }
        }
        if (lane < D) {
          atomicAdd(acc + (size_t)seg * 2 * D + lane, a.sum);
          atomicAdd(acc + (size_t)seg * 2 * D + D + lane, a.sq);
        }
        a.sum = 0ull;
        a.sq = 0ull;
      };
      // Every warp first walks its fixed share of the CTA's units; what the fixed shares leave is handed out by guided
      // self-scheduling: a warp takes (remaining / 2 warps-worth, at most kMaxGrab, at least kMinGrab) consecutive units.
      int u = (tid >> 5) * static_units, uend = u + static_units;
#ifdef SMM_LL_TU
      if ((tid >> 5) == kStaticWarps) u = uend = 0;  // the service warp only takes from the dynamic hand-out
#endif
      for (;;) {
        if (u >= uend) {
          int start = 0, g = 0;
          if (lane == 0) {
            const int rem = total_units - *(volatile int *)&sm.next_unit;
            constexpr int kTwice = 2 * (kPersistThreads / 32);
            g = rem > kTwice * kMaxGrab ? kMaxGrab : (rem > kTwice * kMinGrab ? rem / kTwice : kMinGrab);
            start = atomicAdd(&sm.next_unit, g);
          }
          start = __shfl_sync(0xffffffffu, start, 0);
          g = __shfl_sync(0xffffffffu, g, 0);
          u = start;
          uend = start + g < total_units ? start + g : total_units;
        }
        int s = cur;
        if (u < total_units) {
          if (s < 0 || u >= sm.seg_unit0[s + 1]) {
            s = s < 0 ? 0 : s + 1;
            while (u >= sm.seg_unit0[s + 1]) ++s;
          }
        } else {
          s = -1;
        }
        if (s != cur) {
          if (cur >= 0) fold(cur);
          cur = s;
          if (s >= 0) {
            j0 = sm.seg_j0[s];
            j1 = sm.seg_j1[s];
            jfull = j1 < n_full ? j1 : n_full;
            p = lane_on ? pp_seg[(size_t)s * D + k] : 0.0;
            cx.c2 = sm.seg_c2[s];
            ktag = (uint32_t)k | ((uint32_t)s << 10);
          }
        }
        if (s < 0) break;
        // the part of the grabbed range that lies in this segment: units [u, ue)
        const int ue = uend < sm.seg_unit0[s + 1] ? uend : sm.seg_unit0[s + 1];
        if (pb.obj != SMM_OBJ_FAILS) {  // every lane of the warp walks the steps (ballots inside add_block)
          const int ju = j0 + (u - sm.seg_unit0[s]) * unit_j;   // first block of the range
          const int jue = j0 + (ue - sm.seg_unit0[s]) * unit_j;  // one past its last block
          const int n_steps = (ue - u) * kUnitSteps;
          const int jb = ju + jo;
          if (all_on && jue <= jfull) {
            for (int q = 0; q < n_steps;) {  // the hot loop lives out of line; it comes back when 32 deferred blocks have gathered
              const SimRet r = sim_steps_full(keys_s, cx.ztab, cx.q, cx.c2, cx.c3, p, pb.magic_sum, pb.magic_sq,
                                              (uint32_t)(jb + q * rows), (uint32_t)rows, n_steps - q, ktag | 0x300u, qn, a.sum,
                                              a.sq);
              a.sum = r.sum;
              a.sq = r.sq;
              qn = r.qn;
              q += r.done;
              zig_relieve(pb, cx, qn);
            }
          } else {
            for (int q = 0; q < n_steps;) {
              const SimRet r = sim_steps_masked(keys_s, cx.ztab, cx.q, cx.c2, cx.c3, p, pb.magic_sum, pb.magic_sq,
                                                (uint32_t)(jb + q * rows), (uint32_t)rows, n_steps - q, ktag, qn, a.sum, a.sq,
                                                lane_on, (uint32_t)jfull, 3);
              a.sum = r.sum;
              a.sq = r.sq;
              qn = r.qn;
              q += r.done;
              zig_relieve(pb, cx, qn);
            }
          }
          if (n_tail && n_full < j1 && ju <= n_full && n_full < jue) {  // S not a multiple of 3: the last, partial block
            const bool on = lane_on && jb <= n_full && (n_full - jb) % rows == 0;
            const SimRet r = sim_steps_masked(keys_s, cx.ztab, cx.q, cx.c2, cx.c3, p, pb.magic_sum, pb.magic_sq, (uint32_t)n_full,
                                              0u, 1, ktag, qn, a.sum, a.sq, on, (uint32_t)n_full + 1u, n_tail);
            a.sum = r.sum;
            a.sq = r.sq;
            qn = r.qn;
            zig_relieve(pb, cx, qn);
          }
        }
        u = ue;
      }
      zig_flush(pb, cx, qn);  // the corrections of every segment this warp touched
    }
    PHASE_STAMP((b * 2 + (it & 1)) * 2, 1);  // warp 0 left the work loop
    __syncthreads();
    // The CTA's sums of every chain it touched are complete: one warp per segment publishes them; the warp that brings
    // the last partial of a chain finishes the chain (moments, distance, accept/reject, trace, record, completion tag).
    for (int sp = tid >> 5; sp < n_seg; sp += kPersistThreads / 32)
      if (warp_publish_segment(pb, st, sm, sp, it, fused, part_len, max_seg, acc, pp_seg, fscratch, fs_len, kFlow,
                               have_ex ? exch : nullptr) &&
          lane == 0)
        atomicAdd(&sm.n_finished, 1u);
    PHASE_STAMP((b * 2 + (it & 1)) * 2, 2);  // warp 0 done (incl. chain finalisation)
#ifdef SMM_LL_TU
    if (kFlow && it == iter0 + n_iters - 1) {  // (exchange_mode 3: only the launch's last iteration is published here)
#else
    if (kFlow) {  // one fence and one counter update per CTA tell every rank which chains are complete
#endif
      __syncthreads();
      if (tid == 0)  // (debug rows behind the per-chain rows: two per CTA, by iteration parity)
        publish_completions(pb, st, sm.n_finished,
                            st.phase_ts ? st.phase_ts + ((size_t)G * 4 + L + (size_t)b * 2 + (it & 1)) * 4 : nullptr);
    }
    if (!kFlow) {
      if (!grid_barrier(pb, st, gen, fused, seq)) return;
      PHASE_STAMP((b * 2 + (it & 1)) * 2, 3);
    }
  }
  // ---- exchange of the last iteration of this launch ----
  const int pit = iter0 + n_iters - 1;
  if (kFlow && n_owned > 0 && !wait_all_done(pb, st, done_base + (unsigned long long)N * (unsigned)n_iters)) return;
  if (n_owned > 0 && N > 1 && pit >= 2)
    persistent_exchange(pb, st, pit, fused, val, own, exch, sij, soff, sm.nlev, false, true, mi);
  if (!kFlow && b == 0 && tid == 0) *st.sync_seq = seq;
}

// ------------------------------------------------------------------------------------------------
// batched bare objective: grid (n_split, B)  (evaluateObjective, mprob.jl:175-205)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads) objective_kernel(DevProblem pb, const double *params, int noseed,
                                                                 uint32_t rep0, int n_split, int part_len,
                                                                 double *partials, unsigned *arrive, double *value,
                                                                 double *moments, int *status) {
  __shared__ EvalSmem sm;
  __shared__ unsigned long long s_zigtab[kZigBufEntries];
  const int bi = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  const Grp g{tid, (int)blockDim.x, 0};
  load_logtab(sm.logtab);
  load_zigtab(s_zigtab);
  for (int k = tid; k < pb.P; k += blockDim.x) sm.pp[k] = params[(size_t)bi * pb.P + k];
  __syncthreads();
  pb.noseed = noseed;
  double *part_base = partials + (size_t)bi * n_split * part_len;
  if (pb.obj == SMM_OBJ_FAILS) {
    if (split != 0) return;
  } else {
    if (pb.obj == SMM_OBJ_NORM_SLOW) slow_spin(pb.slow_seconds);
    const int nb = zig_blocks(pb.S);
    const int j0 = (int)(((long long)nb * split) / n_split), j1 = (int)(((long long)nb * (split + 1)) / n_split);
    simulate_static(pb, g, s_zigtab, sm.pp, j0, j1, (uint32_t)bi, rep0 + (uint32_t)bi, sm.red, sm.zfix, sm.zq,
                    part_base + (size_t)split * part_len);
    if (n_split > 1) {
      if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(arrive + bi, 1u);
        sm.is_last = (prev == (unsigned)(n_split - 1));
        if (sm.is_last) arrive[bi] = 0u;
        __threadfence();
      }
      __syncthreads();
      if (!sm.is_last) return;
    }
  }
  group_finalize(pb, g, fin_scratch(sm), part_base, n_split, part_len);
  if (tid == 0) {
    value[bi] = sm.value[0];
    status[bi] = sm.flags[1];
  }
  for (int k = tid; k < pb.M; k += blockDim.x) moments[(size_t)bi * pb.M + k] = sm.mom[k];
}

// ------------------------------------------------------------------------------------------------
// Pairs[iter] + level schedule, one CTA per iteration.
// dynamic smem: cand[n_s] pi[n_s] pj[n_s] lvl[n_s] cnt[n_s+2] last[N]   (all 32-bit)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPairThreads) bgp_pairs_kernel(DevProblem pb, DevState st, int iter0, int n_iters,
                                                                 int n_s) {
  extern __shared__ unsigned smem_u[];
  const int N = pb.N, tid = threadIdx.x, nthr = blockDim.x;
  const int it_idx = blockIdx.x;
  if (it_idx >= n_iters) return;
  const int iter = iter0 + it_idx;
  unsigned *cand = smem_u;
  int *pi = (int *)(cand + n_s), *pj = pi + n_s, *lvl = pj + n_s, *cnt = lvl + n_s, *last = cnt + n_s + 2;
  __shared__ int s_min, s_found;
  const unsigned n_all = (unsigned)((unsigned long long)N * (N - 1) / 2);
  for (int t = tid; t < n_s; t += nthr) cand[t] = smm_pair_candidate(pb.seed_algo, (uint32_t)iter, (uint32_t)t, 0u, n_all);
  __syncthreads();
  // resolve duplicates in slot order: slot t keeps its first candidate not among slots < t
  int start = 0;
  for (;;) {
    if (tid == 0) s_min = 0x7fffffff;
    __syncthreads();
    for (int t = start + tid; t < n_s; t += nthr) {
      const unsigned v = cand[t];
      bool dup = false;
      for (int s = 0; s < t && !dup; ++s) dup = (cand[s] == v);
      if (dup) atomicMin(&s_min, t);
    }
    __syncthreads();
    const int tstar = s_min;
    if (tstar == 0x7fffffff) break;
    for (unsigned a = 1;; ++a) {
      const unsigned v = smm_pair_candidate(pb.seed_algo, (uint32_t)iter, (uint32_t)tstar, a, n_all);
      __syncthreads();
      if (tid == 0) s_found = 0;
      __syncthreads();
      bool dup = false;
      for (int s = tid; s < tstar; s += nthr) dup = dup || (cand[s] == v);
      if (dup) s_found = 1;
      __syncthreads();
      if (!s_found) {
        if (tid == 0) cand[tstar] = v;
        break;
      }
    }
    __syncthreads();
    start = tstar + 1;  // slots <= tstar are final
  }
  for (int t = tid; t < n_s; t += nthr) {
    uint32_t i, j;
    smm_pair_unrank(cand[t], &i, &j);
    pi[t] = (int)i;
    pj[t] = (int)j;
  }
  for (int i = tid; i < N; i += nthr) last[i] = 0;
  for (int i = tid; i < n_s + 2; i += nthr) cnt[i] = 0;
  __syncthreads();
  if (tid == 0) {
    int maxl = 0;
    for (int t = 0; t < n_s; ++t) {
      const int i = pi[t], j = pj[t];
      const int l = 1 + max(last[i], last[j]);
      last[i] = l;
      last[j] = l;
      lvl[t] = l;
      cnt[l] += 1;
      maxl = max(maxl, l);
    }
    // cnt[l] = #pairs of level l (1-based levels) -> off[l-1] = start of level l, off[maxl] = n_s
    int *off = st.sched_off + (size_t)it_idx * (n_s + 1);
    int run = 0;
    for (int l = 1; l <= maxl; ++l) {
      const int n_l = cnt[l];
      off[l - 1] = run;
      cnt[l] = run;
      run += n_l;
    }
    off[maxl] = run;
    st.sched_nlev[it_idx] = maxl;
    // stable placement (order inside a level is irrelevant: its pairs are disjoint)
    int *ij = st.sched_ij + (size_t)it_idx * n_s * 2;
    for (int t = 0; t < n_s; ++t) {
      const int l = lvl[t];
      const int pos = cnt[l]++;
      ij[2 * pos] = pi[t];
      ij[2 * pos + 1] = pj[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------------
// zig == 0: out[2 n] Box-Muller pairs; zig != 0: out[3 n] ziggurat triples of blocks (j, k, c2, c3), j < n
__global__ void debug_normals_kernel(uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int n_blocks, int zig,
                                     double *out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_blocks) return;
  const smm_u32x4 r = smm_philox4x32_10((uint32_t)j, k, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  if (zig) {
    double z[3];
    smm_zig_triple(r, z);
    out[3 * j] = z[0];
    out[3 * j + 1] = z[1];
    out[3 * j + 2] = z[2];
  } else {
    double z0, z1;
    smm_normal_pair(r, &z0, &z1);
    out[2 * j] = z0;
    out[2 * j + 1] = z1;
  }
}

// RNG-only roofline: Philox + Box-Muller + the two accumulations, nothing else
__global__ void __launch_bounds__(kEvalThreads) rng_throughput_kernel(long long n_per_thread, double *out) {
  __shared__ smm_logent tab[1 << SMM_LOG_BITS];
  load_logtab(tab);
  __syncthreads();
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  double sum = 0.0, sq = 0.0;
  for (long long j = 0; j < n_per_thread; ++j) {
    double z0, z1;
    smm_normal_pair_tab(smm_philox4x32_10((uint32_t)j, gid, 0u, 0u, 1234u, 0u), tab, &z0, &z1);
    sum = __dadd_rn(sum, z0);
    sq = __fma_rn(z0, z0, sq);
    sum = __dadd_rn(sum, z1);
    sq = __fma_rn(z1, z1, sq);
  }
  out[2 * gid] = sum;
  out[2 * gid + 1] = sq;
}

// grid-barrier latency alone: n barriers back to back, several implementations (debug aid)
__global__ void __launch_bounds__(kPersistThreads, 1) barrier_bench_kernel(DevProblem pb, DevState st, int variant,
                                                                           int n) {
  unsigned gen = ld_volatile_u32(&st.bar->gen);
  unsigned long long seq = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    if (variant == 0) {
      if (!grid_barrier(pb, st, gen, false, seq)) return;
    } else if (variant == 1) {  // no fences at all (lower bound: one RED + two polls)
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned target = ++gen;
        if (blockIdx.x == 0) {
          while (ld_relaxed_gpu(&st.bar->arrive) < gridDim.x - 1) {}
          st_relaxed_gpu(&st.bar->arrive, 0u);
          st_relaxed_gpu(&st.bar->gen, target);
        } else {
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(&st.bar->arrive), "r"(1u) : "memory");
          while ((int)(ld_relaxed_gpu(&st.bar->gen) - target) < 0) {}
        }
      }
      __syncthreads();
    } else if (variant == 2) {  // release arrive, relaxed everything else
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned target = ++gen;
        if (blockIdx.x == 0) {
          while (ld_relaxed_gpu(&st.bar->arrive) < gridDim.x - 1) {}
          st_relaxed_gpu(&st.bar->arrive, 0u);
          fence_acq_rel_gpu();
          st_relaxed_gpu(&st.bar->gen, target);
        } else {
          red_release_gpu(&st.bar->arrive, 1u);
          while ((int)(ld_relaxed_gpu(&st.bar->gen) - target) < 0) {}
        }
      }
      __syncthreads();
    } else {  // classic: __threadfence + atomicAdd + volatile polls + __threadfence
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned target = ++gen;
        __threadfence();
        if (blockIdx.x == 0) {
          while (ld_volatile_u32(&st.bar->arrive) < gridDim.x - 1) {}
          st.bar->arrive = 0;
          __threadfence();
          *(volatile unsigned *)&st.bar->gen = target;
        } else {
          atomicAdd(&st.bar->arrive, 1u);
          while ((int)(ld_volatile_u32(&st.bar->gen) - target) < 0) {}
        }
        __threadfence();
      }
      __syncthreads();
    }
  }
}

// the simulate inner loop alone (add_block with the handle's keys and accumulators), static or dynamic unit
// distribution, any CTA size: the ceiling the evaluation kernels are measured against
// kBound: the launch bound the variant is compiled for (1024 -> 64 registers per thread, 768 -> 80, 512 -> 128), to
// measure what the register budget of the persistent kernel's CTA shape costs
template <int kBound>
__global__ void __launch_bounds__(kBound) sim_throughput_kernel(DevProblem pb, int n_per_thread, int dyn, double *out) {
  __shared__ unsigned long long ztab[kZigBufEntries];
  extern __shared__ __align__(16) uint32_t zq[];  // [blockDim.x / 32][kZigQWords]
  __shared__ unsigned long long fix[2 * SMM_MAX_PARAMS];
  __shared__ double pvec[SMM_MAX_PARAMS];
  __shared__ __align__(16) uint32_t keys[20];
  __shared__ int ctr;
  load_zigtab(ztab);
  store_keys(pb, keys);
  if (threadIdx.x == 0) ctr = 0;
  const int lane = threadIdx.x & 31, D = pb.P;
  for (int e = threadIdx.x; e < D; e += blockDim.x) pvec[e] = 0.25 * (double)e;
  for (int e = threadIdx.x; e < 2 * D; e += blockDim.x) fix[e] = 0ull;
  __syncthreads();
  const uint32_t k = (uint32_t)(lane % D);
  Acc a{0ull, 0ull};
  const double p = pvec[k];
  ZigCtx cx = zig_ctx(ztab, zq + (size_t)(threadIdx.x >> 5) * kZigQWords, pb.S);
  cx.pvec = smem_addr(pvec);
  cx.fix = smem_addr(fix);
  cx.D = D;
  cx.c3 = SMM_STREAM_SIM << 28;
  const uint32_t kpack3 = k | 0x300u, keys_s = smem_addr(keys);
  int qn = 0;
  const int variant = dyn >> 2;
  dyn &= 1;
  auto steps = [&](uint32_t j, uint32_t dj, int n) {
    for (int q = 0; q < n;) {
      const uint32_t jq = j + (uint32_t)q * dj;
#define SMM_TPUT_CALL(V) sim_steps_full<V>(keys_s, cx.ztab, cx.q, cx.c2, cx.c3, p, pb.magic_sum, pb.magic_sq, jq, dj, n - q, kpack3, qn, a.sum, a.sq)
      const SimRet r = variant == 0 ? SMM_TPUT_CALL(0) : variant == 1 ? SMM_TPUT_CALL(1) : variant == 2 ? SMM_TPUT_CALL(2) : SMM_TPUT_CALL(3);
#undef SMM_TPUT_CALL
      a.sum = r.sum;
      a.sq = r.sq;
      qn = r.qn;
      q += r.done;
      zig_relieve(pb, cx, qn);
    }
  };
  if (!dyn) {
    const uint32_t j0 = (blockIdx.x * blockDim.x + threadIdx.x) / D * (uint32_t)n_per_thread;
    steps(j0, 1u, n_per_thread);
  } else {
    const int rows = 32 / D, unit = rows * kTputSteps, jo = lane / D;
    const int jmax = n_per_thread * (blockDim.x / D) / unit * unit;
    int next = 0;
    if (lane == 0) next = atomicAdd(&ctr, unit);
    next = __shfl_sync(0xffffffffu, next, 0);
    while (next < jmax) {
      const int base = next + jo;
      if (lane == 0) next = atomicAdd(&ctr, unit);
      steps((uint32_t)base + blockIdx.x * 1000003u, (uint32_t)rows, kTputSteps);
      next = __shfl_sync(0xffffffffu, next, 0);
    }
  }
  zig_flush(pb, cx, qn);
  __syncthreads();
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  out[2 * gid] = __longlong_as_double((long long)(a.sum + fix[k]));
  out[2 * gid + 1] = __longlong_as_double((long long)(a.sq + fix[D + k]));
}

// ------------------------------------------------------------------------------------------------
// dynamic-panel objective (C4): its kernels use the device functions above
// ------------------------------------------------------------------------------------------------
#ifndef SMM_LL_TU
#include "smm_panel.cuh"
#endif

// ------------------------------------------------------------------------------------------------
// launchers (called from smm_api.cu)
// ------------------------------------------------------------------------------------------------
size_t pairs_smem_bytes(int N, int n_s) { return sizeof(unsigned) * ((size_t)5 * n_s + 2 + N); }
size_t exch_smem_bytes(int N) { return sizeof(double) * (size_t)N + sizeof(unsigned short) * 2 * (size_t)N; }
// persistent kernel: exchange arrays (u16 part rounded up to whole doubles) + per-segment pp / sums / finalisation scratch
size_t persist_smem_bytes(int N, int D, int M, int cta_seg) {
  const int n_s = N < 3 ? (N > 1 ? N - 1 : 0) : N;
  return sizeof(double) * ((size_t)N + (N + 1) / 2 + (n_s + 1) + (size_t)cta_seg * (D + 2 * D + 2 * D + M + 4)) +
         sizeof(double) * ((size_t)kPropCand + N) + sizeof(uint32_t) * (size_t)(kPersistThreads / 32) * kZigQWords + kPropCand;
}

cudaError_t configure_kernels(int N, int n_s) {
  cudaError_t e = cudaFuncSetAttribute(bgp_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pairs_smem_bytes(N, n_s));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(bgp_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exch_smem_bytes(N));
  if (e != cudaSuccess) return e;
  return cudaSuccess;
}
cudaError_t configure_persistent(int N, int D, int M, int cta_seg) {
  cudaError_t e = cudaFuncSetAttribute(bgp_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)persist_smem_bytes(N, D, M, cta_seg));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(bgp_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)persist_smem_bytes(N, D, M, cta_seg));
}

int eval_max_blocks_per_sm() {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bgp_eval_kernel, kEvalThreads, 0);
  return n;
}
int persistent_max_blocks_per_sm(int N, int D, int M, int cta_seg) {
  int n = 0;
  int m = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bgp_persistent_kernel<false>, kPersistThreads,
                                                persist_smem_bytes(N, D, M, cta_seg));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, bgp_persistent_kernel<true>, kPersistThreads,
                                                persist_smem_bytes(N, D, M, cta_seg));
  return n < m ? n : m;
}
int persistent_unit_blocks(int D) { return (32 / D) * kUnitSteps; }
int persistent_max_cta_seg() { return kMaxCtaSeg; }

void launch_eval(const DevProblem &pb, const DevState &st, int iter, int n_split, int part_len, cudaStream_t s) {
  dim3 grid(n_split, pb.L);
  bgp_eval_kernel<<<grid, kEvalThreads, 0, s>>>(pb, st, iter, n_split, part_len);
}
void launch_pairs(const DevProblem &pb, const DevState &st, int iter0, int n_iters, int n_s, cudaStream_t s) {
  bgp_pairs_kernel<<<n_iters, kPairThreads, pairs_smem_bytes(pb.N, n_s), s>>>(pb, st, iter0, n_iters, n_s);
}
void launch_exchange(const DevProblem &pb, const DevState &st, int iter, int sched_idx, int n_s, cudaStream_t s) {
  bgp_exchange_kernel<<<1, kExchThreads, exch_smem_bytes(pb.N), s>>>(pb, st, iter, sched_idx, n_s);
}
cudaError_t launch_persistent(const DevProblem &pb, const DevState &st, int iter0, int n_iters, int sched_iter0,
                              int n_s, int part_len, int max_seg, int cta_seg, int grid, bool flow,
                              unsigned long long done_base, cudaStream_t s) {
  DevProblem pbc = pb;
  DevState stc = st;
  void *args[] = {&pbc, &stc, &iter0, &n_iters, &sched_iter0, &n_s, &part_len, &max_seg, &cta_seg, &done_base};
  void *fn = flow ? (void *)bgp_persistent_kernel<true> : (void *)bgp_persistent_kernel<false>;
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPersistThreads), args,
                                     persist_smem_bytes(pb.N, pb.P, pb.M, cta_seg), s);
}
void launch_objective(const DevProblem &pb, const double *params, int B, int noseed, uint32_t rep0, int n_split,
                      int part_len, double *partials, unsigned *arrive, double *value, double *moments, int *status,
                      cudaStream_t s) {
  dim3 grid(n_split, B);
  objective_kernel<<<grid, kEvalThreads, 0, s>>>(pb, params, noseed, rep0, n_split, part_len, partials, arrive, value,
                                                 moments, status);
}
void launch_debug_normals(uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int n_pairs, int zig, double *out,
                          cudaStream_t s) {
  debug_normals_kernel<<<(n_pairs + 255) / 256, 256, 0, s>>>(seed, k, c2, c3, n_pairs, zig, out);
}
void launch_rng_throughput(long long n_per_thread, int blocks, double *out, cudaStream_t s) {
  rng_throughput_kernel<<<blocks, kEvalThreads, 0, s>>>(n_per_thread, out);
}
cudaError_t launch_barrier_bench(const DevProblem &pb, const DevState &st, int variant, int n, int grid, cudaStream_t s) {
  DevProblem pbc = pb;
  DevState stc = st;
  void *args[] = {&pbc, &stc, &variant, &n};
  return cudaLaunchCooperativeKernel((void *)barrier_bench_kernel, dim3(grid), dim3(kPersistThreads), args, 0, s);
}
void launch_sim_throughput(const DevProblem &pb, int n_per_thread, int blocks, int threads, int dyn, double *out,
                           cudaStream_t s) {
  const size_t dyn_smem = sizeof(uint32_t) * (threads / 32) * kZigQWords;
  // dyn: 0 = static split, 1 = unit queue; +2 = the variant compiled for the smallest launch bound >= threads
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    kern<<<blocks, threads, dyn_smem, s>>>(pb, n_per_thread, (dyn & 1) | (dyn >> 2 << 2), out);
  };
  if ((dyn & 2) && threads <= 512)
    go(sim_throughput_kernel<512>);
  else if ((dyn & 2) && threads <= 768)
    go(sim_throughput_kernel<768>);
  else
    go(sim_throughput_kernel<1024>);
}

#ifdef SMM_LL_TU
}  // namespace ll
#endif
}  // namespace smm
