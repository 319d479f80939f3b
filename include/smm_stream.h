/* smm_stream.h -- NORMATIVE definition of the counter-indexed random streams of the BGP hot path.
 *
 * The reference draws its randomness from Julia's global RNG and an unseedable RandomDevice
 * (/root/reference/src/SMM.jl:59-60, src/mopt/AlgoBGP.jl:85,404,656, src/mopt/ObjExamples.jl:74-79),
 * none of which is reproducible outside one Julia build.  Parity is therefore defined on the
 * reference's ALGORITHM as a pure function of four injected streams:
 *
 *   Zsim [row k, draw s]                 standard normals of the model simulator
 *   Zprop[chain, iter, attempt, param k] standard normals of the truncated random-walk proposal
 *   Uacc [chain, iter]                   the pre-drawn Metropolis uniforms  (BGPChain.probs_acc)
 *   Pairs[iter][t]                       the exchange-move pair sample
 *
 * Every element is a pure function of (seed, indices): Philox4x32-10 keyed by the seed, the indices in
 * the 128-bit counter, then a fixed fp64 transform that uses only + - * fma / sqrt and therefore
 * yields BIT-IDENTICAL doubles on the host (oracle, Julia CPU objfunc) and on the device (kernels).
 * This header is included by the CUDA kernels (smm_jl_b200/csrc) and by the CPU oracle (oracle/);
 * it contains no algorithm of the BGP sampler itself.
 *
 * Compile host code with -ffp-contract=off (and -mfma for speed); device code uses the _rn
 * intrinsics so nvcc's -fmad setting cannot change results.
 */
#ifndef SMM_STREAM_H
#define SMM_STREAM_H

#include <stdint.h>
#include <string.h>
#include <math.h>
#include "smm_stream_tables.h"

#if defined(__CUDACC__)
#define SMM_HD __host__ __device__ __forceinline__
#else
#define SMM_HD static inline
#endif

/* stream ids: the top 4 bits of counter word 3; the low 28 bits carry the iteration */
#define SMM_STREAM_SIM 1u
#define SMM_STREAM_PROP 2u
#define SMM_STREAM_ACC 3u
#define SMM_STREAM_PAIR 4u
#define SMM_ITER_MASK 0x0FFFFFFFu

/* ---- exactly-rounded primitive ops (never contracted) ------------------------------------ */
#if defined(__CUDA_ARCH__)
#define SMM_MUL(a, b) __dmul_rn((a), (b))
#define SMM_ADD(a, b) __dadd_rn((a), (b))
#define SMM_SUB(a, b) __dsub_rn((a), (b))
#define SMM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SMM_SQRT(a) __dsqrt_rn((a))
#else
#define SMM_MUL(a, b) ((a) * (b))
#define SMM_ADD(a, b) ((a) + (b))
#define SMM_SUB(a, b) ((a) - (b))
#define SMM_FMA(a, b, c) fma((a), (b), (c))
#define SMM_SQRT(a) sqrt((a))
#endif

/* ---- Philox4x32-10 (Salmon et al., SC'11; same constants as Random123 / cuRAND) ---------- */
typedef struct smm_u32x4 {
  uint32_t x, y, z, w;
} smm_u32x4;

#define SMM_PHILOX_M0 0xD2511F53u
#define SMM_PHILOX_M1 0xCD9E8D57u
#define SMM_PHILOX_W0 0x9E3779B9u
#define SMM_PHILOX_W1 0xBB67AE85u

SMM_HD void smm_mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo) {
#if defined(__CUDA_ARCH__)
  *lo = a * b;
  *hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *lo = (uint32_t)p;
  *hi = (uint32_t)(p >> 32);
#endif
}

SMM_HD smm_u32x4 smm_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                   uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    smm_mulhilo(SMM_PHILOX_M0, c0, &hi0, &lo0);
    smm_mulhilo(SMM_PHILOX_M1, c2, &hi1, &lo1);
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += SMM_PHILOX_W0;
    k1 += SMM_PHILOX_W1;
  }
  smm_u32x4 out;
  out.x = c0;
  out.y = c1;
  out.z = c2;
  out.w = c3;
  return out;
}

/* ---- bit casts --------------------------------------------------------------------------- */
SMM_HD double smm_bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, sizeof d);
  return d;
#endif
}
SMM_HD uint64_t smm_double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t b;
  memcpy(&b, &d, sizeof b);
  return b;
#endif
}

/* 52 random mantissa bits from two words: a supplies the high 32, b its top 20 */
SMM_HD uint64_t smm_mant52(uint32_t a, uint32_t b) { return ((uint64_t)a << 20) | (uint64_t)(b >> 12); }

/* uniform on [0,1) with 52-bit resolution (Julia's rand(Float64) convention) */
SMM_HD double smm_u01(uint32_t a, uint32_t b) {
  return SMM_SUB(smm_bits_to_double(0x3FF0000000000000ull | smm_mant52(a, b)), 1.0);
}

/* ---- constants: host arrays, and on the device the constant bank (so that DFMA takes the
 * coefficient as a c[bank][offset] operand instead of materialising it in registers) ------------- */
typedef struct smm_logent {
  double inv, nlnc;
} smm_logent;

static const smm_logent SMM_LOGTAB_HOST[1 << SMM_LOG_BITS] = SMM_LOG_TABLE;
static const double SMM_SIN_HOST[SMM_SIN_DEG + 1] = SMM_SIN_COEFS;
static const double SMM_COS_HOST[SMM_COS_DEG + 1] = SMM_COS_COEFS;
static const double SMM_LOGQ_HOST[SMM_LOGQ_DEG + 1] = SMM_LOGQ_COEFS;
/* ziggurat table (see smm_zig_fast): one 8-byte entry per layer = the bit pattern of the double W'[i] * 2^-32, whose
 * low 12 mantissa bits ARE the fast-accept bound KH[i] = floor(2^12 W'[i+1] / W'[i]) on the top 12 bits of the
 * 32-bit uniform.  W'[i] (= 2^32 * entry) is the outer edge of layer i (virtual width of the base strip for i = 0). */
typedef uint64_t smm_zigent;
static const smm_zigent SMM_ZIGTAB_HOST[SMM_ZIG_LAYERS] = SMM_ZIG_TABLE;
static const double SMM_ZIGF_HOST[SMM_ZIG_LAYERS + 1] = SMM_ZIG_F;
static const double SMM_EXP_HOST[SMM_EXP_DEG + 1] = SMM_EXP_COEFS;
#if defined(__CUDACC__)
static __device__ const smm_zigent SMM_ZIGTAB_DEV[SMM_ZIG_LAYERS] = SMM_ZIG_TABLE;
static __device__ const double SMM_ZIGF_DEV[SMM_ZIG_LAYERS + 1] = SMM_ZIG_F;
static __constant__ double SMM_EXP_DEV[SMM_EXP_DEG + 1] = SMM_EXP_COEFS;
static __device__ const smm_logent SMM_LOGTAB_DEV[1 << SMM_LOG_BITS] = SMM_LOG_TABLE;
static __constant__ double SMM_SIN_DEV[SMM_SIN_DEG + 1] = SMM_SIN_COEFS;
static __constant__ double SMM_COS_DEV[SMM_COS_DEG + 1] = SMM_COS_COEFS;
static __constant__ double SMM_LOGQ_DEV[SMM_LOGQ_DEG + 1] = SMM_LOGQ_COEFS;
#endif
#if defined(__CUDA_ARCH__)
#define SMM_SIN_C(i) SMM_SIN_DEV[i]
#define SMM_COS_C(i) SMM_COS_DEV[i]
#define SMM_LOGQ_C(i) SMM_LOGQ_DEV[i]
#define SMM_EXP_C(i) SMM_EXP_DEV[i]
#define SMM_ZIGF(i) SMM_ZIGF_DEV[i]
#else
#define SMM_SIN_C(i) SMM_SIN_HOST[i]
#define SMM_COS_C(i) SMM_COS_HOST[i]
#define SMM_LOGQ_C(i) SMM_LOGQ_HOST[i]
#define SMM_EXP_C(i) SMM_EXP_HOST[i]
#define SMM_ZIGF(i) SMM_ZIGF_HOST[i]
#endif

SMM_HD const smm_logent *smm_logtab(void) {
#if defined(__CUDA_ARCH__)
  return SMM_LOGTAB_DEV;
#else
  return SMM_LOGTAB_HOST;
#endif
}
SMM_HD const smm_zigent *smm_zigtab(void) {
#if defined(__CUDA_ARCH__)
  return SMM_ZIGTAB_DEV;
#else
  return SMM_ZIGTAB_HOST;
#endif
}

/* -log(u) of a normal double u in (0,1); abs error ~1e-16 (tests/test_stream.py).
 * u = 2^e * m, m in [1,2); bucket j = top SMM_LOG_BITS bits of m; r = m*INV[j] - 1 (one fma);
 * -log(u) = (e * -ln2 + NLNC[j]) + (r^2 * NQ(r) - r).  The end buckets have INV = 1 and 1/2 exactly,
 * so the result is accurate to the last bits on both sides of u = 1. */
SMM_HD double smm_neglog01(double u, const smm_logent *tab) {
  const uint64_t b = smm_double_to_bits(u);
  const uint32_t hi = (uint32_t)(b >> 32);
  const int e = (int)((hi >> 20) & 0x7FFu) - 1023;
  const uint32_t j = (hi >> (20 - SMM_LOG_BITS)) & ((1u << SMM_LOG_BITS) - 1u);
  const double m = smm_bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  const smm_logent t = tab[j];
  const double r = SMM_FMA(m, t.inv, -1.0);
  double p = SMM_LOGQ_C(SMM_LOGQ_DEG);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = SMM_LOGQ_DEG - 1; i >= 0; --i) p = SMM_FMA(p, r, SMM_LOGQ_C(i));
  const double r2 = SMM_MUL(r, r);
  const double nl1p = SMM_FMA(r2, p, -r);                     /* -log(1+r) */
  const double base = SMM_FMA((double)e, SMM_NLN2, t.nlnc);   /* -(e ln2 + ln c_j); 0 near u = 1 */
  return SMM_ADD(base, nl1p);
}

/* Box-Muller on one Philox block r = (x, y, z, w):
 *   A = 52 bits of (x, y) forced odd    -> u1 = 1 - A/2^52 in [2^-52, 1 - 2^-52]   (radius)
 *   B = 52 bits of (z, w): low 49 bits  -> g in [0,1), first-octant angle (pi/4) g
 *                          bits 49,50,51 -> a uniformly random symmetry of the octant tiling:
 *                                           swap (x<->y), sign of z0, sign of z1
 *   z0, z1 = sqrt(-log u1) * { sqrt2 cos, sqrt2 sin }((pi/4) g), swapped / negated as drawn.
 * A uniform point of the first octant under a uniform element of the dihedral group is uniform on the
 * circle, so (z0, z1) are independent standard normals. */
SMM_HD void smm_normal_pair_tab(smm_u32x4 r, const smm_logent *tab, double *z0, double *z1) {
  const double d1 = smm_bits_to_double(0x3FF0000000000000ull | smm_mant52(r.x, r.y) | 1ull);
  const double u1 = SMM_SUB(2.0, d1); /* exact */
  const double rad = SMM_SQRT(smm_neglog01(u1, tab));
  const uint64_t rem = smm_mant52(r.z, r.w) & ((1ull << 49) - 1ull);
  const double g = SMM_SUB(smm_bits_to_double(0x3FF0000000000000ull | (rem << 3)), 1.0); /* exact */
  const double w = SMM_MUL(g, g);
  double ps = SMM_SIN_C(SMM_SIN_DEG);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = SMM_SIN_DEG - 1; i >= 0; --i) ps = SMM_FMA(ps, w, SMM_SIN_C(i));
  double pc = SMM_COS_C(SMM_COS_DEG);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = SMM_COS_DEG - 1; i >= 0; --i) pc = SMM_FMA(pc, w, SMM_COS_C(i));
  const double sn = SMM_MUL(g, ps);
  const int swap = (r.z >> 29) & 1u;
  uint64_t cb = smm_double_to_bits(swap ? sn : pc);
  uint64_t sb = smm_double_to_bits(swap ? pc : sn);
  cb ^= (uint64_t)((r.z << 1) & 0x80000000u) << 32; /* bit 30 of z -> sign of z0 */
  sb ^= (uint64_t)(r.z & 0x80000000u) << 32;        /* bit 31 of z -> sign of z1 */
  *z0 = SMM_MUL(rad, smm_bits_to_double(cb));
  *z1 = SMM_MUL(rad, smm_bits_to_double(sb));
}
SMM_HD void smm_normal_pair(smm_u32x4 r, double *z0, double *z1) {
  smm_normal_pair_tab(r, smm_logtab(), z0, z1);
}

/* ---- ziggurat normals: the simulator stream of the MvNormal objectives ------------------------
 *
 * Julia's randn -- what the reference's rand(MvNormal(..), ns) (ObjExamples.jl:78) bottoms out in -- is a ziggurat;
 * so is this, restated on counter-indexed bits so that it is a pure function.  ONE Philox block (x, y, z, w) yields
 * THREE normals: draw t = 0, 1, 2 takes the 32-bit uniform word u = (x, y, z)[t] and the 10-bit select field
 * s = smm_zig_select(w, t) = [sign : 1][layer i : 9]  (w bits 3..12 | 13..22 | {0, 23..31}; bits 1, 2 are unused):
 *
 *     x = u * (W'[i] 2^-32)                      (u converted exactly; one fp64 multiply; 512 layers)
 *     FAST PATH (99.2 % of draws): accept when the top 12 bits of u are below KH[i], i.e.  u < KH[i] 2^20
 *       (x strictly inside the layer's core: two integer ops, one 8-byte table load, two fp64 ops)
 *     otherwise (smm_zig_slow):
 *       i >= 1: wedge test  f(W'[i]) + U (f(W'[i+1]) - f(W'[i])) < exp(-x^2/2); on rejection a NEW candidate (u, s)
 *       i == 0: x < R is still inside the base strip; beyond R -> Marsaglia's tail: repeat xt = -log(U1)/R,
 *               yt = -log(U2) until 2 yt > xt^2, return R + xt
 *     The extra uniforms and candidates come from auxiliary blocks Philox(u, s, n, TAG; ZIG key), n = 1, 2, ... of the
 *     draw's ORIGINAL (u, s), in the order the state machine consumes them -- a deterministic function of (u, s) alone.
 *
 * exp and log are the fma-only kernels of this header, so host and device agree to the bit. */
#define SMM_ZIG_TAG 0x5A494732u
#define SMM_ZIG_KEY0 0x736D6D5Au
#define SMM_ZIG_KEY1 0x69676733u
#define SMM_ZIG_MAX_AUX 4096u
#define SMM_ZIG_PER_BLOCK 3
#define SMM_ZIG_SEL_MASK ((2u << SMM_ZIG_LAYER_BITS) - 1u) /* sign + layer */

/* exp(t), t in [-700, 0]: t = n ln2 + r, |r| <= ln2/2 (+ rounding), Taylor degree 13, scaled by 2^n */
SMM_HD double smm_exp_neg(double t) {
  const double shift = 6755399441055744.0; /* 1.5 * 2^52: adding it rounds to the nearest integer */
  const double nf = SMM_SUB(SMM_FMA(t, SMM_LOG2E, shift), shift);
  double r = SMM_FMA(nf, -SMM_LN2_HI, t);
  r = SMM_FMA(nf, -SMM_LN2_LO, r);
  double p = SMM_EXP_C(SMM_EXP_DEG);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = SMM_EXP_DEG - 1; i >= 0; --i) p = SMM_FMA(p, r, SMM_EXP_C(i));
  const int n = (int)nf;
  return SMM_MUL(p, smm_bits_to_double((uint64_t)(n + 1023) << 52));
}

/* select field of draw t (0, 1, 2) in the block's fourth word */
SMM_HD uint32_t smm_zig_select(uint32_t w, int t) {
  return (t == 0 ? (w >> 3) : t == 1 ? (w >> 13) : ((w >> 23) | (w << 9))) & SMM_ZIG_SEL_MASK;
}

/* (double)u, exactly, without an integer conversion: 2^52 + u has u in its low mantissa word */
SMM_HD double smm_u32_to_double(uint32_t u) {
  return SMM_SUB(smm_bits_to_double(0x4330000000000000ull | (uint64_t)u), 4503599627370496.0);
}

/* the fast path: candidate value and whether it is accepted outright */
SMM_HD double smm_zig_fast(uint32_t u, uint32_t s, const smm_zigent *tab, int *ok) {
  const smm_zigent e = tab[s & (SMM_ZIG_LAYERS - 1u)];
  const double x = SMM_MUL(smm_u32_to_double(u), smm_bits_to_double(e));
  *ok = u < ((uint32_t)e << 20);
  return smm_bits_to_double(smm_double_to_bits(x) | ((uint64_t)(s >> SMM_ZIG_LAYER_BITS) << 63));
}

/* u in [2^-52, 1 - 2^-52] from two words (odd 52-bit integer: never 0 or 1) */
SMM_HD double smm_u01_open(uint32_t a, uint32_t b) {
  return SMM_SUB(2.0, smm_bits_to_double(0x3FF0000000000000ull | smm_mant52(a, b) | 1ull));
}

/* everything after a failed fast test of candidate (u0, s0) */
SMM_HD double smm_zig_slow(uint32_t u0, uint32_t s0, const smm_zigent *tab, const smm_logent *logtab) {
  uint32_t u = u0, s = s0, n = 0;
  for (;;) {
    const uint32_t i = s & (SMM_ZIG_LAYERS - 1u);
    const uint64_t sign = (uint64_t)(s >> SMM_ZIG_LAYER_BITS) << 63;
    const smm_zigent e = tab[i];
    const double x = SMM_MUL(smm_u32_to_double(u), smm_bits_to_double(e));
    if (u < ((uint32_t)e << 20)) return smm_bits_to_double(smm_double_to_bits(x) | sign);
    if (i == 0u) {
      if (x < SMM_ZIG_R) return smm_bits_to_double(smm_double_to_bits(x) | sign); /* base strip */
      for (;;) {                                                                   /* tail beyond R */
        ++n;
        const smm_u32x4 r = smm_philox4x32_10(u0, s0, n, SMM_ZIG_TAG, SMM_ZIG_KEY0, SMM_ZIG_KEY1);
        const double xt = SMM_MUL(smm_neglog01(smm_u01_open(r.x, r.y), logtab), SMM_ZIG_RINV);
        const double yt = smm_neglog01(smm_u01_open(r.z, r.w), logtab);
        if (SMM_ADD(yt, yt) > SMM_MUL(xt, xt) || n >= SMM_ZIG_MAX_AUX)
          return smm_bits_to_double(smm_double_to_bits(SMM_ADD(SMM_ZIG_R, xt)) | sign);
      }
    }
    ++n;
    const smm_u32x4 r = smm_philox4x32_10(u0, s0, n, SMM_ZIG_TAG, SMM_ZIG_KEY0, SMM_ZIG_KEY1);
    const double f_lo = SMM_ZIGF(i), f_hi = SMM_ZIGF(i + 1u);
    const double y = SMM_FMA(smm_u01(r.z, r.w), SMM_SUB(f_hi, f_lo), f_lo);
    if (y < smm_exp_neg(SMM_MUL(SMM_MUL(x, x), -0.5)) || n >= SMM_ZIG_MAX_AUX)
      return smm_bits_to_double(smm_double_to_bits(x) | sign);
    u = r.x; /* rejected: start over with a fresh candidate */
    s = r.y & SMM_ZIG_SEL_MASK;
  }
}

SMM_HD double smm_zig_normal_tab(uint32_t u, uint32_t s, const smm_zigent *tab, const smm_logent *logtab) {
  int ok;
  const double z = smm_zig_fast(u, s, tab, &ok);
  return ok ? z : smm_zig_slow(u, s, tab, logtab);
}
/* the three normals of one Philox block */
SMM_HD void smm_zig_triple(smm_u32x4 r, double *z) {
  z[0] = smm_zig_normal_tab(r.x, smm_zig_select(r.w, 0), smm_zigtab(), smm_logtab());
  z[1] = smm_zig_normal_tab(r.y, smm_zig_select(r.w, 1), smm_zigtab(), smm_logtab());
  z[2] = smm_zig_normal_tab(r.z, smm_zig_select(r.w, 2), smm_zigtab(), smm_logtab());
}

/* ---- the four streams ----------------------------------------------------------------------- */

/* The block of Zsim[k, .] number j: ziggurat objectives take Zsim[k, 3j .. 3j+2] from it (smm_zig_triple), the
 * dynamic panel two Box-Muller normals (smm_normal_pair).  With common random numbers (noseed == 0, the reference's
 * Random.seed!(1234) at ObjExamples.jl:74) every evaluation sees the same block; with noseed the
 * block is additionally indexed by (eval_uid, rep). */
SMM_HD smm_u32x4 smm_sim_block(uint64_t seed_sim, uint32_t j, uint32_t k, int noseed, uint32_t eval_uid,
                               uint32_t rep) {
  const uint32_t c2 = noseed ? eval_uid : 0u;
  const uint32_t c3 = (SMM_STREAM_SIM << 28) | (noseed ? (rep & SMM_ITER_MASK) : 0u);
  return smm_philox4x32_10(j, k, c2, c3, (uint32_t)seed_sim, (uint32_t)(seed_sim >> 32));
}

/* Zprop[chain, iter, attempt, 2*kpair] and [.., 2*kpair+1]  (chain 0-based global id, iter 1-based) */
SMM_HD smm_u32x4 smm_prop_block(uint64_t seed_algo, uint32_t chain, uint32_t iter, uint32_t attempt,
                                uint32_t kpair) {
  return smm_philox4x32_10(attempt, kpair, chain, (SMM_STREAM_PROP << 28) | (iter & SMM_ITER_MASK),
                           (uint32_t)seed_algo, (uint32_t)(seed_algo >> 32));
}

/* Uacc[chain, iter] in [0,1) */
SMM_HD double smm_acc_uniform(uint64_t seed_algo, uint32_t chain, uint32_t iter) {
  const smm_u32x4 r = smm_philox4x32_10(0u, 0u, chain, (SMM_STREAM_ACC << 28) | (iter & SMM_ITER_MASK),
                                        (uint32_t)seed_algo, (uint32_t)(seed_algo >> 32));
  return smm_u01(r.x, r.y);
}

/* candidate rank in [0, n_pairs) for slot t of Pairs[iter], redraw number `attempt` */
SMM_HD uint32_t smm_pair_candidate(uint64_t seed_algo, uint32_t iter, uint32_t t, uint32_t attempt,
                                   uint32_t n_pairs) {
  const smm_u32x4 r = smm_philox4x32_10(t, attempt, 0u, (SMM_STREAM_PAIR << 28) | (iter & SMM_ITER_MASK),
                                        (uint32_t)seed_algo, (uint32_t)(seed_algo >> 32));
  const uint64_t v = ((uint64_t)r.x << 32) | (uint64_t)r.y;
#if defined(__CUDA_ARCH__)
  return (uint32_t)__umul64hi(v, (uint64_t)n_pairs);
#else
  return (uint32_t)(((unsigned __int128)v * (unsigned __int128)n_pairs) >> 64);
#endif
}

/* rank q (0-based) -> pair (i<j), 0-based, in the order of the reference's comprehension
 * `[(i,j) for i in 1:N, j in 1:N if i<j]` (AlgoBGP.jl:653): i runs fastest, so rank = j(j-1)/2 + i */
SMM_HD void smm_pair_unrank(uint32_t q, uint32_t *i, uint32_t *j) {
  /* j = floor((1+sqrt(1+8q))/2), fixed up with integer arithmetic */
  uint32_t jj = (uint32_t)((1.0 + SMM_SQRT((double)(1.0 + 8.0 * (double)q))) * 0.5);
  while ((uint64_t)jj * (jj - 1u) / 2u > q) --jj;
  while ((uint64_t)(jj + 1u) * jj / 2u <= q) ++jj;
  *j = jj;
  *i = q - (uint32_t)((uint64_t)jj * (jj - 1u) / 2u);
}

#endif /* SMM_STREAM_H */
