# SMMStreams.jl -- Julia port of include/smm_stream.h: the counter-indexed random streams of the B200 BGP path.
#
# Every element of the four streams (Zsim, Zprop, Uacc, Pairs) is Philox4x32-10(seed; indices) followed by a fixed
# Float64 transform that uses only + - * fma / sqrt, so this file, the C header (CPU oracle) and the CUDA kernels
# produce BIT-IDENTICAL doubles.  The numpy re-derivation oracle/oracle_np.py is the model of this port; function
# names follow the header (smm_ prefix dropped).  Constants: julia/smm_stream_tables.jl (generated from the header's
# own literals by tools/gen_julia_tables.py).
#
# NOT EXECUTED in the build environment (no Julia in the image).  julia/parity_harness.jl `--check-port` compares
# every stream of this file against the values dumped by tools/dump_streams.py from the C implementation.
module SMMStreams

export philox4x32_10, u01, neglog01, normal_pair, exp_neg, zig_select, zig_normal, zig_triple,
       sim_block, sim_normals, prop_normal, acc_uniform, pair_sample, pair_unrank

include(joinpath(@__DIR__, "smm_stream_tables.jl"))

const STREAM_SIM = 0x00000001
const STREAM_PROP = 0x00000002
const STREAM_ACC = 0x00000003
const STREAM_PAIR = 0x00000004
const ITER_MASK = 0x0fffffff

const PHILOX_M0 = 0xD2511F53
const PHILOX_M1 = 0xCD9E8D57
const PHILOX_W0 = 0x9E3779B9
const PHILOX_W1 = 0xBB67AE85

# ---- Philox4x32-10 (smm_stream.h: smm_philox4x32_10) --------------------------------------------------------------
@inline function mulhilo(a::UInt32, b::UInt32)
    p = UInt64(a) * UInt64(b)
    return (p >> 32) % UInt32, p % UInt32          # hi, lo
end

function philox4x32_10(c0::UInt32, c1::UInt32, c2::UInt32, c3::UInt32, k0::UInt32, k1::UInt32)
    for _ in 1:10
        hi0, lo0 = mulhilo(PHILOX_M0, c0)
        hi1, lo1 = mulhilo(PHILOX_M1, c2)
        c0, c1, c2, c3 = hi1 ⊻ c1 ⊻ k0, lo1, hi0 ⊻ c3 ⊻ k1, lo0
        k0 += PHILOX_W0                               # UInt32 arithmetic wraps
        k1 += PHILOX_W1
    end
    return (c0, c1, c2, c3)
end
philox(c0, c1, c2, c3, seed::UInt64) =
    philox4x32_10(c0 % UInt32, c1 % UInt32, c2 % UInt32, c3 % UInt32, seed % UInt32, (seed >> 32) % UInt32)

# ---- bit helpers ----------------------------------------------------------------------------------------------------
@inline bits2f(b::UInt64) = reinterpret(Float64, b)
@inline f2bits(d::Float64) = reinterpret(UInt64, d)
@inline mant52(a::UInt32, b::UInt32) = (UInt64(a) << 20) | UInt64(b >> 12)
"uniform on [0,1) with 52-bit resolution (smm_u01)"
@inline u01(a::UInt32, b::UInt32) = bits2f(0x3FF0000000000000 | mant52(a, b)) - 1.0
"u in [2^-52, 1 - 2^-52] (smm_u01_open)"
@inline u01_open(a::UInt32, b::UInt32) = 2.0 - bits2f(0x3FF0000000000000 | mant52(a, b) | 0x0000000000000001)
"(double)u, exactly (smm_u32_to_double)"
@inline u32_to_double(u::UInt32) = bits2f(0x4330000000000000 | UInt64(u)) - 4503599627370496.0

# ---- -log(u), u in (0,1) (smm_neglog01) ------------------------------------------------------------------------------
function neglog01(u::Float64)
    b = f2bits(u)
    hi = (b >> 32) % UInt32
    e = Int((hi >> 20) & 0x000007ff) - 1023
    j = Int((hi >> (20 - SMM_LOG_BITS)) & ((UInt32(1) << SMM_LOG_BITS) - UInt32(1)))
    m = bits2f((b & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000)
    r = fma(m, SMM_LOG_INV[j + 1], -1.0)
    p = SMM_LOGQ_COEFS[SMM_LOGQ_DEG + 1]
    for i in (SMM_LOGQ_DEG - 1):-1:0
        p = fma(p, r, SMM_LOGQ_COEFS[i + 1])
    end
    r2 = r * r
    nl1p = fma(r2, p, -r)
    base = fma(Float64(e), SMM_NLN2, SMM_LOG_NLNC[j + 1])
    return base + nl1p
end

# ---- Box-Muller on one Philox block (smm_normal_pair_tab) -----------------------------------------------------------
function normal_pair(r::NTuple{4,UInt32})
    x, y, z, w = r
    d1 = bits2f(0x3FF0000000000000 | mant52(x, y) | 0x0000000000000001)
    u1 = 2.0 - d1
    rad = sqrt(neglog01(u1))
    rem = mant52(z, w) & ((UInt64(1) << 49) - UInt64(1))
    g = bits2f(0x3FF0000000000000 | (rem << 3)) - 1.0
    wv = g * g
    ps = SMM_SIN_COEFS[SMM_SIN_DEG + 1]
    for i in (SMM_SIN_DEG - 1):-1:0
        ps = fma(ps, wv, SMM_SIN_COEFS[i + 1])
    end
    pc = SMM_COS_COEFS[SMM_COS_DEG + 1]
    for i in (SMM_COS_DEG - 1):-1:0
        pc = fma(pc, wv, SMM_COS_COEFS[i + 1])
    end
    sn = g * ps
    swap = ((z >> 29) & 0x00000001) != 0
    cb = f2bits(swap ? sn : pc)
    sb = f2bits(swap ? pc : sn)
    cb ⊻= UInt64((z << 1) & 0x80000000) << 32          # bit 30 of z -> sign of z0
    sb ⊻= UInt64(z & 0x80000000) << 32                 # bit 31 of z -> sign of z1
    return rad * bits2f(cb), rad * bits2f(sb)
end

# ---- ziggurat (smm_exp_neg, smm_zig_select, smm_zig_fast, smm_zig_slow) ---------------------------------------------
const ZIG_TAG = 0x5A494732
const ZIG_KEY0 = 0x736D6D5A
const ZIG_KEY1 = 0x69676733
const ZIG_MAX_AUX = 0x00001000
const ZIG_SEL_MASK = (UInt32(2) << SMM_ZIG_LAYER_BITS) - UInt32(1)
const LAYER_MASK = UInt32(SMM_ZIG_LAYERS - 1)

function exp_neg(t::Float64)
    shift = 6755399441055744.0
    nf = fma(t, SMM_LOG2E, shift) - shift
    r = fma(nf, -SMM_LN2_HI, t)
    r = fma(nf, -SMM_LN2_LO, r)
    p = SMM_EXP_COEFS[SMM_EXP_DEG + 1]
    for i in (SMM_EXP_DEG - 1):-1:0
        p = fma(p, r, SMM_EXP_COEFS[i + 1])
    end
    n = Int(nf)                                          # nf is integral by construction
    return p * bits2f(UInt64(n + 1023) << 52)
end

"select field (sign, layer) of draw t = 0, 1, 2 in the block's fourth word"
function zig_select(w::UInt32, t::Int)
    v = t == 0 ? (w >> 3) : t == 1 ? (w >> 13) : ((w >> 23) | (w << 9))
    return v & ZIG_SEL_MASK
end

@inline withsign(x::Float64, s::UInt32) = bits2f(f2bits(x) | (UInt64(s >> SMM_ZIG_LAYER_BITS) << 63))

function zig_slow(u0::UInt32, s0::UInt32)
    u, s, n = u0, s0, UInt32(0)
    while true
        i = Int(s & LAYER_MASK)
        e = SMM_ZIG_TABLE[i + 1]
        x = u32_to_double(u) * bits2f(e)
        u < ((e % UInt32) << 20) && return withsign(x, s)
        if i == 0
            x < SMM_ZIG_R && return withsign(x, s)                         # base strip
            while true                                                       # tail beyond R
                n += UInt32(1)
                r = philox4x32_10(u0, s0, n, ZIG_TAG, ZIG_KEY0, ZIG_KEY1)
                xt = neglog01(u01_open(r[1], r[2])) * SMM_ZIG_RINV
                yt = neglog01(u01_open(r[3], r[4]))
                if (yt + yt > xt * xt) || n >= ZIG_MAX_AUX
                    return withsign(SMM_ZIG_R + xt, s)
                end
            end
        end
        n += UInt32(1)
        r = philox4x32_10(u0, s0, n, ZIG_TAG, ZIG_KEY0, ZIG_KEY1)
        f_lo, f_hi = SMM_ZIG_F[i + 1], SMM_ZIG_F[i + 2]
        y = fma(u01(r[3], r[4]), f_hi - f_lo, f_lo)
        if y < exp_neg((x * x) * -0.5) || n >= ZIG_MAX_AUX
            return withsign(x, s)
        end
        u = r[1]                                                             # rejected: fresh candidate
        s = r[2] & ZIG_SEL_MASK
    end
end

function zig_normal(u::UInt32, s::UInt32)
    e = SMM_ZIG_TABLE[Int(s & LAYER_MASK) + 1]
    if u < ((e % UInt32) << 20)                                              # fast path (99.2 % of draws)
        return withsign(u32_to_double(u) * bits2f(e), s)
    end
    return zig_slow(u, s)
end

"the three normals of one Philox block (smm_zig_triple)"
zig_triple(r::NTuple{4,UInt32}) =
    (zig_normal(r[1], zig_select(r[4], 0)), zig_normal(r[2], zig_select(r[4], 1)), zig_normal(r[3], zig_select(r[4], 2)))

# ---- the four streams ---------------------------------------------------------------------------------------------------
"block j of Zsim[k, .] (smm_sim_block); k, j, uid 0-based"
function sim_block(seed_sim::UInt64, j::Integer, k::Integer; noseed::Bool = false, uid::Integer = 0, rep::Integer = 0)
    c2 = noseed ? (uid % UInt32) : UInt32(0)
    c3 = (STREAM_SIM << 28) | (noseed ? ((rep % UInt32) & ITER_MASK) : UInt32(0))
    return philox(j, k, c2, c3, seed_sim)
end

"""
    sim_normals(seed_sim, k, S; noseed=false, uid=0, rep=0, transform=:zig)

Zsim[k, 0:S-1] (row k 0-based): ziggurat, three per block, for the MvNormal objectives; Box-Muller, two per block,
for the dynamic panel (`transform = :bm`)  -- oracle/smm_oracle.cpp::fill_normals_row.
"""
function sim_normals(seed_sim::UInt64, k::Integer, S::Integer; noseed::Bool = false, uid::Integer = 0, rep::Integer = 0,
                     transform::Symbol = :zig)
    out = Vector{Float64}(undef, S)
    if transform == :zig
        j = 0
        while 3j < S
            z = zig_triple(sim_block(seed_sim, j, k; noseed = noseed, uid = uid, rep = rep))
            for t in 0:2
                3j + t < S && (out[3j + t + 1] = z[t + 1])
            end
            j += 1
        end
    else
        j = 0
        while 2j < S
            z0, z1 = normal_pair(sim_block(seed_sim, j, k; noseed = noseed, uid = uid, rep = rep))
            out[2j + 1] = z0
            2j + 1 < S && (out[2j + 2] = z1)
            j += 1
        end
    end
    return out
end

"Zprop[chain, iter, attempt, k]: chain, attempt, k 0-based; iter 1-based (smm_prop_block)"
function prop_normal(seed_algo::UInt64, chain::Integer, iter::Integer, attempt::Integer, k::Integer)
    r = philox(attempt, k >> 1, chain, (STREAM_PROP << 28) | ((iter % UInt32) & ITER_MASK), seed_algo)
    z0, z1 = normal_pair(r)
    return (k & 1) == 1 ? z1 : z0
end

"Uacc[chain, iter] in [0,1): chain 0-based, iter 1-based (smm_acc_uniform) -- BGPChain.probs_acc"
function acc_uniform(seed_algo::UInt64, chain::Integer, iter::Integer)
    r = philox(0, 0, chain, (STREAM_ACC << 28) | ((iter % UInt32) & ITER_MASK), seed_algo)
    return u01(r[1], r[2])
end

"rank q (0-based) -> pair (i < j), 0-based, i fastest as in `[(i,j) for i in 1:N, j in 1:N if i<j]` (AlgoBGP.jl:653)"
function pair_unrank(q::Integer)
    jj = floor(Int, (1.0 + sqrt(1.0 + 8.0 * Float64(q))) * 0.5)
    while jj * (jj - 1) ÷ 2 > q
        jj -= 1
    end
    while (jj + 1) * jj ÷ 2 <= q
        jj += 1
    end
    return q - jj * (jj - 1) ÷ 2, jj
end

"Pairs[iter]: the ordered list of 1-based (i, j) tuples `sample(props, N, replace=false)` is replaced by"
function pair_sample(seed_algo::UInt64, iter::Integer, N::Integer)
    n_all = UInt128(N * (N - 1) ÷ 2)
    n_s = N < 3 ? N - 1 : N
    chosen = Int[]
    for t in 0:(n_s - 1)
        a = 0
        q = 0
        while true
            r = philox(t, a, 0, (STREAM_PAIR << 28) | ((iter % UInt32) & ITER_MASK), seed_algo)
            v = (UInt64(r[1]) << 32) | UInt64(r[2])
            q = Int((UInt128(v) * n_all) >> 64)
            a += 1
            q in chosen || break
        end
        push!(chosen, q)
    end
    return [(p[1] + 1, p[2] + 1) for p in pair_unrank.(chosen)]
end

end # module
