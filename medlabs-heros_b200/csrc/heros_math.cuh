// heros_math.cuh -- scalar building blocks shared by every kernel of the engine (and by the few host-side
// set-up routines of the library).  All floating point is IEEE binary64 with individually rounded operations:
// the library is compiled with -fmad=false so that a*b+c is never contracted, which makes + - * / sqrt bit-identical
// to the JVM's strictfp-style arithmetic the reference runs on.
//
// Reference formulas (line numbers in /root/reference/src/main/java/eu/heros/disease/):
//   Covid19TransmissionArea.java:155-172,182      area model
//   Covid19TransmissionDistance.java:137-143,155-165,173   distance model
//   Covid19Progression.java:182-346               phase state machine
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace heros {

// disease phases in the reference's registration order (Covid19Progression.java:130-137)
enum : uint32_t { PH_S = 0, PH_E = 1, PH_IA = 2, PH_IS = 3, PH_H = 4, PH_ICU = 5, PH_D = 6, PH_R = 7, N_PHASE = 8 };

HD bool phase_is_ill(uint32_t ph) { return ph - 1u < 5u; }        // DiseaseState.ILL: E, I(A), I(S), H, ICU
HD bool phase_is_susceptible(uint32_t ph) { return ph == PH_S; }  // DiseaseState.SUSCEPTIBLE

// ---- bit casts ------------------------------------------------------------------------------------------------
HD uint64_t f64_bits(double x)
{
#ifdef __CUDA_ARCH__
    return (uint64_t) __double_as_longlong(x);
#else
    uint64_t b; __builtin_memcpy(&b, &x, 8); return b;
#endif
}
HD double bits_f64(uint64_t b)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long) b);
#else
    double x; __builtin_memcpy(&x, &b, 8); return x;
#endif
}
HD uint32_t f32_bits(float x)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    uint32_t b; __builtin_memcpy(&b, &x, 4); return b;
#endif
}
HD float bits_f32(uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float x; __builtin_memcpy(&x, &b, 4); return x;
#endif
}

// ---- counter-based random numbers: Philox4x32-10 ----------------------------------------------------------------
// key = the 64-bit scenario seed, counter = (person, event-time bits lo, hi, stream tag): every draw of the model is
// a pure function of (seed, person, event) and never of the order in which threads, windows or GPUs get to it.
struct Philox4 { uint32_t x, y, z, w; };

HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t) a * b;
    lo = (uint32_t) p;
    hi = (uint32_t) (p >> 32);
#endif
}

HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r)
    {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c0, h0, l0);
        mulhilo(0xCD9E8D57u, c2, h1, l1);
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

constexpr uint32_t TAG_TRANSMIT = 0u;      // infection draw, keyed by the evaluation time
constexpr uint32_t TAG_SEEDING = 0x200u;   // initial-infection selection hash
HD uint32_t tag_progression(uint32_t phase_entered, uint32_t draw) { return 0x100u | (phase_entered << 4) | draw; }

// U(0,1) in [0,1) with 53 random bits (replaces model.getU01().draw())
HD double u01(uint64_t seed, uint32_t person, uint64_t bits, uint32_t tag)
{
    Philox4 r = philox4x32_10(person, (uint32_t) bits, (uint32_t) (bits >> 32), tag, (uint32_t) seed,
                              (uint32_t) (seed >> 32));
    uint64_t m = (((uint64_t) r.x << 32) | r.y) >> 11;
    return (double) m * 0x1.0p-53;
}

// ---- exp: argument reduction x = k ln2 + r and a degree-5 minimax correction, the classic fdlibm scheme that
// java.lang.StrictMath.exp specifies; < 1 ulp.  Written with plain IEEE operations only, so host, device and the
// CPU oracle agree to the last bit.
HD double hexp(double x)
{
    const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
                 P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    const uint64_t bx = f64_bits(x);
    const uint32_t hx = (uint32_t) (bx >> 32) & 0x7fffffffu;
    const bool neg = (bx >> 63) != 0;
    double hi = 0.0, lo = 0.0;
    int k = 0;
    if (hx >= 0x40862E42u)
    {
        if (hx >= 0x7ff00000u)
        {
            if ((bx & 0x000fffffffffffffull) != 0) return x + x;
            return neg ? 0.0 : x;
        }
        if (x > 7.09782712893383973096e+02) return bits_f64(0x7ff0000000000000ull);
        if (x < -7.45133219101941108420e+02) return 0.0;
    }
    if (hx > 0x3fd62e42u)
    {
        if (hx < 0x3FF0A2B2u)
        {
            if (neg) { hi = x + ln2hi; lo = -ln2lo; k = -1; }
            else     { hi = x - ln2hi; lo = ln2lo;  k = 1; }
        }
        else
        {
            k = (int) (invln2 * x + (neg ? -0.5 : 0.5));
            const double t = (double) k;
            hi = x - t * ln2hi;
            lo = t * ln2lo;
        }
        x = hi - lo;
    }
    else if (hx < 0x3e300000u)
        return 1.0 + x;
    const double t = x * x;
    const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return bits_f64(f64_bits(y) + ((uint64_t) (int64_t) k << 52));
    return bits_f64(f64_bits(y) + ((uint64_t) (int64_t) (k + 1000) << 52)) * 0x1.0p-1000;
}

HD double hsqrt(double x)
{
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}

// java.lang.Math.max: NaN if either operand is NaN (Covid19TransmissionDistance.java:142)
HD double java_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return (f64_bits(a) >> 63) ? b : a;
    return a > b ? a : b;
}

// ---- order-independent summation: 128-bit fixed point (64 fraction bits below 2^-16, i.e. a resolution of 2^-80).
// The exposure sum of an evaluation runs over the ill occupants of a sub-location, which the bucketing delivers in
// ticket order (any order).  Instead of sorting them into a canonical order before every sum, the large sub-locations
// add the terms as integers: exact, associative and commutative, so the sum does not depend on the order, the thread
// layout or the sharding.  A term keeps all its 53 bits down to 2^-27 and is off by less than 2^-80 below that -- far
// inside the 1e-12 relative tolerance of the probabilities.  Terms of 2^40 and more (or non-finite ones: NaN areas)
// do not fit and are added to a plain double beside it (their result is +-inf / NaN whatever the order).
struct ExactSum
{
    long long hi;           // units of 2^-16
    unsigned long long lo;  // units of 2^-80
    double loose;           // terms that do not fit the fixed-point range
};
HD ExactSum exact_zero() { return ExactSum{0, 0ull, 0.0}; }
HD void exact_add(ExactSum& a, double t)
{
    if (!(t > -0x1p40 && t < 0x1p40)) { a.loose += t; return; }
    const double s = t * 65536.0;  // exact
#ifdef __CUDA_ARCH__
    const double f = floor(s);
#else
    const double f = __builtin_floor(s);
#endif
    const unsigned long long lo = (unsigned long long) ((s - f) * 18446744073709551616.0);  // (s - f) in [0, 1), exact
    const unsigned long long sum = a.lo + lo;
    a.hi += (long long) f + (sum < lo ? 1 : 0);
    a.lo = sum;
}
HD void exact_merge(ExactSum& a, long long hi, unsigned long long lo, double loose)
{
    const unsigned long long sum = a.lo + lo;
    a.hi += hi + (sum < lo ? 1 : 0);
    a.lo = sum;
    a.loose += loose;
}
HD double exact_value(const ExactSum& a)
{
    return ((double) a.hi * 0x1.0p-16 + (double) a.lo * 0x1.0p-80) + a.loose;
}

// ---- model parameters (already converted to hours, Covid19TransmissionArea.java:63-68 / Distance:59-68) ----------
struct AreaParams { double contagiousness, beta, t_e_min, t_e_mode, t_e_max, threshold; };
struct DistParams { double L, I, C, v_max, v_0, r, psi, alpha, mu, threshold; };

// infectiousness triangle of one ill occupant, sub-location branch (Covid19TransmissionArea.java:167-172)
HD double area_contribution(const AreaParams& a, double te)
{
    if (te >= a.t_e_min && te < a.t_e_mode) return (te - a.t_e_min) / (a.t_e_mode - a.t_e_min);
    if (te >= a.t_e_mode && te <= a.t_e_max) return (a.t_e_max - te) / (a.t_e_max - a.t_e_mode);
    return 0.0;
}
// ... and its total-location variant, sign and shape quirks included (Covid19TransmissionArea.java:216-221)
HD double area_contribution_total(const AreaParams& a, double te)
{
    if (te >= a.t_e_min && te < a.t_e_mode) return te / (a.t_e_mode - a.t_e_min);
    if (te >= a.t_e_mode && te <= a.t_e_max) return 1.0 - te / (a.t_e_max - a.t_e_mode);
    return 0.0;
}
// factor of the sub-location branch; denom = correctionFactorArea * (totalSurface / nSub) (TArea:155-156)
HD double area_factor(const AreaParams& a, double duration, double denom)
{
    return -a.beta * a.contagiousness * duration / denom;
}
HD double area_factor_total(const AreaParams& a, double duration, double denom)  // TArea:205, no minus sign
{
    return a.beta * a.contagiousness * duration / denom;
}

// viral load of one ill occupant (Covid19TransmissionDistance.java:154-159); > 0 <=> contagious
HD double dist_viral_load(const DistParams& d, double t)
{
    if (t >= d.L && t < d.I) return d.v_max * ((t - d.L) / (d.I - d.L));
    if (t >= d.I && t < d.I + d.C) return d.v_max * ((d.I + d.C - t) / d.C);
    return 0.0;
}
// sigma(max(Delta, psi)) * alpha * (1-mu)^2 (Covid19TransmissionDistance.java:138-143); area already divided
HD double dist_factor(const DistParams& d, double psi, double mu, double area, int head_count)
{
    const double Delta = hsqrt(area / (double) head_count);
    const double sigma = 1.0 - 1.0 / (1.0 + hexp(-3.0 * (java_max(Delta, psi) - 1.5)));
    return sigma * d.alpha * (1.0 - mu) * (1.0 - mu);
}
// one summand of TDist:163-164
HD double dist_term(const DistParams& d, double factor, double duration, double v_t)
{
    const double Pt = 1 / (1 + hexp(-d.r * (v_t - d.v_0)));
    return factor * duration * Pt;
}

// ---- progression parameters (Covid19Progression.java:143-173) ---------------------------------------------------
enum : int { F_ASYMPTOMATIC = 0, F_S2H = 1, F_H2ICU = 2, F_H2D = 3, F_ICU2D = 4, N_FRACTION = 5 };
enum : int { P_INC_A = 0, P_INC_S = 1, P_A2R = 2, P_S2H = 3, P_S2R = 4, P_H2ICU = 5, P_H2D = 6, P_H2R = 7,
             P_ICU2D = 8, P_ICU2R = 9, N_PERIOD = 10 };
enum : int { DIST_CONSTANT = 0, DIST_UNIFORM = 1, DIST_TRIANGULAR = 2 };

struct Duration { int32_t kind; int32_t pad; double a, b, c; };  // days; Triangular(a = min, b = mode, c = max)

struct ProgParams
{
    double fraction[N_FRACTION][2][128];  // ConditionalProbability tabulated by [female][age]
    Duration period[N_PERIOD];
};

// DurationDistribution(dist, TimeUnit.DAY).getDuration(): a draw in days, returned in hours
HD double sample_duration_hours(const Duration& d, double u)
{
    double days;
    if (d.kind == DIST_TRIANGULAR)
    {
        const double mn = d.a, mode = d.b, mx = d.c;
        if (u <= (mode - mn) / (mx - mn)) days = mn + hsqrt((mode - mn) * (mx - mn) * u);
        else days = mx - hsqrt((mx - mn) * (mx - mode) * (1.0 - u));
    }
    else if (d.kind == DIST_UNIFORM)
        days = d.a + (d.b - d.a) * u;
    else
        days = d.a;
    return days * 24.0;
}

// ---- packed per-agent word ("view") --------------------------------------------------------------------------------
// bits  0..3  phase             current disease phase
//       4..7  tphase            phase at the start of the window being processed (what transmission sees)
//       8..11 next_phase        phase entered at next_time
//       12    female
//       13    ends              the agent stops being ill inside the current window, at ill_end[agent]
//       16..23 age
//       24..31 role
//       32..63 exposure time, float bits (ConstructHerosModel.java:745: exposure time is a float)
HD uint32_t view_phase(uint64_t v) { return (uint32_t) v & 0xFu; }
HD uint32_t view_tphase(uint64_t v) { return ((uint32_t) v >> 4) & 0xFu; }
HD uint32_t view_next_phase(uint64_t v) { return ((uint32_t) v >> 8) & 0xFu; }
HD uint32_t view_female(uint64_t v) { return ((uint32_t) v >> 12) & 1u; }
HD uint32_t view_ends(uint64_t v) { return ((uint32_t) v >> 13) & 1u; }
HD uint32_t view_age(uint64_t v) { return ((uint32_t) v >> 16) & 0xFFu; }
HD uint32_t view_role(uint64_t v) { return ((uint32_t) v >> 24) & 0xFFu; }
HD float view_exposure(uint64_t v) { return bits_f32((uint32_t) (v >> 32)); }
HD uint64_t view_set(uint64_t v, uint32_t shift, uint32_t mask, uint32_t val)
{
    return (v & ~((uint64_t) mask << shift)) | ((uint64_t) (val & mask) << shift);
}
HD uint64_t view_set_exposure(uint64_t v, float e) { return (v & 0xFFFFFFFFull) | ((uint64_t) f32_bits(e) << 32); }

// One step of the state machine: the agent has just entered `phase` at time `now`; returns the next phase and the
// absolute time of the next transition (Covid19Progression.java:189-199, 227-313).  Terminal phases return +inf.
HD void progression_next(const ProgParams& pp, uint64_t seed, uint32_t person, uint32_t age, uint32_t female,
                         uint32_t phase, double now, uint32_t& next_phase, double& next_time)
{
    const double(*fr)[2][128] = pp.fraction;
    const uint32_t a = age > 127u ? 127u : age;
    int period;
    switch (phase)
    {
    case PH_E:
        if (u01(seed, person, 0, tag_progression(PH_E, 0)) < fr[F_ASYMPTOMATIC][female][a]) { next_phase = PH_IA; period = P_INC_A; }
        else { next_phase = PH_IS; period = P_INC_S; }
        break;
    case PH_IA:
        next_phase = PH_R; period = P_A2R;
        break;
    case PH_IS:
        if (u01(seed, person, 0, tag_progression(PH_IS, 0)) < fr[F_S2H][female][a]) { next_phase = PH_H; period = P_S2H; }
        else { next_phase = PH_R; period = P_S2R; }
        break;
    case PH_H:
        if (u01(seed, person, 0, tag_progression(PH_H, 0)) < fr[F_H2ICU][female][a]) { next_phase = PH_ICU; period = P_H2ICU; }
        else if (u01(seed, person, 0, tag_progression(PH_H, 2)) < fr[F_H2D][female][a]) { next_phase = PH_D; period = P_H2D; }
        else { next_phase = PH_R; period = P_H2R; }
        break;
    case PH_ICU:
        if (u01(seed, person, 0, tag_progression(PH_ICU, 0)) < fr[F_ICU2D][female][a]) { next_phase = PH_D; period = P_ICU2D; }
        else { next_phase = PH_R; period = P_ICU2R; }
        break;
    default:
        next_phase = phase;
        next_time = bits_f64(0x7ff0000000000000ull);
        return;
    }
    const double ud = u01(seed, person, 0, tag_progression(phase, 1));
    next_time = now + sample_duration_hours(pp.period[period], ud);
}

}  // namespace heros
