/*
 * heros_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See heros_oracle.h.
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared  (oracle/Makefile).  -ffp-contract=off keeps every
 * + - * / sqrt an individually rounded IEEE-754 binary64 operation, exactly like the JVM.
 *
 * Citations "TArea:n", "TDist:n", "Prog:n", "DPol:n", "Construct:n" are line numbers in
 *   /root/reference/src/main/java/eu/heros/disease/Covid19TransmissionArea.java
 *   /root/reference/src/main/java/eu/heros/disease/Covid19TransmissionDistance.java
 *   /root/reference/src/main/java/eu/heros/disease/Covid19Progression.java
 *   /root/reference/src/main/java/eu/heros/policy/DiseasePolicy.java
 *   /root/reference/src/main/java/eu/heros/factory/ConstructHerosModel.java
 *
 * Arithmetic that lives in un-vendored dependencies (medlabs:2.1.3, dsol 4.2.1) and is therefore DEFINED here
 * (SURVEY.md 8c): the U(0,1) stream (-> counter-based Philox4x32-10 keyed by seed, person and event time, as the
 * north star prescribes), DistTriangular/DistUniform inverse CDFs, DurationDistribution (days*24 = hours),
 * ConditionalProbability (tabulated by gender and age), the movement driver (contract rows C1-C8).
 *
 * PARITY STATUS: EXACT PARITY UNPINNED BY THE REFERENCE.  The reference has no tests, golden vectors or fixtures for
 * this path and cannot run here (no JVM; medlabs / DSOL jars not vendored), so the exact outputs of this oracle
 * (event lists, probabilities) are checked against nothing the reference produced.  What pins it (tests/test_oracle.py):
 * Random123 Philox known answers, libm within 1 ulp, the Javadoc worked example TArea:112-121, hand-derived boundary
 * cases of the formulas, the parameter files of the reference (tests/golden/params_*.json), and the reference-produced
 * epidemic curves of "seed comparison.xlsx" and "out-flip-lockdown-no-10-20-comparison.xlsx"
 * (tests/golden/seed_comparison.json, policy_comparison.json) for the observable consequences of the driver contract
 * and of the progression: latent period; the empirical CDF of the incubation period of the 100 initial cases
 * (Kolmogorov-Smirnov against the triangular inverse CDF in hours and against this oracle's draws); the asymptomatic
 * split; the lower ends of the sojourn supports (first Hospitalized / ICU / Dead / Recovered); exposures only at
 * movement events, none at night; the share of exposures in the first hour after 06:00 (which tells the area model
 * made those curves); the Sunday dips of days 6 and 13; runs of one seed coinciding until the morning of a policy's
 * day.  These are statistical / structural pins, not bit-exact vectors.
 */
#include "heros_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 constants)     */
/* ------------------------------------------------------------------------------------------------------------ */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round)
    {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
        uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* stream tags (word 3 of the counter) */
#define TAG_TRANSMIT 0u
#define TAG_RATE 0x500u /* draw of a rate-based location (LocationProbBased), keyed by (person, exit time) */
#define ROLE_REFERENCE 7 /* Worker: the reference group of the satellite person types (Construct:303-308) */
#define TAG_PROG(phase, draw) (0x100u | ((uint32_t) (phase) << 4) | (uint32_t) (draw))

/* U(0,1) with 53 random bits, in [0,1): stands in for model.getU01().draw() (TArea:190, TDist:181, Prog:189). */
double orc_u01(uint64_t seed, uint32_t person, uint64_t bits, uint32_t tag)
{
    uint32_t ctr[4] = {person, (uint32_t) bits, (uint32_t) (bits >> 32), tag};
    uint32_t key[2] = {(uint32_t) seed, (uint32_t) (seed >> 32)};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    uint64_t m = (((uint64_t) r[0] << 32) | r[1]) >> 11;
    return (double) m * 0x1.0p-53;
}

static uint64_t dbl_bits(double x) { uint64_t b; memcpy(&b, &x, 8); return b; }
static double bits_dbl(uint64_t b) { double x; memcpy(&x, &b, 8); return x; }

/* ------------------------------------------------------------------------------------------------------------ */
/* exp(): the fdlibm algorithm (= java.lang.StrictMath.exp), restated.  Error < 1 ulp.                           */
/*   x = k*ln2 + r, |r| <= 0.5 ln2;  exp(r) = 1 + r + r*c/(2-c),  c = r - r^2*(P1 + r^2*(P2 + ...))             */
/* ------------------------------------------------------------------------------------------------------------ */

double orc_exp(double x)
{
    static const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10,
                        invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
                        P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                        P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    uint64_t bx = dbl_bits(x);
    uint32_t hx = (uint32_t) (bx >> 32) & 0x7fffffffu;
    int neg = (int) (bx >> 63);
    double hi = 0.0, lo = 0.0;
    int k = 0;

    if (hx >= 0x40862E42u) /* |x| >= 709.78 or non-finite */
    {
        if (hx >= 0x7ff00000u)
        {
            if (((bx & 0x000fffffffffffffull) != 0)) return x + x; /* NaN */
            return neg ? 0.0 : x;                                  /* exp(+-inf) */
        }
        if (x > 7.09782712893383973096e+02) return HUGE_VAL;
        if (x < -7.45133219101941108420e+02) return 0.0;
    }
    if (hx > 0x3fd62e42u) /* |x| > 0.5 ln2 */
    {
        if (hx < 0x3FF0A2B2u) /* |x| < 1.5 ln2 */
        {
            if (neg) { hi = x + ln2hi; lo = -ln2lo; k = -1; }
            else     { hi = x - ln2hi; lo = ln2lo;  k = 1; }
        }
        else
        {
            k = (int) (invln2 * x + (neg ? -0.5 : 0.5));
            double t = (double) k;
            hi = x - t * ln2hi;
            lo = t * ln2lo;
        }
        x = hi - lo;
    }
    else if (hx < 0x3e300000u) /* |x| < 2^-28 */
    {
        return 1.0 + x;
    }
    double t = x * x;
    double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021)
        return bits_dbl(dbl_bits(y) + ((uint64_t) (int64_t) k << 52));
    return bits_dbl(dbl_bits(y) + ((uint64_t) (int64_t) (k + 1000) << 52)) * 0x1.0p-1000;
}

/* java.lang.Math.max(double,double): NaN if either argument is NaN (TDist:142) */
double orc_java_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return signbit(a) ? b : a;
    return a > b ? a : b;
}

/* DurationDistribution(dist, TimeUnit.DAY).getDuration(): days -> hours (contract C2).  Inverse CDFs as in DSOL's
 * DistTriangular / DistUniform / DistConstant. */
double orc_sample_duration_hours(const orc_duration* d, double u)
{
    double days;
    if (d->kind == ORC_DIST_TRIANGULAR)
    {
        double min = d->a, mode = d->b, max = d->c;
        if (u <= (mode - min) / (max - min))
            days = min + sqrt((mode - min) * (max - min) * u);
        else
            days = max - sqrt((max - min) * (max - mode) * (1.0 - u));
    }
    else if (d->kind == ORC_DIST_UNIFORM)
        days = d->a + (d->b - d->a) * u;
    else
        days = d->a;
    return days * 24.0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* infectPeople, restated                                                                                        */
/* ------------------------------------------------------------------------------------------------------------ */

static int is_ill(uint8_t ph) { return ph >= ORC_E && ph <= ORC_ICU; } /* DiseaseState.ILL, Prog:131-135 */
static int is_susceptible(uint8_t ph) { return ph == ORC_S; }          /* Prog:130 */

typedef struct
{
    const int32_t* ids;
    int n;
    const uint8_t* phase;
    const float* exposure;
    int by_id; /* 1: phase[ids[k]] (personMap.get(id)); 0: phase[k] */
} person_set;

#define PH(s, k) ((s)->phase[(s)->by_id ? (s)->ids[k] : (k)])
#define EX(s, k) ((s)->exposure[(s)->by_id ? (s)->ids[k] : (k)])

typedef struct
{
    int32_t* infectious; int32_t n_infectious;
    int32_t* infected;   int32_t n_infected;
    double p;
    int64_t draws;
} record;

/* TArea:133-248.  prm holds the converted parameters of TArea:63-68. */
typedef struct { double contagiousness, beta, t_e_min, t_e_mode, t_e_max, threshold; } area_prm;
typedef struct { double L, I, C, v_max, v_0, r, psi, alpha, mu, threshold; } dist_prm;

static area_prm area_convert(const orc_area_props* p)
{
    area_prm a;
    a.contagiousness = p->contagiousness;              /* TArea:63 */
    a.beta = p->beta;                                  /* TArea:64 */
    a.t_e_min = p->t_e_min_days * 24.0;                /* TArea:65 */
    a.t_e_mode = p->t_e_mode_days * 24.0;              /* TArea:66 */
    a.t_e_max = p->t_e_max_days * 24.0;                /* TArea:67 */
    a.threshold = p->calculation_threshold_s / 3600.0; /* TArea:68 */
    return a;
}

static dist_prm dist_convert(const orc_dist_props* p)
{
    dist_prm d;
    d.L = p->L_days * 24.0; /* TDist:59 */
    d.I = p->I_days * 24.0; /* TDist:60 */
    d.C = p->C_days * 24.0; /* TDist:61 */
    d.v_max = p->v_max; d.v_0 = p->v_0; d.r = p->r; d.psi = p->psi; d.alpha = p->alpha; d.mu = p->mu;
    d.threshold = p->calculation_threshold_s / 3600.0; /* TDist:68 */
    return d;
}

static int infect_area(const area_prm* prm, uint64_t seed, int infect_sub, double corr, float total_surface,
                       int n_sublocations, const person_set* sub, const person_set* all, double now,
                       double duration, record* rec)
{
    rec->n_infectious = 0; rec->n_infected = 0; rec->p = NAN;
    if (duration < prm->threshold) /* TArea:138 */
        return 0;
    double area = (double) total_surface; /* TArea:144, float getTotalSurfaceM2() widened */
    const uint64_t now_bits = dbl_bits(now);

    if (infect_sub || n_sublocations < 2) /* TArea:149 */
    {
        area /= (double) n_sublocations;                                             /* TArea:155 */
        double factor = -prm->beta * prm->contagiousness * duration / (corr * area); /* TArea:156 */
        if (factor == 0.0) return 1;                                                 /* TArea:157 */
        double sum = 0.0;
        for (int k = 0; k < sub->n; ++k) /* TArea:162-177 */
        {
            if (is_ill(PH(sub, k)))
            {
                double te = now - (double) EX(sub, k); /* TArea:167 */
                double contribution = 0.0;
                if (te >= prm->t_e_min && te < prm->t_e_mode)
                    contribution += (te - prm->t_e_min) / (prm->t_e_mode - prm->t_e_min);
                else if (te >= prm->t_e_mode && te <= prm->t_e_max)
                    contribution += (prm->t_e_max - te) / (prm->t_e_max - prm->t_e_mode);
                sum += contribution;
                if (rec->infectious) rec->infectious[rec->n_infectious] = sub->ids[k];
                rec->n_infectious++; /* TArea:175: every ill occupant is listed */
            }
        }
        if (sum == 0.0) return 1;              /* TArea:178 */
        double p = 1.0 - orc_exp(factor * sum); /* TArea:182 */
        rec->p = p;
        for (int k = 0; k < sub->n; ++k) /* TArea:184-195 */
        {
            if (is_susceptible(PH(sub, k)))
            {
                rec->draws++;
                if (orc_u01(seed, (uint32_t) sub->ids[k], now_bits, TAG_TRANSMIT) < p) /* TArea:190 */
                {
                    if (rec->infected) rec->infected[rec->n_infected] = sub->ids[k];
                    rec->n_infected++;
                }
            }
        }
    }
    else
    {
        double factor = prm->beta * prm->contagiousness * duration / (corr * area); /* TArea:205, no minus sign */
        if (factor == 0.0) return 1;
        double sum = 0.0;
        for (int k = 0; k < all->n; ++k) /* TArea:211-226 */
        {
            if (is_ill(PH(all, k)))
            {
                double te = now - (double) EX(all, k);
                double contribution = 0.0;
                if (te >= prm->t_e_min && te < prm->t_e_mode)
                    contribution += te / (prm->t_e_mode - prm->t_e_min); /* TArea:219 */
                else if (te >= prm->t_e_mode && te <= prm->t_e_max)
                    contribution += 1.0 - te / (prm->t_e_max - prm->t_e_mode); /* TArea:221 */
                sum += contribution;
                if (rec->infectious) rec->infectious[rec->n_infectious] = all->ids[k];
                rec->n_infectious++;
            }
        }
        if (sum == 0.0) return 1;
        double p = 1.0 - orc_exp(factor * sum); /* TArea:231 */
        rec->p = p;
        for (int k = 0; k < all->n; ++k) /* TArea:234-245 */
        {
            if (is_susceptible(PH(all, k)))
            {
                rec->draws++;
                if (orc_u01(seed, (uint32_t) all->ids[k], now_bits, TAG_TRANSMIT) < p)
                {
                    if (rec->infected) rec->infected[rec->n_infected] = all->ids[k];
                    rec->n_infected++;
                }
            }
        }
    }
    return 1;
}

/* TDist:115-248 */
static int infect_dist(const dist_prm* prm, uint64_t seed, int infect_sub, float total_surface, int n_sublocations,
                       const person_set* sub, const person_set* all, double now, double duration, record* rec)
{
    rec->n_infectious = 0; rec->n_infected = 0; rec->p = NAN;
    if (duration < prm->threshold) /* TDist:120 */
        return 0;
    double area = (double) total_surface; /* TDist:126 */
    const uint64_t now_bits = dbl_bits(now);
    const person_set* scan;
    if (infect_sub || n_sublocations < 2) /* TDist:131 */
    {
        area /= (double) n_sublocations; /* TDist:137 */
        scan = sub;
    }
    else
        scan = all; /* TDist:196: un-divided area, sub-location head count, all persons scanned */

    double Delta = sqrt(area / (double) sub->n);                                                  /* TDist:138,196 */
    double sigma = 1.0 - 1.0 / (1.0 + orc_exp(-3.0 * (orc_java_max(Delta, prm->psi) - 1.5)));     /* TDist:142,200 */
    double factor = sigma * prm->alpha * (1.0 - prm->mu) * (1.0 - prm->mu);                       /* TDist:143,201 */
    if (factor == 0.0) return 1;                                                                  /* TDist:144 */
    double sum = 0.0;
    for (int k = 0; k < scan->n; ++k) /* TDist:149-168 / 207-226 */
    {
        if (is_ill(PH(scan, k)))
        {
            double v_t = 0.0;
            double t = now - (double) EX(scan, k); /* TDist:155 */
            if (t >= prm->L && t < prm->I)
                v_t = prm->v_max * ((t - prm->L) / (prm->I - prm->L)); /* TDist:157 */
            else if (t >= prm->I && t < prm->I + prm->C)
                v_t = prm->v_max * ((prm->I + prm->C - t) / prm->C); /* TDist:159 */
            if (v_t > 0)
            {
                double Pt = 1 / (1 + orc_exp(-prm->r * (v_t - prm->v_0))); /* TDist:163 */
                sum += factor * duration * Pt;                            /* TDist:164 */
                if (rec->infectious) rec->infectious[rec->n_infectious] = scan->ids[k];
                rec->n_infectious++; /* TDist:165: only persons with v_t > 0 */
            }
        }
    }
    if (sum == 0.0) return 1;      /* TDist:169 */
    double p = 1.0 - orc_exp(-sum); /* TDist:173 */
    rec->p = p;
    for (int k = 0; k < scan->n; ++k) /* TDist:175-186 / 234-245 */
    {
        if (is_susceptible(PH(scan, k)))
        {
            rec->draws++;
            if (orc_u01(seed, (uint32_t) scan->ids[k], now_bits, TAG_TRANSMIT) < p) /* TDist:181 */
            {
                if (rec->infected) rec->infected[rec->n_infected] = scan->ids[k];
                rec->n_infected++;
            }
        }
    }
    return 1;
}

int orc_infect_people_area(const orc_area_props* props, uint64_t seed, int infect_in_sublocation,
                           double correction_factor_area, float total_surface_m2, int n_sublocations, int n_sub,
                           const int32_t* sub_ids, const uint8_t* sub_phase, const float* sub_exposure, int n_all,
                           const int32_t* all_ids, const uint8_t* all_phase, const float* all_exposure, double now,
                           double duration, int32_t* infectious_out, int32_t* n_infectious, int32_t* infected_out,
                           int32_t* n_infected, double* p_out)
{
    area_prm prm = area_convert(props);
    person_set sub = {sub_ids, n_sub, sub_phase, sub_exposure, 0};
    person_set all = {all_ids, n_all, all_phase, all_exposure, 0};
    record rec = {infectious_out, 0, infected_out, 0, NAN, 0};
    int calc = infect_area(&prm, seed, infect_in_sublocation, correction_factor_area, total_surface_m2,
                           n_sublocations, &sub, &all, now, duration, &rec);
    *n_infectious = rec.n_infectious; *n_infected = rec.n_infected; *p_out = rec.p;
    return calc;
}

int orc_infect_people_dist(const orc_dist_props* props, uint64_t seed, int infect_in_sublocation,
                           float total_surface_m2, int n_sublocations, int n_sub, const int32_t* sub_ids,
                           const uint8_t* sub_phase, const float* sub_exposure, int n_all, const int32_t* all_ids,
                           const uint8_t* all_phase, const float* all_exposure, double now, double duration,
                           int32_t* infectious_out, int32_t* n_infectious, int32_t* infected_out,
                           int32_t* n_infected, double* p_out)
{
    dist_prm prm = dist_convert(props);
    person_set sub = {sub_ids, n_sub, sub_phase, sub_exposure, 0};
    person_set all = {all_ids, n_all, all_phase, all_exposure, 0};
    record rec = {infectious_out, 0, infected_out, 0, NAN, 0};
    int calc = infect_dist(&prm, seed, infect_in_sublocation, total_surface_m2, n_sublocations, &sub, &all, now,
                           duration, &rec);
    *n_infectious = rec.n_infectious; *n_infected = rec.n_infected; *p_out = rec.p;
    return calc;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Sequential simulation: the medlabs driver contract of SURVEY.md 3.2                                           */
/* ------------------------------------------------------------------------------------------------------------ */

typedef struct { double t; int32_t person; uint32_t key; uint8_t kind; } move_ev; /* kind: 0 exit, 1 enter */
typedef struct { double t; int32_t person; } heap_ev;
typedef struct { double t; int which; double value; } policy_ev; /* which: 0 psi, 1 mu */
typedef struct { int32_t person, loc; int16_t sub; double t, p; } infection_ev;
typedef struct { int32_t person; uint8_t phase; double t; } transition_ev;

struct orc_sim
{
    int model, trigger;
    uint64_t seed;
    area_prm area;
    dist_prm dist;
    orc_prog_props prog;
    double now;

    int n_types; uint8_t* type_infect_sub; double* type_corr;
    int n_loc; uint8_t* loc_type; int16_t* loc_nsub; float* loc_area; uint32_t* loc_base; uint32_t n_keys;
    int32_t* key_loc;
    int n_person; int8_t* age; uint8_t* female; uint8_t* phase; float* exposure; double* next_time;
    uint8_t* next_phase; int32_t* pos_in_set;
    uint32_t* open_key; double* open_tin; /* the stay a person was reported to be in with t_out = +inf */

    /* occupancy: one growable id array per sub-location key (C5) */
    int32_t** set_ids; int32_t* set_n; int32_t* set_cap; double* last_calc;

    move_ev* ev; int64_t n_ev, cap_ev, ev_head; int ev_sorted;
    heap_ev* heap; int64_t n_heap, cap_heap;
    policy_ev* pol; int n_pol, cap_pol, pol_head;
    infection_ev* inf; int64_t n_inf, cap_inf;
    transition_ev* tr; int64_t n_tr, cap_tr;
    int64_t counts[ORC_NPHASE];
    int64_t n_eval, n_draws;
    int32_t* scratch_a; int32_t* scratch_b; int32_t* scratch_all; int64_t cap_scratch;

    /* rate-based locations (medlabs LocationProbBased, Construct:382-388): see rate_exit() */
    uint8_t* role; int64_t n_ref;
    int64_t* ref_exposed; int64_t n_ref_days; /* persons of the reference group exposed per simulated day */
    uint8_t* loc_rate_kind; double* loc_rate_factor; double* loc_rate; /* kind: 0 none, 1 rate-based */
    double* stay_tin; /* when the person entered the place it is in */
};

orc_sim* orc_sim_create(int model, int trigger, uint64_t seed)
{
    orc_sim* s = (orc_sim*) calloc(1, sizeof(orc_sim));
    s->model = model; s->trigger = trigger; s->seed = seed;
    return s;
}

void orc_sim_destroy(orc_sim* s)
{
    if (!s) return;
    for (uint32_t k = 0; k < s->n_keys; ++k) free(s->set_ids[k]);
    free(s->set_ids); free(s->set_n); free(s->set_cap); free(s->last_calc);
    free(s->type_infect_sub); free(s->type_corr);
    free(s->loc_type); free(s->loc_nsub); free(s->loc_area); free(s->loc_base); free(s->key_loc);
    free(s->age); free(s->female); free(s->phase); free(s->exposure); free(s->next_time); free(s->next_phase);
    free(s->pos_in_set); free(s->open_key); free(s->open_tin);
    free(s->ev); free(s->heap); free(s->pol); free(s->inf); free(s->tr);
    free(s->scratch_a); free(s->scratch_b); free(s->scratch_all);
    free(s->role); free(s->ref_exposed); free(s->loc_rate_kind); free(s->loc_rate_factor); free(s->loc_rate); free(s->stay_tin);
    free(s);
}

void orc_sim_set_area(orc_sim* s, const orc_area_props* p) { s->area = area_convert(p); }
void orc_sim_set_dist(orc_sim* s, const orc_dist_props* p) { s->dist = dist_convert(p); }
void orc_sim_set_prog(orc_sim* s, const orc_prog_props* p) { s->prog = *p; }

void orc_sim_set_location_types(orc_sim* s, int n, const uint8_t* infect_sub, const double* corr)
{
    s->n_types = n;
    s->type_infect_sub = (uint8_t*) malloc(n); memcpy(s->type_infect_sub, infect_sub, n);
    s->type_corr = (double*) malloc(n * sizeof(double)); memcpy(s->type_corr, corr, n * sizeof(double));
}

void orc_sim_set_locations(orc_sim* s, int n, const uint8_t* type, const int16_t* n_sub, const float* total_area)
{
    s->n_loc = n;
    s->loc_type = (uint8_t*) malloc(n); memcpy(s->loc_type, type, n);
    s->loc_nsub = (int16_t*) malloc(n * 2); memcpy(s->loc_nsub, n_sub, n * 2);
    s->loc_area = (float*) malloc(n * 4); memcpy(s->loc_area, total_area, n * 4);
    s->loc_base = (uint32_t*) malloc((n + 1) * 4);
    uint32_t base = 0;
    for (int i = 0; i < n; ++i)
    {
        s->loc_base[i] = base;
        base += (uint32_t) (n_sub[i] > 0 ? n_sub[i] : 1); /* a location with 0 sub-locations still owns one key */
    }
    s->loc_base[n] = base;
    s->n_keys = base;
    s->key_loc = (int32_t*) malloc(base * 4);
    for (int i = 0; i < n; ++i)
        for (uint32_t k = s->loc_base[i]; k < s->loc_base[i + 1]; ++k) s->key_loc[k] = i;
    s->set_ids = (int32_t**) calloc(base, sizeof(int32_t*));
    s->set_n = (int32_t*) calloc(base, 4);
    s->set_cap = (int32_t*) calloc(base, 4);
    s->last_calc = (double*) calloc(base, sizeof(double));
}

void orc_sim_set_persons(orc_sim* s, int n, const int8_t* age, const uint8_t* female)
{
    s->n_person = n;
    s->age = (int8_t*) malloc(n); memcpy(s->age, age, n);
    s->female = (uint8_t*) malloc(n); memcpy(s->female, female, n);
    s->phase = (uint8_t*) calloc(n, 1);            /* Construct:746-747: everybody starts Susceptible */
    s->exposure = (float*) calloc(n, 4);           /* Construct:745: setExposureTime(0.0f) */
    s->next_time = (double*) malloc(n * 8);
    s->next_phase = (uint8_t*) calloc(n, 1);
    s->pos_in_set = (int32_t*) malloc(n * 4);
    s->open_key = (uint32_t*) malloc(n * 4);
    s->open_tin = (double*) malloc(n * 8);
    s->stay_tin = (double*) calloc(n, 8);
    for (int i = 0; i < n; ++i) { s->next_time[i] = INFINITY; s->pos_in_set[i] = -1; s->open_key[i] = 0xFFFFFFFFu; s->open_tin[i] = 0.0; }
    memset(s->counts, 0, sizeof(s->counts));
    s->counts[ORC_S] = n;
}

int orc_sim_set_parameter(orc_sim* s, const char* name, double value, double at_time)
{
    int which;
    if (strcmp(name, "psi") == 0) which = 0;      /* TDist:254 */
    else if (strcmp(name, "mu") == 0) which = 1;  /* TDist:258 */
    else return -1;                               /* TDist:263 / TArea:254 */
    if (s->model != ORC_MODEL_DIST) return -1;    /* TArea.setParameter rejects every name */
    if (s->n_pol == s->cap_pol)
    {
        s->cap_pol = s->cap_pol ? 2 * s->cap_pol : 16;
        s->pol = (policy_ev*) realloc(s->pol, s->cap_pol * sizeof(policy_ev));
    }
    /* keep sorted by time, stable (DSOL fires equal-time events in scheduling order) */
    int i = s->n_pol++;
    while (i > s->pol_head && s->pol[i - 1].t > at_time) { s->pol[i] = s->pol[i - 1]; --i; }
    s->pol[i].t = at_time; s->pol[i].which = which; s->pol[i].value = value;
    return 0;
}

/* --- transition heap (stands in for the DSOL event list; one pending event per person) --- */
static int heap_less(const heap_ev* a, const heap_ev* b) { return a->t < b->t || (a->t == b->t && a->person < b->person); }

static void heap_push(orc_sim* s, double t, int32_t person)
{
    if (s->n_heap == s->cap_heap)
    {
        s->cap_heap = s->cap_heap ? 2 * s->cap_heap : 1024;
        s->heap = (heap_ev*) realloc(s->heap, s->cap_heap * sizeof(heap_ev));
    }
    int64_t i = s->n_heap++;
    heap_ev e = {t, person};
    while (i > 0)
    {
        int64_t parent = (i - 1) / 2;
        if (!heap_less(&e, &s->heap[parent])) break;
        s->heap[i] = s->heap[parent];
        i = parent;
    }
    s->heap[i] = e;
}

static heap_ev heap_pop(orc_sim* s)
{
    heap_ev top = s->heap[0];
    heap_ev last = s->heap[--s->n_heap];
    int64_t i = 0;
    for (;;)
    {
        int64_t c = 2 * i + 1;
        if (c >= s->n_heap) break;
        if (c + 1 < s->n_heap && heap_less(&s->heap[c + 1], &s->heap[c])) ++c;
        if (!heap_less(&s->heap[c], &last)) break;
        s->heap[i] = s->heap[c];
        i = c;
    }
    if (s->n_heap > 0) s->heap[i] = last;
    return top;
}

static double frac(const orc_sim* s, int which, int32_t person)
{
    int a = s->age[person];
    if (a < 0) a = 0;
    return s->prog.fraction[which][s->female[person] ? 1 : 0][a];
}

static void schedule(orc_sim* s, int32_t person, double now, double dt_hours, uint8_t next)
{
    s->next_time[person] = now + dt_hours; /* scheduleEventRel: absolute time = now + dt (Prog:192,232,...) */
    s->next_phase[person] = next;
    heap_push(s, s->next_time[person], person);
}

static void record_transition(orc_sim* s, int32_t person, uint8_t phase, double t)
{
    if (s->n_tr == s->cap_tr)
    {
        s->cap_tr = s->cap_tr ? 2 * s->cap_tr : 1024;
        s->tr = (transition_ev*) realloc(s->tr, s->cap_tr * sizeof(transition_ev));
    }
    s->tr[s->n_tr].person = person; s->tr[s->n_tr].phase = phase; s->tr[s->n_tr].t = t; s->n_tr++;
}

/* Prog:182-201 */
static void expose(orc_sim* s, int32_t person, double now)
{
    if (s->role && s->role[person] == ROLE_REFERENCE)
    {
        int64_t day = (int64_t) floor(now / 24.0);
        if (day >= s->n_ref_days)
        {
            int64_t n = 2 * (day + 1) + 16;
            s->ref_exposed = (int64_t*) realloc(s->ref_exposed, (size_t) n * sizeof(int64_t));
            for (int64_t k = s->n_ref_days; k < n; ++k) s->ref_exposed[k] = 0;
            s->n_ref_days = n;
        }
        s->ref_exposed[day]++;
    }
    s->counts[s->phase[person]]--; /* Prog:184 */
    s->phase[person] = ORC_E;      /* Prog:185 */
    s->counts[ORC_E]++;            /* Prog:186 */
    record_transition(s, person, ORC_E, now);
    double u = orc_u01(s->seed, (uint32_t) person, 0, TAG_PROG(ORC_E, 0));
    double ud = orc_u01(s->seed, (uint32_t) person, 0, TAG_PROG(ORC_E, 1));
    if (u < frac(s, ORC_F_ASYMPTOMATIC, person)) /* Prog:189 */
        schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_INC_A], ud), ORC_IA);
    else
        schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_INC_S], ud), ORC_IS);
}

/* Prog:208-346 */
static void change_phase(orc_sim* s, int32_t person, uint8_t next, double now)
{
    s->counts[s->phase[person]]--; /* Prog:211 */
    s->next_time[person] = INFINITY;
    const uint32_t id = (uint32_t) person;
    if (next == ORC_E)
    {
        /* Prog:217-221: expose() decrements the old phase a second time.  Unreachable in normal flow; the
         * quirk is kept by re-incrementing nothing here. */
        expose(s, person, now);
        return;
    }
    s->phase[person] = next;
    s->counts[next]++;
    record_transition(s, person, next, now);
    double ud = orc_u01(s->seed, id, 0, TAG_PROG(next, 1));
    switch (next)
    {
    case ORC_IA: /* Prog:227-235 */
        schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_A2R], ud), ORC_R);
        break;
    case ORC_IS: /* Prog:241-253 */
        if (orc_u01(s->seed, id, 0, TAG_PROG(ORC_IS, 0)) < frac(s, ORC_F_S2H, person))
            schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_S2H], ud), ORC_H);
        else
            schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_S2R], ud), ORC_R);
        break;
    case ORC_H: /* Prog:259-288 */
        if (orc_u01(s->seed, id, 0, TAG_PROG(ORC_H, 0)) < frac(s, ORC_F_H2ICU, person))
            schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_H2ICU], ud), ORC_ICU);
        else if (orc_u01(s->seed, id, 0, TAG_PROG(ORC_H, 2)) < frac(s, ORC_F_H2D, person))
            schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_H2D], ud), ORC_D);
        else
            schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_H2R], ud), ORC_R);
        break;
    case ORC_ICU: /* Prog:294-313 */
        if (orc_u01(s->seed, id, 0, TAG_PROG(ORC_ICU, 0)) < frac(s, ORC_F_ICU2D, person))
            schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_ICU2D], ud), ORC_D);
        else
            schedule(s, person, now, orc_sample_duration_hours(&s->prog.period[ORC_P_ICU2R], ud), ORC_R);
        break;
    default: /* Recovered Prog:319-324, Dead Prog:330-336: terminal */
        break;
    }
}

void orc_sim_expose(orc_sim* s, int n, const int32_t* person, const double* at_time)
{
    for (int i = 0; i < n; ++i)
    {
        if (!is_susceptible(s->phase[person[i]])) continue; /* Construct:1119 */
        s->exposure[person[i]] = (float) at_time[i];        /* contract C4 (0.0f for the t=0 seeds) */
        expose(s, person[i], at_time[i]);
    }
}

void orc_sim_add_visits(orc_sim* s, int64_t n, const int32_t* person, const int32_t* loc, const int16_t* sub,
                        const double* t_in, const double* t_out)
{
    if (s->n_ev + 2 * n > s->cap_ev)
    {
        s->cap_ev = 2 * (s->n_ev + 2 * n);
        s->ev = (move_ev*) realloc(s->ev, s->cap_ev * sizeof(move_ev));
    }
    for (int64_t i = 0; i < n; ++i)
    {
        if (!(t_out[i] > t_in[i])) continue; /* a stay must have positive length */
        uint32_t key = s->loc_base[loc[i]] + (uint32_t) sub[i];
        if (t_out[i] < INFINITY)
        {
            /* a stay reported earlier as still open (same person, place and t_in) is being closed: its enter event
             * already exists, only the exit is new (the contract of heros_submit_visits) */
            if (s->open_key[person[i]] == key && s->open_tin[person[i]] == t_in[i])
                s->open_key[person[i]] = 0xFFFFFFFFu;
            else
            {
                move_ev a = {t_in[i], person[i], key, 1};
                s->ev[s->n_ev++] = a;
            }
            move_ev b = {t_out[i], person[i], key, 0};
            s->ev[s->n_ev++] = b;
        }
    }
    for (int64_t i = 0; i < n; ++i)
    {
        if (!(t_out[i] > t_in[i]) || t_out[i] < INFINITY) continue;
        uint32_t key = s->loc_base[loc[i]] + (uint32_t) sub[i];
        move_ev a = {t_in[i], person[i], key, 1};
        s->ev[s->n_ev++] = a;
        s->open_key[person[i]] = key;
        s->open_tin[person[i]] = t_in[i];
    }
    s->ev_sorted = 0;
}

static int move_cmp(const void* pa, const void* pb)
{
    const move_ev* a = (const move_ev*) pa;
    const move_ev* b = (const move_ev*) pb;
    if (a->t != b->t) return a->t < b->t ? -1 : 1;
    if (a->kind != b->kind) return a->kind < b->kind ? -1 : 1; /* exits before enters at equal time */
    /* movements at an identical time: by sub-location, then by person.  The order only matters for the total-location
     * branch, where evaluations of different sub-locations of one location draw for the same persons. */
    if (a->key != b->key) return a->key < b->key ? -1 : 1;
    if (a->person != b->person) return a->person < b->person ? -1 : 1;
    return 0;
}

static void set_add(orc_sim* s, uint32_t key, int32_t person)
{
    if (s->set_n[key] == s->set_cap[key])
    {
        s->set_cap[key] = s->set_cap[key] ? 2 * s->set_cap[key] : 4;
        s->set_ids[key] = (int32_t*) realloc(s->set_ids[key], s->set_cap[key] * 4);
    }
    s->pos_in_set[person] = s->set_n[key];
    s->set_ids[key][s->set_n[key]++] = person;
}

static void set_remove(orc_sim* s, uint32_t key, int32_t person)
{
    int32_t pos = s->pos_in_set[person];
    if (pos < 0 || pos >= s->set_n[key] || s->set_ids[key][pos] != person)
    {
        pos = -1; /* a person with overlapping stays: fall back to a scan */
        for (int32_t k = 0; k < s->set_n[key]; ++k)
            if (s->set_ids[key][k] == person) { pos = k; break; }
        if (pos < 0) return;
    }
    int32_t last = s->set_ids[key][--s->set_n[key]];
    s->set_ids[key][pos] = last;
    s->pos_in_set[last] = pos;
    s->pos_in_set[person] = -1;
}

static void ensure_scratch(orc_sim* s, int64_t n)
{
    if (n <= s->cap_scratch) return;
    s->cap_scratch = 2 * n + 64;
    s->scratch_a = (int32_t*) realloc(s->scratch_a, s->cap_scratch * 4);
    s->scratch_b = (int32_t*) realloc(s->scratch_b, s->cap_scratch * 4);
    s->scratch_all = (int32_t*) realloc(s->scratch_all, s->cap_scratch * 4);
}

/* The medlabs DiseaseTransmission base-class step (contract C4, C6): duration since the last calculated call of
 * this sub-location, call infectPeople with the occupants present BEFORE the movement, update lastCalculation when
 * calculated, apply infections. */
static void evaluate(orc_sim* s, uint32_t key, double now)
{
    int32_t loc = s->key_loc[key];
    int infect_sub = s->type_infect_sub[s->loc_type[loc]];
    int nsub = s->loc_nsub[loc];
    double duration = now - s->last_calc[key];
    person_set sub = {s->set_ids[key], s->set_n[key], s->phase, s->exposure, 1};
    person_set all = sub;
    if (!(infect_sub || nsub < 2))
    {
        /* location.getAllPersonIds() (TArea:211): the union of the location's sub-location sets */
        int64_t tot = 0;
        for (uint32_t k = s->loc_base[loc]; k < s->loc_base[loc + 1]; ++k) tot += s->set_n[k];
        ensure_scratch(s, tot);
        int64_t m = 0;
        for (uint32_t k = s->loc_base[loc]; k < s->loc_base[loc + 1]; ++k)
            for (int32_t j = 0; j < s->set_n[k]; ++j) s->scratch_all[m++] = s->set_ids[k][j];
        all.ids = s->scratch_all; all.n = (int) tot;
    }
    else
        ensure_scratch(s, s->set_n[key]);
    record rec = {s->scratch_a, 0, s->scratch_b, 0, NAN, 0};
    int calculated;
    if (s->model == ORC_MODEL_AREA)
        calculated = infect_area(&s->area, s->seed, infect_sub, s->type_corr[s->loc_type[loc]], s->loc_area[loc],
                                 nsub, &sub, &all, now, duration, &rec);
    else
        calculated = infect_dist(&s->dist, s->seed, infect_sub, s->loc_area[loc], nsub, &sub, &all, now, duration,
                                 &rec);
    s->n_draws += rec.draws;
    if (!calculated) return;
    s->n_eval++;
    s->last_calc[key] = now;
    for (int32_t k = 0; k < rec.n_infected; ++k)
    {
        int32_t person = rec.infected[k];
        if (s->n_inf == s->cap_inf)
        {
            s->cap_inf = s->cap_inf ? 2 * s->cap_inf : 1024;
            s->inf = (infection_ev*) realloc(s->inf, s->cap_inf * sizeof(infection_ev));
        }
        infection_ev e = {person, loc, (int16_t) (key - s->loc_base[loc]), now, rec.p};
        s->inf[s->n_inf++] = e;
        s->exposure[person] = (float) now; /* contract C3/C4 */
        expose(s, person, now);
    }
}

void orc_sim_set_roles(orc_sim* s, int n, const uint8_t* role)
{
    free(s->role);
    s->role = (uint8_t*) malloc((size_t) (n > 0 ? n : 1));
    memcpy(s->role, role, (size_t) n);
    s->n_ref = 0;
    for (int i = 0; i < n; ++i) s->n_ref += role[i] == ROLE_REFERENCE;
}

/* The rows of data/<city>/epidemiology/infection_rates.csv (Construct:274-295): location id, infection_rate_factor,
 * infection_rate (-1 = not given). */
void orc_sim_set_rate_locations(orc_sim* s, int n, const int32_t* loc, const double* factor, const double* rate)
{
    free(s->loc_rate_kind); free(s->loc_rate_factor); free(s->loc_rate);
    s->loc_rate_kind = (uint8_t*) calloc((size_t) s->n_loc, 1);
    s->loc_rate_factor = (double*) calloc((size_t) s->n_loc, sizeof(double));
    s->loc_rate = (double*) calloc((size_t) s->n_loc, sizeof(double));
    for (int i = 0; i < n; ++i)
    {
        s->loc_rate_kind[loc[i]] = 1;
        s->loc_rate_factor[loc[i]] = factor[i];
        s->loc_rate[loc[i]] = rate[i] >= 0.0 ? rate[i] : -1.0;
    }
}

/* A person leaves a rate-based location (medlabs LocationProbBased; the class is not in the reference tree, SURVEY.md C16
 * INFERRED -- this is the convention of DESIGN.md 3.5, stated once here and once in csrc/transmit.cu:k_rate_locations):
 * no occupancy-based evaluation; a susceptible person who leaves at time t after a stay of d hours is exposed iff
 * U(seed; person, t, TAG_RATE) < 1 - exp(-lambda d / 24), lambda = infection_rate (per day) where the file gives one, else
 * infection_rate_factor x the share of the reference group (Worker) exposed during the previous simulated day. */
static void rate_exit(orc_sim* s, const move_ev* e, int32_t loc)
{
    if (!is_susceptible(s->phase[e->person])) return;
    double lambda = s->loc_rate[loc];
    if (!(lambda >= 0.0))
    {
        int64_t day = (int64_t) floor(e->t / 24.0);
        double share = 0.0;
        if (day > 0 && s->n_ref > 0 && day - 1 < s->n_ref_days) share = (double) s->ref_exposed[day - 1] / (double) s->n_ref;
        lambda = s->loc_rate_factor[loc] * share;
    }
    if (!(lambda > 0.0)) return;
    double x = lambda * ((e->t - s->stay_tin[e->person]) / 24.0);
    double p = 1.0 - orc_exp(-x);
    s->n_draws++;
    if (orc_u01(s->seed, (uint32_t) e->person, dbl_bits(e->t), TAG_RATE) < p)
    {
        if (s->n_inf == s->cap_inf)
        {
            s->cap_inf = s->cap_inf ? 2 * s->cap_inf : 1024;
            s->inf = (infection_ev*) realloc(s->inf, s->cap_inf * sizeof(infection_ev));
        }
        infection_ev ev = {e->person, loc, (int16_t) (e->key - s->loc_base[loc]), e->t, p};
        s->inf[s->n_inf++] = ev;
        s->exposure[e->person] = (float) e->t; /* contract C3/C4 */
        expose(s, e->person, e->t);
    }
}

void orc_sim_run(orc_sim* s, double t_until)
{
    if (!s->ev_sorted)
    {
        qsort(s->ev + s->ev_head, (size_t) (s->n_ev - s->ev_head), sizeof(move_ev), move_cmp);
        s->ev_sorted = 1;
    }
    for (;;)
    {
        double t_move = s->ev_head < s->n_ev ? s->ev[s->ev_head].t : INFINITY;
        double t_tr = s->n_heap > 0 ? s->heap[0].t : INFINITY;
        double t_pol = s->pol_head < s->n_pol ? s->pol[s->pol_head].t : INFINITY;
        /* equal times: policy events (scheduled at construction), then phase changes, then movements */
        if (t_pol <= t_tr && t_pol <= t_move)
        {
            if (!(t_pol < t_until)) break;
            policy_ev* p = &s->pol[s->pol_head++];
            s->now = p->t;
            if (p->which == 0) s->dist.psi = p->value; else s->dist.mu = p->value; /* DPol:33-40 */
        }
        else if (t_tr <= t_move)
        {
            if (!(t_tr < t_until)) break;
            heap_ev e = heap_pop(s);
            if (s->next_time[e.person] != e.t) continue; /* stale entry */
            s->now = e.t;
            change_phase(s, e.person, s->next_phase[e.person], e.t);
        }
        else
        {
            if (!(t_move < t_until)) break;
            move_ev* e = &s->ev[s->ev_head++];
            s->now = e->t;
            int32_t loc = s->key_loc[e->key];
            int rate_based = s->loc_rate_kind && s->loc_rate_kind[loc];
            if (e->kind == 0)
            {
                if (rate_based) rate_exit(s, e, loc);
                else evaluate(s, e->key, e->t); /* leaving person still in the set */
                set_remove(s, e->key, e->person);
            }
            else
            {
                if (!rate_based && s->trigger == ORC_TRIGGER_ENTER_AND_EXIT) evaluate(s, e->key, e->t); /* entering one not yet */
                set_add(s, e->key, e->person);
                if (s->stay_tin) s->stay_tin[e->person] = e->t;
            }
        }
    }
    s->now = t_until;
}

int64_t orc_sim_n_infections(const orc_sim* s) { return s->n_inf; }

void orc_sim_get_infections(const orc_sim* s, int32_t* person, int32_t* loc, int16_t* sub, double* time, double* p)
{
    for (int64_t i = 0; i < s->n_inf; ++i)
    {
        person[i] = s->inf[i].person; loc[i] = s->inf[i].loc; sub[i] = s->inf[i].sub;
        time[i] = s->inf[i].t; p[i] = s->inf[i].p;
    }
}

int64_t orc_sim_n_transitions(const orc_sim* s) { return s->n_tr; }

void orc_sim_get_transitions(const orc_sim* s, int32_t* person, uint8_t* phase, double* time)
{
    for (int64_t i = 0; i < s->n_tr; ++i)
    {
        person[i] = s->tr[i].person; phase[i] = s->tr[i].phase; time[i] = s->tr[i].t;
    }
}

void orc_sim_get_phase_counts(const orc_sim* s, int64_t counts[8]) { memcpy(counts, s->counts, sizeof(s->counts)); }

void orc_sim_get_agent_state(const orc_sim* s, uint8_t* phase, float* exposure, double* next_time,
                             uint8_t* next_phase)
{
    memcpy(phase, s->phase, s->n_person);
    memcpy(exposure, s->exposure, (size_t) s->n_person * 4);
    memcpy(next_time, s->next_time, (size_t) s->n_person * 8);
    memcpy(next_phase, s->next_phase, s->n_person);
}

int64_t orc_sim_n_evaluations(const orc_sim* s) { return s->n_eval; }
int64_t orc_sim_n_draws(const orc_sim* s) { return s->n_draws; }
double orc_sim_last_calc(const orc_sim* s, int32_t loc, int16_t sub) { return s->last_calc[s->loc_base[loc] + (uint32_t) sub]; }
