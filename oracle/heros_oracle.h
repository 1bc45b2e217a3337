/*
 * heros_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C, single-threaded, event-by-event restatement of the hot path of
 * averbraeck/medlabs-heros:
 *     eu/heros/disease/Covid19TransmissionArea.java      (TArea)
 *     eu/heros/disease/Covid19TransmissionDistance.java  (TDist)
 *     eu/heros/disease/Covid19Progression.java           (Prog)
 *     eu/heros/policy/DiseasePolicy.java                 (DPol)
 * plus a minimal deterministic driver standing in for the un-vendored
 * medlabs:2.1.3 engine (contract rows C1-C8).
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests -- the
 * reference has no tests, golden vectors or fixtures for this path and it
 * cannot be executed in this container (no JVM, medlabs/DSOL jars not
 * vendored).  The oracle is pinned only by (a) the worked example in the
 * reference's Javadoc (TArea:112-121), (b) the Philox4x32-10 known-answer
 * vectors of Random123, (c) libm cross-checks of its exp(), (d) hand-derived
 * boundary cases of both transmission models and of the total-location
 * branch, (e) the transition table, sojourn supports and age-band fractions
 * of Prog:208-346 with the alpha parameters, (f) three observable properties
 * of the only numbers the reference publishes for this path, the epidemic
 * curves of "seed comparison.xlsx" (tests/golden/seed_comparison.json) --
 * see tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libheros_gpu.so)
 * never links, loads or calls it.
 */
#ifndef HEROS_ORACLE_H
#define HEROS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* disease phases in registration order, Prog:130-137 */
enum { ORC_S = 0, ORC_E = 1, ORC_IA = 2, ORC_IS = 3, ORC_H = 4, ORC_ICU = 5, ORC_D = 6, ORC_R = 7, ORC_NPHASE = 8 };

enum { ORC_MODEL_AREA = 0, ORC_MODEL_DIST = 1 };
enum { ORC_TRIGGER_EXIT_ONLY = 0, ORC_TRIGGER_ENTER_AND_EXIT = 1 };
enum { ORC_DIST_CONSTANT = 0, ORC_DIST_UNIFORM = 1, ORC_DIST_TRIANGULAR = 2 };

/* TArea:63-68 -- raw property values (days / seconds), converted by the oracle */
typedef struct {
    double contagiousness, beta, t_e_min_days, t_e_mode_days, t_e_max_days, calculation_threshold_s;
} orc_area_props;

/* TDist:59-68 */
typedef struct {
    double L_days, I_days, C_days, v_max, v_0, r, psi, alpha, mu, calculation_threshold_s;
} orc_dist_props;

typedef struct { int32_t kind; double a, b, c; } orc_duration; /* days; Triangular(a=min,b=mode,c=max) */

/* Prog:143-173.  Fractions are tabulated as [female][age 0..127] */
enum { ORC_F_ASYMPTOMATIC = 0, ORC_F_S2H = 1, ORC_F_H2ICU = 2, ORC_F_H2D = 3, ORC_F_ICU2D = 4, ORC_NFRAC = 5 };
enum { ORC_P_INC_A = 0, ORC_P_INC_S = 1, ORC_P_A2R = 2, ORC_P_S2H = 3, ORC_P_S2R = 4, ORC_P_H2ICU = 5,
       ORC_P_H2D = 6, ORC_P_H2R = 7, ORC_P_ICU2D = 8, ORC_P_ICU2R = 9, ORC_NPERIOD = 10 };
typedef struct {
    double fraction[ORC_NFRAC][2][128];
    orc_duration period[ORC_NPERIOD];
} orc_prog_props;

/* ---- primitives (exported for known-answer tests) ---- */
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double   orc_u01(uint64_t seed, uint32_t person, uint64_t bits, uint32_t tag);
double   orc_exp(double x);
double   orc_java_max(double a, double b);
double   orc_sample_duration_hours(const orc_duration* d, double u);

/* ---- single-call restatement of infectPeople (TArea:133-248 / TDist:115-248) ----
 * sub_*  : the TIntSet personsInSublocation handed to infectPeople
 * all_*  : location.getAllPersonIds() (only read by the total-location branch)
 * returns 1 when InfectionRecord.calculated, 0 otherwise.
 * p_out  : pInfection (NaN when the call returned before computing it)
 */
int orc_infect_people_area(const orc_area_props* props, uint64_t seed,
                           int infect_in_sublocation, double correction_factor_area, float total_surface_m2,
                           int n_sublocations,
                           int n_sub, const int32_t* sub_ids, const uint8_t* sub_phase, const float* sub_exposure,
                           int n_all, const int32_t* all_ids, const uint8_t* all_phase, const float* all_exposure,
                           double now, double duration,
                           int32_t* infectious_out, int32_t* n_infectious, int32_t* infected_out, int32_t* n_infected,
                           double* p_out);

int orc_infect_people_dist(const orc_dist_props* props, uint64_t seed,
                           int infect_in_sublocation, float total_surface_m2, int n_sublocations,
                           int n_sub, const int32_t* sub_ids, const uint8_t* sub_phase, const float* sub_exposure,
                           int n_all, const int32_t* all_ids, const uint8_t* all_phase, const float* all_exposure,
                           double now, double duration,
                           int32_t* infectious_out, int32_t* n_infectious, int32_t* infected_out, int32_t* n_infected,
                           double* p_out);

/* ---- sequential event-by-event simulation (driver contract C1-C8) ---- */
typedef struct orc_sim orc_sim;

orc_sim* orc_sim_create(int model, int trigger, uint64_t seed);
void     orc_sim_destroy(orc_sim*);
void     orc_sim_set_area(orc_sim*, const orc_area_props*);
void     orc_sim_set_dist(orc_sim*, const orc_dist_props*);
void     orc_sim_set_prog(orc_sim*, const orc_prog_props*);
void     orc_sim_set_location_types(orc_sim*, int n, const uint8_t* infect_sub, const double* correction_factor_area);
void     orc_sim_set_locations(orc_sim*, int n, const uint8_t* type, const int16_t* n_sub, const float* total_area);
void     orc_sim_set_persons(orc_sim*, int n, const int8_t* age, const uint8_t* female);
/* DiseasePolicy (DPol:29-43): name is "psi" or "mu"; returns 0 ok, -1 unrecognised */
int      orc_sim_set_parameter(orc_sim*, const char* name, double value, double at_time);
/* rate-based locations (medlabs LocationProbBased, Construct:382-388; the convention of DESIGN.md 3.5) */
void     orc_sim_set_roles(orc_sim*, int n, const uint8_t* role);
void     orc_sim_set_rate_locations(orc_sim*, int n, const int32_t* loc, const double* factor, const double* rate);
void     orc_sim_expose(orc_sim*, int n, const int32_t* person, const double* at_time);
void     orc_sim_add_visits(orc_sim*, int64_t n, const int32_t* person, const int32_t* loc, const int16_t* sub,
                            const double* t_in, const double* t_out);
/* process every event with time < t_until */
void     orc_sim_run(orc_sim*, double t_until);

int64_t  orc_sim_n_infections(const orc_sim*);
void     orc_sim_get_infections(const orc_sim*, int32_t* person, int32_t* loc, int16_t* sub, double* time, double* p);
int64_t  orc_sim_n_transitions(const orc_sim*);
void     orc_sim_get_transitions(const orc_sim*, int32_t* person, uint8_t* phase, double* time);
void     orc_sim_get_phase_counts(const orc_sim*, int64_t counts[8]);
void     orc_sim_get_agent_state(const orc_sim*, uint8_t* phase, float* exposure, double* next_time, uint8_t* next_phase);
int64_t  orc_sim_n_evaluations(const orc_sim*);   /* infectPeople calls that returned calculated=true */
int64_t  orc_sim_n_draws(const orc_sim*);         /* U01 draws made by infectPeople */
double   orc_sim_last_calc(const orc_sim*, int32_t loc, int16_t sub);

#ifdef __cplusplus
}
#endif
#endif
