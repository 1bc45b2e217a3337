/*
 * heros_gpu.h -- C ABI of libheros_gpu.so, the B200 (sm_100a) engine for the data-parallel hot path of the
 * MedLabs/HEROS Covid-19 model: per-movement disease transmission over sub-location occupancy and the
 * disease-phase progression of all agents.
 *
 * This is the drop-in boundary.  The reference exposes this path as Java abstract-method plug-ins of
 * nl.tudelft.simulation:medlabs:2.1.3; a Java host binds the functions below through the Panama Foreign Function &
 * Memory API (java.lang.foreign.Linker.downcallHandle) -- see INTEGRATION.md for the binding class.  Each entry
 * point names the reference interface it replaces (paths relative to /root/reference/src/main/java/eu/heros/).
 *
 * Conventions
 *   - every function returns an int32 status: 0 = HEROS_OK, < 0 = error; heros_last_error() returns a description.
 *     No C++ exception and no exit() crosses the ABI.
 *   - all times are fp64 simulation hours (the DSOL simulator time unit, model/HerosApplication.java:170);
 *     person / location ids are int32 indices 0..n-1 in the order of heros_set_persons / heros_set_locations
 *     (plus the id bases of heros_set_id_bases when the population is sharded over several engines).
 *   - host buffers are caller-owned and only read (or written, for outputs) during the call; the engine owns all
 *     device memory.  Output buffers are caller-allocated with a capacity; the element count is returned.
 *   - a handle is not thread-safe (matches the single DSOL simulator thread); distinct handles are independent
 *     (one per replication / seed / GPU).
 *   - there is no CPU fallback: heros_create fails with HEROS_ERR_CUDA when no CUDA device is usable.
 */
#ifndef HEROS_GPU_H
#define HEROS_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEROS_ABI_VERSION 1

typedef struct heros_engine* heros_handle;

enum heros_status {
    HEROS_OK = 0,
    HEROS_ERR_INVALID = -1,      /* bad argument / call order */
    HEROS_ERR_CUDA = -2,         /* CUDA runtime error, or no device */
    HEROS_ERR_UNSUPPORTED = -3,  /* configuration outside what the engine reproduces exactly */
    HEROS_ERR_CAPACITY = -4,     /* caller buffer too small */
    HEROS_ERR_UNKNOWN_PARAMETER = -5
};

/* generic.diseasePropertiesModel (factory/ConstructHerosModel.java:137-139) */
enum heros_model { HEROS_MODEL_AREA = 0, HEROS_MODEL_DISTANCE = 1 };

/* which movements make the medlabs engine call infectPeople (driver contract C8) */
enum heros_trigger { HEROS_TRIGGER_EXIT_ONLY = 0, HEROS_TRIGGER_ENTER_AND_EXIT = 1 };

/* disease phases, registration order of disease/Covid19Progression.java:130-137 */
enum heros_phase {
    HEROS_SUSCEPTIBLE = 0, HEROS_EXPOSED = 1, HEROS_INFECTED_ASYMPTOMATIC = 2, HEROS_INFECTED_SYMPTOMATIC = 3,
    HEROS_HOSPITALIZED = 4, HEROS_ICU = 5, HEROS_DEAD = 6, HEROS_RECOVERED = 7, HEROS_N_PHASES = 8
};

typedef struct heros_config {
    uint32_t abi_version;  /* HEROS_ABI_VERSION */
    int32_t model;         /* heros_model */
    int32_t trigger;       /* heros_trigger */
    int32_t device;        /* CUDA device ordinal */
    uint64_t seed;         /* generic.Seed; key of every Philox draw */
    double window_hours;   /* batch window; must be <= the latent period (t_e_min / L) of the transmission model */
    uint32_t flags;        /* HEROS_FLAG_* */
    uint32_t reserved;
} heros_config;

/* count every calculated evaluation in heros_step_stats.evaluations, also in sub-locations where nobody can be
 * infected (there the engine otherwise only advances lastCalculation and skips the walk over the evaluations) */
#define HEROS_FLAG_EXACT_STATS 1u
/* (formerly a development aid of the group kernels; ignored) */
#define HEROS_FLAG_PHASE_CLOCKS 2u
/* the device buffers handed to heros_submit_visits_device are complete when the call is made (nothing that is still
 * enqueued on the engine stream writes them): the engine may then take them in on its side stream, beside the
 * transmission stage of a window that is still running (heros_step_async) */
#define HEROS_FLAG_RESIDENT_INPUTS 4u

/* covidT_area.* exactly as in the .properties file (days / seconds); converted as disease/Covid19TransmissionArea.java:63-68 */
typedef struct heros_area_params {
    double contagiousness, beta, t_e_min, t_e_mode, t_e_max, calculation_threshold;
} heros_area_params;

/* covidT_dist.*; converted as disease/Covid19TransmissionDistance.java:59-68 */
typedef struct heros_dist_params {
    double L, I, C, v_max, v_0, r, psi, alpha, mu, calculation_threshold;
} heros_dist_params;

enum heros_distribution { HEROS_DIST_CONSTANT = 0, HEROS_DIST_UNIFORM = 1, HEROS_DIST_TRIANGULAR = 2 };
typedef struct heros_duration { int32_t kind; int32_t reserved; double a, b, c; } heros_duration; /* days */

/* covidP.* (disease/Covid19Progression.java:143-173).  ConditionalProbability values are tabulated by the host as
 * fraction[which][female][age 0..127]; which / period indices below. */
enum { HEROS_F_ASYMPTOMATIC = 0, HEROS_F_SYMPTOMATIC_TO_HOSPITALIZED = 1, HEROS_F_HOSPITALIZED_TO_ICU = 2,
       HEROS_F_HOSPITALIZED_TO_DEAD = 3, HEROS_F_ICU_TO_DEAD = 4, HEROS_N_FRACTIONS = 5 };
enum { HEROS_P_INCUBATION_ASYMPTOMATIC = 0, HEROS_P_INCUBATION_SYMPTOMATIC = 1, HEROS_P_ASYMPTOMATIC_TO_RECOVERED = 2,
       HEROS_P_SYMPTOMATIC_TO_HOSPITALIZED = 3, HEROS_P_SYMPTOMATIC_TO_RECOVERED = 4, HEROS_P_HOSPITALIZED_TO_ICU = 5,
       HEROS_P_HOSPITALIZED_TO_DEAD = 6, HEROS_P_HOSPITALIZED_TO_RECOVERED = 7, HEROS_P_ICU_TO_DEAD = 8,
       HEROS_P_ICU_TO_RECOVERED = 9, HEROS_N_PERIODS = 10 };
typedef struct heros_prog_params {
    double fraction[HEROS_N_FRACTIONS][2][128];
    heros_duration period[HEROS_N_PERIODS];
} heros_prog_params;

typedef struct heros_step_stats {
    int64_t windows;          /* windows processed by this call */
    int64_t visits;           /* stays bucketed (sum over windows) */
    int64_t sublocations;     /* occupied sub-location segments (sum over windows) */
    int64_t evaluations;      /* infectPeople evaluations with calculated == true */
    int64_t draws;            /* Philox infection draws */
    int64_t infections;       /* new infections applied */
    int64_t transitions;      /* disease-phase transitions applied (incl. exposures) */
    int64_t gpu_launches;     /* kernels launched by this call */
    double gpu_ms;            /* device time of this call (CUDA events on the engine stream) */
    double transmit_ms;       /* of which: transmission kernels (small + big segments) */
    int64_t near_ties;        /* draws with |u - p| < 1e-13 (a flip could differ from the fp64 CPU result) */
    /* device time per stage, summed over the windows of the call (CUDA events on the engine stream) */
    double progress_ms;       /* k_window_open: disease-phase state machine + tickets of open / carried-over stays */
    double bucket_ms;         /* k_scan_lookback + k_bucket_scatter (+ retire) */
    double transmit_small_ms; /* k_transmit_stream: streams every stay of the window once (the hot path runs beside it) */
    double transmit_big_ms;   /* from the end of k_transmit_stream to the end of the stage: sub-warp kernels, the rest of
                                 the hot path */
    double resolve_ms;        /* earliest-draw-wins + expose */
    double retire_ms;         /* 0 (retire is part of k_bucket_scatter) */
    int64_t big_sublocations; /* sub-locations of > 32 stays (the hot path) */
    int64_t big_stays;        /* ... and their stays */
    double bucket_count_ms;   /* 0 (tickets are taken by k_submit / k_window_open) */
    double bucket_scan_ms, bucket_scatter_ms; /* the two parts of bucket_ms */
    /* the hot path (sub-locations of > 32 stays: k_hot_filter, k_hot_chain, k_hot_eval) runs on auxiliary streams, side by
     * side with k_transmit_stream and the sub-warp kernels: time from the fork (after the scatter) until [0] the hot path
     * as a whole and [1] the chain kernel of its small class finished ([2], [3]: 0), and until the 32-lane sub-warp kernel
     * finished (it starts after k_transmit_stream) */
    double group_done_ms[4], sub32_done_ms;
} heros_step_stats;

/* ---- lifecycle ---- */
int32_t heros_abi_version(void);
int32_t heros_create(const heros_config* cfg, heros_handle* out);
int32_t heros_destroy(heros_handle h);
const char* heros_last_error(heros_handle h); /* h may be NULL: error of the last failed heros_create */

/* ---- static inputs (factory/ConstructHerosModel.java: readLocationTypeTable :230-269, readLocationTable :300-395,
 *      readPersonTable :450-758).  Column meaning as in the CSV files; total_area = area * nb_sublocations as float
 *      (:381).  Optional arrays may be NULL. ---- */
int32_t heros_set_location_types(heros_handle h, int32_t n, const uint8_t* infect_in_sublocation,
                                 const double* correction_factor_area, const uint8_t* cap_constrained,
                                 const double* cap_persons_per_m2, const double* size_factor);
int32_t heros_set_locations(heros_handle h, int32_t n, const uint8_t* type, const int16_t* n_sublocations,
                            const float* total_area_m2, const float* lat, const float* lon);
int32_t heros_set_persons(heros_handle h, int32_t n, const int8_t* age, const uint8_t* female, const uint8_t* role,
                          const int32_t* home_location, const int16_t* home_sublocation,
                          const int32_t* work_location);
/* Rate-based locations = the medlabs LocationProbBased objects the reference builds for the rows of
 * data/<city>/epidemiology/infection_rates.csv (factory/ConstructHerosModel.java:274-295, 382-388: the satellite towns and
 * countries of the commuter roles), constructor arguments (infectionRateFactor, infectionRate, referenceGroupMap,
 * Covid19Progression.exposed).  location: ids of heros_set_locations; exactly one of rate_factor[i] / rate[i] is >= 0 (the
 * file holds -1 in the other column).  No occupancy-based transmission takes place in such a location; a susceptible
 * person who leaves it at time t after a stay of d hours is exposed iff U(seed; person, t) < 1 - exp(-lambda d / 24) with
 * lambda = rate (per day), or rate_factor x the share of the reference group (social role 7, Worker:
 * ConstructHerosModel.java:303-308) that was exposed during the previous simulated day.  The class itself is inside the
 * un-vendored medlabs jar: this rule is a stated convention, shared with the oracle.  With a factor-driven
 * location a window must not span a midnight (HEROS_ERR_UNSUPPORTED otherwise).  n = 0 removes the table. */
int32_t heros_set_rate_locations(heros_handle h, int32_t n, const int32_t* location, const double* rate_factor,
                                 const double* rate);

/* ---- model parameters ---- */
int32_t heros_set_transmission_area(heros_handle h, const heros_area_params* p);     /* Covid19TransmissionArea(model) */
int32_t heros_set_transmission_distance(heros_handle h, const heros_dist_params* p); /* Covid19TransmissionDistance(model) */
int32_t heros_set_progression(heros_handle h, const heros_prog_params* p);           /* Covid19Progression(model) */
/* DiseaseTransmission.setParameter(name, value) as fired by policy/DiseasePolicy.java:29-43 at simulator time
 * at_time.  Names "psi", "mu" (distance model only); anything else returns HEROS_ERR_UNKNOWN_PARAMETER, like the
 * reference's "Unrecognized variable name" message (Covid19TransmissionDistance.java:263, ...Area.java:254). */
int32_t heros_set_parameter(heros_handle h, const char* name, double value, double at_time);

/* ---- dynamic inputs ---- */
/* DiseaseProgression.expose(person, exposed) (Covid19Progression.java:182-201) for persons that are Susceptible:
 * initial seeding (ConstructHerosModel.java:1109-1128) and probability-based locations.  exposure time = (float) at_time. */
int32_t heros_expose(heros_handle h, int32_t n, const int32_t* person, const double* at_time);
/* Deterministic replacement of infectPersons (ConstructHerosModel.java:1109-1128): exposes at t = 0 the n_infected
 * eligible persons (Susceptible, min_age <= age <= max_age) with the smallest Philox hash of (seed, person).
 * selected_out (may be NULL) receives the chosen ids in ascending order. */
int32_t heros_seed_infections(heros_handle h, int32_t n_infected, int32_t min_age, int32_t max_age,
                              int32_t* selected_out);
/* Stays of persons in sub-locations: what Location.addPerson / removePerson of the medlabs engine record.  A stay
 * still open at submission time has t_out = +inf and is re-submitted, with the same t_in, once it is closed.  A stay
 * must be submitted before heros_step passes its t_in.  t_out <= t_in stays are ignored.  Entries whose person,
 * location or sub-location id lies outside the tables, or whose times are NaN / negative, are rejected on the device:
 * the next heros_step / heros_window_finish carries its window out without them and then returns HEROS_ERR_INVALID
 * with their number in heros_last_error (once; the engine stays usable). */
int32_t heros_submit_visits(heros_handle h, int64_t n, const int32_t* person, const int32_t* location,
                            const int16_t* sublocation, const double* t_in, const double* t_out);
/* The two halves of heros_submit_visits, for a host that wants the copy to start before it enqueues the window in
 * front of it:  heros_submit_visits_begin(day d + 1) starts the host -> device copies on the engine's copy stream and
 * returns at once (the buffers must stay untouched until the matching _end);  heros_step_async(day d) can then be
 * enqueued while the stays are already crossing PCIe;  heros_submit_visits_end() takes the stays in piece by piece
 * behind their copies and returns when the caller's buffers have been read.  One submission at a time. */
int32_t heros_submit_visits_begin(heros_handle h, int64_t n, const int32_t* person, const int32_t* location,
                                  const int16_t* sublocation, const double* t_in, const double* t_out);
int32_t heros_submit_visits_end(heros_handle h);
/* same, from device memory of the engine's GPU (multi-GPU exchange buffers, device-side schedule generators).  The
 * buffers are read asynchronously on the engine's stream (heros_get_stream): they must stay unchanged until the next
 * call that synchronises (heros_step, heros_window_finish, any heros_get_*). */
int32_t heros_submit_visits_device(heros_handle h, int64_t n, const int32_t* person, const int32_t* location,
                                   const int16_t* sublocation, const double* t_in, const double* t_out);

/* ---- device-side activity schedule (the producer of the stays).  Replaces, for the locators the
 *      engine implements, the per-person activity loop of the medlabs engine: week patterns "0_<phase>_<role>"
 *      (model/HerosModel.java:484-504; Mon-Thu use the Monday pattern, factory/ConstructHerosModel.java:931-940), activity
 *      rows (:763-942: duration / travel / until-fixed-time activities with Constant / Uniform / Triangular durations),
 *      locators (:944-1072: 0 Home, 1 Work / School, 2 Current, 3 Nearest(type choice), 4 Random(type choice) among the
 *      candidates within the maximum distance; both search from the place the person IS in, as the reference's
 *      NearestLocator / RandomLocator(new CurrentLocator(), ...) do (:997, 1012); + 0x10: the capacity-constrained form).  The phase that selects a day pattern is the one
 *      held at midnight; travel time is spent outside every location.  Draws are Philox numbers keyed by
 *      (seed, person, day, activity), so the stays do not depend on the device or the thread layout.
 *      heros_advance_schedule(day) produces the stays of simulated day `day` for every person directly in HBM (what
 *      heros_submit_visits would have received); heros_step then processes the day.  Capacity-constrained locators: see
 *      heros_set_capacity_diversion; LocationPolicy closures: heros_set_closure_policy. ---- */
typedef struct heros_activity {
    int32_t kind;      /* 0 duration activity, 1 travel, 2 until fixed time */
    int32_t locator;   /* 0 home, 1 work / school, 2 current, 3 nearest, 4 random, 5 travel-mode locator; | 0x10 capacity-constrained */
    int32_t choice;    /* index of the location-type choice table, -1 none */
    int32_t dist;      /* heros_distribution of the duration (hours) */
    double min, mode, max, until;
} heros_activity;
/* pattern_start / pattern_count: [phase 8][role 16][day type 4 = Mon-Thu, Fri, Sat, Sun] -> slice of activities[];
 * choice table c = entries choice_start[c] .. choice_start[c+1]) of (choice_type, choice_cum = cumulative weights);
 * nearest / n_cand: [location][type], cand: [location][type][n_candidates] (-1 = none; only rows of home buildings are
 * read); work_sublocation: [person].  Like HerosModel.checkChangeActivityPattern (model/HerosModel.java:490-501), which
 * throws when a person type has no name or the week pattern "0_<phase>_<role>" is missing, the call fails with
 * HEROS_ERR_INVALID if a social role present in the population is outside 1..15 or lacks the pattern of a phase other
 * than Dead. */
int32_t heros_set_schedule(heros_handle h, int32_t n_activities, const heros_activity* activities,
                           const int32_t* pattern_start, const int32_t* pattern_count, int32_t n_choices,
                           const int32_t* choice_start, const int32_t* choice_type, const double* choice_cum,
                           int32_t n_types, int32_t accommodation_type, int32_t n_candidates, const int32_t* nearest,
                           const int32_t* n_cand, const int32_t* cand, const int16_t* work_sublocation, uint64_t seed);
int32_t heros_advance_schedule(heros_handle h, int32_t day);
/* The range checks heros_set_schedule applies to its tables before any kernel indexes with them (host code, no device
 * needed): homes and work places inside the location table, sub-locations inside their location, nearest / candidate
 * entries inside the location table (-1 = none), choice tables naming known types.  HEROS_ERR_INVALID with the reason
 * in message (may be NULL). */
int32_t heros_validate_schedule_tables(int32_t n_locations, const int16_t* n_sub, int32_t n_persons, const int32_t* home_loc,
                                       const int16_t* home_sub, const int32_t* work_loc, const int16_t* work_sub,
                                       int32_t n_types, int32_t n_candidates, const int32_t* nearest, const int32_t* n_cand,
                                       const int32_t* cand, int32_t n_choices, const int32_t* choice_start,
                                       const int32_t* choice_type, int32_t accommodation_type, char* message,
                                       int32_t message_capacity);
/* = LocationType.setClosurePolicy(fractionOpen, fractionActivities, alternativeLocationType, reportAsLocationName) as
 * policy/LocationPolicy.java:39-46 calls it at the policy's event time (rows of data/<city>/policies/\*.csv; the report
 * name only labels output and is not needed here).  The rule itself is inside the un-vendored medlabs jar; the
 * convention of this engine (and of scenario.ScheduleGenerator, bit for bit): a location of the type is open iff its
 * uniform keyed by (seed, location) is < fraction_open -- the same locations stay open from day to day and a smaller
 * fraction keeps a subset; an activity (Work / School, Nearest, Random locators) whose location is open takes place
 * iff its uniform keyed by (seed, person, day, activity) is < fraction_activities; otherwise the person goes home.
 * alternative_type must be the Accommodation type (the person's home; every shipped policy row) or location_type
 * itself (= reopened, fractions taken as 1); anything else: HEROS_ERR_UNSUPPORTED.  Applies to the days advanced
 * after the call; call it between heros_step(24 d) and heros_advance_schedule(d) for a policy at day d. */
int32_t heros_set_closure_policy(heros_handle h, int32_t location_type, double fraction_open, double fraction_activities,
                                 int32_t alternative_type);
/* Capacity-constrained locators (NearestLocatorCap / RandomLocatorCap and their Choice forms,
 * factory/ConstructHerosModel.java:998-1020; capConstrained / capPersonsPerM2 / sizeFactor of locationtypes.csv): activities
 * whose locator carries the flag 0x10 send the share divert[l] of their visits to location l -- decided by a keyed
 * uniform per (person, day, activity) -- to the next candidate of the same search (the next-nearest place of the type;
 * the next of the random candidates).  The rule of the medlabs classes is not in the reference tree; this is the
 * convention of DESIGN.md 3.4, where the shares follow from the peak occupancy of an unconstrained day
 * (scenario.capacity_diversion).  n_locations = 0 or NULL: no location sheds anything. */
int32_t heros_set_capacity_diversion(heros_handle h, int32_t n_locations, const double* divert);
/* = what HerosModel.locationDump (model/HerosModel.java:266-299) writes every simulated hour: the number of persons in
 * every location (Location.getAllPersonIds().size(), :280) and in every sub-location
 * (DiseaseTransmission.getNrPersonsInSublocation(location, index), :284), sampled at times[0..n_times) (ascending,
 * hours) from the stays the engine holds for the coming window -- call it after heros_submit_visits /
 * heros_advance_schedule and before heros_step of that window.  A stay counts at t iff t_in <= t < t_out (the
 * movements of time t have taken place).  persons_per_location: [n_times][n_locations]; persons_per_sublocation:
 * [n_times][n_sublocations] with the sub-locations in location-major order, location l owning max(n_sub[l], 1)
 * consecutive entries; either may be NULL.  The sizes are checked against the tables of heros_set_locations. */
int32_t heros_sample_occupancy(heros_handle h, int32_t n_times, const double* times, int64_t n_locations,
                               int32_t* persons_per_location, int64_t n_sublocations, int32_t* persons_per_sublocation);
/* debugging / tests: every stay the engine currently holds (pending closed stays, then the open stay of every person
 * with t_out = +inf), in no particular order */
int32_t heros_debug_get_stays(heros_handle h, int64_t capacity, int32_t* person, int32_t* location, int16_t* sublocation,
                              double* t_in, double* t_out, int64_t* n_out);

/* debugging / profiling: when every kernel of a window actually ran.  enable != 0 starts (or restarts) the collection;
 * start_us / end_us (32 entries each, may be NULL) receive, per kernel id (HEROS_TL_*), the start of its first thread block
 * and the end of its last one in microseconds after the start of the window's first kernel, averaged over the windows
 * since the collection was started (-1: the kernel did not run); n_windows the number of windows.  The kernels of the
 * transmission stage run side by side on several streams: this is the picture no stream-ordered event can give. */
int32_t heros_debug_timeline(heros_handle h, int32_t enable, double* start_us, double* end_us, int64_t* n_windows);

/* ---- the hot path: advance simulated time to t_until in windows of cfg.window_hours: bucket stays by
 *      sub-location, evaluate infectPeople for every movement event of the window, draw, resolve (earliest successful
 *      draw wins), run the disease-phase state machine.  stats may be NULL. ---- */
int32_t heros_step(heros_handle h, double t_until, heros_step_stats* stats);
/* The same without waiting for the device: the windows up to t_until are enqueued and the call returns.  The host can
 * then hand over the stays of the NEXT window (heros_submit_visits copies them on its own stream while the windows run)
 * or advance its own simulator, and collects the result with heros_wait -- which waits for everything enqueued, fills
 * the statistics of the windows since the previous heros_wait / heros_step and reports their errors.  Every heros_get_*
 * also waits.  This is how a host whose activity engine produces day d + 1 while day d is processed keeps PCIe and GPU
 * busy at the same time. */
int32_t heros_step_async(heros_handle h, double t_until);
int32_t heros_wait(heros_handle h, heros_step_stats* stats);

/* ---- multi-GPU: one engine per GPU owns a shard of persons and locations.  Every person / location id that
 *      crosses this ABI is GLOBAL: the engine owns persons [person_base, person_base + n) and locations
 *      [location_base, location_base + n).  With the default bases (0, 0) nothing changes for a single engine.
 *      Draws are keyed by the global person id and exposure sums are formed in (person, t_in) order, so the results do
 *      not depend on how the population is sharded.
 *      A window then runs in three phases, and the host exchanges data between them (NCCL all-to-all):
 *        heros_window_begin            disease-phase state machine up to t_until
 *        heros_export_agent_words      packed agent words of the persons whose stays go to other shards   -> all-to-all
 *        heros_submit_guest_visits_device   stays of other shards' persons in this shard's locations (this window only)
 *        heros_window_transmit         bucketing + every evaluation + draws
 *        heros_export_guest_candidates successful draws on guests                                         -> all-to-all
 *        heros_import_candidates_device     ... received by the shard that owns the person
 *        heros_window_finish           earliest draw wins, expose, retire
 *      heros_step(t) == begin / transmit / finish per window.  All pointers of these calls are DEVICE pointers. ---- */
int32_t heros_set_id_bases(heros_handle h, int32_t person_base, int32_t location_base);
int32_t heros_window_begin(heros_handle h, double t_until);
int32_t heros_export_agent_words(heros_handle h, int64_t n, const int32_t* person, uint64_t* word_out, double* ill_end_out);
int32_t heros_submit_guest_visits_device(heros_handle h, int64_t n, const int32_t* person, const uint64_t* word,
                                         const double* ill_end, const int32_t* location, const int16_t* sublocation,
                                         const double* t_in, const double* t_out);
int32_t heros_window_transmit(heros_handle h);
int32_t heros_export_guest_candidates(heros_handle h, int64_t capacity, int32_t* person, int32_t* location,
                                      int16_t* sublocation, double* time, double* p, int64_t* n_out);
int32_t heros_import_candidates_device(heros_handle h, int64_t n, const int32_t* person, const int32_t* location,
                                       const int16_t* sublocation, const double* time, const double* p);
int32_t heros_window_finish(heros_handle h, heros_step_stats* stats);
/* ... without waiting for the device (the result is collected by heros_wait) */
int32_t heros_window_finish_async(heros_handle h);
/* A window that cannot be completed (an exchange call failed between its phases) is carried out with what it has
 * received so far, so that the engine is idle again instead of refusing every further call.  No-op when idle. */
int32_t heros_window_abort(heros_handle h);

/* The same exchange as FIXED-CAPACITY BLOCKS, one per peer, so that the host can run both all-to-alls of a window as
 * equal-split collectives on the engine's stream with nothing to negotiate and no host synchronisation before
 * heros_window_finish.  All block pointers are DEVICE pointers to n_blocks consecutive blocks of
 * heros_*_block_bytes(capacity) bytes (block d belongs to peer / shard d; capacity a multiple of 4):
 *   guest block     u64 count | pad | t_in f64[C] | t_out f64[C] | word u64[C] | ill_end f64[C] | person i32[C] |
 *                   location i32[C] | sublocation i16[C]
 *   candidate block u64 count | pad | time f64[C] | p f64[C] | person i32[C] | location i32[C] | sublocation i16[C]
 *   heros_pack_guest_blocks_device        rows (device columns, sorted by destination; row_offsets: n_blocks + 1 HOST
 *                                         integers) + the packed agent words of the persons -> one block per destination
 *   heros_submit_guest_blocks_device      the blocks received from the peers -> guest stays of the window
 *   heros_export_candidate_blocks_device  successful draws on guests -> the block of the shard that owns the person
 *                                         (person_bases: n_blocks + 1 HOST integers, first global person id per shard)
 *   heros_import_candidate_blocks_device  the blocks received from the peers -> candidates of the window
 * A block that overflows is reported by heros_window_finish (HEROS_ERR_CAPACITY). */
int64_t heros_guest_block_bytes(int64_t capacity);
int64_t heros_candidate_block_bytes(int64_t capacity);
int32_t heros_pack_guest_blocks_device(heros_handle h, int32_t n_blocks, const int64_t* row_offsets, const int32_t* person,
                                       const int32_t* location, const int16_t* sublocation, const double* t_in,
                                       const double* t_out, int64_t capacity, void* blocks_out);
int32_t heros_submit_guest_blocks_device(heros_handle h, int32_t n_blocks, int64_t capacity, const void* blocks);
int32_t heros_export_candidate_blocks_device(heros_handle h, int32_t n_blocks, const int32_t* person_bases,
                                             int64_t capacity, void* blocks_out);
int32_t heros_import_candidate_blocks_device(heros_handle h, int32_t n_blocks, int64_t capacity, const void* blocks);

/* The block exchange WITHOUT any collective call: the shards of one node map each other's receive regions (CUDA IPC over
 * NVLink / NVSwitch) and the pack / export kernels store every row straight into the memory of the shard it is meant
 * for; the kernel's last block raises a flag at every peer, the receiver waits for the flags of all peers with a
 * one-block kernel.  Everything is enqueued on the engine's stream; the host is not in the loop.
 *   heros_exchange_create   allocates this shard's receive region (rank of n_peers, block capacities as above) and
 *                           returns its 64-byte CUDA IPC handle (and the device pointer, for peers in the same process)
 *   heros_exchange_connect  ipc_handles: n_peers x 64 bytes gathered from all shards (own entry ignored);
 *                           local_regions (optional, may be NULL): device pointers of peers living in THIS process
 *   per window:  heros_window_begin -> heros_window_send_guests (rows sorted by destination, as for
 *                heros_pack_guest_blocks_device) -> heros_window_recv_guests -> heros_window_transmit ->
 *                heros_window_send_candidates (person_bases: n_peers + 1 HOST integers) -> heros_window_recv_candidates
 *                -> heros_window_finish.  A peer that does not answer within ~4 s fails the window (HEROS_ERR_CAPACITY). */
int32_t heros_exchange_create(heros_handle h, int32_t rank, int32_t n_peers, int64_t guest_capacity,
                              int64_t candidate_capacity, void* ipc_handle_out, void** region_out);
int32_t heros_exchange_connect(heros_handle h, const void* ipc_handles, void* const* local_regions);
/* optional: only the shards in source_mask (bit s = shard s) ever send guests to this shard, so only their guest flags
 * are awaited (default: all).  With a mask the candidate flags of ALL peers are awaited instead of only those of the
 * shards this one sent guests to: that wait is then what keeps a shard from running two windows ahead of a peer. */
int32_t heros_exchange_set_sources(heros_handle h, uint64_t source_mask);
/* the seven calls of a window in one (begin .. finish), for hosts that do nothing in between */
int32_t heros_window_run_peers(heros_handle h, double t_until, const int64_t* row_offsets, const int32_t* person,
                               const int32_t* location, const int16_t* sublocation, const double* t_in, const double* t_out,
                               const int32_t* person_bases, heros_step_stats* stats);
/* the same without waiting for the device (heros_window_finish_async at the end): windows can be enqueued back to back,
 * heros_wait collects them.  The peers hold each other back through the flags, never by more than one window. */
int32_t heros_window_run_peers_async(heros_handle h, double t_until, const int64_t* row_offsets, const int32_t* person,
                                     const int32_t* location, const int16_t* sublocation, const double* t_in,
                                     const double* t_out, const int32_t* person_bases);
int32_t heros_window_send_guests(heros_handle h, const int64_t* row_offsets, const int32_t* person, const int32_t* location,
                                 const int16_t* sublocation, const double* t_in, const double* t_out);
int32_t heros_window_recv_guests(heros_handle h);
int32_t heros_window_send_candidates(heros_handle h, const int32_t* person_bases);
int32_t heros_window_recv_candidates(heros_handle h);

/* ---- outputs ---- */
/* infection events (what the medlabs engine applies from InfectionRecord.infectedPersons), ordered by (time, person);
 * p = pInfection of the evaluation.  first = index of the first event wanted. */
int32_t heros_get_infections(heros_handle h, int64_t first, int64_t capacity, int32_t* person, int32_t* location,
                             int16_t* sublocation, double* time, double* p, int64_t* n_out);
/* The same events in the order the device logged them: all events of a window lie behind those of the windows before it,
 * inside a window the order is arbitrary (and may differ from run to run; the SET of events does not).  For a host that
 * applies the day's events and does not care about their order: no ordering pass, the columns are copied as they are.
 * `first` counts in log order here, in (time, person) order in heros_get_infections -- use one of the two consistently. */
int32_t heros_get_infections_unordered(heros_handle h, int64_t first, int64_t capacity, int32_t* person, int32_t* location,
                                       int16_t* sublocation, double* time, double* p, int64_t* n_out);
/* For a host that keeps windows enqueued (heros_step_async / heros_window_run_peers_async) and consumes results one window
 * behind: the compartment counts at the end of window number `window` (0-based since heros_set_persons / a snapshot load)
 * and the number of infection events the log held then.  Waits for THAT window only -- the windows enqueued behind it keep
 * running -- and reads on a side stream.  heros_get_infections_range then fetches log rows [first, first + n) of finished
 * windows, in log order (see heros_get_infections_unordered), without waiting for anything. */
int32_t heros_window_results(heros_handle h, int64_t window, int64_t counts[8], int64_t* infections_logged);
int32_t heros_get_infections_range(heros_handle h, int64_t first, int64_t n, int32_t* person, int32_t* location,
                                   int16_t* sublocation, double* time, double* p);
/* disease-phase trajectory: every phase entered (incl. Exposed), ordered by (time, person) */
int32_t heros_get_transitions(heros_handle h, int64_t first, int64_t capacity, int32_t* person, uint8_t* phase,
                              double* time, int64_t* n_out);
/* DiseasePhase person counters (addPerson / removePerson, Covid19Progression.java:184-186,211) at the current time */
int32_t heros_get_phase_counts(heros_handle h, int64_t counts[8]);
/* the same counters at the end of every window processed so far: counts[w*8 + phase], times[w] = window end */
int32_t heros_get_count_series(heros_handle h, int64_t capacity_windows, double* times, int64_t* counts,
                               int64_t* n_windows_out);
int32_t heros_get_agent_state(heros_handle h, uint8_t* phase, float* exposure_time, double* next_time,
                              uint8_t* next_phase);
/* debugging / restore: overwrite phase and exposure time of the listed persons (next transition re-drawn at `now`) */
int32_t heros_set_agent_state(heros_handle h, int32_t n, const int32_t* person, const uint8_t* phase,
                              const float* exposure_time, double now);
int32_t heros_get_last_calculation(heros_handle h, int32_t location, int16_t sublocation, double* out);
int32_t heros_get_time(heros_handle h, double* now);
/* the CUDA stream (cudaStream_t) every kernel and copy of this engine is ordered on: hosts that time the engine with
 * CUDA events, or that chain their own device work (exchange buffers of a sharded run), use it */
int32_t heros_get_stream(heros_handle h, void** stream_out);
/* HEROS_FLAG_PHASE_CLOCKS: clocks[class * 16 + phase], 64 entries, of the last window */
int32_t heros_get_phase_clocks(heros_handle h, uint64_t* clocks);

/* ---- checkpoint / resume.  The blob holds the dynamic state of an idle engine (no window in
 *      progress): agent words, next transitions, open stays, lastCalculation of every sub-location, the pending stays
 *      with their tickets, both event logs, the counters and compartment series, the parameter-change schedule
 *      (heros_set_parameter) and the closure table (heros_set_closure_policy), the engine time.  To resume, create an
 *      engine with the same model and seed, repeat the static inputs (heros_set_location_types / _locations / _persons,
 *      the transmission and progression parameters, heros_set_schedule if it is used) and call heros_snapshot_load: the
 *      run then continues bit for bit as the original would have.  Engines connected to a peer exchange are refused
 *      (HEROS_ERR_UNSUPPORTED); a blob taken with other tables, model or seed is refused (HEROS_ERR_INVALID). ---- */
int32_t heros_snapshot_size(heros_handle h, int64_t* bytes_out);
int32_t heros_snapshot_save(heros_handle h, int64_t capacity, void* blob, int64_t* bytes_out);
int32_t heros_snapshot_load(heros_handle h, int64_t bytes, const void* blob);

/* ---- 1:1 shim of DiseaseTransmission.infectPeople(location, personsInSublocation, duration)
 *      (Covid19TransmissionArea.java:133-248 / Covid19TransmissionDistance.java:115-248), evaluated on the GPU against
 *      the engine's current agent state with simulator time `now`.  all_persons = location.getAllPersonIds() (only
 *      read by the total-location branch; may be NULL when that branch is not taken).  Does not modify any state:
 *      like the reference method it only returns the InfectionRecord. ---- */
int32_t heros_infect_people(heros_handle h, int32_t location, int32_t n_persons, const int32_t* persons_in_sublocation,
                            int32_t n_all, const int32_t* all_persons, double now, double duration,
                            int32_t* calculated, int32_t* infectious_out, int32_t* n_infectious,
                            int32_t* infected_out, int32_t* n_infected, double* p_infection);

#ifdef __cplusplus
}
#endif
#endif /* HEROS_GPU_H */
