// schedule.cu -- device-side activity schedule (row A9 / SURVEY.md 8f row 1): heros_set_schedule,
// heros_advance_schedule.
#include "engine_state.cuh"

namespace heros {
namespace {

// ---- activity-schedule advancement ------------------------------------------------------------------------------------
// One simulated day of the activity loop the medlabs engine runs per person (week pattern "0_<phase>_<role>",
// model/HerosModel.java:484-504; activity rows and locators, factory/ConstructHerosModel.java:763-1072): one thread per
// person walks the day pattern selected by the phase held at midnight, draws durations and destinations from
// counter-based Philox numbers keyed by (seed, person, day, activity) and turns every move into a closed stay in the
// pool (with its ticket, like k_submit) and a new open stay.  The stays of a thread are collected in local memory and
// written out with one atomic per block.
constexpr int MAX_DAY_ACTIVITIES = 32;
constexpr uint32_t TAG_SCHEDULE = 0x300u;


struct ScheduleTables
{
    const Activity* acts; const int32_t* pat_start; const int32_t* pat_count;  // [phase 8][role 16][day type 4]
    const int32_t* choice_start; const int32_t* choice_type; const double* choice_cum;
    int n_types, n_candidates, accommodation_type;
    const int32_t* nearest; const int32_t* n_cand; const int32_t* cand;  // per location: [type], [type], [type][candidate]
    const int32_t* home_loc; const int16_t* home_sub; const int32_t* work_loc; const int16_t* work_sub;
    const uint32_t* loc_base; const int16_t* loc_nsub; uint32_t n_loc;
    const uint32_t* key_loc;  // location of every sub-location key (the anchor of Nearest / Random is the CURRENT place)
    const double* divert;     // per location: share of the visits of a *Cap locator that goes to the next candidate (or nullptr)
    uint64_t seed;
    // closure policy (policy/LocationPolicy.java:39-46): per location type {fraction of locations open, fraction of
    // activities that still take place}; nullptr while every type is fully open
    const double2* closure; const uint8_t* loc_type;
};

// Does the activity (person a, day * 64 + i) at location dl fall to the closure policy of the location's type?  The
// location is open iff its keyed uniform (seed, location) < fractionOpen; an activity in an open location takes place
// iff its keyed uniform (seed, person, day, activity) < fractionActivities (same rule as scenario.ScheduleGenerator).
__device__ __forceinline__ bool closed_by_policy(const ScheduleTables& S, uint32_t a, uint32_t day_act, int32_t dl, int ty)
{
    const double2 f = S.closure[ty];
    if (!(f.x < 1.0) && !(f.y < 1.0)) return false;
    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t) (S.seed >> 32);
    const double u_loc = (double) philox4x32_10((uint32_t) dl, 0u, 2u, TAG_SCHEDULE, k0, k1).x * 0x1.0p-32;
    const double u_act = (double) philox4x32_10(a, day_act, 1u, TAG_SCHEDULE, k0, k1).x * 0x1.0p-32;
    return !(u_loc < f.x && u_act < f.y);
}

__device__ __forceinline__ double activity_duration(const Activity& a, double u)
{
    if (a.dist == 2)  // Triangular(min, mode, max)
    {
        if (a.mx <= a.mn) return a.mn;
        const double fc = (a.mode - a.mn) / (a.mx - a.mn);
        if (u <= fc) return a.mn + hsqrt((a.mode - a.mn) * (a.mx - a.mn) * u);
        return a.mx - hsqrt((a.mx - a.mn) * (a.mx - a.mode) * (1.0 - u));
    }
    if (a.dist == 1) return a.mn + (a.mx - a.mn) * u;  // Uniform(min, max)
    return a.mode;
}

// The stays that end today are collected per block in shared memory (the stays of 256 persons: about 1 100, at most
// 256 x 16) and written to the pool together: one atomic per block reserves the slots, the writes are coalesced.
constexpr int STAGE_CAP = 2560;
struct StagedStays
{
    uint32_t person[STAGE_CAP], key[STAGE_CAP];
    double tin[STAGE_CAP], tout[STAGE_CAP];
    uint32_t n;         // staged so far (may exceed STAGE_CAP: the excess went straight to the pool)
    uint32_t spilled;   // ... that many
};

// The day of one person: every stay that ends goes to the staging area, the open stay of the person is updated.
__device__ __forceinline__ void walk_day(const ScheduleTables& S, uint32_t a, const uint64_t* view, uint32_t* open_key,
                                         double* open_tin, StagedStays& stage, const PoolArrays& P,
                                         uint32_t pool_cap, uint32_t* count, Counters* ctr, int day, int day_type)
{
    const uint64_t v = view[a];
    const uint32_t ph = view_phase(v), role = view_role(v);
    const double t0 = 24.0 * day, t1 = 24.0 * (day + 1);
    uint32_t cur_key = open_key[a];
    double cur_tin = open_tin[a];
    if (cur_tin != cur_tin) cur_key = KEY_NONE;  // nobody in the slot
    const uint32_t key0 = cur_key;
    const double tin0 = cur_tin;
    auto leave = [&](double t_leave) {  // the current stay ends
        if (cur_key != KEY_NONE && t_leave > cur_tin)
        {
            const uint32_t j = atomicAdd(&stage.n, 1u);
            if (j < (uint32_t) STAGE_CAP)
            {
                stage.person[j] = a; stage.key[j] = cur_key; stage.tin[j] = cur_tin; stage.tout[j] = t_leave;
            }
            else  // the staging area is full (cannot happen with <= 10 stays per person and day): straight to the pool
            {
                atomicAdd(&stage.spilled, 1u);
                atomicAdd(&ctr->n_sched, 1ull);
                const unsigned long long slot = atomicAdd(&ctr->n_fresh, 1ull);
                if (slot < pool_cap)
                {
                    P.person[slot] = a; P.key[slot] = cur_key;
                    P.tin[slot] = cur_tin; P.tout[slot] = t_leave;
                    P.rank[slot] = atomicAdd(&count[cur_key], 1u);
                }
                else
                    atomicOr(&ctr->overflow, 32u);
            }
        }
    };
    int n_act = 0, start = 0;
    if (ph == PH_D)  // a dead person leaves the model
    {
        leave(t0);
        cur_key = KEY_NONE;
    }
    else if (role >= 1 && role <= 15)
    {
        const int idx = ((int) ph * 16 + (int) role) * 4 + day_type;
        n_act = S.pat_count[idx];
        start = S.pat_start[idx];
    }
    const int32_t home_l = S.home_loc[a];
    const uint32_t home_key = S.loc_base[home_l] + (uint32_t) S.home_sub[a];
    double t = t0;
    for (int i = 0; i < n_act; ++i)
    {
        Activity act = S.acts[start + i];
        const bool capped = (act.locator & 0x10) != 0;  // NearestLocatorCap / RandomLocatorCap and their Choice forms
        act.locator &= 0xF;
        // destination
        uint32_t dest = cur_key;  // CurrentLocator, and whoever has nowhere else to go
        Philox4 r{0, 0, 0, 0};
        const bool needs_draw = act.dist != 0 || ((act.locator == 3 || act.locator == 4) && act.choice >= 0);
        if (needs_draw)
            r = philox4x32_10(a, (uint32_t) (day * 64 + i), 0u, TAG_SCHEDULE, (uint32_t) S.seed, (uint32_t) (S.seed >> 32));
        if (act.locator == 0)
            dest = home_key;
        else if (act.locator == 1)
        {
            const int32_t wl = S.work_loc[a];
            if (wl >= 0)
            {
                dest = S.loc_base[wl] + (uint32_t) S.work_sub[a];
                if (S.closure && closed_by_policy(S, a, (uint32_t) (day * 64 + i), wl, S.loc_type[wl])) dest = home_key;
            }
        }
        else if ((act.locator == 3 || act.locator == 4) && act.choice >= 0)
        {
            const double u_type = (double) r.z * 0x1.0p-32;
            const double u_cand = (double) (r.w >> 16) * 0x1.0p-16, u_sub = (double) (r.w & 0xFFFFu) * 0x1.0p-16;
            const int c0 = S.choice_start[act.choice], c1 = S.choice_start[act.choice + 1];
            const double x = u_type * S.choice_cum[c1 - 1];
            int pick = 0;  // number of cumulative weights <= x (searchsorted, side = right)
            while (pick < c1 - c0 && S.choice_cum[c0 + pick] <= x) ++pick;
            if (pick > c1 - c0 - 1) pick = c1 - c0 - 1;
            const int ty = S.choice_type[c0 + pick];
            // NearestLocator / RandomLocator(new CurrentLocator(), ...) (ConstructHerosModel.java:997, 1012): the search
            // starts from the place the person is in; somebody on the way (in no location) searches from home
            const size_t anchor = cur_key != KEY_NONE ? (size_t) S.key_loc[cur_key] : (size_t) home_l;
            int32_t dl;
            const int n = S.n_cand[anchor * S.n_types + ty];
            long long j = 0;
            if (act.locator == 3)  // Nearest(type)
                dl = S.nearest[anchor * S.n_types + ty];
            else  // Random(type) within the maximum distance
            {
                j = (long long) (u_cand * (double) n);
                if (j > n - 1) j = n - 1;
                if (j < 0) j = 0;
                dl = n > 0 ? S.cand[(anchor * S.n_types + ty) * S.n_candidates + j] : -1;
            }
            if (capped && S.divert && dl >= 0 && n >= 2)
            {
                // capacity-constrained form: a keyed share of the visits to an over-subscribed place goes to the next
                // candidate of the search (scenario.capacity_diversion: the rule itself is inside medlabs)
                const double share = S.divert[dl];
                if (share > 0.0)
                {
                    const double u_cap = (double) philox4x32_10(a, (uint32_t) (day * 64 + i), 4u, TAG_SCHEDULE, (uint32_t) S.seed,
                                                                (uint32_t) (S.seed >> 32)).x * 0x1.0p-32;
                    if (u_cap < share)
                    {
                        const long long j2 = act.locator == 3 ? 1 : (j + 1) % n;
                        const int32_t alt = S.cand[(anchor * S.n_types + ty) * S.n_candidates + (j2 < S.n_candidates ? j2 : S.n_candidates - 1)];
                        if (alt >= 0) dl = alt;
                    }
                }
            }
            if (ty == S.accommodation_type || dl < 0)
                dest = home_key;
            else if (S.closure && closed_by_policy(S, a, (uint32_t) (day * 64 + i), dl, ty))
                dest = home_key;
            else
            {
                const int nsub = max((int) S.loc_nsub[dl], 1);
                int ds = (int) (u_sub * (double) nsub);
                if (ds > nsub - 1) ds = nsub - 1;
                dest = S.loc_base[dl] + (uint32_t) ds;
            }
        }
        const bool move = dest != cur_key;
        const double u_dur = (double) ((((uint64_t) r.x << 32) | r.y) >> 11) * 0x1.0p-53;
        const double dur = activity_duration(act, u_dur);
        if (act.kind == 1)  // travel: leave now, arrive after the travel time (spent outside every location)
        {
            if (move)
            {
                leave(t);
                const double t_arr = t + dur;
                const bool arrived = t_arr < t1;
                cur_key = arrived ? dest : KEY_NONE;
                cur_tin = t_arr;
                t = t_arr;
                if (!arrived) break;
            }
        }
        else
        {
            if (move) { leave(t); cur_key = dest; cur_tin = t; }
            if (act.kind == 0) t = t + dur;
            else { const double until = t0 + act.until; t = t > until ? t : until; }
            if (!(t < t1)) break;
        }
    }
    if (cur_key != key0 || (cur_key != KEY_NONE && f64_bits(cur_tin) != f64_bits(tin0)))
    {
        open_key[a] = cur_key;
        open_tin[a] = cur_key != KEY_NONE ? cur_tin : bits_f64(OPEN_CLOSED);
    }
}

__global__ void __launch_bounds__(TPB) k_advance_day(const ScheduleTables S, uint32_t n_agents, const uint32_t* order,
                                                     const uint64_t* view, uint32_t* open_key, double* open_tin,
                                                     PoolArrays P, uint32_t pool_cap, uint32_t* count,
                                                     Counters* ctr, int day, int day_type)
{
    extern __shared__ __align__(16) unsigned char smem[];
    StagedStays& stage = *reinterpret_cast<StagedStays*>(smem);
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) { stage.n = 0; stage.spilled = 0; }
    __syncthreads();
    // persons are visited grouped by social role: the threads of a warp then walk the same day pattern (the phase that
    // also selects it is Susceptible for most) instead of 32 different ones
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < n_agents)
        walk_day(S, order[slot], view, open_key, open_tin, stage, P, pool_cap, count, ctr, day, day_type);
    __syncthreads();
    const uint32_t n = min(stage.n, (uint32_t) STAGE_CAP);
    // the block's slots in the fresh pool, whose cursor lives on the device
    if (threadIdx.x == 0)
    {
        s_base = n ? atomicAdd(&ctr->n_fresh, (unsigned long long) n) : 0ull;
        if (n) atomicAdd(&ctr->n_sched, (unsigned long long) n);
    }
    __syncthreads();
    // four staged stays per thread and round: their tickets are taken together (four atomics in flight)
    for (uint32_t j0 = threadIdx.x; j0 < n; j0 += blockDim.x * 4)
    {
        uint32_t key[4], ticket[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const uint32_t j = j0 + q * blockDim.x;
            key[q] = (j < n && s_base + j < pool_cap) ? stage.key[j] : KEY_NONE;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) ticket[q] = key[q] != KEY_NONE ? atomicAdd(&count[key[q]], 1u) : KEY_NONE;  // for the next window
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const uint32_t j = j0 + q * blockDim.x;
            if (j >= n) continue;
            const unsigned long long at = s_base + j;
            if (at >= pool_cap) { atomicOr(&ctr->overflow, 32u); continue; }
            P.person[at] = stage.person[j];
            P.key[at] = key[q];
            P.tin[at] = stage.tin[j];
            P.tout[at] = stage.tout[j];
            P.rank[at] = ticket[q];
        }
    }
}

}  // namespace
}  // namespace heros

using namespace heros;
using namespace heros_impl;

extern "C" {

int32_t heros_validate_schedule_tables(int32_t n_locations, const int16_t* n_sub, int32_t n_persons, const int32_t* home_loc,
                                       const int16_t* home_sub, const int32_t* work_loc, const int16_t* work_sub,
                                       int32_t n_types, int32_t n_candidates, const int32_t* nearest, const int32_t* n_cand,
                                       const int32_t* cand, int32_t n_choices, const int32_t* choice_start,
                                       const int32_t* choice_type, int32_t accommodation_type, char* message,
                                       int32_t message_capacity)
{
    auto bad = [&](const char* why) {
        if (message && message_capacity > 0) snprintf(message, (size_t) message_capacity, "%s", why);
        return (int32_t) HEROS_ERR_INVALID;
    };
    if (n_locations <= 0 || n_persons < 0 || n_types <= 0 || n_candidates < 0 || n_choices < 0 || !n_sub || !nearest || !n_cand ||
        (n_candidates > 0 && !cand) || (n_persons > 0 && (!home_loc || !home_sub || !work_loc || !work_sub)) ||
        (n_choices > 0 && (!choice_start || !choice_type)))
        return bad("null or empty schedule table");
    const int64_t n_loc = n_locations;
    auto subs_of = [&](int64_t loc) { return (int) (n_sub[loc] > 1 ? n_sub[loc] : 1); };
    for (int64_t a = 0; a < n_persons; ++a)
    {
        const int64_t hl = home_loc[a], wl = work_loc[a];
        if (hl < 0 || hl >= n_loc || home_sub[a] < 0 || home_sub[a] >= subs_of(hl))
            return bad("a person's home location / sub-location lies outside the location table");
        if (wl >= n_loc) return bad("a person's work location lies outside the location table");
        if (wl >= 0 && (work_sub[a] < 0 || work_sub[a] >= subs_of(wl)))
            return bad("a person's work sub-location lies outside the work location");
    }
    const size_t cells = (size_t) n_locations * (size_t) n_types;
    for (size_t i = 0; i < cells; ++i)
    {
        if (nearest[i] >= n_loc) return bad("nearest-location table points outside the location table");
        if (n_cand[i] < 0 || n_cand[i] > n_candidates) return bad("candidate count outside 0..n_candidates");
        for (int j = 0; j < n_cand[i]; ++j)
            if (cand[i * (size_t) n_candidates + (size_t) j] >= n_loc) return bad("candidate table points outside the location table");
    }
    const size_t n_entries = n_choices ? (size_t) choice_start[n_choices] : 0;
    for (size_t i = 0; i < n_entries; ++i)
        if (choice_type[i] < 0 || choice_type[i] >= n_types) return bad("choice table names an unknown location type");
    if (accommodation_type < 0 || accommodation_type >= n_types) return bad("unknown accommodation type");
    return HEROS_OK;
}

int32_t heros_set_schedule(heros_handle h, int32_t n_activities, const heros_activity* activities,
                           const int32_t* pattern_start, const int32_t* pattern_count, int32_t n_choices,
                           const int32_t* choice_start, const int32_t* choice_type, const double* choice_cum,
                           int32_t n_types, int32_t accommodation_type, int32_t n_candidates, const int32_t* nearest,
                           const int32_t* n_cand, const int32_t* cand, const int16_t* work_sublocation, uint64_t seed)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->n_agents == 0 || h->n_loc == 0) return fail(h, HEROS_ERR_INVALID, "locations and persons must be set first");
    if (!h->d_home_loc.p) return fail(h, HEROS_ERR_INVALID, "heros_set_persons was called without home / work locations");
    if (n_activities <= 0 || !activities || !pattern_start || !pattern_count || n_choices < 0 || !choice_start ||
        n_types <= 0 || n_candidates < 0 || !nearest || !n_cand || (n_candidates > 0 && !cand) || !work_sublocation)
        return fail(h, HEROS_ERR_INVALID, "bad schedule tables");
    static_assert(sizeof(heros_activity) == sizeof(Activity), "activity layout");
    int max_acts = 0;
    for (int i = 0; i < 8 * 16 * 4; ++i)
    {
        if (pattern_count[i] < 0 || pattern_start[i] < 0 || pattern_start[i] + pattern_count[i] > n_activities)
            return fail(h, HEROS_ERR_INVALID, "day pattern outside the activity table");
        max_acts = std::max(max_acts, pattern_count[i]);
    }
    if (max_acts > MAX_DAY_ACTIVITIES) return fail(h, HEROS_ERR_UNSUPPORTED, "more than 32 activities in a day pattern");
    for (int i = 0; i < n_activities; ++i)
        if (activities[i].choice >= n_choices) return fail(h, HEROS_ERR_INVALID, "choice table index out of range");
    {
        // HerosModel.checkChangeActivityPattern (model/HerosModel.java:490-501) throws when a person's type has no name or
        // the week pattern "0_<phase>_<role>" does not exist: every role present in the population needs the (Monday)
        // pattern of every phase but Dead
        bool present[256] = {};
        for (uint8_t r : h->h_role) present[r] = true;
        for (int r = 0; r < 256; ++r)
        {
            if (!present[r]) continue;
            char buf[160];
            if (r < 1 || r > 15)
            {
                snprintf(buf, sizeof buf, "Person type name of social role %d not found", r);
                return fail(h, HEROS_ERR_INVALID, buf);
            }
            for (int ph = 0; ph < 8; ++ph)
                if (ph != (int) PH_D && pattern_count[(ph * 16 + r) * 4] <= 0)
                {
                    snprintf(buf, sizeof buf, "Week pattern key 0_<phase %d>_<role %d> not found in week pattern map", ph, r);
                    return fail(h, HEROS_ERR_INVALID, buf);
                }
        }
    }
    for (int i = 0; i < n_choices; ++i)
        if (choice_start[i + 1] <= choice_start[i]) return fail(h, HEROS_ERR_INVALID, "empty choice table");
    {
        // every index the kernel follows is checked here, once, on the host: a table that points outside the location
        // table is refused instead of being read out of bounds on the device
        if (h->h_home_loc.size() != h->n_agents || h->h_home_sub.size() != h->n_agents || h->h_work_loc.size() != h->n_agents)
            return fail(h, HEROS_ERR_INVALID, "heros_set_persons was called without home / work locations");
        char why[160] = "";
        if (heros_validate_schedule_tables((int32_t) h->n_loc, h->h_loc_nsub.data(), (int32_t) h->n_agents, h->h_home_loc.data(),
                                           h->h_home_sub.data(), h->h_work_loc.data(), work_sublocation, n_types, n_candidates,
                                           nearest, n_cand, cand, n_choices, choice_start, choice_type, accommodation_type,
                                           why, (int32_t) sizeof why) != HEROS_OK)
            return fail(h, HEROS_ERR_INVALID, why);
    }
    cudaSetDevice(h->cfg.device);
    const size_t N = h->n_agents, L = h->n_loc, n_ch = n_choices ? (size_t) choice_start[n_choices] : 0;
    cudaStream_t s = h->stream;
    CU(h, h->s_acts.ensure(n_activities)); CU(h, h->s_pat_start.ensure(512)); CU(h, h->s_pat_count.ensure(512));
    CU(h, h->s_choice_start.ensure((size_t) n_choices + 1)); CU(h, h->s_choice_type.ensure(n_ch + 1));
    CU(h, h->s_choice_cum.ensure(n_ch + 1)); CU(h, h->s_nearest.ensure(L * n_types)); CU(h, h->s_ncand.ensure(L * n_types));
    CU(h, h->s_cand.ensure(L * n_types * (size_t) std::max(n_candidates, 1))); CU(h, h->d_work_sub.ensure(N));
    CU(h, cudaMemcpyAsync(h->s_acts.p, activities, (size_t) n_activities * sizeof(Activity), cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->s_pat_start.p, pattern_start, 512 * 4, cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->s_pat_count.p, pattern_count, 512 * 4, cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->s_choice_start.p, choice_start, ((size_t) n_choices + 1) * 4, cudaMemcpyHostToDevice, s));
    if (n_ch)
    {
        CU(h, cudaMemcpyAsync(h->s_choice_type.p, choice_type, n_ch * 4, cudaMemcpyHostToDevice, s));
        CU(h, cudaMemcpyAsync(h->s_choice_cum.p, choice_cum, n_ch * 8, cudaMemcpyHostToDevice, s));
    }
    CU(h, cudaMemcpyAsync(h->s_nearest.p, nearest, L * n_types * 4, cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->s_ncand.p, n_cand, L * n_types * 4, cudaMemcpyHostToDevice, s));
    if (n_candidates > 0)
        CU(h, cudaMemcpyAsync(h->s_cand.p, cand, L * n_types * (size_t) n_candidates * 4, cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->d_work_sub.p, work_sublocation, N * 2, cudaMemcpyHostToDevice, s));
    {
        std::vector<uint32_t> order(N);  // counting sort of the persons by role, stable
        size_t first[257] = {};
        for (size_t a = 0; a < N; ++a) ++first[(size_t) h->h_role[a] + 1];
        for (int r = 0; r < 256; ++r) first[r + 1] += first[r];
        for (size_t a = 0; a < N; ++a) order[first[h->h_role[a]]++] = (uint32_t) a;
        CU(h, h->s_order.ensure(N));
        CU(h, cudaMemcpyAsync(h->s_order.p, order.data(), N * 4, cudaMemcpyHostToDevice, s));
        CU(h, cudaStreamSynchronize(s));
    }
    CU(h, cudaStreamSynchronize(s));
    {
        // room the fresh pool needs for one day: a stay can only end at an activity whose locator may name another place
        // (everything but the Current locator), plus one for a person who dies; per person the longest such pattern
        size_t per_role[16] = {};
        for (int r = 1; r < 16; ++r)
            for (int ph = 0; ph < 8; ++ph)
                for (int d = 0; d < 4; ++d)
                {
                    const int idx = (ph * 16 + r) * 4 + d;
                    size_t moves = 1;
                    for (int i = 0; i < pattern_count[idx]; ++i) moves += (activities[pattern_start[idx] + i].locator & 0xF) != 2 ? 1 : 0;
                    per_role[r] = std::max(per_role[r], moves);
                }
        size_t room = 0;
        for (uint8_t r : h->h_role) room += r < 16 ? per_role[r] : 1;
        h->sched_room = room;
    }
    h->sched_max_acts = max_acts; h->sched_n_types = n_types; h->sched_n_cand = n_candidates;
    h->sched_accommodation = accommodation_type; h->sched_seed = seed;
    h->have_schedule = true;
    return HEROS_OK;
}

int32_t heros_set_capacity_diversion(heros_handle h, int32_t n_locations, const double* divert)
{
    if (!h) return HEROS_ERR_INVALID;
    if (!h->have_schedule) return fail(h, HEROS_ERR_INVALID, "heros_set_schedule has not been called");
    cudaSetDevice(h->cfg.device);
    CU(h, cudaStreamSynchronize(h->stream));
    if (n_locations == 0 || !divert) { h->s_divert.release(); return HEROS_OK; }
    if ((uint32_t) n_locations != h->n_loc) return fail(h, HEROS_ERR_INVALID, "one share per location of heros_set_locations");
    bool any = false;
    for (int32_t i = 0; i < n_locations; ++i)
    {
        if (!(divert[i] >= 0.0 && divert[i] <= 1.0)) return fail(h, HEROS_ERR_INVALID, "a share must lie in [0, 1]");
        any = any || divert[i] > 0.0;
    }
    if (!any) { h->s_divert.release(); return HEROS_OK; }
    CU(h, h->s_divert.ensure((size_t) n_locations));
    CU(h, cudaMemcpy(h->s_divert.p, divert, (size_t) n_locations * 8, cudaMemcpyHostToDevice));
    return HEROS_OK;
}

int32_t heros_set_closure_policy(heros_handle h, int32_t location_type, double fraction_open, double fraction_activities,
                                 int32_t alternative_type)
{
    if (!h) return HEROS_ERR_INVALID;
    if (!h->have_schedule) return fail(h, HEROS_ERR_INVALID, "heros_set_schedule has not been called");
    if (location_type < 0 || location_type >= h->n_types || alternative_type < 0 || alternative_type >= h->n_types)
        return fail(h, HEROS_ERR_INVALID, "location type id out of range");
    if (alternative_type == location_type)
        fraction_open = fraction_activities = 1.0;  // "Workplace -> Workplace": the type is open again
    else if (alternative_type != h->sched_accommodation)
        return fail(h, HEROS_ERR_UNSUPPORTED, "the alternative of a closed location type must be Accommodation (the person's home) or the type itself");
    if (!(fraction_open >= 0.0 && fraction_open <= 1.0) || !(fraction_activities >= 0.0 && fraction_activities <= 1.0))
        return fail(h, HEROS_ERR_INVALID, "fractions must lie in [0, 1]");
    cudaSetDevice(h->cfg.device);
    if (h->h_closure.size() != (size_t) h->n_types * 2) h->h_closure.assign((size_t) h->n_types * 2, 1.0);
    h->h_closure[(size_t) location_type * 2] = fraction_open;
    h->h_closure[(size_t) location_type * 2 + 1] = fraction_activities;
    bool active = false;
    for (double f : h->h_closure) active = active || f < 1.0;
    cudaStream_t s = h->stream;
    if (!h->s_loc_type.p)
    {
        CU(h, h->s_loc_type.ensure(h->n_loc));
        CU(h, cudaMemcpyAsync(h->s_loc_type.p, h->h_loc_type.data(), h->n_loc, cudaMemcpyHostToDevice, s));
    }
    CU(h, h->s_closure.ensure(h->n_types));
    // the table is read by the k_advance_day launches that follow on this stream
    CU(h, cudaMemcpyAsync(h->s_closure.p, h->h_closure.data(), (size_t) h->n_types * 16, cudaMemcpyHostToDevice, s));
    CU(h, cudaStreamSynchronize(s));
    h->closure_active = active;
    return HEROS_OK;
}

int32_t heros_advance_schedule(heros_handle h, int32_t day)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (!h->have_schedule) return fail(h, HEROS_ERR_INVALID, "heros_set_schedule has not been called");
    if (h->win_state != 0) return fail(h, HEROS_ERR_INVALID, "a window opened by heros_window_begin is still in progress");
    if (day < 0 || 24.0 * day < h->t_now) return fail(h, HEROS_ERR_INVALID, "the day lies before the engine time");
    if (h->person_base != 0 || h->location_base != 0)
        return fail(h, HEROS_ERR_UNSUPPORTED, "the device-side schedule runs on an unsharded engine");
    cudaSetDevice(h->cfg.device);
    // every activity of a day can end a stay: the fresh pool gets room for all of them (its fill count lives on the device)
    const size_t room = h->sched_room;
    if ((uint64_t) h->carry_ub + h->fresh_ub + room + h->n_agents >= 0xFFFFFFF0ull) return fail(h, HEROS_ERR_CAPACITY, "more than 2^32 stays in one window");
    rc = ensure_fresh(h, h->fresh_ub + room);
    if (rc != HEROS_OK) return rc;
    if ((rc = join_submissions(h)) != HEROS_OK) return rc;  // stays submitted by the host come first
    ScheduleTables S{};
    S.acts = h->s_acts.p; S.pat_start = h->s_pat_start.p; S.pat_count = h->s_pat_count.p;
    S.choice_start = h->s_choice_start.p; S.choice_type = h->s_choice_type.p; S.choice_cum = h->s_choice_cum.p;
    S.n_types = h->sched_n_types; S.n_candidates = h->sched_n_cand; S.accommodation_type = h->sched_accommodation;
    S.nearest = h->s_nearest.p; S.n_cand = h->s_ncand.p; S.cand = h->s_cand.p;
    S.home_loc = h->d_home_loc.p; S.home_sub = h->d_home_sub.p; S.work_loc = h->d_work_loc.p; S.work_sub = h->d_work_sub.p;
    S.loc_base = h->d_loc_base.p; S.loc_nsub = h->d_loc_nsub.p; S.n_loc = h->n_loc; S.key_loc = h->d_key_loc.p; S.seed = h->sched_seed;
    S.divert = h->s_divert.p;
    S.closure = h->closure_active ? h->s_closure.p : nullptr; S.loc_type = h->s_loc_type.p;
    static const int DAY_TYPE[7] = {0, 0, 0, 0, 1, 2, 3};  // Mon-Thu share the Monday pattern (ConstructHerosModel.java:931-940)
    CU(h, cudaMemsetAsync(&h->d_ctr.p->n_sched, 0, sizeof(unsigned long long), h->stream));
    static bool configured[64] = {};
    if (h->cfg.device >= 0 && h->cfg.device < 64 && !configured[h->cfg.device])
    {
        CU(h, cudaFuncSetAttribute(k_advance_day, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(StagedStays)));
        configured[h->cfg.device] = true;
    }
    k_advance_day<<<blocks_for(h->n_agents), TPB, sizeof(StagedStays), h->stream>>>(
        S, h->n_agents, h->s_order.p, h->d_view.p, h->d_open_key.p, h->d_open_tin.p, fresh_arrays(h),
        (uint32_t) std::min<size_t>(h->fresh_person.n, 0xFFFFFFFFu), h->bk_count.p, h->d_ctr.p, day, DAY_TYPE[day % 7]);
    CU(h, cudaGetLastError());
    // nothing is read back: the window kernels take the fill count from the device.  An average day ends about four
    // stays per person; the bound used for the grids of the next window is the room reserved above.
    h->fresh_ub += room;
    h->ctr_current = false;
    h->launches_total += 1;
    return HEROS_OK;
}

}  // extern "C"
