// transmit.cu -- the transmission kernels: every infectPeople evaluation of one time window, for all occupied
// sub-locations at once.
//
// Replaces, per window, the event-by-event calls
//     DiseaseTransmission.infectPeople(location, personsInSublocation, duration)
// of /root/reference/src/main/java/eu/heros/disease/Covid19TransmissionArea.java:133-196 and
// Covid19TransmissionDistance.java:115-187 (sub-location branch), driven as the medlabs engine drives them
// (SURVEY.md 3.2, C5-C8): an evaluation at every movement event t of a sub-location whose time since the last
// calculated evaluation is >= the threshold, over the occupants S(t-) = { stays : t_in < t <= t_out }.
//
// Input: the window's stays sorted by sub-location key (one contiguous segment per occupied sub-location).
//   k_transmit_stream : a thread per sub-location key walks its (contiguous, 32-byte) records once and finds who is
//                       there and the two latest event times.  Where nobody can be infected (no ill or no susceptible
//                       occupant -- the vast majority) only lastCalculation has to advance; the other sub-locations of
//                       <= 32 stays are queued for the sub-warp kernels.
//   k_transmit_sub    : 8 or 32 lanes per queued segment (<= 8 / 9..32 stays): the lanes walk the chain of calculated
//                       evaluations together (REDUX), then every lane computes ONE evaluation of the chain -- the long
//                       dependent fp64 chains (sqrt, exp) of up to 32 evaluations run side by side -- then lane = stay
//                       for the draws
// The hot path, sub-locations of > 32 stays (supermarkets, parks, lecture halls ...), in two steps so that the latency
// of the largest sub-location is not the latency of the stage:
//   k_hot_chain       : one thread block per hot sub-location: the same early-out first; where evaluations are needed
//                       the ill and the susceptible stays are copied into a partitioned segment, the movement events are
//                       ordered in shared memory (counting sort into time buckets), the successor of every event is
//                       found (the earliest later event not closer than the threshold: three buckets to look at), the
//                       chain of calculated evaluations is materialised by pointer doubling and one row per evaluation
//                       (time, duration, head count) plus one work item per 32 evaluations is left
//   k_hot_eval        : one thread block per work item, spread over the whole GPU: the ILL stays of the sub-location are
//                       streamed through shared memory by the bulk-copy engine (cp.async.bulk + mbarrier, two tiles in
//                       flight); lane = evaluation; every ill stay overlapping the item is broadcast and each lane adds
//                       its term to a 128-bit fixed-point sum in registers (order-independent, no atomics); then the
//                       threads take the SUSCEPTIBLE stays, one each, for the Philox draws
//   k_transmit_total  : one block per location of the total-location branch (TArea:198-246 / TDist:189-246)
// Output: candidate infections (person, key, time, p); engine.cu keeps the earliest per person.
#include <cstdlib>

#include "engine.cuh"

namespace heros {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int THREAD_SLOW_MAX = 8;  // largest segment handled by 8 lanes (four per warp) in k_transmit_sub

__device__ __forceinline__ double pos_inf() { return bits_f64(0x7ff0000000000000ull); }
__device__ __forceinline__ double neg_inf() { return bits_f64(0xfff0000000000000ull); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// psi / mu in force at simulator time t (policy/DiseasePolicy.java:29-43 fires before movements of the same time)
__device__ __forceinline__ void policy_at(const WindowCtx& c, double t, double& psi, double& mu)
{
    psi = c.dist.psi;
    mu = c.dist.mu;
    for (int i = 0; i < c.n_policy; ++i)
    {
        if (!(c.policy[i].t <= t)) break;
        if (c.policy[i].which == 0) psi = c.policy[i].value; else mu = c.policy[i].value;
    }
}

// A successful draw.  Only the earliest one of a person counts (engine.cu: k_resolve), so a candidate that is later
// than one already seen for the same (own) person is dropped here; equal times are kept (the lowest sub-location wins).
__device__ __forceinline__ void push_candidate(const WindowCtx& c, uint32_t person, uint32_t key, double t, double p)
{
    const uint32_t local = person - c.person_base;
    if (local < c.n_agents)
    {
        const unsigned long long tb = (unsigned long long) f64_bits(t);  // times are >= 0: their bit patterns order as integers
        if (atomicMin(&c.best_t[local], tb) < tb) return;
    }
    const unsigned long long slot = atomicAdd(&c.ctr->n_cand, 1ull);
    if (slot < c.cand_cap)
    {
        // global sub-location id: (global location id << 16) | sub-location index
        const uint32_t loc = c.key_loc[key];
        c.cand_person[slot] = person;
        c.cand_gkey[slot] = ((uint64_t) (uint32_t) (c.location_base + (int32_t) loc) << 16) | (uint64_t) (key - c.loc_first_key[loc]);
        c.cand_t[slot] = t;
        c.cand_p[slot] = p;
    }
    else
        atomicOr(&c.ctr->overflow, 1u);
}

// is the stay's person ill at time t, as infectPeople would see it (person.getDiseasePhase().isIll())?
__device__ __forceinline__ bool ill_at(const WindowCtx& c, const StayRec& r, double t)
{
    if (!phase_is_ill(view_tphase(r.view))) return false;
    if (view_ends(r.view))  // recovers / dies inside this window
        return t < ((r.pad & REC_GUEST) ? c.guest_ill_end[r.pad & ~REC_GUEST] : c.ill_end[r.person - c.person_base]);
    return true;
}

struct Local { unsigned long long evaluations = 0, draws = 0, near = 0; };

__device__ __forceinline__ void draw_and_push(const WindowCtx& c, Local& st, uint32_t person, uint32_t key, double t,
                                              double p)
{
    const double u = u01(c.seed, person, f64_bits(t), TAG_TRANSMIT);
    st.draws++;
    if (fabs(u - p) < 1e-13) st.near++;
    if (u < p) push_candidate(c, person, key, t, p);
}

__device__ __forceinline__ void flush_stats(const WindowCtx& c, Local st, int lane)
{
    st.evaluations = (unsigned long long) warp_sum((double) st.evaluations);
    st.draws = (unsigned long long) warp_sum((double) st.draws);
    st.near = (unsigned long long) warp_sum((double) st.near);
    if (lane == 0)
    {
        if (st.evaluations) atomicAdd(&c.ctr->evaluations, st.evaluations);
        if (st.draws) atomicAdd(&c.ctr->draws, st.draws);
        if (st.near) atomicAdd(&c.ctr->near_ties, st.near);
    }
}

// the two latest distinct event times seen so far
struct Top2
{
    double m1, m2;
    __device__ __forceinline__ void init() { m1 = neg_inf(); m2 = neg_inf(); }
    __device__ __forceinline__ void add(double e)
    {
        if (e > m1) { m2 = m1; m1 = e; }
        else if (e < m1 && e > m2) m2 = e;
    }
    __device__ __forceinline__ void merge(double o1, double o2)
    {
        add(o1);
        add(o2);
    }
};

// With no evaluation to compute, lastCalculation after the window is the latest event time whenever that event is at
// least a threshold after the event (or carried-in lastCalculation) before it: whatever the chain did earlier, its
// last link is then that event.  Returns false when the full chain has to be walked.
__device__ __forceinline__ bool advance_only(const Top2& t, double L0, double thr, double& L_out)
{
    const double prev = t.m2 > L0 ? t.m2 : L0;
    if (t.m1 - prev < thr) return false;
    L_out = t.m1;
    return true;
}

// p of one evaluation from the exposure sum (TArea:178-182 / TDist:169-173); false: the call returned before drawing
template <int MODEL>
__device__ __forceinline__ bool probability(double factor, double sum, double& p)
{
    if (sum == 0.0) return false;
    p = MODEL == HEROS_MODEL_AREA ? 1.0 - hexp(factor * sum) : 1.0 - hexp(-sum);
    return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// one thread per sub-location key
// ------------------------------------------------------------------------------------------------------------------
namespace {

// canonical order of the stays of a sub-location: by (person, t_in).  Exposure sums are accumulated in this order, so
// probabilities do not depend on the order in which stays were submitted, exchanged between GPUs or scattered.
__device__ __forceinline__ bool stay_before(uint32_t pa, double ta, uint32_t pb, double tb)
{
    return pa < pb || (pa == pb && ta < tb);
}

}  // namespace


// ------------------------------------------------------------------------------------------------------------------
// one thread per sub-location key
// ------------------------------------------------------------------------------------------------------------------
// A record is one 32-byte sector and the records of consecutive keys are contiguous in bucket order, so the threads of
// a warp walk neighbouring sectors: every sector that is fetched is used completely, and the independent loads of the
// (unrolled) walk are in flight together.  Measured against a warp-cooperative form (coalesced 1 KB loads + a segmented
// scan over the lanes to find who is in which sub-location: 0.105 ms at Hague size) this form takes a third of the time:
// the stream is bound by instruction issue, not by sectors, and the thread form issues the fewest.
// Segments of > 32 stays (the hot path) and of total-branch locations are skipped without being read.
template <int MODEL>
__global__ void __launch_bounds__(256) k_transmit_stream(const WindowCtx c)
{
    TlStamp tl_(c.ctr, TL_STREAM);
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    const bool exact = (c.flags & HEROS_FLAG_EXACT_STATS) != 0;
    unsigned occupied = 0;

    for (uint32_t key = blockIdx.x * blockDim.x + threadIdx.x; key < c.n_keys; key += gridDim.x * blockDim.x)
    {
        const uint32_t a = c.seg_start[key], b = c.seg_start[key + 1];
        const uint32_t n = b - a;
        if (n == 0) continue;
        ++occupied;
        if (n > 32) continue;  // queued for the hot path by the prefix-sum kernel
        if (c.key_total && c.key_total[key]) continue;  // evaluated per location by k_transmit_total
        if (!HEROS_CHECK(c.ctr, a <= b && b <= c.rec_cap)) continue;
        const StayRec* rec = c.rec + a;
        // who is here, and the two latest events
        bool any_ill = false, any_sus = false;
        Top2 top;
        top.init();
#pragma unroll 4
        for (uint32_t j = 0; j < n; ++j)
        {
            const StayRec r = ld_rec(rec + j);
            const uint32_t tph = view_tphase(r.view);
            any_ill |= phase_is_ill(tph);
            any_sus |= phase_is_susceptible(tph);
            if (r.tout >= c.T0 && r.tout < c.T1) top.add(r.tout);
            if (both && r.tin >= c.T0 && r.tin < c.T1) top.add(r.tin);
        }
        if (!(top.m1 > neg_inf())) continue;  // no movement in this window
        const bool need_eval = any_ill && any_sus;
        if (!need_eval && !exact)
        {
            double L;
            if (advance_only(top, c.last_calc[key], thr, L))
            {
                c.last_calc[key] = L;
                continue;
            }
        }
        // everything else: the sub-warp kernels (8 lanes per sub-location of <= 8 stays, 32 lanes otherwise)
        const int which = n > THREAD_SLOW_MAX ? 1 : 0;
        const unsigned long long slot = atomicAdd(&c.ctr->n_sub[which], 1ull);
        if (slot < c.sub_cap[which]) c.sub_seg[which][slot] = key;
        else atomicOr(&c.ctr->overflow, 2u);
    }
    occupied = __reduce_add_sync(FULL, occupied);
    if ((threadIdx.x & 31) == 0 && occupied)
    {
        atomicAdd(&c.ctr->n_seg, (unsigned long long) occupied);
        atomicAdd(&c.ctr->seg_total, (unsigned long long) occupied);
    }
}
// ------------------------------------------------------------------------------------------------------------------
// S = 8 or 32 lanes per queued sub-location (<= 8 / 9..32 stays); lane i of the group holds stay i (canonical order)
// ------------------------------------------------------------------------------------------------------------------
template <int MODEL, int S>
__global__ void __launch_bounds__(256, 4) k_transmit_sub(const WindowCtx c)
{
    TlStamp tl_(c.ctr, S == 8 ? TL_SUB8 : TL_SUB32);
    constexpr int GPW = 32 / S;  // groups per warp
    __shared__ StayRec s_rec[8][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane / S, gl = lane % S, g0 = g * S;
    const unsigned gmask = S == 32 ? FULL : (((1u << S) - 1u) << g0);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    constexpr int WHICH = S == 8 ? 0 : 1;
    const uint32_t n_list = (uint32_t) min((unsigned long long) c.sub_cap[WHICH], c.ctr->n_sub[WHICH]);
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    StayRec* my = &s_rec[wib][g0];
    Local st;

    for (uint32_t w0 = warp * GPW; w0 < n_list; w0 += n_warps * GPW)
    {
        const uint32_t w = w0 + g;
        const bool valid = w < n_list;
        uint32_t key = 0, a = 0, n = 0;
        if (valid)
        {
            key = c.sub_seg[WHICH][w];
            a = c.seg_start[key];
            n = c.seg_start[key + 1] - a;
            if (!HEROS_CHECK(c.ctr, key < c.n_keys && n <= (uint32_t) S && a + n <= c.rec_cap)) n = 0;
        }
        const bool have = (uint32_t) gl < n;
        StayRec r;
        r.tin = pos_inf(); r.tout = neg_inf(); r.view = 0; r.person = 0xFFFFFFFFu; r.pad = 0;
        if (have) r = c.rec[a + gl];
        {
            // canonical order: lane <- stay with that rank by (person, t_in); empty lanes sort last
            int rank = 0;
#pragma unroll
            for (int j = 0; j < S; ++j)
            {
                const uint32_t pj = __shfl_sync(FULL, r.person, g0 + j);
                const double tj = __shfl_sync(FULL, r.tin, g0 + j);
                rank += (stay_before(pj, tj, r.person, r.tin) || (pj == r.person && tj == r.tin && j < gl)) ? 1 : 0;
            }
            __syncwarp();
            my[rank] = r;
            __syncwarp();
            r = my[gl];  // my[0..S) keeps the group's stays for the evaluations below
        }
        const uint32_t tph = view_tphase(r.view);
        const bool here = r.person != 0xFFFFFFFFu;
        const bool susceptible = here && phase_is_susceptible(tph);
        const unsigned ill_b = __ballot_sync(FULL, here && phase_is_ill(tph)), sus_b = __ballot_sync(FULL, susceptible);
        const bool need_eval = (ill_b & gmask) != 0 && (sus_b & gmask) != 0;
        LocParam lp;
        lp.area_sub = lp.denom = 1.0;
        double L = 0.0;
        if (valid) { lp = c.loc_param[c.key_loc[key]]; L = c.last_calc[key]; }
        const double L_in = L;
        // movement events of this window that can trigger an evaluation
        double e_in = (both && r.tin >= c.T0 && r.tin < c.T1) ? r.tin : pos_inf();
        double e_out = (r.tout >= c.T0 && r.tout < c.T1) ? r.tout : pos_inf();

        for (;;)
        {
            // 1. the lanes walk the next (up to) S links of the chain together: the next evaluation is the earliest
            //    event whose time since the last calculation is not below the threshold.  Lane k keeps evaluation k.
            double my_now = pos_inf(), my_dur = 0.0;
            int n_ev = 0;
#pragma unroll 1
            for (int k = 0; k < S; ++k)
            {
                const double c_in = !(e_in - L < thr) ? e_in : pos_inf();
                const double c_out = !(e_out - L < thr) ? e_out : pos_inf();
                const double cand = c_in < c_out ? c_in : c_out;
                const uint64_t cb = f64_bits(cand);
                const uint32_t hi = (uint32_t) (cb >> 32), lo = (uint32_t) cb;
                const uint32_t mhi = __reduce_min_sync(gmask, hi);
                const uint32_t mlo = __reduce_min_sync(gmask, hi == mhi ? lo : 0xffffffffu);
                const double now = bits_f64(((uint64_t) mhi << 32) | mlo);
                const bool got = now < pos_inf();
                if (__all_sync(FULL, !got)) break;
                if (got)
                {
                    if (gl == k) { my_now = now; my_dur = now - L; }
                    L = now;
                    ++n_ev;
                    if (e_in <= now) e_in = pos_inf();
                    if (e_out <= now) e_out = pos_inf();
                }
            }
            if (gl == 0) st.evaluations += (unsigned long long) n_ev;
            if (!__any_sync(FULL, n_ev > 0)) break;
            if (__any_sync(FULL, need_eval && n_ev > 0))
            {
                // 2. every lane computes its own evaluation over the stays of its group (TArea:155-182 / TDist:137-173)
                double p = 0.0;
                if (need_eval && my_now < pos_inf())
                {
                    double factor;
                    if (MODEL == HEROS_MODEL_AREA)
                        factor = area_factor(c.area, my_dur, lp.denom);  // TArea:156
                    else
                    {
                        int head_count = 0;  // personsInSublocation.size()
#pragma unroll
                        for (int j = 0; j < S; ++j) head_count += (my[j].tin < my_now && my_now <= my[j].tout) ? 1 : 0;
                        double psi, mu;
                        policy_at(c, my_now, psi, mu);
                        factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);  // TDist:138-143
                    }
                    if (factor != 0.0)  // TArea:157 / TDist:144
                    {
                        double sum = 0.0;
#pragma unroll 1
                        for (int j = 0; j < S; ++j)
                        {
                            if (!((ill_b >> (g0 + j)) & 1u)) continue;
                            const StayRec q = my[j];
                            if (q.tin < my_now && my_now <= q.tout && ill_at(c, q, my_now))
                            {
                                const double te = my_now - (double) view_exposure(q.view);
                                if (MODEL == HEROS_MODEL_AREA)
                                    sum += area_contribution(c.area, te);  // TArea:167-172
                                else
                                {
                                    const double v_t = dist_viral_load(c.dist, te);                  // TDist:155-159
                                    if (v_t > 0) sum += dist_term(c.dist, factor, my_dur, v_t);      // TDist:161-164
                                }
                            }
                        }
                        if (!(probability<MODEL>(factor, sum, p) && p > 0)) p = 0.0;
                    }
                }
                // 3. draws: lane = stay again, against the evaluations of the batch it was present at
                const unsigned pos = __ballot_sync(FULL, p > 0.0);
#pragma unroll 1
                for (int k = 0; k < S; ++k)
                {
                    unsigned at_k = 0;  // any group whose k-th evaluation has p > 0 ?
#pragma unroll
                    for (int q = 0; q < GPW; ++q) at_k |= (pos >> (q * S + k)) & 1u;
                    if (!at_k) continue;
                    const double now_k = __shfl_sync(FULL, my_now, g0 + k);
                    const double p_k = __shfl_sync(FULL, p, g0 + k);
                    if (p_k > 0 && susceptible && r.tin < now_k && now_k <= r.tout)
                        draw_and_push(c, st, r.person, key, now_k, p_k);  // TArea:190 / TDist:181
                }
            }
            if (__all_sync(FULL, n_ev < S)) break;
        }
        if (valid && gl == 0 && L != L_in) c.last_calc[key] = L;
        __syncwarp();  // the staged stays are overwritten by the next sub-location
    }
    flush_stats(c, st, lane);
}

// ------------------------------------------------------------------------------------------------------------------
// block-wide building blocks of the hot path
// ------------------------------------------------------------------------------------------------------------------
namespace {

// G = 32: the group is a warp (several groups per block, no block barrier is ever used); G > 32: the whole block
template <int G> __device__ __forceinline__ void gsync() { if (G == 32) __syncwarp(); else __syncthreads(); }
template <int G> __device__ __forceinline__ int grank() { return G == 32 ? (int) (threadIdx.x & 31) : (int) threadIdx.x; }

// sum of v over the group, returned to every thread.  s_w: 33 ints of shared memory per block (unused for G = 32)
template <int G>
__device__ __forceinline__ int group_sum(int v, int* s_w)
{
    v = __reduce_add_sync(FULL, v);
    if (G == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = (threadIdx.x & 31) < (G >> 5) ? s_w[threadIdx.x & 31] : 0;
    return __reduce_add_sync(FULL, t);
}

// out[i] = sum_{k<i} f(k) for i < n (exclusive), returns the total.  All threads of the group must call.  `out` may be
// the storage f reads from (every thread reads an element before it overwrites it, chunks are private).
template <int G, class F, class T>
__device__ int group_exclusive_scan(int n, F f, T* out, int* s_w)
{
    const int tid = grank<G>();
    const int chunk = (n + G - 1) / G;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int local = 0;
    for (int i = lo; i < hi; ++i) local += f(i);
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int v = __shfl_up_sync(FULL, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    int warp_off = 0, total;
    if (G == 32)
        total = __shfl_sync(FULL, incl, 31);
    else
    {
        __syncthreads();
        if ((tid & 31) == 31) s_w[tid >> 5] = incl;
        __syncthreads();
        // every warp scans the (at most 32) warp totals
        const int wt = (tid & 31) < (G >> 5) ? s_w[tid & 31] : 0;
        int wi = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(FULL, wi, o);
            if ((tid & 31) >= o) wi += v;
        }
        total = __shfl_sync(FULL, wi, 31);
        warp_off = __shfl_sync(FULL, wi - wt, tid >> 5);
    }
    int run = warp_off + incl - local;
    for (int i = lo; i < hi; ++i)
    {
        const int v = f(i);
        out[i] = (T) run;
        run += v;
    }
    gsync<G>();
    return total;
}

// out[i] = min_{k >= i} f(k) for i < n (f returns an int <= INT_MAX).  All threads of the group must call.
template <int G, class F, class T>
__device__ void group_suffix_min(int n, F f, T* out, int* s_w)
{
    const int tid = grank<G>();
    const int chunk = (n + G - 1) / G;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int local = 0x7fffffff;
    for (int i = lo; i < hi; ++i) local = min(local, f(i));
    int incl = local;  // min over the lanes >= mine of my warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int v = __shfl_down_sync(FULL, incl, o);
        if ((tid & 31) + o < 32) incl = min(incl, v);
    }
    int after = 0x7fffffff;  // min over the warps above mine
    if (G > 32)
    {
        __syncthreads();
        if ((tid & 31) == 0) s_w[tid >> 5] = incl;
        __syncthreads();
        const int wv = (tid & 31) < (G >> 5) ? s_w[tid & 31] : 0x7fffffff;
        int ws = wv;  // suffix-min over the warp values
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_down_sync(FULL, ws, o);
            if ((tid & 31) + o < 32) ws = min(ws, v);
        }
        const int nxt = __shfl_sync(FULL, ws, min((tid >> 5) + 1, 31));
        after = (tid >> 5) + 1 < (G >> 5) ? nxt : 0x7fffffff;
    }
    const int above_in_warp = __shfl_down_sync(FULL, incl, 1);
    int run = after;
    if ((tid & 31) < 31) run = min(run, above_in_warp);
    for (int i = hi - 1; i >= lo; --i)
    {
        run = min(run, f(i));
        out[i] = (T) run;
    }
    gsync<G>();
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// the hot path: sub-locations of > 32 stays
// ------------------------------------------------------------------------------------------------------------------

namespace {

template <int CLS> struct HotCfg;
template <> struct HotCfg<0> { static constexpr int T = 256, M_CAP = HOT_M_CAP0, NB_CAP = HOT_NB_CAP0; using idx_t = uint16_t; };
template <> struct HotCfg<1> { static constexpr int T = 512, M_CAP = HOT_M_CAP1, NB_CAP = HOT_NB_CAP; using idx_t = uint16_t; };
template <> struct HotCfg<2> { static constexpr int T = 512, M_CAP = 0, NB_CAP = HOT_NB_CAP; using idx_t = uint32_t; };

template <int CLS> __host__ __device__ constexpr size_t hot_smem_bytes()
{
    using Cfg = HotCfg<CLS>;
    return CLS == HOT_HUGE ? 0
                           : (size_t) Cfg::M_CAP * 8 + 5 * 4 * (size_t) (Cfg::NB_CAP + 2) +
                                 3 * sizeof(typename Cfg::idx_t) * (size_t) (Cfg::M_CAP + 1) + (size_t) Cfg::M_CAP + 64;
}

}  // namespace

// The chain of calculated evaluations of one hot sub-location.  From an evaluation at time e the next one is the
// earliest later candidate event t with !(t - e < threshold) (candidates: exits, and enters when the trigger is
// ENTER_AND_EXIT).  The movement events of the window are put into NB time buckets (counting sort; the order inside a
// bucket is never needed).  With x = e + threshold falling into bucket b, every event of a bucket >= b + 2 qualifies and
// no event of a bucket <= b - 2 does (a bucket is many orders of magnitude wider than the rounding of t - e), so the
// successor is the earliest qualifying event of the buckets b - 1 .. b + 1, tested with the reference's own
// comparison, or else the earliest candidate of the next bucket that holds one.  The chain start, succ[start], ... is
// materialised in O(log R) rounds of pointer doubling.  Head counts of the distance model (TDist:137: the size of the
// set S(t-)) = stays carried in + enters - exits before t: a bucket prefix plus the earlier events of t's own bucket.
template <int MODEL, int CLS>
__global__ void __launch_bounds__(HotCfg<CLS>::T) k_hot_chain(const WindowCtx c)
{
    TlStamp tl_(c.ctr, TL_CHAIN0 + CLS);
    using Cfg = HotCfg<CLS>;
    using idx_t = typename Cfg::idx_t;
    constexpr int G = Cfg::T;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_w[33];
    __shared__ unsigned long long s_base[2];
    __shared__ double s_top[2][32];
    __shared__ uint32_t s_pstart[2][PART_SLOTS + 1], s_pcur[2][PART_SLOTS];
    __shared__ unsigned long long s_pmax[2][PART_SLOTS];
    const int tid = threadIdx.x;
    const uint32_t n_list = (uint32_t) min((unsigned long long) c.hot_cap, c.ctr->n_hot[CLS]);
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    const bool need_enters = both || MODEL == HEROS_MODEL_DISTANCE;
    const bool exact = (c.flags & HEROS_FLAG_EXACT_STATS) != 0;
    Local st;
    if (blockIdx.x == 0 && tid == 0 && n_list) atomicAdd(&c.ctr->hot_total, (unsigned long long) n_list);

    __shared__ uint32_t s_next;
    for (;;)
    {
        // the next sub-location of the class: handed out one at a time (their cost differs by orders of magnitude)
        __syncthreads();
        if (tid == 0) s_next = (uint32_t) min(atomicAdd(&c.ctr->chain_next[CLS], 1ull), 0xFFFFFFFFull);
        __syncthreads();
        const uint32_t w = s_next;
        if (w >= n_list) break;
        const uint32_t key = c.hot_seg[CLS][w];
        if (key == KEY_NONE) continue;  // no scratch left for it (the overflow bit is set)
        const uint32_t a = c.seg_start[key];
        const int n = (int) (c.seg_start[key + 1] - a);
        if (!HEROS_CHECK(c.ctr, key < c.n_keys && n > 32 && a + (uint32_t) n <= c.rec_cap &&
                                    (CLS == HOT_HUGE || (need_enters ? 2 * n : n) <= (int) HotCfg<CLS>::M_CAP)))
            continue;
        const StayRec* rec = c.rec + a;
        const double L0 = c.last_calc[key];
        // 0. who is here, and the two latest movement events.  Where nobody can be infected (the majority while the
        //    prevalence is low) only lastCalculation has to advance; that usually follows from those two times alone.
        int n_ill = 0, n_sus = 0;
        {
            Top2 top;
            top.init();
            for (int v = tid; v < n; v += G)
            {
                const StayRec r = ld_rec(rec + v);
                const uint32_t tph = view_tphase(r.view);
                n_ill += phase_is_ill(tph) ? 1 : 0;
                n_sus += phase_is_susceptible(tph) ? 1 : 0;
                if (r.tout >= c.T0 && r.tout < c.T1) top.add(r.tout);
                if (both && r.tin >= c.T0 && r.tin < c.T1) top.add(r.tin);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const double o1 = __shfl_xor_sync(FULL, top.m1, o), o2 = __shfl_xor_sync(FULL, top.m2, o);
                top.merge(o1, o2);
            }
            __syncthreads();  // the previous sub-location is finished with the shared arrays
            if ((tid & 31) == 0) { s_top[0][tid >> 5] = top.m1; s_top[1][tid >> 5] = top.m2; }
            n_ill = group_sum<G>(n_ill, s_w);
            n_sus = group_sum<G>(n_sus, s_w);
            __syncthreads();
#pragma unroll 1
            for (int q = 0; q < (G >> 5); ++q) top.merge(s_top[0][q], s_top[1][q]);
            if (!(top.m1 > neg_inf())) continue;  // no movement in this window (the same in every thread)
            if (!(n_ill > 0 && n_sus > 0) && !exact)
            {
                double L;
                if (advance_only(top, L0, thr, L))
                {
                    if (tid == 0) c.last_calc[key] = L;
                    continue;
                }
            }
        }
        const bool need_eval = n_ill > 0 && n_sus > 0;
        uint32_t part_idx = 0xFFFFFFFFu;
        if (need_eval)
        {
            // The evaluation kernel only ever looks at the ill and the susceptible stays: a partitioned copy of the segment
            // -- the ill ones from the front, the susceptible ones behind the others at the back -- each list ordered by the
            // time slot in which the stay began (a counting sort over PART_SLOTS slots; the order inside a slot is free:
            // exposure sums are formed in fixed point, draws are keyed by person and time), with the latest t_out per slot.
            for (int i = tid; i < 2 * (PART_SLOTS + 1); i += G) s_pstart[i / (PART_SLOTS + 1)][i % (PART_SLOTS + 1)] = 0u;
            for (int i = tid; i < 2 * PART_SLOTS; i += G) s_pmax[i / PART_SLOTS][i % PART_SLOTS] = 0ull;
            if (tid == 0) s_base[0] = atomicAdd(&c.ctr->n_part, 1ull);
            __syncthreads();
            const double pscale = (double) PART_SLOTS / (c.T1 - c.T0);
            auto pslot = [&](double tin) -> int {
                const double f = floor((tin - c.T0) * pscale);
                return f < 0.0 ? 0 : (f > (double) (PART_SLOTS - 1) ? PART_SLOTS - 1 : (int) f);
            };
            for (int v = tid; v < n; v += G)
            {
                const StayRec r = ld_rec(rec + v);
                const uint32_t tph = view_tphase(r.view);
                const int kind = phase_is_ill(tph) ? 0 : (phase_is_susceptible(tph) ? 1 : -1);
                if (kind < 0) continue;
                const int b = pslot(r.tin);
                atomicAdd(&s_pstart[kind][b + 1], 1u);
                // t_out >= t_in >= 0 (or +inf): the bit patterns order as integers; 0 = no stay in the slot
                atomicMax(&s_pmax[kind][b], (unsigned long long) f64_bits(r.tout));
            }
            __syncthreads();
            if (tid < 2)  // 32 slots: one thread per list
            {
                uint32_t run = 0;
                for (int b = 0; b <= PART_SLOTS; ++b) { run += s_pstart[tid][b]; s_pstart[tid][b] = run; }
            }
            __syncthreads();
            for (int i = tid; i < 2 * PART_SLOTS; i += G) s_pcur[i / PART_SLOTS][i % PART_SLOTS] = s_pstart[i / PART_SLOTS][i % PART_SLOTS];
            __syncthreads();
            StayRec* part = c.rec_part + a;
            for (int v = tid; v < n; v += G)
            {
                const StayRec r = ld_rec(rec + v);
                const uint32_t tph = view_tphase(r.view);
                const int kind = phase_is_ill(tph) ? 0 : (phase_is_susceptible(tph) ? 1 : -1);
                if (kind < 0) continue;
                const int slot = (int) atomicAdd(&s_pcur[kind][pslot(r.tin)], 1u);
                const int at = kind == 0 ? slot : n - n_sus + slot;
                if (HEROS_CHECK(c.ctr, slot >= 0 && slot < (kind == 0 ? n_ill : n_sus))) st_rec(part + at, r);
            }
            const unsigned long long pi = s_base[0];
            if (pi < c.part_cap)
            {
                part_idx = (uint32_t) pi;
                PartTable& T = c.part_table[pi];
                for (int i = tid; i < 2 * (PART_SLOTS + 1); i += G) T.start[i / (PART_SLOTS + 1)][i % (PART_SLOTS + 1)] = s_pstart[i / (PART_SLOTS + 1)][i % (PART_SLOTS + 1)];
                for (int i = tid; i < 2 * PART_SLOTS; i += G)
                {
                    const unsigned long long m = s_pmax[i / PART_SLOTS][i % PART_SLOTS];
                    T.max_tout[i / PART_SLOTS][i % PART_SLOTS] = m ? bits_f64(m) : neg_inf();
                }
            }
            else if (tid == 0)
                atomicOr(&c.ctr->overflow, 2u);
        }
        const int slots = need_enters ? 2 * n : n;
        int NB = 32;
        while (NB * 2 < slots && NB < Cfg::NB_CAP) NB <<= 1;
        const double scale = (double) NB / (c.T1 - c.T0);
        auto bucket_raw = [&](double t) -> int {
            const double f = floor((t - c.T0) * scale);
            return f < -4.0 ? -4 : (f > (double) (NB + 4) ? NB + 4 : (int) f);
        };
        auto bucket_of = [&](double t) { return min(max(bucket_raw(t), 0), NB - 1); };
        // scratch: ev_t f64[M] | bstart u32[NB + 2] | bcur i32[NB + 2] | bsign i32[NB + 2] | bminc u32[NB + 2] | nextc u32[NB + 2] |
        //          pa idx[M + 1] | pb idx[M + 1] | node idx[M + 1] | ev_k u8[M]
        const size_t M = CLS == HOT_HUGE ? (size_t) slots : (size_t) Cfg::M_CAP;
        unsigned char* base = CLS == HOT_HUGE ? c.scratch + c.huge_off[w] : smem;
        double* ev_t = (double*) base;
        uint32_t* bstart = (uint32_t*) (ev_t + M);
        int32_t* bcur = (int32_t*) (bstart + NB + 2);
        int32_t* bsign = bcur + NB + 2;
        uint32_t* bminc = (uint32_t*) (bsign + NB + 2);
        uint32_t* nextc = bminc + NB + 2;
        idx_t* pa = (idx_t*) (nextc + NB + 2);
        idx_t* pb = pa + (M + 1);
        idx_t* node = pb + (M + 1);
        uint8_t* ev_k = (uint8_t*) (node + (M + 1));

        __syncthreads();  // the previous sub-location is finished with the arrays
        // 1. histogram of the window's movement events over the time buckets
        for (int i = tid; i < NB + 2; i += G) { bstart[i] = 0; bsign[i] = 0; }
        __syncthreads();
        int carried = 0;
        for (int v = tid; v < n; v += G)
        {
            const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(rec + v));
            const double tin = bits_f64(((uint64_t) h0.y << 32) | h0.x), tout = bits_f64(((uint64_t) h0.w << 32) | h0.z);
            carried += tin < c.T0 ? 1 : 0;
            if (tout >= c.T0 && tout < c.T1)
            {
                const int bk = bucket_of(tout);
                atomicAdd(&bstart[bk], 1u);
                if (MODEL == HEROS_MODEL_DISTANCE) atomicAdd(&bsign[bk], -1);
            }
            if (need_enters && tin >= c.T0 && tin < c.T1)
            {
                const int bk = bucket_of(tin);
                atomicAdd(&bstart[bk], 1u);
                if (MODEL == HEROS_MODEL_DISTANCE) atomicAdd(&bsign[bk], 1);
            }
        }
        if (MODEL == HEROS_MODEL_DISTANCE) carried = group_sum<G>(carried, s_w);
        __syncthreads();
        // 2. bucket offsets
        const int m_ev = group_exclusive_scan<G>(NB, [&](int i) { return (int) bstart[i]; }, bstart, s_w);
        if (tid == 0) { bstart[NB] = (uint32_t) m_ev; bstart[NB + 1] = (uint32_t) m_ev; }
        for (int i = tid; i < NB; i += G) bcur[i] = (int32_t) bstart[i];
        __syncthreads();
        // 3. the events into their buckets (the order inside a bucket is irrelevant to every result)
        for (int v = tid; v < n; v += G)
        {
            const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(rec + v));
            const double tin = bits_f64(((uint64_t) h0.y << 32) | h0.x), tout = bits_f64(((uint64_t) h0.w << 32) | h0.z);
            if (tout >= c.T0 && tout < c.T1)
            {
                const int pos = atomicAdd(&bcur[bucket_of(tout)], 1);
                if (!HEROS_CHECK(c.ctr, pos >= 0 && pos < m_ev && (size_t) pos < M)) continue;
                ev_t[pos] = tout;
                ev_k[pos] = 0;
            }
            if (need_enters && tin >= c.T0 && tin < c.T1)
            {
                const int pos = atomicAdd(&bcur[bucket_of(tin)], 1);
                if (!HEROS_CHECK(c.ctr, pos >= 0 && pos < m_ev && (size_t) pos < M)) continue;
                ev_t[pos] = tin;
                ev_k[pos] = 1;
            }
        }
        __syncthreads();
        const uint32_t END = (uint32_t) m_ev;
        auto is_cand = [&](int j) { return both || ev_k[j] == 0; };
        // 4. per bucket: its earliest candidate; occupancy change before the bucket (distance model)
        for (int b = tid; b < NB; b += G)
        {
            uint32_t best = END;
            double best_t = pos_inf();
            for (int j = (int) bstart[b]; j < (int) bstart[b + 1]; ++j)
                if (is_cand(j) && ev_t[j] < best_t) { best_t = ev_t[j]; best = (uint32_t) j; }
            bminc[b] = best;
        }
        __syncthreads();
        if (MODEL == HEROS_MODEL_DISTANCE)
            group_exclusive_scan<G>(NB, [&](int i) { return bsign[i]; }, bcur, s_w);
        group_suffix_min<G>(NB, [&](int b) { return bminc[b] != END ? b : NB; }, nextc, s_w);
        if (tid == 0) { nextc[NB] = (uint32_t) NB; nextc[NB + 1] = (uint32_t) NB; }
        __syncthreads();
        // successor of an evaluation at time e: the earliest candidate t > e with !(t - e < thr)
        auto successor = [&](double e) -> uint32_t {
            const int b = bucket_raw(e + thr);
            uint32_t best = END;
            double best_t = pos_inf();
            const int b_lo = max(b - 1, 0), b_hi = min(b + 1, NB - 1);
            if (b_lo <= b_hi)
                for (int j = (int) bstart[b_lo]; j < (int) bstart[b_hi + 1]; ++j)
                {
                    const double t = ev_t[j];
                    if (is_cand(j) && t > e && !(t - e < thr) && t < best_t) { best_t = t; best = (uint32_t) j; }
                }
            if (best != END) return best;
            const int nb = (int) nextc[min(max(b + 2, 0), NB)];
            return nb < NB ? bminc[nb] : END;
        };
        // 5. next pointers, chain start, pointer doubling
        for (int i = tid; i <= m_ev; i += G) pa[i] = (idx_t) ((i < m_ev && is_cand(i)) ? successor(ev_t[i]) : END);
        const uint32_t start = successor(L0);
        int r_cap = m_ev;
        if (thr > 0)
        {
            const double bound = (c.T1 - c.T0) / thr + 2.0;
            if (bound < (double) r_cap) r_cap = (int) bound;
        }
        for (int r = tid; r < r_cap; r += G) node[r] = (idx_t) start;
        __syncthreads();
        {
            idx_t* qa = pa;
            idx_t* qb = pb;
            for (int k = 0; (1 << k) < r_cap; ++k)
            {
                for (int r = tid; r < r_cap; r += G)
                    if ((r >> k) & 1) node[r] = qa[node[r]];
                for (int i = tid; i <= m_ev; i += G) qb[i] = qa[qa[i]];
                __syncthreads();
                idx_t* t = qa; qa = qb; qb = t;
            }
        }
        int R;  // chain length: node[r] != END for r < R
        {
            int lo = 0, hi = r_cap;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((uint32_t) node[mid] == END) hi = mid; else lo = mid + 1; }
            R = lo;
        }
        if (R == 0) continue;
        if (tid == 0)
        {
            st.evaluations += (unsigned long long) R;
            c.last_calc[key] = ev_t[node[R - 1]];
        }
        if (!need_eval || part_idx == 0xFFFFFFFFu) continue;
        // 6. one row per evaluation (time, duration, head count), one work item per 32 rows
        const int n_chunks = (R + 31) >> 5;
        if (tid == 0)
        {
            s_base[0] = atomicAdd(&c.ctr->tab_top, (unsigned long long) R);
            s_base[1] = atomicAdd(&c.ctr->n_items, (unsigned long long) n_chunks);
        }
        __syncthreads();
        const unsigned long long tab0 = s_base[0], item0 = s_base[1];
        if (tab0 + (unsigned long long) R > c.tab_cap || item0 + (unsigned long long) n_chunks > c.item_cap)
        {
            if (tid == 0) atomicOr(&c.ctr->overflow, 2u);
            continue;
        }
        for (int r = tid; r < R; r += G)
        {
            const double now = ev_t[node[r]];
            c.tab_now[tab0 + r] = now;
            c.tab_dur[tab0 + r] = now - (r ? ev_t[node[r - 1]] : L0);
            int head_count = 0;
            if (MODEL == HEROS_MODEL_DISTANCE)
            {
                // occupants before any event of time `now`: bucket prefix + the earlier events of the bucket
                const int bk = bucket_of(now);
                head_count = carried + bcur[bk];
                for (int j = (int) bstart[bk]; j < (int) bstart[bk + 1]; ++j)
                    if (ev_t[j] < now) head_count += ev_k[j] ? 1 : -1;
            }
            c.tab_head[tab0 + r] = (uint32_t) head_count;
        }
        for (int ch = tid; ch < n_chunks; ch += G)
        {
            c.item_key[item0 + ch] = key;
            c.item_tab[item0 + ch] = tab0 + 32ull * (unsigned long long) ch;
            c.item_cnt[item0 + ch] = (uint32_t) min(32, R - 32 * ch);
            c.item_nill[item0 + ch] = (uint32_t) n_ill;
            c.item_nsus[item0 + ch] = (uint32_t) n_sus;
            c.item_part[item0 + ch] = part_idx;
        }
    }
    flush_stats(c, st, threadIdx.x & 31);
}

namespace {

// ---- bulk asynchronous copies global -> shared (the TMA unit's 1-D form) completing on an mbarrier ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

constexpr int EVAL_T = 128;      // threads of an evaluation block (4 warps)
constexpr int EVAL_TILE = 128;   // records per tile (4 KB): every warp takes a quarter of every tile
constexpr int EVAL_STAGES = 2;   // tiles in flight

}  // namespace

// One thread block per work item = up to 32 consecutive calculated evaluations of one hot sub-location.
// 1. Exposure sums, evaluation-parallel: lane k of every warp stands for evaluation k.  The ILL stays of the sub-location
//    (the front of its partitioned copy) are streamed through shared memory, tile by tile, by the bulk-copy engine
//    (cp.async.bulk + mbarrier, two tiles in flight); those that overlap the item's time span are broadcast lane to lane
//    and every lane adds the term of its evaluation (TArea:167-172 / TDist:155-164) to a 128-bit fixed-point sum in
//    registers -- exact integer additions, so neither the order of the stays nor the split over the four warps can change
//    the result; the four partial sums are merged in shared memory, p follows (TArea:182 / TDist:173).
// 2. Draws, stay-parallel: every thread takes SUSCEPTIBLE stays (the back of the partitioned copy, coalesced 32-byte
//    records) and draws against the evaluations with p > 0 it was present at (TArea:190 / TDist:181).
template <int MODEL>
__global__ void __launch_bounds__(EVAL_T) k_hot_eval(const WindowCtx c)
{
    TlStamp tl_(c.ctr, TL_EVAL);
    __shared__ __align__(128) StayRec s_tile[EVAL_STAGES][EVAL_TILE];
    __shared__ __align__(8) uint64_t s_bar[EVAL_STAGES];
    __shared__ long long s_hi[EVAL_T / 32][32];
    __shared__ unsigned long long s_lo[EVAL_T / 32][32];
    __shared__ double s_loose[EVAL_T / 32][32];
    __shared__ double s_now[32], s_p[32];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const unsigned long long n_items = min(c.ctr->n_items, c.item_cap);
    uint32_t parity[EVAL_STAGES];
#pragma unroll
    for (int s = 0; s < EVAL_STAGES; ++s) parity[s] = 0u;
    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < EVAL_STAGES; ++s) mbar_init(&s_bar[s], 1u);
        mbar_fence_init();
    }
    __syncthreads();
    Local st;

    // streams the records [0, n) through the tiles; `body(r, v)` is called by every lane with the record of its slot
    // (all lanes of a warp call together; v may be >= n: then r is undefined and must be ignored)
    auto for_each_record = [&](const StayRec* rec, int n, auto&& body) {
        const int n_tiles = (n + EVAL_TILE - 1) / EVAL_TILE;
        auto issue = [&](int j) {
            const int s = j % EVAL_STAGES;
            const uint32_t bytes = (uint32_t) min(EVAL_TILE, n - j * EVAL_TILE) * (uint32_t) sizeof(StayRec);
            mbar_expect_tx(&s_bar[s], bytes);
            bulk_g2s(&s_tile[s][0], rec + (size_t) j * EVAL_TILE, bytes, &s_bar[s]);
        };
        if (tid == 0)
            for (int j = 0; j < EVAL_STAGES && j < n_tiles; ++j) issue(j);
        for (int j = 0; j < n_tiles; ++j)
        {
            const int s = j % EVAL_STAGES;
            mbar_wait(&s_bar[s], parity[s]);
            parity[s] ^= 1u;
            const int slot = lane * (EVAL_T / 32) + wib;  // neighbouring records go to different warps: few ill stays are shared out
            const int v = j * EVAL_TILE + slot;
            StayRec r;
            {
                const uint4* q = reinterpret_cast<const uint4*>(&s_tile[s][slot]);
                const uint4 h0 = q[0], h1 = q[1];
                r.tin = bits_f64(((uint64_t) h0.y << 32) | h0.x); r.tout = bits_f64(((uint64_t) h0.w << 32) | h0.z);
                r.view = ((uint64_t) h1.y << 32) | h1.x; r.person = h1.z; r.pad = h1.w;
            }
            body(r, v);
            __syncthreads();  // everybody is done with the tile: its buffer can be filled again
            if (tid == 0 && j + EVAL_STAGES < n_tiles) issue(j + EVAL_STAGES);
        }
    };

    __shared__ unsigned long long s_item;
    for (;;)
    {
        // the next work item: handed out one at a time (a corner shop at night, a supermarket at noon)
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(&c.ctr->eval_next, 1ull);
        __syncthreads();
        const unsigned long long it = s_item;
        if (it >= n_items) break;
        const uint32_t key = c.item_key[it];
        const unsigned long long tab = c.item_tab[it];
        const int cnt = (int) c.item_cnt[it];
        const int n_ill = (int) c.item_nill[it], n_sus = (int) c.item_nsus[it];
        const uint32_t a = c.seg_start[key];
        const int n = (int) (c.seg_start[key + 1] - a);
        const StayRec* part = c.rec_part + a;  // ill stays [0, n_ill), susceptible stays [n - n_sus, n)
        if (!HEROS_CHECK(c.ctr, n_ill >= 0 && n_sus >= 0 && n_ill + n_sus <= n && a + (uint32_t) n <= c.rec_cap && cnt >= 1 && cnt <= 32 &&
                                    tab + (unsigned long long) cnt <= c.tab_cap && (unsigned long long) c.item_part[it] < c.part_cap))
            continue;
        const LocParam lp = c.loc_param[c.key_loc[key]];
        const bool active = lane < cnt;
        const double now = active ? c.tab_now[tab + lane] : pos_inf();
        const double duration = active ? c.tab_dur[tab + lane] : 0.0;
        double factor = 0.0;
        if (active)
        {
            if (MODEL == HEROS_MODEL_AREA)
                factor = area_factor(c.area, duration, lp.denom);  // TArea:156
            else
            {
                double psi, mu;
                policy_at(c, now, psi, mu);
                factor = dist_factor(c.dist, psi, mu, lp.area_sub, (int) c.tab_head[tab + lane]);  // TDist:138-143
            }
        }
        const double t_first = __shfl_sync(FULL, now, 0), t_last = __shfl_sync(FULL, now, cnt - 1);
        // the part of the two time-ordered lists this item can meet: from the first slot that still holds somebody present
        // at t_first to the slot of t_last (every lane looks at one slot)
        int ill_lo = 0, ill_hi = 0, sus_lo = 0, sus_hi = 0;
        {
            const PartTable& T = c.part_table[c.item_part[it]];
            const double f = floor((t_last - c.T0) * ((double) PART_SLOTS / (c.T1 - c.T0)));
            const int b_hi = f < 0.0 ? 0 : (f > (double) (PART_SLOTS - 1) ? PART_SLOTS - 1 : (int) f);
#pragma unroll
            for (int kind = 0; kind < 2; ++kind)
            {
                const unsigned live = __ballot_sync(FULL, lane <= b_hi && T.max_tout[kind][lane] >= t_first);
                int lo = 0, hi = 0;
                if (live)
                {
                    lo = (int) T.start[kind][__ffs(live) - 1];
                    hi = (int) T.start[kind][b_hi + 1];
                }
                if (kind == 0) { ill_lo = lo; ill_hi = hi; } else { sus_lo = lo; sus_hi = hi; }
            }
        }

        // 1. exposure sums
        ExactSum acc = exact_zero();
        for_each_record(part + ill_lo, ill_hi - ill_lo, [&](const StayRec& r, int v) {
            const bool cand = v < ill_hi - ill_lo && r.tin < t_last && r.tout >= t_first;
            unsigned m = __ballot_sync(FULL, cand);
            while (m)
            {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                StayRec q;
                q.tin = __shfl_sync(FULL, r.tin, src); q.tout = __shfl_sync(FULL, r.tout, src);
                q.view = __shfl_sync(FULL, r.view, src); q.person = __shfl_sync(FULL, r.person, src);
                q.pad = __shfl_sync(FULL, r.pad, src);
                if (factor != 0.0 && q.tin < now && now <= q.tout && ill_at(c, q, now))  // TArea:157 / TDist:144
                {
                    const double te = now - (double) view_exposure(q.view);
                    if (MODEL == HEROS_MODEL_AREA)
                        exact_add(acc, area_contribution(c.area, te));  // TArea:167-172
                    else
                    {
                        const double v_t = dist_viral_load(c.dist, te);                       // TDist:155-159
                        if (v_t > 0) exact_add(acc, dist_term(c.dist, factor, duration, v_t));  // TDist:161-164
                    }
                }
            }
        });
        s_hi[wib][lane] = acc.hi; s_lo[wib][lane] = acc.lo; s_loose[wib][lane] = acc.loose;
        __syncthreads();
        ExactSum tot = exact_zero();
#pragma unroll
        for (int w = 0; w < EVAL_T / 32; ++w) exact_merge(tot, s_hi[w][lane], s_lo[w][lane], s_loose[w][lane]);
        double p = 0.0;
        if (!(active && factor != 0.0 && probability<MODEL>(factor, exact_value(tot), p) && p > 0)) p = 0.0;
        const unsigned pos = __ballot_sync(FULL, p > 0.0);  // the same in every warp
        if (wib == 0) { s_now[lane] = now; s_p[lane] = p; }
        __syncthreads();  // the partial sums are read (the arrays can be written by the next item); s_now / s_p are set

        // 2. draws
        if (pos != 0u)
        {
            const StayRec* sus = part + (n - n_sus) + sus_lo;
            const int n_scan = sus_hi - sus_lo;
            constexpr int U = 4;  // records in flight per thread
            for (int v0 = tid; v0 < n_scan; v0 += U * EVAL_T)
            {
                double tin[U], tout[U];
                uint32_t person[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const int v = v0 + u * EVAL_T;
                    tin[u] = pos_inf(); tout[u] = neg_inf(); person[u] = 0;
                    if (v < n_scan)
                    {
                        const uint4* q = reinterpret_cast<const uint4*>(sus + v);
                        const uint4 h0 = __ldg(q);
                        tin[u] = bits_f64(((uint64_t) h0.y << 32) | h0.x); tout[u] = bits_f64(((uint64_t) h0.w << 32) | h0.z);
                        person[u] = __ldg(reinterpret_cast<const uint32_t*>(q + 1) + 2);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    if (!(tin[u] < t_last && tout[u] >= t_first)) continue;
                    unsigned m = pos;
                    while (m)
                    {
                        const int k = __ffs(m) - 1;
                        m &= m - 1;
                        const double now_k = s_now[k];
                        if (tin[u] < now_k && now_k <= tout[u]) draw_and_push(c, st, person[u], key, now_k, s_p[k]);
                    }
                }
            }
        }
        __syncthreads();  // s_now / s_p are read
    }
    flush_stats(c, st, lane);
}
// ------------------------------------------------------------------------------------------------------------------
// total-location branch (Covid19TransmissionArea.java:198-246, Covid19TransmissionDistance.java:189-246): one block per
// location with infectInSublocation = false and >= 2 sub-locations.  The movement events of every sub-location drive
// their own chain of calculated evaluations (lastCalculation is per sub-location), but every evaluation scans the
// persons of ALL sub-locations of the location (location.getAllPersonIds()) with the un-divided area, the formulas of
// that branch (sign and shape quirks included) and, in the distance model, the head count of the sub-location.
// The stays of a location are contiguous in bucket order (its keys are consecutive).  This is a rarely taken path (no
// location of the shipped data takes it), written for exactness rather than speed; the sums are formed in fixed point
// (ExactSum), so they do not depend on the order of the stays.
// ------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ double block_sum_fixed(double v, double* s_red)
{
    __syncthreads();
    s_red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int) threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    return s_red[0];
}

// order-independent sum of the threads' fixed-point partial sums
__device__ __forceinline__ double block_sum_exact(ExactSum v, long long* s_hi, unsigned long long* s_lo, double* s_loose)
{
    __syncthreads();
    s_hi[threadIdx.x] = v.hi; s_lo[threadIdx.x] = v.lo; s_loose[threadIdx.x] = v.loose;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int) threadIdx.x < o)
        {
            ExactSum a{s_hi[threadIdx.x], s_lo[threadIdx.x], s_loose[threadIdx.x]};
            exact_merge(a, s_hi[threadIdx.x + o], s_lo[threadIdx.x + o], s_loose[threadIdx.x + o]);
            s_hi[threadIdx.x] = a.hi; s_lo[threadIdx.x] = a.lo; s_loose[threadIdx.x] = a.loose;
        }
        __syncthreads();
    }
    return exact_value(ExactSum{s_hi[0], s_lo[0], s_loose[0]});
}

__device__ __forceinline__ double block_min_time(double v, double* s_red)
{
    __syncthreads();
    s_red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int) threadIdx.x < o) { const double w = s_red[threadIdx.x + o]; if (w < s_red[threadIdx.x]) s_red[threadIdx.x] = w; }
        __syncthreads();
    }
    return s_red[0];
}

}  // namespace

template <int MODEL>
__global__ void __launch_bounds__(256) k_transmit_total(const WindowCtx c)
{
    TlStamp tl_(c.ctr, TL_TOTAL);
    __shared__ double s_red[256];
    __shared__ long long s_hi[256];
    __shared__ unsigned long long s_lo[256];
    __shared__ unsigned long long s_off;
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    const int tid = threadIdx.x;
    Local st;
    for (uint32_t li = blockIdx.x; li < c.n_total_loc; li += gridDim.x)
    {
        const TotalLoc L = c.total_loc[li];
        if (c.key_total[L.first_key] & KEYFLAG_RATE) continue;  // a rate-based location: k_rate_locations
        const uint32_t a0 = c.seg_start[L.first_key], a1 = c.seg_start[L.end_key];
        const uint32_t n_all = a1 - a0;
        if (n_all == 0) continue;
        const StayRec* all = c.rec + a0;
        // the sub-location key of every stay of the location (global scratch)
        if (tid == 0)
        {
            const unsigned long long off = atomicAdd(&c.ctr->scratch_top, (unsigned long long) n_all * 4 + 16);
            s_off = off + (unsigned long long) n_all * 4 + 16 <= c.scratch_cap ? off : ~0ull;
            if (s_off == ~0ull) atomicOr(&c.ctr->overflow, 2u);
        }
        __syncthreads();
        if (s_off == ~0ull) continue;
        uint32_t* skey = reinterpret_cast<uint32_t*>(c.scratch + ((s_off + 15) & ~15ull));
        for (uint32_t key = L.first_key; key < L.end_key; ++key)
            for (uint32_t i = c.seg_start[key] - a0 + tid; i < c.seg_start[key + 1] - a0; i += 256) skey[i] = key;
        __syncthreads();
        for (uint32_t key = L.first_key; key < L.end_key; ++key)
        {
            const uint32_t b0 = c.seg_start[key] - a0, b1 = c.seg_start[key + 1] - a0;  // the sub-location's own stays
            if (b1 == b0) continue;
            double Lc = c.last_calc[key];
            const double L_in = Lc;
            double done = neg_inf();
            for (;;)
            {
                // next calculated evaluation of this sub-location: its earliest movement event after the last one
                // whose time since the last calculation is not below the threshold
                double cand = pos_inf();
                for (uint32_t i = b0 + tid; i < b1; i += 256)
                {
                    const double tout = all[i].tout, tin = all[i].tin;
                    if (tout >= c.T0 && tout < c.T1 && tout > done && !(tout - Lc < thr) && tout < cand) cand = tout;
                    if (both && tin >= c.T0 && tin < c.T1 && tin > done && !(tin - Lc < thr) && tin < cand) cand = tin;
                }
                const double now = block_min_time(cand, s_red);
                if (!(now < pos_inf())) break;
                // The evaluation is triggered by the first movement of this sub-location at `now`: an exit if there is
                // one (exits are handled before enters of the same time).  Who of the OTHER sub-locations is in
                // location.getAllPersonIds() at that moment follows from the order of simultaneous movements (by
                // sub-location key): a person who left another sub-location at exactly `now` is still there iff that
                // sub-location comes later; a person who entered one at exactly `now` is already there iff the
                // trigger is an enter and that sub-location comes earlier.
                double any_exit = 0.0;
                for (uint32_t i = b0 + tid; i < b1; i += 256) any_exit += all[i].tout == now ? 1.0 : 0.0;
                const bool by_exit = block_sum_fixed(any_exit, s_red) > 0.0;
                auto present = [&](const StayRec& q, uint32_t kq) {
                    if (kq == key) return q.tin < now && now <= q.tout;
                    if (by_exit) return q.tin < now && (q.tout > now || (q.tout == now && kq > key));
                    return q.tout > now && (q.tin < now || (q.tin == now && kq < key));
                };
                const double duration = now - Lc;
                Lc = now;
                done = now;
                if (tid == 0) st.evaluations++;
                double factor;
                if (MODEL == HEROS_MODEL_AREA)
                    factor = area_factor_total(c.area, duration, L.denom_total);  // TArea:205
                else
                {
                    double head = 0.0;  // personsInSublocation.size(): the sub-location's own head count (TDist:196)
                    for (uint32_t i = b0 + tid; i < b1; i += 256) head += (all[i].tin < now && now <= all[i].tout) ? 1.0 : 0.0;
                    const int head_count = (int) block_sum_fixed(head, s_red);
                    double psi, mu;
                    policy_at(c, now, psi, mu);
                    factor = dist_factor(c.dist, psi, mu, L.area_total, head_count);  // TDist:196-201
                }
                if (factor == 0.0) continue;
                ExactSum part = exact_zero();  // order-independent: the stays are in ticket order
                for (uint32_t r = tid; r < n_all; r += 256)  // every person of the location (TArea:211 / TDist:207)
                {
                    const StayRec q = all[r];
                    if (present(q, skey[r]) && ill_at(c, q, now))
                    {
                        const double te = now - (double) view_exposure(q.view);
                        if (MODEL == HEROS_MODEL_AREA)
                            exact_add(part, area_contribution_total(c.area, te));  // TArea:216-221
                        else
                        {
                            const double v_t = dist_viral_load(c.dist, te);
                            if (v_t > 0) exact_add(part, dist_term(c.dist, factor, duration, v_t));
                        }
                    }
                }
                const double sum = block_sum_exact(part, s_hi, s_lo, s_red);
                double p;
                if (!probability<MODEL>(factor, sum, p) || !(p > 0)) continue;  // p = 1 - exp(+factor * sum) in the area model
                for (uint32_t i = tid; i < n_all; i += 256)
                {
                    const StayRec q = all[i];
                    if (present(q, skey[i]) && phase_is_susceptible(view_tphase(q.view)))
                        draw_and_push(c, st, q.person, key, now, p);  // TArea:234-245 / TDist:234-245
                }
            }
            if (tid == 0 && Lc != L_in) c.last_calc[key] = Lc;
        }
        __syncthreads();
    }
    flush_stats(c, st, tid & 31);
}


// ------------------------------------------------------------------------------------------------------------------
// rate-based locations (medlabs LocationProbBased, factory/ConstructHerosModel.java:382-388; data/*/epidemiology/
// infection_rates.csv): the satellite towns and countries of the commuter roles.  The class lives in the un-vendored
// medlabs jar (INFERRED); the rule implemented here -- and, statement for statement, in the oracle -- is
// the convention of DESIGN.md 3.5: no occupancy-based evaluation in such a location; a susceptible person who leaves it
// at time t after a stay of d hours is exposed iff U(seed; person, t, TAG_RATE) < 1 - exp(-lambda d / 24), with
// lambda = infection_rate (per day) where the file gives one, else infection_rate_factor x the share of the reference
// group (Worker) that was exposed during the previous simulated day.  One block per location.
// ------------------------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(256) k_rate_locations(const WindowCtx c)
{
    TlStamp tl_(c.ctr, TL_RATE);
    Local st;
    const long long day = (long long) floor(c.T0 / 24.0);
    double ref_share = 0.0;
    if (day > 0 && c.ctr->n_ref > 0)
        ref_share = (double) c.ctr->ref_exposed[(day - 1) & 3] / (double) c.ctr->n_ref;
    for (uint32_t li = blockIdx.x; li < c.n_rate_loc; li += gridDim.x)
    {
        const RateLoc L = c.rate_loc[li];
        const double lambda = L.rate >= 0.0 ? L.rate : L.factor * ref_share;
        if (!(lambda > 0.0)) continue;
        const uint32_t a0 = c.seg_start[L.first_key], a1 = c.seg_start[L.end_key];
        for (uint32_t i = a0 + threadIdx.x; i < a1; i += blockDim.x)
        {
            const StayRec r = ld_rec(c.rec + i);
            if (!(r.tout >= c.T0 && r.tout < c.T1) || !phase_is_susceptible(view_tphase(r.view))) continue;
            const double x = lambda * ((r.tout - r.tin) / 24.0);
            const double p = 1.0 - hexp(-x);
            const double u = u01(c.seed, r.person, f64_bits(r.tout), TAG_RATE);
            st.draws++;
            if (u < p)
            {
                // the sub-location key of the record: the segment it lies in
                uint32_t key = L.first_key;
                while (key + 1 < L.end_key && c.seg_start[key + 1] <= i) ++key;
                push_candidate(c, r.person, key, r.tout, p);
            }
        }
    }
    st.draws = (unsigned long long) warp_sum((double) st.draws);
    if ((threadIdx.x & 31) == 0 && st.draws)
    {
        atomicAdd(&c.ctr->draws, st.draws);
        atomicAdd(&c.ctr->rate_draws, st.draws);
    }
}

// host-side launcher ----------------------------------------------------------------------------------------------------
// The hot path (its work lists were filled by the prefix-sum kernel) starts at once on the auxiliary streams, side by
// side with the kernel that streams the small sub-locations and the sub-warp kernels it feeds (they touch disjoint
// sub-locations); the main stream joins them.
static int tune(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <int MODEL>
static void launch_model(const WindowCtx& c, int n_sm, const TransmitStreams& ts, int& launches)
{
    // thread blocks per SM of the kernels that run side by side (development knobs; the defaults are the measured best)
    static const int stream_bpsm = tune("HEROS_TUNE_STREAM_BPSM", 32), chain0_bpsm = tune("HEROS_TUNE_CHAIN0_BPSM", 8),
                     eval_bpsm = tune("HEROS_TUNE_EVAL_BPSM", 8);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem0 = hot_smem_bytes<0>(), smem1 = hot_smem_bytes<1>();
    if (dev >= 0 && dev < 64 && !configured[dev])
    {
        cudaFuncSetAttribute(k_hot_chain<MODEL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem0);
        cudaFuncSetAttribute(k_hot_chain<MODEL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
        configured[dev] = true;
    }
    cudaStream_t s = ts.main;
    cudaEventRecord(ts.fork, s);
    cudaStreamWaitEvent(ts.aux[0], ts.fork, 0);
    cudaStreamWaitEvent(ts.aux[2], ts.fork, 0);
    k_hot_chain<MODEL, 1><<<n_sm, HotCfg<1>::T, smem1, ts.aux[0]>>>(c);
    k_hot_chain<MODEL, 2><<<n_sm, HotCfg<2>::T, 0, ts.aux[0]>>>(c);
    k_hot_chain<MODEL, 0><<<n_sm * chain0_bpsm, HotCfg<0>::T, smem0, ts.aux[2]>>>(c);
    cudaEventRecord(ts.join[2], ts.aux[2]);
    cudaStreamWaitEvent(ts.aux[0], ts.join[2], 0);
    k_hot_eval<MODEL><<<n_sm * eval_bpsm, EVAL_T, 0, ts.aux[0]>>>(c);
    cudaEventRecord(ts.join[0], ts.aux[0]);

    k_transmit_stream<MODEL><<<(int) min((c.n_keys + 255u) / 256u, (uint32_t) (n_sm * stream_bpsm)), 256, 0, s>>>(c);
    cudaEventRecord(ts.mid, s);
    cudaStreamWaitEvent(ts.aux[1], ts.mid, 0);
    k_transmit_sub<MODEL, 32><<<n_sm * 8, 256, 0, ts.aux[1]>>>(c);
    cudaEventRecord(ts.join[1], ts.aux[1]);
    k_transmit_sub<MODEL, 8><<<n_sm * 8, 256, 0, s>>>(c);
    cudaStreamWaitEvent(s, ts.join[0], 0);
    cudaStreamWaitEvent(s, ts.join[1], 0);
    launches += 7;
    if (c.n_total_loc)
    {
        k_transmit_total<MODEL><<<(int) min((uint32_t) (n_sm * 4), c.n_total_loc), 256, 0, s>>>(c);
        ++launches;
    }
    if (c.n_rate_loc)
    {
        k_rate_locations<MODEL><<<(int) min((uint32_t) (n_sm * 4), c.n_rate_loc), 256, 0, s>>>(c);
        ++launches;
    }
}

void launch_transmit(const WindowCtx& c, int n_sm, const TransmitStreams& ts, int& launches)
{
    if (c.model == HEROS_MODEL_AREA) launch_model<HEROS_MODEL_AREA>(c, n_sm, ts, launches);
    else launch_model<HEROS_MODEL_DISTANCE>(c, n_sm, ts, launches);
}

}  // namespace heros
