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
//   k_transmit_thread : one THREAD per segment of <= 32 stays streams its records once.  Where nobody can be infected
//                       (no ill or no susceptible occupant -- the vast majority) only lastCalculation has to advance,
//                       which needs just the two latest event times; the other segments are queued for the sub-warp
//                       kernels, segments of > 32 stays for the group kernels.
//   k_transmit_sub    : 8 or 32 lanes per queued segment (<= 8 / 9..32 stays): the lanes walk the chain of calculated
//                       evaluations together (REDUX), then every lane computes ONE evaluation of the chain -- the long
//                       dependent fp64 chains (sqrt, exp) of up to 32 evaluations run side by side -- then lane = stay
//                       for the draws
//   k_group_filter    : one warp per queued segment of 33..512 stays: the same early-out, so that the group kernels of
//                       the two small classes only get the segments with work for a whole thread block
//   k_transmit_group  : one block per segment of > 32 stays, five classes by its number of events (queued by the
//                       prefix-sum kernel): events bucketed by time in shared memory, the chain of calculated
//                       evaluations by successor search + pointer doubling, exposure sums in 128-bit fixed point
//                       (order-independent; pair by pair with integer atomics when many occupants are ill), then one
//                       draw item per susceptible stay = a run of the segment's table of evaluations with p > 0
//   k_transmit_total  : one block per location of the total-location branch (TArea:198-246 / TDist:189-246)
//   k_draw            : expands the draw items into Philox draws, spread evenly over the whole GPU
// Output: candidate infections (person, key, time, p); engine.cu keeps the earliest per person.
#include "engine.cuh"

namespace heros {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PAIR_MIN_ILL = 48;    // more ill stays than this in a sub-location: exposure sums pair by pair
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

__device__ __forceinline__ void push_candidate(const WindowCtx& c, uint32_t person, uint32_t key, double t, double p)
{
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

template <int MODEL>
__global__ void __launch_bounds__(256) k_transmit_thread(const WindowCtx c)
{
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
        if (n > 32) continue;  // queued for the group kernels by the prefix-sum kernel
        if (c.key_total && c.key_total[key]) continue;  // evaluated per location by k_transmit_total
        const StayRec* rec = c.rec + a;
        // who is here, and the two latest events
        bool any_ill = false, any_sus = false;
        Top2 top;
        top.init();
#pragma unroll 4
        for (uint32_t j = 0; j < n; ++j)
        {
            const StayRec r = rec[j];
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
    if ((threadIdx.x & 31) == 0 && occupied) atomicAdd(&c.ctr->n_seg, (unsigned long long) occupied);
}

// ------------------------------------------------------------------------------------------------------------------
// S = 8 or 32 lanes per queued sub-location (<= 8 / 9..32 stays); lane i of the group holds stay i (canonical order)
// ------------------------------------------------------------------------------------------------------------------
template <int MODEL, int S>
__global__ void __launch_bounds__(256, 4) k_transmit_sub(const WindowCtx c)
{
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

// The Philox draws of the group kernels' items: pair q belongs to the last item whose first pair number is <= q (the
// items were reserved in ascending pair order) and is the (q - first pair)-th table entry from the item's first one.
__global__ void __launch_bounds__(256) k_draw(const WindowCtx c)
{
    const unsigned long long packed = c.ctr->n_draw;
    const unsigned long long n_items = min(packed >> DRAW_ITEM_SHIFT, c.item_cap), n_pairs = packed & DRAW_PAIR_MASK;
    Local st;
    if (n_items)
        for (unsigned long long q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += (unsigned long long) gridDim.x * blockDim.x)
        {
            unsigned long long lo = 0, hi = n_items - 1;  // item_pair[0] == 0 <= q
            while (lo < hi)
            {
                const unsigned long long mid = (lo + hi + 1) >> 1;
                if (c.item_pair[mid] <= q) lo = mid; else hi = mid - 1;
            }
            const unsigned long long e = c.item_tab[lo] + (q - c.item_pair[lo]);
            draw_and_push(c, st, c.item_person[lo], c.item_key[lo], c.tab_t[e], c.tab_p[e]);
        }
    flush_stats(c, st, threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------------------------------------
// one thread group (a warp or a block) per sub-location of > 32 stays
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

// per class: group size, groups per block, capacity in events and in kept probabilities
template <int CLS> struct GroupCfg;
template <> struct GroupCfg<0> { static constexpr int G = 128, PER_BLOCK = 1, S_CAP = 256, P_CAP = 256, NB_CAP = 512; using idx_t = uint16_t; };
template <> struct GroupCfg<1> { static constexpr int G = 256, PER_BLOCK = 1, S_CAP = 1024, P_CAP = 1024, NB_CAP = 2048; using idx_t = uint16_t; };
template <> struct GroupCfg<2> { static constexpr int G = 512, PER_BLOCK = 1, S_CAP = BLOCK_S_CAP, P_CAP = BLOCK_P_CAP, NB_CAP = 2048; using idx_t = uint16_t; };
template <> struct GroupCfg<3> { static constexpr int G = 1024, PER_BLOCK = 1, S_CAP = BLOCK_S_CAP3, P_CAP = BLOCK_P_CAP, NB_CAP = 2048; using idx_t = uint16_t; };
template <> struct GroupCfg<4> { static constexpr int G = 1024, PER_BLOCK = 1, S_CAP = 0, P_CAP = 0, NB_CAP = 65536; using idx_t = uint32_t; };

// bytes of scratch of one group: 17 per event + 24 per kept evaluation (probability + fixed-point exposure sum) + 12
// per bucket (budgeted 16; + slack)
template <int CLS> __host__ __device__ constexpr size_t group_smem_bytes()
{
    return (size_t) GroupCfg<CLS>::S_CAP * 17 + (size_t) GroupCfg<CLS>::P_CAP * 24 + (size_t) GroupCfg<CLS>::NB_CAP * 16 + 256;
}

}  // namespace

constexpr int FILTERED_CLASSES = 2;

// One warp per queued sub-location of classes 0 and 1 (33 .. 512 stays): who is here and the two latest events.  Where nobody can
// be infected only lastCalculation has to advance (usually most of them); the others are passed on to the group kernel
// of their class, which then runs whole thread blocks only where there is work for them.
template <int MODEL>
__global__ void __launch_bounds__(256) k_group_filter(const WindowCtx c)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    const bool exact = (c.flags & HEROS_FLAG_EXACT_STATS) != 0;
    uint32_t first[N_CLS + 1];
    first[0] = 0;
    for (int k = 0; k < N_CLS; ++k) first[k + 1] = first[k] + (uint32_t) min((unsigned long long) c.blk_cap, c.ctr->n_blk[k]);
    for (uint32_t i = warp; i < first[FILTERED_CLASSES]; i += n_warps)
    {
        int cls = 0;
        while (i >= first[cls + 1]) ++cls;
        const uint32_t w = i - first[cls];
        const uint32_t key = c.blk_seg[cls][w];
        if (key == KEY_NONE) continue;
        const uint32_t a = c.seg_start[key], n = c.seg_start[key + 1] - a;
        const StayRec* rec = c.rec + a;
        int ill = 0, sus = 0;
        Top2 top;
        top.init();
        for (uint32_t v = lane; v < n; v += 32)
        {
            const StayRec r = rec[v];
            const uint32_t tph = view_tphase(r.view);
            ill |= phase_is_ill(tph);
            sus |= phase_is_susceptible(tph);
            if (r.tout >= c.T0 && r.tout < c.T1) top.add(r.tout);
            if (both && r.tin >= c.T0 && r.tin < c.T1) top.add(r.tin);
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            const double o1 = __shfl_xor_sync(FULL, top.m1, o), o2 = __shfl_xor_sync(FULL, top.m2, o);
            top.merge(o1, o2);
        }
        const bool need_eval = __any_sync(FULL, ill) && __any_sync(FULL, sus);
        if (!(top.m1 > neg_inf())) continue;  // no movement in this window
        if (!need_eval && !exact)
        {
            double L;
            if (advance_only(top, c.last_calc[key], thr, L))
            {
                if (lane == 0) c.last_calc[key] = L;
                continue;
            }
        }
        if (lane == 0)
        {
            const unsigned long long slot = atomicAdd(&c.ctr->n_run[cls], 1ull);  // <= n_blk[cls] <= blk_cap
            c.run_seg[cls][slot] = key;
        }
    }
}

// The chain of calculated evaluations needs, from an evaluation at time e, the earliest later candidate event t with
// !(t - e < threshold).  Events are put in time order by a counting sort into time buckets (no wider than the threshold
// where affordable) followed by a rank sort inside every bucket; the successor of an event is then a binary search over
// the <= 3 buckets around e + threshold; the chain start, next[start], next[next[start]] ... is materialised in
// O(log n) rounds of pointer doubling.  Classes 0-2 keep everything in shared memory, class 3 (anything larger) in global scratch.
template <int MODEL, int CLS>
__global__ void __launch_bounds__(GroupCfg<CLS>::G * GroupCfg<CLS>::PER_BLOCK, 1024 / (GroupCfg<CLS>::G * GroupCfg<CLS>::PER_BLOCK))
k_transmit_group(const WindowCtx c)  // <= 64 registers per thread: the register file, not the work, limits how many groups run side by side
{
    using Cfg = GroupCfg<CLS>;
    using idx_t = typename Cfg::idx_t;
    constexpr int G = Cfg::G;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_w[33];
    __shared__ double s_red[2 * 32];
    __shared__ unsigned long long s_tab;
    const int tid = grank<G>();
    const int group_in_block = G == 32 ? (int) (threadIdx.x >> 5) : 0;
    // classes 0 and 1 (thousands of sub-locations, most of which only need lastCalculation advanced) come through
    // k_group_filter; the few large ones of classes 2-4 start at once and sort that out themselves (step 0)
    constexpr bool FILTERED = CLS < FILTERED_CLASSES;
    const uint32_t n_list = (uint32_t) min((unsigned long long) c.blk_cap, FILTERED ? c.ctr->n_run[CLS] : c.ctr->n_blk[CLS]);
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    const bool need_enters = both || MODEL == HEROS_MODEL_DISTANCE;
    const bool exact = (c.flags & HEROS_FLAG_EXACT_STATS) != 0;
    const bool dbg = (c.flags & HEROS_FLAG_PHASE_CLOCKS) != 0;
    Local st;
    long long clk = 0;
#define PHASE_CLOCK(k) do { if (dbg && tid == 0) { const long long t_ = clock64(); atomicMax(&c.ctr->dbg_clk[CLS][k], (unsigned long long) (t_ - clk)); clk = t_; } } while (0)

    for (uint32_t w = blockIdx.x * Cfg::PER_BLOCK + group_in_block; w < n_list; w += gridDim.x * Cfg::PER_BLOCK)
    {
        const uint32_t key = FILTERED ? c.run_seg[CLS][w] : c.blk_seg[CLS][w];
        if (key == KEY_NONE) continue;
        const uint32_t a = c.seg_start[key], b = c.seg_start[key + 1];
        const int n = (int) (b - a);
        const StayRec* rec = c.rec + a;
        const double L0 = c.last_calc[key];

        if (dbg) clk = clock64();
        // 0. who is here and the two latest events (k_group_filter has already dealt with the sub-locations that only
        //    need lastCalculation advanced)
        bool need_eval;
        {
            int ill = 0, sus = 0;
            Top2 top;
            top.init();
            for (int v = tid; v < n; v += G)
            {
                const StayRec r = rec[v];
                const uint32_t tph = view_tphase(r.view);
                ill |= phase_is_ill(tph);
                sus |= phase_is_susceptible(tph);
                if (r.tout >= c.T0 && r.tout < c.T1) top.add(r.tout);
                if (both && r.tin >= c.T0 && r.tin < c.T1) top.add(r.tin);
            }
            for (int o = 16; o > 0; o >>= 1)
            {
                const double o1 = __shfl_xor_sync(FULL, top.m1, o), o2 = __shfl_xor_sync(FULL, top.m2, o);
                top.merge(o1, o2);
            }
            int flags = (__any_sync(FULL, ill) ? 1 : 0) | (__any_sync(FULL, sus) ? 2 : 0);
            if (G > 32)
            {
                __syncthreads();  // the previous segment is finished with the shared arrays
                if ((tid & 31) == 0)
                {
                    s_red[2 * (tid >> 5)] = top.m1;
                    s_red[2 * (tid >> 5) + 1] = top.m2;
                    s_w[tid >> 5] = flags;
                }
                __syncthreads();
                const bool in = (tid & 31) < (G >> 5);
                top.m1 = in ? s_red[2 * (tid & 31)] : neg_inf();
                top.m2 = in ? s_red[2 * (tid & 31) + 1] : neg_inf();
                flags = in ? s_w[tid & 31] : 0;
                for (int o = 16; o > 0; o >>= 1)
                {
                    const double o1 = __shfl_xor_sync(FULL, top.m1, o), o2 = __shfl_xor_sync(FULL, top.m2, o);
                    top.merge(o1, o2);
                    flags |= __shfl_xor_sync(FULL, flags, o);
                }
                __syncthreads();
            }
            need_eval = flags == 3;
            if (!(top.m1 > neg_inf())) continue;  // no movement in this window
            if (!need_eval && !exact)
            {
                double L;
                if (advance_only(top, L0, thr, L))
                {
                    if (tid == 0) c.last_calc[key] = L;
                    continue;
                }
            }
        }
        gsync<G>();
        PHASE_CLOCK(0);
        const LocParam lp = c.loc_param[c.key_loc[key]];

        const int slots = need_enters ? 2 * n : n;
        // buckets: about four events per bucket, and where affordable no wider than the threshold, so that the
        // successor search below looks at a handful of events even when they come in bursts
        int NB = 32;
        {
            const double want = thr > 0 ? (c.T1 - c.T0) / thr : 0.0;
            while ((NB * 4 < slots || (double) NB < want) && NB < Cfg::NB_CAP) NB <<= 1;
        }
        const double scale = (double) NB / (c.T1 - c.T0);
        auto bucket_raw = [&](double t) { return (int) floor((t - c.T0) * scale); };
        auto bucket_of = [&](double t) { return min(max(bucket_raw(t), 0), NB - 1); };
        // scratch: acc_hi u64[P] | acc_lo u64[P] | ev_t f64[slots] | pr f64[P] | bstart u32[NB+1] | bcur i32[NB] | bsign i32[NB] |
        //          ja idx[slots+1] | jb idx[slots+1] | node idx[slots] | ev_k u8[slots]                    (P = P_CAP or slots)
        unsigned char* base = CLS == CLS_HUGE ? c.scratch + c.huge_off[w] : smem + (size_t) group_in_block * group_smem_bytes<CLS>();
        const size_t P = CLS == CLS_HUGE ? (size_t) slots : (size_t) Cfg::P_CAP;
        unsigned long long* acc_hi = (unsigned long long*) base;  // fixed-point exposure sum of every evaluation (ExactSum)
        unsigned long long* acc_lo = acc_hi + P;
        double* ev_t = (double*) (acc_lo + P);
        double* pr = ev_t + slots;
        uint32_t* bstart = (uint32_t*) (pr + P);
        int32_t* bcur = (int32_t*) (bstart + NB + 1);
        int32_t* bsign = bcur + NB;
        idx_t* ja = (idx_t*) (bsign + NB);
        idx_t* jb = ja + (slots + 1);
        idx_t* node = jb + (slots + 1);
        uint8_t* ev_k = (uint8_t*) (node + slots);

        // 1. histogram of the window's movement events over NB time buckets
        for (int i = tid; i <= NB; i += G) bstart[i] = 0;
        for (int i = tid; i < NB; i += G) bsign[i] = 0;
        gsync<G>();
        for (int i = tid; i < slots; i += G)
        {
            const int v = need_enters ? (i >> 1) : i;
            const int kind = need_enters ? (i & 1) : 0;
            const double t = kind ? rec[v].tin : rec[v].tout;
            if (t >= c.T0 && t < c.T1)
            {
                const int bk = bucket_of(t);
                atomicAdd(&bstart[bk], 1u);
                if (MODEL == HEROS_MODEL_DISTANCE) atomicAdd(&bsign[bk], kind ? 1 : -1);
            }
        }
        gsync<G>();
        PHASE_CLOCK(1);
        // 2. bucket offsets
        const int m_ev = group_exclusive_scan<G>(NB, [&](int i) { return (int) bstart[i]; }, bstart, s_w);
        if (tid == 0) bstart[NB] = (uint32_t) m_ev;
        for (int i = tid; i < NB; i += G) bcur[i] = (int32_t) bstart[i];
        gsync<G>();
        PHASE_CLOCK(2);
        // 3. scatter the events into their buckets (order inside a bucket is irrelevant to every result)
        for (int i = tid; i < slots; i += G)
        {
            const int v = need_enters ? (i >> 1) : i;
            const int kind = need_enters ? (i & 1) : 0;
            const double t = kind ? rec[v].tin : rec[v].tout;
            if (t >= c.T0 && t < c.T1)
            {
                const int pos = atomicAdd(&bcur[bucket_of(t)], 1);
                ev_t[pos] = t;
                ev_k[pos] = (uint8_t) kind;
            }
        }
        gsync<G>();
        PHASE_CLOCK(3);
        const idx_t END = (idx_t) m_ev;
        auto is_cand = [&](int j) { return both || ev_k[j] == 0; };
        // 4. sort inside every bucket: each event counts the events of its bucket that precede it and moves to that
        //    rank.  Buckets hold a handful of events, so this is a few compares per event, all events in parallel.
        if (CLS == CLS_HUGE)
        {
            double* tmp_t = pr;                  // f64[slots]
            uint8_t* tmp_k = (uint8_t*) node;    // idx[slots] >= slots bytes
            for (int i = tid; i < m_ev; i += G) { tmp_t[i] = ev_t[i]; tmp_k[i] = ev_k[i]; }
            gsync<G>();
            for (int i = tid; i < m_ev; i += G)
            {
                const double t = tmp_t[i];
                const int bk = bucket_of(t);
                const int lo = (int) bstart[bk], hi = (int) bstart[bk + 1];
                int rank = 0;
                for (int q = lo; q < hi; ++q) rank += (tmp_t[q] < t || (tmp_t[q] == t && q < i)) ? 1 : 0;
                ev_t[lo + rank] = t;
                ev_k[lo + rank] = tmp_k[i];
            }
        }
        else
        {
            constexpr int MAXE = Cfg::S_CAP / G;  // events per thread
            double e_t[MAXE > 0 ? MAXE : 1];
            int e_dst[MAXE > 0 ? MAXE : 1];
            uint8_t e_k[MAXE > 0 ? MAXE : 1];
#pragma unroll
            for (int k = 0; k < MAXE; ++k)
            {
                const int i = tid + k * G;
                e_dst[k] = -1;
                if (i < m_ev)
                {
                    const double t = ev_t[i];
                    const int bk = bucket_of(t);
                    const int lo = (int) bstart[bk], hi = (int) bstart[bk + 1];
                    int rank = 0;
                    for (int q = lo; q < hi; ++q) rank += (ev_t[q] < t || (ev_t[q] == t && q < i)) ? 1 : 0;
                    e_t[k] = t;
                    e_k[k] = ev_k[i];
                    e_dst[k] = lo + rank;
                }
            }
            gsync<G>();
#pragma unroll
            for (int k = 0; k < MAXE; ++k)
                if (e_dst[k] >= 0) { ev_t[e_dst[k]] = e_t[k]; ev_k[e_dst[k]] = e_k[k]; }
        }
        gsync<G>();
        int carried = 0;
        if (MODEL == HEROS_MODEL_DISTANCE)
        {
            group_exclusive_scan<G>(NB, [&](int i) { return bsign[i]; }, bcur, s_w);  // occupancy change before bucket
            for (int v = tid; v < n; v += G) carried += (rec[v].tin < c.T0);
            carried = group_sum<G>(carried, s_w);
        }
        gsync<G>();
        PHASE_CLOCK(4);
        // next candidate at or after every event (kept in jb until the pointer doubling needs that table)
        idx_t* nextc = jb;
        group_suffix_min<G>(m_ev, [&](int i) { return is_cand(i) ? i : m_ev; }, nextc, s_w);
        // successor of an evaluation at time e: the earliest candidate t > e with !(t - e < thr)
        auto successor = [&](double e) -> int {
            int b0 = bucket_raw(e + thr) - 1;
            b0 = min(max(b0, 0), NB);
            // the events are in time order; the first one that qualifies lies in buckets b0 .. b0 + 2 or, if none of those
            // does, is the first event after them (everything later is more than a threshold away)
            int lo = (int) bstart[b0], hi = (int) bstart[min(b0 + 3, NB)];
            while (lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                const double t = ev_t[mid];
                if (t > e && !(t - e < thr)) hi = mid; else lo = mid + 1;
            }
            return lo < m_ev ? (int) nextc[lo] : m_ev;  // enter events are not candidates in EXIT_ONLY mode
        };
        // 5. next pointers, chain start, pointer doubling
        for (int i = tid; i <= m_ev; i += G) ja[i] = (i < m_ev && is_cand(i)) ? (idx_t) successor(ev_t[i]) : END;
        const int start = successor(L0);
        int r_cap = m_ev;
        if (thr > 0)
        {
            const double bound = (c.T1 - c.T0) / thr + 2.0;
            if (bound < (double) r_cap) r_cap = (int) bound;
        }
        for (int r = tid; r < r_cap; r += G) node[r] = (idx_t) start;
        gsync<G>();
        PHASE_CLOCK(5);
        {
            idx_t* pa = ja;
            idx_t* pb = jb;
            for (int k = 0; (1 << k) < r_cap; ++k)
            {
                for (int r = tid; r < r_cap; r += G)
                    if ((r >> k) & 1) node[r] = pa[node[r]];
                for (int i = tid; i <= m_ev; i += G) pb[i] = pa[pa[i]];
                gsync<G>();
                idx_t* t = pa; pa = pb; pb = t;
            }
        }
        PHASE_CLOCK(6);
        int R;  // chain length: node[r] != END for r < R
        {
            int lo = 0, hi = r_cap;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (node[mid] == END) hi = mid; else lo = mid + 1; }
            R = lo;
        }
        if (tid == 0)
        {
            st.evaluations += (unsigned long long) R;
            if (R > 0) c.last_calc[key] = ev_t[node[R - 1]];
        }
        if (!need_eval || R == 0) { gsync<G>(); continue; }
        // 6. list of the stays whose person can be ill (ja as scan output; the doubling tables are dead) ...
        gsync<G>();
        const int n_ill = group_exclusive_scan<G>(n, [&](int v) { return phase_is_ill(view_tphase(rec[v].view)) ? 1 : 0; },
                                                  ja, s_w);
        for (int v = tid; v < n; v += G)
            if (phase_is_ill(view_tphase(rec[v].view))) jb[ja[v]] = (idx_t) v;
        gsync<G>();
        PHASE_CLOCK(7);
        // 7. duration, exposure sum and probability of every evaluation.
        //    a) the factor of every evaluation (head count, sqrt, exp: a long dependent fp64 chain, one per thread);
        //    b) the exposure sums, PAIR by pair: every ill stay finds the run of evaluations it was present at, the warp
        //       deals out the (ill stay, evaluation) pairs evenly and every lane adds its term to the evaluation's
        //       fixed-point sum (ExactSum) with integer atomics -- exact, so the order in which the terms arrive does
        //       not matter, and no lane idles while another one works through a term;
        //    c) the probability of every evaluation.
        if (n_ill <= PAIR_MIN_ILL)
        {
            // few ill stays (the usual case while the prevalence is low): one evaluation per thread -- or 2^k lanes per
            // evaluation when threads outnumber evaluations -- walking the short list of ill stays
            int lpn = 1;
            while (lpn < 32 && 2 * lpn * R <= G) lpn <<= 1;
            const int sub = tid & (lpn - 1), per_pass = G / lpn;
            for (int r0 = 0; r0 < R; r0 += per_pass)
            {
                const int r = r0 + tid / lpn;
                const bool active = r < R;
                double p = 0.0, factor = 0.0, now = 0.0, duration = 0.0;
                ExactSum acc = exact_zero();
                if (active)
                {
                    now = ev_t[node[r]];
                    duration = now - (r ? ev_t[node[r - 1]] : L0);
                    if (MODEL == HEROS_MODEL_AREA)
                        factor = area_factor(c.area, duration, lp.denom);
                    else
                    {
                        const int bk = bucket_of(now);
                        int head_count = carried + bcur[bk];
                        for (int j = (int) bstart[bk]; j < (int) bstart[bk + 1] && ev_t[j] < now; ++j)
                            head_count += ev_k[j] ? 1 : -1;
                        double psi, mu;
                        policy_at(c, now, psi, mu);
                        factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);
                    }
                    if (factor != 0.0)
                        for (int k = sub; k < n_ill; k += lpn)
                        {
                            const StayRec& s = rec[jb[k]];
                            if (s.tin < now && now <= s.tout && ill_at(c, s, now))
                            {
                                const double te = now - (double) view_exposure(s.view);
                                if (MODEL == HEROS_MODEL_AREA)
                                    exact_add(acc, area_contribution(c.area, te));
                                else
                                {
                                    const double v_t = dist_viral_load(c.dist, te);
                                    if (v_t > 0) exact_add(acc, dist_term(c.dist, factor, duration, v_t));
                                }
                            }
                        }
                }
                for (int o = lpn >> 1; o > 0; o >>= 1)
                    exact_merge(acc, __shfl_xor_sync(FULL, acc.hi, o), __shfl_xor_sync(FULL, acc.lo, o), __shfl_xor_sync(FULL, acc.loose, o));
                if (active && sub == 0)
                {
                    if (!(factor != 0.0 && probability<MODEL>(factor, exact_value(acc), p) && p > 0)) p = 0.0;
                    pr[r] = p;
                }
            }
        }
        else
        {
            for (int r = tid; r < R; r += G)
            {
                const double now = ev_t[node[r]];
                double factor;
                if (MODEL == HEROS_MODEL_AREA)
                    factor = area_factor(c.area, now - (r ? ev_t[node[r - 1]] : L0), lp.denom);
                else
                {
                    // head count = occupants before any event of time `now`: bucket prefix + earlier events of the bucket
                    const int bk = bucket_of(now);
                    int head_count = carried + bcur[bk];
                    for (int j = (int) bstart[bk]; j < (int) bstart[bk + 1] && ev_t[j] < now; ++j)
                        head_count += ev_k[j] ? 1 : -1;
                    double psi, mu;
                    policy_at(c, now, psi, mu);
                    factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);
                }
                pr[r] = factor;
                acc_hi[r] = 0ull;
                acc_lo[r] = 0ull;
            }
            if (tid == 0) s_w[32] = 0;  // set when a term does not fit the fixed-point range (non-finite areas)
            gsync<G>();
            for (int k0 = tid & ~31; k0 < n_ill; k0 += G)
            {
                const int lane = tid & 31;
                const int k = k0 + lane;
                StayRec s;
                s.tin = 0.0; s.tout = 0.0; s.view = 0; s.person = 0; s.pad = 0;
                int lo = 0, cnt = 0;
                if (k < n_ill)
                {
                    s = rec[jb[k]];
                    int l = 0, h = R;  // first evaluation with time > t_in
                    while (l < h) { const int mid = (l + h) >> 1; if (ev_t[node[mid]] > s.tin) h = mid; else l = mid + 1; }
                    lo = l;
                    h = R;             // first evaluation with time > t_out
                    while (l < h) { const int mid = (l + h) >> 1; if (ev_t[node[mid]] > s.tout) h = mid; else l = mid + 1; }
                    cnt = l - lo;
                }
                int incl = cnt;  // inclusive prefix sum of the pair counts over the lanes
    #pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const int t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                for (int p0 = 0; p0 < total; p0 += 32)
                {
                    const int pi = p0 + lane;
                    int src = 0;  // first lane whose inclusive prefix exceeds pi
    #pragma unroll
                    for (int step = 16; step > 0; step >>= 1)
                    {
                        const int probe = __shfl_sync(FULL, incl, min(src + step - 1, 31));
                        if (probe <= pi) src += step;
                    }
                    src = min(src, 31);
                    const int s_incl = __shfl_sync(FULL, incl, src), s_cnt = __shfl_sync(FULL, cnt, src);
                    const int s_lo = __shfl_sync(FULL, lo, src);
                    StayRec q;
                    q.tin = 0.0;
                    q.tout = __shfl_sync(FULL, s.tout, src);
                    q.view = __shfl_sync(FULL, s.view, src);
                    q.person = __shfl_sync(FULL, s.person, src);
                    q.pad = __shfl_sync(FULL, s.pad, src);
                    if (pi < total)
                    {
                        const int r = s_lo + (pi - (s_incl - s_cnt));
                        const double now = ev_t[node[r]];
                        const double factor = pr[r];
                        if (factor != 0.0 && ill_at(c, q, now))  // present by construction: t_in < now <= t_out
                        {
                            const double te = now - (double) view_exposure(q.view);
                            double term = 0.0;
                            if (MODEL == HEROS_MODEL_AREA)
                                term = area_contribution(c.area, te);
                            else
                            {
                                const double v_t = dist_viral_load(c.dist, te);
                                if (v_t > 0) term = dist_term(c.dist, factor, now - (r ? ev_t[node[r - 1]] : L0), v_t);
                            }
                            if (term != 0.0)
                            {
                                ExactSum one = exact_zero();
                                exact_add(one, term);
                                if (one.loose != 0.0 || one.loose != one.loose) s_w[32] = 1;
                                else
                                {
                                    const unsigned long long old = atomicAdd(&acc_lo[r], one.lo);
                                    atomicAdd(&acc_hi[r], (unsigned long long) one.hi + ((old + one.lo < old) ? 1ull : 0ull));
                                }
                            }
                        }
                    }
                }
            }
            gsync<G>();
            const bool loose = s_w[32] != 0;
            for (int r = tid; r < R; r += G)
            {
                const double factor = pr[r];
                double sum;
                if (!loose)
                    sum = exact_value(ExactSum{(long long) acc_hi[r], acc_lo[r], 0.0});
                else
                {
                    // a term outside the fixed-point range (NaN / infinite areas): plain doubles; the result is +-inf or
                    // NaN whatever the order
                    const double now = ev_t[node[r]], duration = now - (r ? ev_t[node[r - 1]] : L0);
                    sum = 0.0;
                    if (factor != 0.0)
                        for (int k = 0; k < n_ill; ++k)
                        {
                            const StayRec& q = rec[jb[k]];
                            if (q.tin < now && now <= q.tout && ill_at(c, q, now))
                            {
                                const double te = now - (double) view_exposure(q.view);
                                if (MODEL == HEROS_MODEL_AREA)
                                    sum += area_contribution(c.area, te);
                                else
                                {
                                    const double v_t = dist_viral_load(c.dist, te);
                                    if (v_t > 0) sum += dist_term(c.dist, factor, duration, v_t);
                                }
                            }
                        }
                }
                double p = 0.0;
                if (!(factor != 0.0 && probability<MODEL>(factor, sum, p) && p > 0)) p = 0.0;
                pr[r] = p;
            }
        }
        gsync<G>();
        if (dbg && tid == 0)
        {
            const long long t_ = clock64();
            const unsigned long long d_ = (unsigned long long) (t_ - clk);
            if (atomicMax(&c.ctr->dbg_clk[CLS][8], d_) < d_)
            {
                c.ctr->dbg_clk[CLS][10] = (unsigned long long) n; c.ctr->dbg_clk[CLS][11] = (unsigned long long) n_ill;
                c.ctr->dbg_clk[CLS][12] = (unsigned long long) R; c.ctr->dbg_clk[CLS][13] = (unsigned long long) m_ev;
                c.ctr->dbg_clk[CLS][14] = (unsigned long long) key;
            }
            clk = t_;
        }
        // 8. draws: every susceptible stay against exactly the evaluations with p > 0 it was present at, (t_in, t_out].
        //    The (stay, evaluation) pairs can be a hundred times the stays, so they are not expanded here: the
        //    evaluations with p > 0 go to a table in global memory (ja = how many precede each evaluation; the tables of
        //    steps 5-6 are dead), every susceptible stay leaves ONE item (person, first table entry, first pair number),
        //    and k_draw expands the items into draws over the whole GPU.  Nothing is left to do when nobody ill was ever
        //    present.
        const int n_pos = group_exclusive_scan<G>(R, [&](int r) { return pr[r] > 0 ? 1 : 0; }, ja, s_w);
        if (n_pos == 0) { gsync<G>(); continue; }
        if (tid == 0)
        {
            ja[R] = (idx_t) n_pos;
            s_tab = atomicAdd(&c.ctr->tab_top, (unsigned long long) n_pos);
        }
        gsync<G>();
        const unsigned long long tab0 = s_tab;
        for (int r = tid; r < R; r += G)
            if (pr[r] > 0)
            {
                const unsigned long long e = tab0 + ja[r];
                if (e < c.tab_cap) { c.tab_t[e] = ev_t[node[r]]; c.tab_p[e] = pr[r]; }
                else atomicOr(&c.ctr->overflow, 2u);
            }
        for (int v0 = tid & ~31; v0 < n; v0 += G)
        {
            const int lane = tid & 31;
            const int v = v0 + lane;
            int lo = 0, cnt = 0;
            uint32_t person = 0;
            if (v < n)
            {
                const StayRec s = rec[v];
                if (phase_is_susceptible(view_tphase(s.view)))
                {
                    person = s.person;
                    int l = 0, h = R;  // first evaluation with time > t_in
                    while (l < h) { const int mid = (l + h) >> 1; if (ev_t[node[mid]] > s.tin) h = mid; else l = mid + 1; }
                    lo = (int) ja[l];
                    h = R;             // first evaluation with time > t_out
                    while (l < h) { const int mid = (l + h) >> 1; if (ev_t[node[mid]] > s.tout) h = mid; else l = mid + 1; }
                    cnt = (int) ja[l] - lo;
                }
            }
            int incl = cnt;  // inclusive prefix sum of the pair counts over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            const unsigned has = __ballot_sync(FULL, cnt > 0);
            if (total == 0) continue;
            // one atomic reserves items and pair numbers together, so that both ascend in the same order
            unsigned long long at = 0;
            if (lane == 0) at = atomicAdd(&c.ctr->n_draw, ((unsigned long long) __popc(has) << DRAW_ITEM_SHIFT) | (unsigned long long) total);
            at = __shfl_sync(FULL, at, 0);
            if (cnt > 0)
            {
                const unsigned long long item = (at >> DRAW_ITEM_SHIFT) + __popc(has & ((1u << lane) - 1u));
                if (item < c.item_cap && tab0 + lo + cnt <= c.tab_cap)
                {
                    c.item_person[item] = person;
                    c.item_key[item] = key;
                    c.item_tab[item] = tab0 + (unsigned long long) lo;
                    c.item_pair[item] = (at & DRAW_PAIR_MASK) + (unsigned long long) (incl - cnt);
                }
                else
                    atomicOr(&c.ctr->overflow, 2u);
            }
        }
        gsync<G>();
        PHASE_CLOCK(9);
    }
#undef PHASE_CLOCK
    flush_stats(c, st, threadIdx.x & 31);
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

// host-side launcher ----------------------------------------------------------------------------------------------------
// k_transmit_thread fills the work lists; the warp kernel and the group kernels then run side by side on the auxiliary
// streams (they touch disjoint sub-locations), and the main stream joins them.
template <int MODEL>
static void launch_model(const WindowCtx& c, int n_sm, const TransmitStreams& ts, int& launches)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem0 = group_smem_bytes<0>() * GroupCfg<0>::PER_BLOCK, smem1 = group_smem_bytes<1>(),
                 smem2 = group_smem_bytes<2>(), smem3 = group_smem_bytes<3>();
    if (dev >= 0 && dev < 64 && !configured[dev])
    {
        cudaFuncSetAttribute(k_transmit_group<MODEL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem0);
        cudaFuncSetAttribute(k_transmit_group<MODEL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
        cudaFuncSetAttribute(k_transmit_group<MODEL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
        cudaFuncSetAttribute(k_transmit_group<MODEL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem3);
        configured[dev] = true;
    }
    // the group kernels (their work lists were filled by the prefix-sum kernel) start at once on the auxiliary streams,
    // side by side with the kernel that streams the small sub-locations and the sub-warp kernels it feeds
    cudaStream_t s = ts.main;
    cudaEventRecord(ts.fork, s);
    for (int k = 0; k < N_AUX; ++k)
        if (k != 3) cudaStreamWaitEvent(ts.aux[k], ts.fork, 0);
    k_transmit_group<MODEL, 2><<<n_sm * 2, GroupCfg<2>::G, smem2, ts.aux[0]>>>(c);
    k_transmit_group<MODEL, 3><<<n_sm, GroupCfg<3>::G, smem3, ts.aux[4]>>>(c);
    k_transmit_group<MODEL, 4><<<n_sm, 1024, 0, ts.aux[4]>>>(c);
    k_group_filter<MODEL><<<n_sm * 2, 256, 0, ts.aux[1]>>>(c);
    cudaEventRecord(ts.filtered, ts.aux[1]);
    cudaStreamWaitEvent(ts.aux[2], ts.filtered, 0);
    k_transmit_group<MODEL, 1><<<n_sm * 4, GroupCfg<1>::G, smem1, ts.aux[1]>>>(c);
    k_transmit_group<MODEL, 0><<<n_sm * 4, GroupCfg<0>::G * GroupCfg<0>::PER_BLOCK, smem0, ts.aux[2]>>>(c);
    k_transmit_thread<MODEL><<<n_sm * 8, 256, 0, s>>>(c);
    cudaEventRecord(ts.mid, s);
    cudaStreamWaitEvent(ts.aux[3], ts.mid, 0);
    k_transmit_sub<MODEL, 32><<<n_sm * 8, 256, 0, ts.aux[3]>>>(c);
    k_transmit_sub<MODEL, 8><<<n_sm * 8, 256, 0, s>>>(c);
    for (int k = 0; k < N_AUX; ++k)
    {
        cudaEventRecord(ts.join[k], ts.aux[k]);
        cudaStreamWaitEvent(s, ts.join[k], 0);
    }
    if (c.n_total_loc)
    {
        k_transmit_total<MODEL><<<(int) min((uint32_t) (n_sm * 4), c.n_total_loc), 256, 0, s>>>(c);
        ++launches;
    }
    k_draw<<<n_sm * 4, 256, 0, s>>>(c);
    launches += 10;
}

void launch_transmit(const WindowCtx& c, int n_sm, const TransmitStreams& ts, int& launches)
{
    if (c.model == HEROS_MODEL_AREA) launch_model<HEROS_MODEL_AREA>(c, n_sm, ts, launches);
    else launch_model<HEROS_MODEL_DISTANCE>(c, n_sm, ts, launches);
}

}  // namespace heros
