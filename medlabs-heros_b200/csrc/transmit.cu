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
//   k_transmit_thread : one THREAD per segment of <= 32 stays.  Where nobody can be infected (no ill or no
//                       susceptible occupant -- the vast majority) only lastCalculation has to advance, which needs
//                       just the two latest event times; otherwise segments of <= 8 stays are evaluated in place and
//                       larger ones are queued for the warp kernel.  Segments of > 32 stays are queued for the block
//                       kernels.
//   k_transmit_warp   : one warp per queued segment (9..32 stays), lane i holds stay i, shuffles + REDUX
//   k_transmit_block  : one block per segment of > 32 stays: bitonic sort of its events in shared memory, the chain of
//                       calculated evaluations by pointer doubling, exposure sums per evaluation, then per stay the
//                       draws of exactly the evaluations it was present at
// Output: candidate infections (person, key, time, p); engine.cu keeps the earliest per person.
#include "engine.cuh"

namespace heros {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int THREAD_SLOW_MAX = 8;  // largest segment a single thread evaluates itself
constexpr int BLOCK_P_CAP = 2048;   // evaluations per segment the shared-memory block kernels keep probabilities for

__device__ __forceinline__ double pos_inf() { return bits_f64(0x7ff0000000000000ull); }
__device__ __forceinline__ double neg_inf() { return bits_f64(0xfff0000000000000ull); }

// min over the warp of non-negative doubles (or +inf): their bit patterns order like unsigned integers
__device__ __forceinline__ double warp_min_time(double v)
{
    const uint64_t b = f64_bits(v);
    const uint32_t hi = (uint32_t) (b >> 32), lo = (uint32_t) b;
    const uint32_t mhi = __reduce_min_sync(FULL, hi);
    const uint32_t mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
    return bits_f64(((uint64_t) mhi << 32) | mlo);
}

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
        c.cand_person[slot] = person;
        c.cand_key[slot] = key;
        c.cand_t[slot] = t;
        c.cand_p[slot] = p;
    }
    else
        atomicOr(&c.ctr->overflow, 1u);
}

// is the stay's person ill at time t, as infectPeople would see it (person.getDiseasePhase().isIll())?
__device__ __forceinline__ bool ill_at(const WindowCtx& c, uint64_t view, uint32_t person, double t)
{
    if (!phase_is_ill(view_tphase(view))) return false;
    if (view_ends(view)) return t < c.ill_end[person];  // recovers / dies inside this window
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
// one thread per sub-location
// ------------------------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(256) k_transmit_thread(const WindowCtx c)
{
    const uint32_t n_seg = (uint32_t) c.ctr->n_seg;
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    const bool exact = (c.flags & HEROS_FLAG_EXACT_STATS) != 0;
    Local st;

    for (uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x; seg < n_seg; seg += gridDim.x * blockDim.x)
    {
        const uint32_t a = c.seg_start[seg], b = c.seg_start[seg + 1];
        const uint32_t n = b - a;
        if (n > 32)
        {
            // block kernels: by the power of two that holds the segment's events and the bound on its evaluations
            const bool need_enters = both || MODEL == HEROS_MODEL_DISTANCE;
            const uint32_t slots = need_enters ? 2 * n : n;
            double r_bound = both ? 2.0 * n : (double) n;
            if (thr > 0 && (c.T1 - c.T0) / thr + 2.0 < r_bound) r_bound = (c.T1 - c.T0) / thr + 2.0;
            const int cls = r_bound > (double) BLOCK_P_CAP ? 2 : (slots <= 1024 ? 0 : (slots <= 8192 ? 1 : 2));
            const unsigned long long slot = atomicAdd(&c.ctr->n_blk[cls], 1ull);
            if (slot < c.blk_cap)
            {
                c.blk_seg[cls][slot] = seg;
                if (cls == 2)
                {
                    unsigned long long mp = 64;
                    while (mp < slots) mp <<= 1;
                    const unsigned long long bytes = mp * 40 + 256;
                    const unsigned long long off = atomicAdd(&c.ctr->scratch_top, bytes);
                    if (off + bytes <= c.scratch_cap) c.huge_off[slot] = off;
                    else { c.blk_seg[cls][slot] = KEY_NONE; atomicOr(&c.ctr->overflow, 2u); }
                }
            }
            else
                atomicOr(&c.ctr->overflow, 2u);
            continue;
        }
        const uint32_t key = c.s_key[a];
        const double L0 = c.last_calc[key];
        // pass 1: who is here, and the two latest events
        bool any_ill = false, any_sus = false;
        Top2 top;
        top.init();
        for (uint32_t j = a; j < b; ++j)
        {
            const uint32_t tph = view_tphase(c.s_view[j]);
            any_ill |= phase_is_ill(tph);
            any_sus |= phase_is_susceptible(tph);
            const double tout = c.s_tout[j];
            if (tout >= c.T0 && tout < c.T1) top.add(tout);
            if (both)
            {
                const double tin = c.s_tin[j];
                if (tin >= c.T0 && tin < c.T1) top.add(tin);
            }
        }
        if (!(top.m1 > neg_inf())) continue;  // no movement in this window
        const bool need_eval = any_ill && any_sus;
        if (!need_eval && !exact)
        {
            double L;
            if (advance_only(top, L0, thr, L))
            {
                c.last_calc[key] = L;
                continue;
            }
        }
        if (n > THREAD_SLOW_MAX)
        {
            const unsigned long long slot = atomicAdd(&c.ctr->n_warp, 1ull);
            if (slot < c.warp_cap) c.warp_seg[slot] = seg;
            else atomicOr(&c.ctr->overflow, 2u);
            continue;
        }
        // walk the chain of calculated evaluations sequentially
        const LocParam lp = c.loc_param[c.key_loc[key]];
        double L = L0, done = neg_inf();
        for (;;)
        {
            double now = pos_inf();
            for (uint32_t j = a; j < b; ++j)
            {
                const double tout = c.s_tout[j];
                if (tout >= c.T0 && tout < c.T1 && tout > done && !(tout - L < thr) && tout < now) now = tout;
                if (both)
                {
                    const double tin = c.s_tin[j];
                    if (tin >= c.T0 && tin < c.T1 && tin > done && !(tin - L < thr) && tin < now) now = tin;
                }
            }
            if (!(now < pos_inf())) break;
            const double duration = now - L;
            L = now;
            done = now;
            st.evaluations++;
            if (!need_eval) continue;
            double factor, sum = 0.0;
            if (MODEL == HEROS_MODEL_AREA)
            {
                factor = area_factor(c.area, duration, lp.denom);  // TArea:156
                if (factor == 0.0) continue;                       // TArea:157
                for (uint32_t j = a; j < b; ++j)
                {
                    const uint64_t v = c.s_view[j];
                    if (c.s_tin[j] < now && now <= c.s_tout[j] && ill_at(c, v, c.s_person[j], now))
                        sum += area_contribution(c.area, now - (double) view_exposure(v));  // TArea:167-174
                }
            }
            else
            {
                int head_count = 0;
                for (uint32_t j = a; j < b; ++j) head_count += (c.s_tin[j] < now && now <= c.s_tout[j]);
                double psi, mu;
                policy_at(c, now, psi, mu);
                factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);  // TDist:138-143
                if (factor == 0.0) continue;                                     // TDist:144
                for (uint32_t j = a; j < b; ++j)
                {
                    const uint64_t v = c.s_view[j];
                    if (c.s_tin[j] < now && now <= c.s_tout[j] && ill_at(c, v, c.s_person[j], now))
                    {
                        const double v_t = dist_viral_load(c.dist, now - (double) view_exposure(v));  // TDist:155-159
                        if (v_t > 0) sum += dist_term(c.dist, factor, duration, v_t);                 // TDist:161-164
                    }
                }
            }
            double p;
            if (!probability<MODEL>(factor, sum, p) || !(p > 0)) continue;
            for (uint32_t j = a; j < b; ++j)
                if (c.s_tin[j] < now && now <= c.s_tout[j] && phase_is_susceptible(view_tphase(c.s_view[j])))
                    draw_and_push(c, st, c.s_person[j], key, now, p);  // TArea:190 / TDist:181
        }
        if (L != L0) c.last_calc[key] = L;
    }
    flush_stats(c, st, threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------------------------------------
// one warp per queued sub-location of 9..32 stays, lane i holds stay i
// ------------------------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(256) k_transmit_warp(const WindowCtx c)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_list = (uint32_t) min((unsigned long long) c.warp_cap, c.ctr->n_warp);
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    Local st;

    for (uint32_t w = warp; w < n_list; w += n_warps)
    {
        const uint32_t seg = c.warp_seg[w];
        const uint32_t a = c.seg_start[seg], b = c.seg_start[seg + 1];
        const uint32_t n = b - a;
        const uint32_t key = c.s_key[a];
        const bool have = (uint32_t) lane < n;
        double tin = pos_inf(), tout = neg_inf();
        uint64_t view = 0;
        uint32_t person = 0;
        if (have)
        {
            tin = c.s_tin[a + lane];
            tout = c.s_tout[a + lane];
            view = c.s_view[a + lane];
            person = c.s_person[a + lane];
        }
        const LocParam lp = c.loc_param[c.key_loc[key]];
        double L = c.last_calc[key];
        const double L_in = L;
        // movement events of this window that can trigger an evaluation
        double e_in = (c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT && tin >= c.T0 && tin < c.T1) ? tin : pos_inf();
        double e_out = (tout >= c.T0 && tout < c.T1) ? tout : pos_inf();
        const bool susceptible = have && phase_is_susceptible(view_tphase(view));
        const bool maybe_ill = have && phase_is_ill(view_tphase(view));
        const bool need_eval = __any_sync(FULL, maybe_ill) && __any_sync(FULL, susceptible);

        for (;;)
        {
            // next evaluation: the earliest event whose time since the last calculation is not below the threshold
            const double c_in = !(e_in - L < thr) ? e_in : pos_inf();
            const double c_out = !(e_out - L < thr) ? e_out : pos_inf();
            const double now = warp_min_time(c_in < c_out ? c_in : c_out);
            if (!(now < pos_inf())) break;
            const double duration = now - L;
            L = now;
            st.evaluations += (lane == 0);
            if (e_in <= now) e_in = pos_inf();
            if (e_out <= now) e_out = pos_inf();
            if (!need_eval) continue;

            const bool in_set = tin < now && now <= tout;  // present during the interval that ends at `now`
            double factor, term = 0.0;
            if (MODEL == HEROS_MODEL_AREA)
            {
                factor = area_factor(c.area, duration, lp.denom);  // TArea:156
                if (factor == 0.0) continue;                       // TArea:157
                if (in_set && maybe_ill && ill_at(c, view, person, now))
                    term = area_contribution(c.area, now - (double) view_exposure(view));  // TArea:167-172
            }
            else
            {
                const int head_count = __popc(__ballot_sync(FULL, in_set));  // personsInSublocation.size()
                double psi, mu;
                policy_at(c, now, psi, mu);
                factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);  // TDist:138-143
                if (factor == 0.0) continue;                                     // TDist:144
                if (in_set && maybe_ill && ill_at(c, view, person, now))
                {
                    const double v_t = dist_viral_load(c.dist, now - (double) view_exposure(view));  // TDist:155-159
                    if (v_t > 0) term = dist_term(c.dist, factor, duration, v_t);                    // TDist:161-164
                }
            }
            double p;
            if (!probability<MODEL>(factor, warp_sum(term), p)) continue;
            if (p > 0 && in_set && susceptible) draw_and_push(c, st, person, key, now, p);  // TArea:190 / TDist:181
        }
        if (lane == 0 && L != L_in) c.last_calc[key] = L;
    }
    flush_stats(c, st, lane);
}

// ------------------------------------------------------------------------------------------------------------------
// one block per sub-location of > 32 stays
// ------------------------------------------------------------------------------------------------------------------
namespace {

// out[i] = sum_{k<i} f(k) for i < n (exclusive), returns the total.  All threads of the block must call.
template <class F, class T>
__device__ int block_exclusive_scan(int n, F f, T* out, int* s_warp /* 32 ints of shared memory */)
{
    const int tid = threadIdx.x, TH = blockDim.x;
    const int chunk = (n + TH - 1) / TH;
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
    __syncthreads();
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    int warp_off = 0, total = 0;
    for (int w = 0; w < (TH >> 5); ++w)
    {
        const int v = s_warp[w];
        if (w < (tid >> 5)) warp_off += v;
        total += v;
    }
    int run = warp_off + incl - local;
    for (int i = lo; i < hi; ++i)
    {
        const int v = f(i);
        out[i] = (T) run;
        run += v;
    }
    __syncthreads();
    return total;
}

// order-preserving map of a double onto uint64 and back
__device__ __forceinline__ uint64_t time_key(double t)
{
    const uint64_t b = f64_bits(t);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_time(uint64_t k)
{
    return bits_f64((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k);
}

}  // namespace

// CLS 0: <= 1024 events in shared memory, 256 threads; CLS 1: <= 8192 events in shared memory, 1024 threads (16-bit
// indices, 20 bytes per event + 16 KB of probabilities); CLS 2: anything larger, in global scratch (40 bytes per event)
template <int CLS> struct BlockIdx { using type = uint16_t; using stype = int16_t; };
template <> struct BlockIdx<2> { using type = uint32_t; using stype = int32_t; };

template <int MODEL, int CLS>
__global__ void __launch_bounds__(CLS == 0 ? 256 : 1024) k_transmit_block(const WindowCtx c)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_warp[32];
    __shared__ double s_red[2 * 32];
    __shared__ int s_flag[2];
    const int tid = threadIdx.x, TH = blockDim.x;
    const uint32_t n_list = (uint32_t) min((unsigned long long) c.blk_cap, c.ctr->n_blk[CLS]);
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
    const bool need_enters = both || MODEL == HEROS_MODEL_DISTANCE;
    const bool exact = (c.flags & HEROS_FLAG_EXACT_STATS) != 0;
    Local st;

    for (uint32_t w = blockIdx.x; w < n_list; w += gridDim.x)
    {
        const uint32_t seg = c.blk_seg[CLS][w];
        if (seg == KEY_NONE) continue;
        const uint32_t a = c.seg_start[seg], b = c.seg_start[seg + 1];
        const int n = (int) (b - a);
        const double* tin = c.s_tin + a;
        const double* tout = c.s_tout + a;
        const uint64_t* view = c.s_view + a;
        const uint32_t* person = c.s_person + a;
        const uint32_t key = c.s_key[a];
        const double L0 = c.last_calc[key];

        // 0. who is here and the two latest events; most segments only need lastCalculation advanced
        bool need_eval;
        {
            int ill = 0, sus = 0;
            Top2 top;
            top.init();
            for (int v = tid; v < n; v += TH)
            {
                const uint32_t tph = view_tphase(view[v]);
                ill |= phase_is_ill(tph);
                sus |= phase_is_susceptible(tph);
                const double to = tout[v];
                if (to >= c.T0 && to < c.T1) top.add(to);
                if (both)
                {
                    const double ti = tin[v];
                    if (ti >= c.T0 && ti < c.T1) top.add(ti);
                }
            }
            for (int o = 16; o > 0; o >>= 1)
            {
                const double o1 = __shfl_xor_sync(FULL, top.m1, o), o2 = __shfl_xor_sync(FULL, top.m2, o);
                top.merge(o1, o2);
            }
            ill = __any_sync(FULL, ill);
            sus = __any_sync(FULL, sus);
            __syncthreads();  // the previous segment is finished with the shared arrays
            if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; }
            __syncthreads();
            if ((tid & 31) == 0)
            {
                s_red[2 * (tid >> 5)] = top.m1;
                s_red[2 * (tid >> 5) + 1] = top.m2;
                if (ill) atomicOr(&s_flag[0], 1);
                if (sus) atomicOr(&s_flag[1], 1);
            }
            __syncthreads();
            top.init();
            for (int k = 0; k < (TH >> 5); ++k) top.merge(s_red[2 * k], s_red[2 * k + 1]);
            need_eval = s_flag[0] && s_flag[1];
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
        __syncthreads();
        const LocParam lp = c.loc_param[c.key_loc[key]];

        const int slots = need_enters ? 2 * n : n;
        int mp = 64;
        while (mp < slots) mp <<= 1;
        // scratch: evt u64[mp] | p f64[BLOCK_P_CAP or mp] | j0, j1 idx[mp + 2] (event -> stay, then the pointer-doubling
        //          tables) | pre[mp] | aidx[mp] | node[mp] | ill[mp]
        using idx_t = typename BlockIdx<CLS>::type;
        using sidx_t = typename BlockIdx<CLS>::stype;
        unsigned char* base = CLS == 2 ? c.scratch + c.huge_off[w] : smem;
        uint64_t* evt = (uint64_t*) base;
        double* pr = (double*) (evt + mp);
        idx_t* j0 = (idx_t*) (pr + (CLS == 2 ? mp : BLOCK_P_CAP));
        idx_t* j1 = j0 + (mp + 2);
        sidx_t* pre = (sidx_t*) (j1 + (mp + 2));
        idx_t* aidx = (idx_t*) (pre + mp);
        idx_t* node = aidx + mp;
        idx_t* illl = node + mp;
        idx_t* evx = j0;

        // 1. movement events of the window: (time, stay << 1 | is_enter); padding sorts last
        for (int i = tid; i < mp; i += TH)
        {
            uint64_t k = ~0ull;
            const int v = need_enters ? (i >> 1) : i;
            const int kind = need_enters ? (i & 1) : 0;
            if (i < slots)
            {
                const double t = kind ? tin[v] : tout[v];
                if (t >= c.T0 && t < c.T1) k = time_key(t);
            }
            evt[i] = k;
            evx[i] = (idx_t) ((v << 1) | kind);
        }
        __syncthreads();
        // 2. bitonic sort by time
        for (int k = 2; k <= mp; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1)
            {
                for (int idx = tid; idx < (mp >> 1); idx += TH)
                {
                    const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));
                    const int q = i | j;
                    const uint64_t ki = evt[i], kq = evt[q];
                    const bool up = (i & k) == 0;
                    if ((ki > kq) == up && ki != kq)
                    {
                        evt[i] = kq; evt[q] = ki;
                        const idx_t xi = evx[i]; evx[i] = evx[q]; evx[q] = xi;
                    }
                }
                __syncthreads();
            }
        int m_ev;
        {
            int lo = 0, hi = mp;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (evt[mid] == ~0ull) hi = mid; else lo = mid + 1; }
            m_ev = lo;
        }
        // 3. occupancy change before every event (enter +1, exit -1) and the occupants carried into the window
        int carried = 0;
        if (MODEL == HEROS_MODEL_DISTANCE)
        {
            block_exclusive_scan(m_ev, [&](int i) { return (evx[i] & 1u) ? 1 : -1; }, pre, s_warp);
            for (int v = tid; v < n; v += TH) carried += (tin[v] < c.T0);
            carried = (int) warp_sum((double) carried);
            if (tid == 0) s_flag[1] = 0;
            __syncthreads();
            if ((tid & 31) == 0) atomicAdd(&s_flag[1], carried);
            __syncthreads();
            carried = s_flag[1];
        }
        // 4. the events that may trigger an evaluation, in time order: aidx[k] = event index of candidate k
        const int mc = block_exclusive_scan(m_ev, [&](int i) { return (both || !(evx[i] & 1u)) ? 1 : 0; }, j1, s_warp);
        for (int i = tid; i < m_ev; i += TH)
            if (both || !(evx[i] & 1u)) aidx[j1[i]] = (idx_t) i;
        __syncthreads();
        auto cand = [&](int k) { return key_time(evt[aidx[k]]); };
        // 5. chain of calculated evaluations: next[i] = first candidate whose time since candidate i is not below
        //    the threshold; the evaluations are start, next[start], next[next[start]], ... (pointer doubling)
        for (int i = tid; i <= mc; i += TH)
        {
            idx_t nx = (idx_t) mc;
            if (i < mc)
            {
                const double ti = cand(i);
                int lo = i + 1, hi = mc;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (!(cand(mid) - ti < thr)) hi = mid; else lo = mid + 1; }
                nx = (idx_t) lo;
            }
            j0[i] = nx;  // the event -> stay table is dead from here on
        }
        int start;
        {
            int lo = 0, hi = mc;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (!(cand(mid) - L0 < thr)) hi = mid; else lo = mid + 1; }
            start = lo;
        }
        int r_cap = mc;
        if (thr > 0)
        {
            const double bound = (c.T1 - c.T0) / thr + 2.0;
            if (bound < (double) r_cap) r_cap = (int) bound;
        }
        for (int r = tid; r < r_cap; r += TH) node[r] = (idx_t) start;
        __syncthreads();
        {
            idx_t* ja = j0;
            idx_t* jb = j1;
            for (int k = 0; (1 << k) < r_cap; ++k)
            {
                for (int r = tid; r < r_cap; r += TH)
                    if ((r >> k) & 1) node[r] = ja[node[r]];
                for (int i = tid; i <= mc; i += TH) jb[i] = ja[ja[i]];
                __syncthreads();
                idx_t* t = ja; ja = jb; jb = t;
            }
        }
        int R;  // chain length: node[r] < mc for r < R
        {
            int lo = 0, hi = r_cap;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int) node[mid] >= mc) hi = mid; else lo = mid + 1; }
            R = lo;
        }
        __syncthreads();
        // from here on node[r] is the EVENT index of evaluation r: its time is key_time(evt[node[r]])
        for (int r = tid; r < R; r += TH) node[r] = aidx[node[r]];
        __syncthreads();
        if (tid == 0)
        {
            st.evaluations += (unsigned long long) R;
            if (R > 0) c.last_calc[key] = key_time(evt[node[R - 1]]);
        }
        if (!need_eval || R == 0) continue;
        // 6. list of the stays whose person can be ill (j1 as scan output; the doubling tables are dead)
        const int n_ill = block_exclusive_scan(n, [&](int v) { return phase_is_ill(view_tphase(view[v])) ? 1 : 0; }, j1,
                                               s_warp);
        for (int v = tid; v < n; v += TH)
            if (phase_is_ill(view_tphase(view[v]))) illl[j1[v]] = (idx_t) v;
        __syncthreads();
        // 7. one thread per evaluation: duration, exposure sum, probability
        for (int r = tid; r < R; r += TH)
        {
            const double now = key_time(evt[node[r]]);
            const double duration = now - (r ? key_time(evt[node[r - 1]]) : L0);
            double p = 0.0, factor, sum = 0.0;
            if (MODEL == HEROS_MODEL_AREA)
            {
                factor = area_factor(c.area, duration, lp.denom);
                if (factor != 0.0)
                    for (int q = 0; q < n_ill; ++q)
                    {
                        const int v = (int) illl[q];
                        if (tin[v] < now && now <= tout[v] && ill_at(c, view[v], person[v], now))
                            sum += area_contribution(c.area, now - (double) view_exposure(view[v]));
                    }
            }
            else
            {
                int lo = 0, hi = m_ev;
                const uint64_t kn = time_key(now);
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (evt[mid] >= kn) hi = mid; else lo = mid + 1; }
                const int head_count = carried + pre[lo];  // lo < m_ev: `now` is itself an event
                double psi, mu;
                policy_at(c, now, psi, mu);
                factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);
                if (factor != 0.0)
                    for (int q = 0; q < n_ill; ++q)
                    {
                        const int v = (int) illl[q];
                        if (tin[v] < now && now <= tout[v] && ill_at(c, view[v], person[v], now))
                        {
                            const double v_t = dist_viral_load(c.dist, now - (double) view_exposure(view[v]));
                            if (v_t > 0) sum += dist_term(c.dist, factor, duration, v_t);
                        }
                    }
            }
            if (!(factor != 0.0 && probability<MODEL>(factor, sum, p) && p > 0)) p = 0.0;
            pr[r] = p;
        }
        __syncthreads();
        // 8. draws: every susceptible stay against exactly the evaluations it was present at, (t_in, t_out]
        for (int v = tid; v < n; v += TH)
        {
            if (!phase_is_susceptible(view_tphase(view[v]))) continue;
            const uint64_t k_in = time_key(tin[v]), k_out = time_key(tout[v]);
            int lo = 0, hi = R;  // first evaluation with time > t_in
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (evt[node[mid]] > k_in) hi = mid; else lo = mid + 1; }
            for (int r = lo; r < R && evt[node[r]] <= k_out; ++r)
                if (pr[r] > 0) draw_and_push(c, st, person[v], key, key_time(evt[node[r]]), pr[r]);
        }
    }
    flush_stats(c, st, tid & 31);
}

// host-side launcher ----------------------------------------------------------------------------------------------------
template <int MODEL>
static void launch_model(const WindowCtx& c, int n_sm, cudaStream_t s, int& launches, cudaEvent_t between)
{
    const size_t smem0 = 1024 * 20 + BLOCK_P_CAP * 8 + 64, smem1 = 8192 * 20 + BLOCK_P_CAP * 8 + 64;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev])
    {
        cudaFuncSetAttribute(k_transmit_block<MODEL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem0);
        cudaFuncSetAttribute(k_transmit_block<MODEL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
        configured[dev] = true;
    }
    k_transmit_thread<MODEL><<<n_sm * 8, 256, 0, s>>>(c);
    k_transmit_warp<MODEL><<<n_sm * 8, 256, 0, s>>>(c);
    cudaEventRecord(between, s);
    k_transmit_block<MODEL, 1><<<n_sm, 1024, smem1, s>>>(c);
    k_transmit_block<MODEL, 2><<<n_sm, 1024, 0, s>>>(c);
    k_transmit_block<MODEL, 0><<<n_sm * 4, 256, smem0, s>>>(c);
    launches += 5;
}

void launch_transmit(const WindowCtx& c, int n_sm, cudaStream_t s, int& launches, cudaEvent_t between)
{
    if (c.model == HEROS_MODEL_AREA) launch_model<HEROS_MODEL_AREA>(c, n_sm, s, launches, between);
    else launch_model<HEROS_MODEL_DISTANCE>(c, n_sm, s, launches, between);
}

}  // namespace heros
