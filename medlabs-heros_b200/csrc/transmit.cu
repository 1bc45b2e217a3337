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
//   k_transmit_small : one warp per segment of <= 32 stays, everything in registers + shuffles
//   k_transmit_big   : one block per larger segment: bitonic sort of the segment's events, the chain of calculated
//                      evaluations by pointer doubling, then parallel exposure sums and Philox draws
// Output: candidate infections (person, key, time, p); resolve.cu keeps the earliest per person.
#include "engine.cuh"

namespace heros {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double pos_inf() { return bits_f64(0x7ff0000000000000ull); }

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

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// small segments: one warp per occupied sub-location, lane i holds stay i
// ------------------------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(256) k_transmit_small(const WindowCtx c)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_seg = (uint32_t) c.ctr->n_seg;
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    Local st;

    for (uint32_t seg = warp; seg < n_seg; seg += n_warps)
    {
        const uint32_t a = c.seg_start[seg], b = c.seg_start[seg + 1];
        const uint32_t n = b - a;
        if (n > 32)
        {
            if (lane == 0)
            {
                // hand the segment to the block kernel, with scratch for its events (padded to a power of two)
                uint32_t mp = 64;
                while (mp < 2 * n) mp <<= 1;
                const unsigned long long off = atomicAdd(&c.ctr->scratch_top, (unsigned long long) mp + 32);
                const unsigned long long slot = atomicAdd(&c.ctr->n_big, 1ull);
                if (off + mp + 32 <= c.scratch_cap && slot < c.big_cap)
                {
                    c.big_seg[slot] = seg;
                    c.big_off[slot] = (uint32_t) off;
                }
                else
                {
                    atomicOr(&c.ctr->overflow, 2u);
                    if (slot < c.big_cap) c.big_seg[slot] = KEY_NONE;
                }
            }
            continue;
        }
        const uint32_t key = c.s_key[a];
        const bool have = (uint32_t) lane < n;
        double tin = pos_inf(), tout = -pos_inf();
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
        const bool any_ill = __any_sync(FULL, maybe_ill);

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
            if (!any_ill) continue;

            const bool in_set = tin < now && now <= tout;  // present during the interval that ends at `now`
            double p;
            if (MODEL == HEROS_MODEL_AREA)
            {
                const double factor = area_factor(c.area, duration, lp.denom);  // TArea:156
                if (factor == 0.0) continue;                                    // TArea:157
                double contribution = 0.0;
                if (in_set && maybe_ill && ill_at(c, view, person, now))
                    contribution = area_contribution(c.area, now - (double) view_exposure(view));  // TArea:167-172
                const double sum = warp_sum(contribution);
                if (sum == 0.0) continue;      // TArea:178
                p = 1.0 - hexp(factor * sum);  // TArea:182
            }
            else
            {
                const int head_count = __popc(__ballot_sync(FULL, in_set));  // personsInSublocation.size()
                double psi, mu;
                policy_at(c, now, psi, mu);
                const double factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);  // TDist:138-143
                if (factor == 0.0) continue;                                                    // TDist:144
                double term = 0.0;
                if (in_set && maybe_ill && ill_at(c, view, person, now))
                {
                    const double v_t = dist_viral_load(c.dist, now - (double) view_exposure(view));  // TDist:155-159
                    if (v_t > 0) term = dist_term(c.dist, factor, duration, v_t);                    // TDist:161-164
                }
                const double sum = warp_sum(term);
                if (sum == 0.0) continue;  // TDist:169
                p = 1.0 - hexp(-sum);      // TDist:173
            }
            if (p > 0 && in_set && susceptible) draw_and_push(c, st, person, key, now, p);  // TArea:190 / TDist:181
        }
        if (lane == 0 && L != L_in) c.last_calc[key] = L;
    }
    // one set of atomics per warp
    st.draws = (unsigned long long) warp_sum((double) st.draws);
    st.near = (unsigned long long) warp_sum((double) st.near);
    if (lane == 0)
    {
        if (st.evaluations) atomicAdd(&c.ctr->evaluations, st.evaluations);
        if (st.draws) atomicAdd(&c.ctr->draws, st.draws);
        if (st.near) atomicAdd(&c.ctr->near_ties, st.near);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// big segments: one block per occupied sub-location
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int BIG_THREADS = 256;

// out[i] = sum_{k<i} f(k) for i < n (exclusive), returns the total.  All threads of the block must call.
template <class F>
__device__ int block_exclusive_scan(int n, F f, int32_t* out, int* s_warp /* 32 ints of shared memory */)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int chunk = (n + T - 1) / T;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int local = 0;
    for (int i = lo; i < hi; ++i) local += f(i);
    // scan of the per-thread totals
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
    for (int w = 0; w < (T >> 5); ++w)
    {
        const int v = s_warp[w];
        if (w < (tid >> 5)) warp_off += v;
        total += v;
    }
    int run = warp_off + incl - local;
    for (int i = lo; i < hi; ++i)
    {
        const int v = f(i);
        out[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

// order-preserving map of a double onto uint64
__device__ __forceinline__ uint64_t time_key(double t)
{
    const uint64_t b = f64_bits(t);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

}  // namespace

template <int MODEL>
__global__ void __launch_bounds__(BIG_THREADS) k_transmit_big(const WindowCtx c)
{
    __shared__ int s_warp[32];
    __shared__ int s_misc[4];
    const int tid = threadIdx.x, T = blockDim.x;
    const uint32_t n_big = (uint32_t) min((unsigned long long) c.big_cap, c.ctr->n_big);
    const double thr = MODEL == HEROS_MODEL_AREA ? c.area.threshold : c.dist.threshold;
    Local st;

    for (uint32_t w = blockIdx.x; w < n_big; w += gridDim.x)
    {
        const uint32_t seg = c.big_seg[w];
        if (seg == KEY_NONE) continue;
        const uint32_t a = c.seg_start[seg], b = c.seg_start[seg + 1];
        const int n = (int) (b - a);
        int mp = 64;
        while (mp < 2 * n) mp <<= 1;
        const size_t off = c.big_off[w];
        uint64_t* evt = c.sc_evt + off;
        uint32_t* evx = c.sc_evx + off;
        int32_t* pre = c.sc_pre + off;
        uint32_t* j0 = c.sc_j0 + off;
        uint32_t* j1 = c.sc_j1 + off;
        double* cand = c.sc_cand + off;
        uint32_t* node = c.sc_node + off;
        double* pr = c.sc_p + off;
        double* du = c.sc_dur + off;
        uint32_t* ill = c.sc_ill + off;
        const double* tin = c.s_tin + a;
        const double* tout = c.s_tout + a;
        const uint64_t* view = c.s_view + a;
        const uint32_t* person = c.s_person + a;
        const uint32_t key = c.s_key[a];
        const LocParam lp = c.loc_param[c.key_loc[key]];
        const double L0 = c.last_calc[key];

        // 1. the window's movement events of this sub-location: (time, stay << 1 | is_enter); padding sorts last
        for (int i = tid; i < mp; i += T)
        {
            uint64_t k = ~0ull;
            const int v = i >> 1;
            if (v < n)
            {
                const double t = (i & 1) ? tin[v] : tout[v];
                if (t >= c.T0 && t < c.T1) k = time_key(t);
            }
            evt[i] = k;
            evx[i] = (uint32_t) i;
        }
        __syncthreads();
        // 2. bitonic sort by time
        for (int k = 2; k <= mp; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1)
            {
                for (int idx = tid; idx < (mp >> 1); idx += T)
                {
                    const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));
                    const int q = i | j;
                    const uint64_t ki = evt[i], kq = evt[q];
                    const bool up = (i & k) == 0;
                    if ((ki > kq) == up && ki != kq)
                    {
                        evt[i] = kq; evt[q] = ki;
                        const uint32_t xi = evx[i]; evx[i] = evx[q]; evx[q] = xi;
                    }
                }
                __syncthreads();
            }
        // number of real events
        int m_ev;
        {
            int lo = 0, hi = mp;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (evt[mid] == ~0ull) hi = mid; else lo = mid + 1; }
            m_ev = lo;
        }
        // 3. occupancy change before every event (enter +1, exit -1) and the occupants carried into the window
        block_exclusive_scan(m_ev, [&](int i) { return (evx[i] & 1u) ? 1 : -1; }, pre, s_warp);
        int base = 0;
        for (int v = tid; v < n; v += T) base += (tin[v] < c.T0);
        base = (int) warp_sum((double) base);
        if (tid == 0) s_misc[0] = 0;
        __syncthreads();
        if ((tid & 31) == 0) atomicAdd(&s_misc[0], base);
        __syncthreads();
        base = s_misc[0];
        __syncthreads();
        // 4. the events that may trigger an evaluation, still sorted by time
        const bool both = c.trigger == HEROS_TRIGGER_ENTER_AND_EXIT;
        const int mc = block_exclusive_scan(m_ev, [&](int i) { return (both || !(evx[i] & 1u)) ? 1 : 0; },
                                            (int32_t*) j1, s_warp);
        for (int i = tid; i < m_ev; i += T)
            if (both || !(evx[i] & 1u))
            {
                const uint32_t x = evx[i];
                cand[j1[i]] = (x & 1u) ? tin[x >> 1] : tout[x >> 1];
            }
        __syncthreads();
        // 5. chain of calculated evaluations: next[i] = first candidate whose time since candidate i is not below
        //    the threshold; the evaluations are start, next[start], next[next[start]], ... (pointer doubling)
        for (int i = tid; i <= mc; i += T)
        {
            uint32_t nx = (uint32_t) mc;
            if (i < mc)
            {
                const double ti = cand[i];
                int lo = i + 1, hi = mc;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (!(cand[mid] - ti < thr)) hi = mid; else lo = mid + 1; }
                nx = (uint32_t) lo;
            }
            j0[i] = nx;
        }
        int start;
        {
            int lo = 0, hi = mc;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (!(cand[mid] - L0 < thr)) hi = mid; else lo = mid + 1; }
            start = lo;
        }
        int r_cap = mc;
        if (thr > 0)
        {
            const double bound = (c.T1 - c.T0) / thr + 2.0;
            if (bound < (double) r_cap) r_cap = (int) bound;
        }
        for (int r = tid; r < r_cap; r += T) node[r] = (uint32_t) start;
        __syncthreads();
        {
            uint32_t* ja = j0;
            uint32_t* jb = j1;
            for (int k = 0; (1 << k) < r_cap; ++k)
            {
                for (int r = tid; r < r_cap; r += T)
                    if ((r >> k) & 1) node[r] = ja[node[r]];
                for (int i = tid; i <= mc; i += T) jb[i] = ja[ja[i]];
                __syncthreads();
                uint32_t* t = ja; ja = jb; jb = t;
            }
        }
        // chain length R: node[r] < mc for r < R
        int R;
        {
            int lo = 0, hi = r_cap;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (node[mid] >= (uint32_t) mc) hi = mid; else lo = mid + 1; }
            R = lo;
        }
        // 6. list of the stays whose person can be ill
        const int n_ill = block_exclusive_scan(n, [&](int v) { return phase_is_ill(view_tphase(view[v])) ? 1 : 0; },
                                               (int32_t*) j0, s_warp);
        for (int v = tid; v < n; v += T)
            if (phase_is_ill(view_tphase(view[v]))) ill[j0[v]] = (uint32_t) v;
        __syncthreads();
        // 7. one thread per evaluation: duration, exposure sum, probability
        for (int r = tid; r < R; r += T)
        {
            const double now = cand[node[r]];
            const double duration = now - (r ? cand[node[r - 1]] : L0);
            double p = 0.0;
            if (n_ill > 0)
            {
                if (MODEL == HEROS_MODEL_AREA)
                {
                    const double factor = area_factor(c.area, duration, lp.denom);
                    if (factor != 0.0)
                    {
                        double sum = 0.0;
                        for (int q = 0; q < n_ill; ++q)
                        {
                            const int v = (int) ill[q];
                            if (tin[v] < now && now <= tout[v] && ill_at(c, view[v], person[v], now))
                                sum += area_contribution(c.area, now - (double) view_exposure(view[v]));
                        }
                        if (sum != 0.0) p = 1.0 - hexp(factor * sum);
                    }
                }
                else
                {
                    int lo = 0, hi = m_ev;
                    const uint64_t kn = time_key(now);
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (evt[mid] >= kn) hi = mid; else lo = mid + 1; }
                    const int head_count = base + (lo < m_ev ? pre[lo] : 0);
                    double psi, mu;
                    policy_at(c, now, psi, mu);
                    const double factor = dist_factor(c.dist, psi, mu, lp.area_sub, head_count);
                    if (factor != 0.0)
                    {
                        double sum = 0.0;
                        for (int q = 0; q < n_ill; ++q)
                        {
                            const int v = (int) ill[q];
                            if (tin[v] < now && now <= tout[v] && ill_at(c, view[v], person[v], now))
                            {
                                const double v_t = dist_viral_load(c.dist, now - (double) view_exposure(view[v]));
                                if (v_t > 0) sum += dist_term(c.dist, factor, duration, v_t);
                            }
                        }
                        if (sum != 0.0) p = 1.0 - hexp(-sum);
                    }
                }
            }
            pr[r] = p;
            du[r] = now;
        }
        if (tid == 0)
        {
            st.evaluations += (unsigned long long) R;
            if (R > 0) c.last_calc[key] = cand[node[R - 1]];
        }
        __syncthreads();
        // 8. draws: every susceptible occupant of every evaluation with p > 0
        const int n_pos = block_exclusive_scan(R, [&](int r) { return pr[r] > 0 ? 1 : 0; }, (int32_t*) j0, s_warp);
        for (int r = tid; r < R; r += T)
            if (pr[r] > 0) j1[j0[r]] = (uint32_t) r;
        __syncthreads();
        const unsigned long long pairs = (unsigned long long) n_pos * (unsigned long long) n;
        for (unsigned long long idx = tid; idx < pairs; idx += T)
        {
            const int r = (int) j1[idx / (unsigned) n];
            const int v = (int) (idx % (unsigned) n);
            const double now = du[r];
            if (tin[v] < now && now <= tout[v] && phase_is_susceptible(view_tphase(view[v])))
                draw_and_push(c, st, person[v], key, now, pr[r]);
        }
        __syncthreads();
    }
    if (st.evaluations) atomicAdd(&c.ctr->evaluations, st.evaluations);
    st.draws = (unsigned long long) warp_sum((double) st.draws);
    st.near = (unsigned long long) warp_sum((double) st.near);
    if ((tid & 31) == 0)
    {
        if (st.draws) atomicAdd(&c.ctr->draws, st.draws);
        if (st.near) atomicAdd(&c.ctr->near_ties, st.near);
    }
}

// host-side launchers --------------------------------------------------------------------------------------------------
void launch_transmit(const WindowCtx& c, int n_sm, cudaStream_t s, int& launches, cudaEvent_t between)
{
    const int grid = n_sm * 8;
    if (c.model == HEROS_MODEL_AREA)
    {
        k_transmit_small<HEROS_MODEL_AREA><<<grid, 256, 0, s>>>(c);
        cudaEventRecord(between, s);
        k_transmit_big<HEROS_MODEL_AREA><<<n_sm * 8, BIG_THREADS, 0, s>>>(c);
    }
    else
    {
        k_transmit_small<HEROS_MODEL_DISTANCE><<<grid, 256, 0, s>>>(c);
        cudaEventRecord(between, s);
        k_transmit_big<HEROS_MODEL_DISTANCE><<<n_sm * 8, BIG_THREADS, 0, s>>>(c);
    }
    launches += 2;
}

}  // namespace heros
