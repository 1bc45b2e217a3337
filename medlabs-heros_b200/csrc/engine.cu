// engine.cu -- libheros_gpu.so: the engine object, the window pipeline and most of the C ABI of include/heros_gpu.h
// (multi-GPU exchange: exchange.cu; device-side activity schedule: schedule.cu; transmission kernels: transmit.cu).
//
// Stays arrive through k_submit (or k_advance_day), which takes every stay's ticket at its sub-location on arrival.
// One window [T0, T1) of simulated time is then processed as
//   1. k_window_open     disease-phase state machine of every agent up to T1 (Covid19Progression.java:182-346),
//                        remembering the phase transmission must see (the phase at T0) and when illness ends; tickets
//                        of the open stay of every agent and of the stays carried over from earlier windows (replaces
//                        the per-location TIntSets the medlabs engine maintains on every addPerson / removePerson)
//   2. k_scan_lookback   bucket offsets = prefix sum over the dense sub-location keys, single pass; files the
//                        sub-locations of > 32 stays into the work lists of the group kernels
//   3. k_bucket_scatter  stays + packed agent words into bucket order; stays that outlive the window move to the other
//                        pool buffer (retire)
//   4. k_transmit_*      transmit.cu: every infectPeople evaluation of the window, Philox draws -> candidates
//   5. k_resolve         earliest successful draw per person wins; expose() the winners
#include "engine_state.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

#include <cooperative_groups.h>

namespace heros {

void launch_transmit(const WindowCtx& c, int n_sm, const TransmitStreams& ts, int& launches);

namespace {

thread_local std::string g_create_error;

// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void log_transition(const AgentArrays& A, uint32_t person, uint32_t phase, double t)
{
    const unsigned long long slot = atomicAdd(&A.ctr->n_tr_log, 1ull);
    if (slot < A.tr_cap)
    {
        A.tr_person[slot] = person;
        A.tr_phase[slot] = (uint8_t) phase;
        A.tr_time[slot] = t;
    }
    else
        atomicOr(&A.ctr->overflow, 4u);
    atomicAdd(&A.ctr->transitions, 1ull);
}

__device__ __forceinline__ void count_move(const AgentArrays& A, uint32_t from, uint32_t to)
{
    atomicAdd((unsigned long long*) &A.ctr->phase_counts[from], (unsigned long long) -1ll);  // removePerson()
    atomicAdd((unsigned long long*) &A.ctr->phase_counts[to], 1ull);                         // addPerson()
}

// Fire the pending changeDiseasePhase events of one agent with time < t_end (Covid19Progression.java:208-346).
// v: packed word (phase / next_phase updated), nt: time of the pending event.  Returns true when the agent left the
// ILL state, at *ill_end.
__device__ __forceinline__ bool run_transitions(const AgentArrays& A, uint32_t person, uint64_t& v, double& nt,
                                                double t_end, double* ill_end)
{
    bool ended = false;
    while (nt < t_end)
    {
        const uint32_t from = view_phase(v), to = view_next_phase(v);
        count_move(A, from, to);
        log_transition(A, person, to, nt);
        v = view_set(v, 0, 0xF, to);
        if (!phase_is_ill(to))  // Recovered / Dead: terminal
        {
            *ill_end = nt;
            ended = true;
            nt = bits_f64(0x7ff0000000000000ull);
            break;
        }
        uint32_t next_phase;
        double next_time;
        progression_next(*A.prog, A.seed, person, view_age(v), view_female(v), to, nt, next_phase, next_time);
        v = view_set(v, 8, 0xF, next_phase);
        nt = next_time;
    }
    return ended;
}

// DiseaseProgression.expose (Covid19Progression.java:182-201) at time t
__device__ __forceinline__ void expose_agent(const AgentArrays& A, uint32_t person, uint64_t& v, double& nt, double t)
{
    count_move(A, view_phase(v), PH_E);
    log_transition(A, person, PH_E, t);
    if (view_role(v) == ROLE_REFERENCE)  // the rate of the rate-based locations follows the reference group, day by day
        atomicAdd(&A.ctr->ref_exposed[((long long) floor(t / 24.0)) & 3], 1ull);
    v = view_set(v, 0, 0xF, PH_E);
    v = view_set_exposure(v, (float) t);
    uint32_t next_phase;
    progression_next(*A.prog, A.seed, person, view_age(v), view_female(v), PH_E, t, next_phase, nt);
    v = view_set(v, 8, 0xF, next_phase);
}

// the disease-phase state machine of one agent up to T1; leaves in the packed word the phase transmission must see
// (the one held at the start of the window) and whether / when the agent stops being ill inside the window
__device__ __forceinline__ void progress_agent(const AgentArrays& A, uint32_t a, double T1)
{
    const uint64_t v0 = A.view[a];
    const uint32_t ph = view_phase(v0);
    uint64_t v = view_set(v0, 4, 0xF, ph);  // transmission of this window sees the phase held at its start
    v = view_set(v, 13, 1, 0);
    if (phase_is_ill(ph))
    {
        double nt = A.next_time[a];
        if (nt < T1)
        {
            double ill_end;
            if (run_transitions(A, A.person_base + a, v, nt, T1, &ill_end))
            {
                A.ill_end[a] = ill_end;
                v = view_set(v, 13, 1, 1);
            }
            A.next_time[a] = nt;
        }
    }
    if (v != v0) A.view[a] = v;
}

// ---- occupancy bucketing: a counting sort over the dense sub-location keys ---------------------------------------------
// Replaces the per-location TIntSets the medlabs engine updates on every addPerson / removePerson.
// 1. tickets          every stay of the window takes a ticket at its sub-location: rank = atomicAdd(count[key], 1).
//                     A submitted stay takes it on arrival (k_submit); k_window_open does it for the open stay of every
//                     agent and for the stays carried over from earlier windows, and runs the disease-phase state machine.
// 2. k_scan_lookback  seg_start = exclusive prefix sum of the counts (n_keys + 1 entries), single pass
// 3. k_bucket_scatter the stay, with the packed agent word of its person, goes to rec[seg_start[key] + rank] (one 256-bit
//                     store); the stays that outlive the window move to the other carry pool on the way (retire).  A
//                     stay that took a ticket on arrival but only begins after the window leaves a dead record in its slot.
// The order inside a bucket is the ticket order; the transmission kernels put a bucket into canonical
// (person, t_in) order wherever an exposure sum is computed, so no result depends on it.
// The fill counts of the pools are read from the device counters: the host never waits for a window to learn them.
// These kernels are bound by the latency of dependent memory operations (load key -> atomic -> store) and by the
// number of L2 sectors they touch, not by DRAM bandwidth, so every thread handles IPT items whose loads and atomics
// are in flight together.
__global__ void __launch_bounds__(TPB) k_window_open(const AgentArrays A, uint32_t n_agents, double T1, PoolArrays P,
                                                     uint32_t pool_blocks, const uint32_t* __restrict__ open_key,
                                                     const double* __restrict__ open_tin, uint32_t* count,
                                                     uint32_t* __restrict__ agent_rank)
{
    TlStamp tl_(A.ctr, TL_OPEN);
    if (blockIdx.x == 0 && threadIdx.x == 0)  // the slot of tomorrow's exposures of the reference group starts from zero
        A.ctr->ref_exposed[((long long) floor(nextafter(T1, 0.0) / 24.0) + 1) & 3] = 0ull;
    if (blockIdx.x < pool_blocks)
    {
        // stays carried over from earlier windows (the newer ones took their ticket when they were submitted)
        const uint32_t n_carried = (uint32_t) A.ctr->n_carry;
        const uint32_t base = blockIdx.x * (TPB * IPT) + threadIdx.x;
        if (blockIdx.x * (TPB * IPT) >= n_carried) return;
        uint32_t key[IPT], r[IPT];
        double tin[IPT];
#pragma unroll
        for (int k = 0; k < IPT; ++k)
        {
            const uint32_t i = base + k * TPB;
            key[k] = KEY_NONE; tin[k] = 0.0;
            if (i < n_carried) { key[k] = P.key[i]; tin[k] = P.tin[i]; }
        }
#pragma unroll
        for (int k = 0; k < IPT; ++k)
            r[k] = (key[k] != KEY_NONE && tin[k] < T1) ? atomicAdd(&count[key[k]], 1u) : KEY_NONE;
#pragma unroll
        for (int k = 0; k < IPT; ++k)
        {
            const uint32_t i = base + k * TPB;
            if (i < n_carried) P.rank[i] = r[k];
        }
    }
    else
    {
        const uint32_t base = (blockIdx.x - pool_blocks) * (TPB * IPT) + threadIdx.x;
        uint32_t key[IPT], r[IPT];
        double tin[IPT];
#pragma unroll
        for (int k = 0; k < IPT; ++k)
        {
            const uint32_t a = base + k * TPB;
            key[k] = KEY_NONE; tin[k] = 0.0;
            if (a < n_agents) { key[k] = open_key[a]; tin[k] = open_tin[a]; }  // a closed slot holds NaN
        }
#pragma unroll
        for (int k = 0; k < IPT; ++k)
            r[k] = (key[k] != KEY_NONE && tin[k] < T1) ? atomicAdd(&count[key[k]], 1u) : KEY_NONE;
#pragma unroll
        for (int k = 0; k < IPT; ++k)
        {
            const uint32_t a = base + k * TPB;
            if (a < n_agents) agent_rank[a] = r[k];
        }
#pragma unroll
        for (int k = 0; k < IPT; ++k)
        {
            const uint32_t a = base + k * TPB;
            if (a < n_agents) progress_agent(A, a, T1);
        }
    }
}

constexpr uint64_t DEAD_VIEW = (uint64_t) PH_R << 4;  // neither ill nor susceptible: the record of a slot nobody is in

// Virtual blocks of TPB * IPT items over four sources: the carry pool, the fresh pool, the open stay of every agent, the
// guests.  The first two counts come from the device counters; the grid walks the virtual blocks with a stride.
__global__ void __launch_bounds__(TPB) k_bucket_scatter(PoolArrays PC, PoolArrays PF, PoolArrays keep, uint32_t keep_cap,
                                                        const uint32_t* __restrict__ open_key,
                                                        const double* __restrict__ open_tin, uint32_t n_agents,
                                                        const uint32_t* __restrict__ agent_rank, GuestArrays G,
                                                        uint32_t person_base, double T1,
                                                        const uint32_t* __restrict__ seg_start,
                                                        const uint64_t* __restrict__ view, StayRec* __restrict__ rec,
                                                        Counters* ctr, uint32_t rec_cap, uint32_t n_keys)
{
    TlStamp tl_(ctr, TL_SCATTER);
    constexpr uint32_t CHUNK = TPB * IPT;
    const uint32_t n_carry = (uint32_t) ctr->n_carry, n_fresh = (uint32_t) ctr->n_fresh;
    const uint32_t vb_carry = (n_carry + CHUNK - 1) / CHUNK, vb_fresh = (n_fresh + CHUNK - 1) / CHUNK,
                   vb_agents = (n_agents + CHUNK - 1) / CHUNK, vb_guests = (G.n + CHUNK - 1) / CHUNK;
    const uint32_t vb_total = vb_carry + vb_fresh + vb_agents + vb_guests;
    const double inf = bits_f64(0x7ff0000000000000ull);
    for (uint32_t vb = blockIdx.x; vb < vb_total; vb += gridDim.x)
    {
        uint32_t r[IPT], key[IPT], person[IPT], pad[IPT], dst[IPT];
        double tin[IPT], tout[IPT];
        uint64_t vw[IPT];
        if (vb < vb_carry + vb_fresh)
        {
            const bool fresh = vb >= vb_carry;
            const PoolArrays& P = fresh ? PF : PC;
            const uint32_t n_pool = fresh ? n_fresh : n_carry;
            const uint32_t base = (fresh ? vb - vb_carry : vb) * CHUNK + threadIdx.x;
#pragma unroll
            for (int k = 0; k < IPT; ++k)
            {
                const uint32_t i = base + k * TPB;
                r[k] = KEY_NONE; key[k] = KEY_NONE; person[k] = 0; tin[k] = 0.0; tout[k] = 0.0; pad[k] = 0;
                if (i < n_pool) { r[k] = P.rank[i]; key[k] = P.key[i]; person[k] = P.person[i]; tin[k] = P.tin[i]; tout[k] = P.tout[i]; }
            }
            // retire: the stays that ended before T1 leave the pool; the others move on (one atomic per warp)
#pragma unroll
            for (int k = 0; k < IPT; ++k)
            {
                const bool kept = key[k] != KEY_NONE && !(tout[k] < T1);
                const unsigned m = __ballot_sync(0xffffffffu, kept);
                if (m)
                {
                    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
                    unsigned long long at = 0;
                    if (lane == leader) at = atomicAdd(&ctr->n_keep, (unsigned long long) __popc(m));
                    at = __shfl_sync(0xffffffffu, at, leader);
                    if (kept)
                    {
                        const unsigned long long j = at + __popc(m & ((1u << lane) - 1u));
                        if (j < keep_cap)
                        {
                            keep.person[j] = person[k]; keep.key[j] = key[k]; keep.tin[j] = tin[k]; keep.tout[j] = tout[k];
                        }
                        else
                            atomicOr(&ctr->overflow, 32u);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < IPT; ++k)
            {
                vw[k] = 0; dst[k] = 0;
                if (r[k] != KEY_NONE)
                {
                    dst[k] = seg_start[key[k]] + r[k];
                    if (tin[k] < T1) { vw[k] = view[person[k]]; person[k] += person_base; }
                    else  // took its ticket on arrival but begins after this window: nobody is in the slot
                    {
                        vw[k] = DEAD_VIEW; person[k] = 0xFFFFFFFFu;
                        tin[k] = inf; tout[k] = bits_f64(0xfff0000000000000ull);
                    }
                }
            }
        }
        else if (vb < vb_carry + vb_fresh + vb_agents)
        {
            const uint32_t base = (vb - vb_carry - vb_fresh) * CHUNK + threadIdx.x;
#pragma unroll
            for (int k = 0; k < IPT; ++k)
            {
                const uint32_t a = base + k * TPB;
                r[k] = KEY_NONE; key[k] = 0; person[k] = person_base + a; tin[k] = 0.0; pad[k] = 0; vw[k] = 0;
                tout[k] = inf;
                if (a < n_agents) { r[k] = agent_rank[a]; key[k] = open_key[a]; tin[k] = open_tin[a]; vw[k] = view[a]; }
            }
#pragma unroll
            for (int k = 0; k < IPT; ++k) dst[k] = r[k] != KEY_NONE ? seg_start[key[k]] + r[k] : 0u;
        }
        else
        {
            const uint32_t base = (vb - vb_carry - vb_fresh - vb_agents) * CHUNK + threadIdx.x;
#pragma unroll
            for (int k = 0; k < IPT; ++k)
            {
                const uint32_t j = base + k * TPB;
                r[k] = KEY_NONE; key[k] = 0; person[k] = 0; tin[k] = 0.0; tout[k] = 0.0; pad[k] = REC_GUEST | j; vw[k] = 0;
                if (j < G.n) { r[k] = G.rank[j]; key[k] = G.key[j]; person[k] = G.gid[j]; tin[k] = G.tin[j]; tout[k] = G.tout[j]; vw[k] = G.view[j]; }
            }
#pragma unroll
            for (int k = 0; k < IPT; ++k) dst[k] = r[k] != KEY_NONE ? seg_start[key[k]] + r[k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < IPT; ++k)
            if (r[k] != KEY_NONE && HEROS_CHECK(ctr, key[k] < n_keys && dst[k] < rec_cap && dst[k] < seg_start[key[k] + 1]))
            {
                StayRec o;
                o.tin = tin[k]; o.tout = tout[k]; o.view = vw[k]; o.person = person[k]; o.pad = pad[k];
                st_rec(rec + dst[k], o);
            }
    }
}

// after the scatter the pools have been consumed: what outlived the window is the carry pool of the next one, the fresh
// pool is empty again (k_submit of the next window may already be waiting for this)
__global__ void k_window_roll(Counters* ctr)
{
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        ctr->n_carry = ctr->n_keep;
        ctr->n_fresh = 0ull;
    }
}

// heros_debug_timeline: the stamps of one window (ns since the start of k_window_open) are added to the running sums
// tl: stamps u64[TL_N][2] | sums i64[TL_N][2] | windows u64 | runs u64[TL_N]
__global__ void k_timeline_fold(unsigned long long* tl)
{
    const int id = threadIdx.x;
    const unsigned long long base = tl[2 * TL_OPEN];
    long long* acc = reinterpret_cast<long long*>(tl + 2 * TL_N);
    __syncwarp();
    if (id < TL_N)
    {
        const unsigned long long t0 = tl[2 * id], t1 = tl[2 * id + 1];
        if (t0 != ~0ull && base != ~0ull)
        {
            acc[2 * id] += (long long) (t0 - base);
            acc[2 * id + 1] += (long long) (t1 - base);
            tl[4 * TL_N + 1 + id] += 1ull;
        }
        tl[2 * id] = ~0ull;
        tl[2 * id + 1] = 0ull;
    }
    if (id == 0) tl[4 * TL_N] += 1ull;
}

// Exclusive prefix sum of count[0..n) into out[0..n], out[n] = total, in ONE pass: tiles are handed out in order, every
// tile publishes its sum and then its inclusive prefix in a 64-bit status word, and looks back over the statuses of the
// tiles before it (decoupled look-back).  The counts are zeroed on the way (the next window starts from zero), and the
// block that finishes last resets the status words.
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = TPB * SCAN_ITEMS;
#define SCAN_SUM (1ull << 32)
#define SCAN_PREFIX (2ull << 32)

__global__ void __launch_bounds__(TPB) k_scan_lookback(uint32_t* count, uint32_t n, uint32_t* out,
                                                       unsigned long long* status, uint32_t n_tiles, Counters* ctr,
                                                       const BigLists big)
{
    TlStamp tl_(ctr, TL_SCAN);
    __shared__ uint32_t s_w[TPB / 32];
    __shared__ uint32_t s_tile, s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->scan_tile, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], local = 0;
    if (base + SCAN_ITEMS <= n)
    {
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q)
        {
            const uint4 w = *reinterpret_cast<const uint4*>(count + base + 4 * q);
            v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
        }
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) *reinterpret_cast<uint4*>(count + base + 4 * q) = make_uint4(0, 0, 0, 0);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
        {
            v[k] = (base + k < n) ? count[base + k] : 0u;
            if (base + k < n) count[base + k] = 0u;
        }
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        local += v[k];
        if (v[k] > 32 && !(big.key_total && big.key_total[base + k]))
            queue_big(big, ctr, base + k, v[k]);  // a few thousand of half a million sub-locations
    }
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (unsigned) o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t warp_off = 0, tile_sum = 0;
#pragma unroll
    for (int w = 0; w < TPB / 32; ++w)
    {
        if (w < (int) (threadIdx.x >> 5)) warp_off += s_w[w];
        tile_sum += s_w[w];
    }
    if (threadIdx.x < 32)  // warp 0 publishes the tile sum and looks back
    {
        volatile unsigned long long* st = status;
        if (threadIdx.x == 0) st[tile] = (tile == 0 ? SCAN_PREFIX : SCAN_SUM) | tile_sum;
        uint32_t excl = 0;
        int look = (int) tile - 1;
        while (look >= 0)
        {
            const int j = look - (int) threadIdx.x;
            unsigned long long w = j >= 0 ? st[j] : SCAN_PREFIX;
            while (__any_sync(0xffffffffu, (w >> 32) == 0)) w = j >= 0 ? st[j] : SCAN_PREFIX;  // not published yet
            const unsigned has_prefix = __ballot_sync(0xffffffffu, (w >> 32) == 2);
            const int stop = has_prefix ? __ffs(has_prefix) - 1 : 32;  // nearest tile whose prefix is complete
            const uint32_t add = ((int) threadIdx.x <= stop && j >= 0) ? (uint32_t) w : 0u;
            excl += __reduce_add_sync(0xffffffffu, add);
            if (has_prefix) break;
            look -= 32;
        }
        if (threadIdx.x == 0)
        {
            if (tile > 0) st[tile] = SCAN_PREFIX | (unsigned long long) (excl + tile_sum);
            s_excl = excl;
        }
    }
    __syncthreads();
    uint32_t run = s_excl + warp_off + incl - local;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (tile == n_tiles - 1 && threadIdx.x == TPB - 1)
    {
        out[n] = run;  // the last thread of the last tile has seen everything
        ctr->n_active = run;
        ctr->visits_total += run;
    }
    // the last block to finish resets the scan state for the next window
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->scan_done, 1u);
    __syncthreads();
    if (s_tile == n_tiles - 1)
    {
        for (uint32_t t = threadIdx.x; t < n_tiles; t += TPB) status[t] = 0ull;
        if (threadIdx.x == 0) { ctr->scan_tile = 0; ctr->scan_done = 0; }
    }
}

// ---- resolve: earliest successful draw wins (ties: lowest sub-location key) -------------------------------------------
constexpr uint32_t RESOLVE_SOLO = 8192;  // up to this many candidates one block resolves everything without grid barriers

// the window is complete: its compartment counts go into the series ring, the window counter moves on
__device__ __forceinline__ void close_window(Counters* ctr, long long* series, unsigned long long n_cand)
{
    const unsigned long long w = ctr->windows;
    long long* row = series + (size_t) (w % SERIES_RING) * 8;
    for (int k = 0; k < 8; ++k) row[k] = *((volatile long long*) &ctr->phase_counts[k]);
    // behind the rows: how many infection events the log holds at the end of every window (heros_window_results)
    series[(size_t) SERIES_RING * 8 + (size_t) (w % SERIES_RING)] = (long long) *((volatile unsigned long long*) &ctr->n_inf_log);
    ctr->cand_total += n_cand;
    ctr->windows = w + 1;
}

// Candidates carry global person ids; the ones of other shards' persons (guests) are skipped here and sent home.
// Four phases separated by barriers: earliest time per person, lowest sub-location among those, expose() the winner
// (exposure time := (float) now, contract C4) and run its state machine to the end of the window, reset the scratch.
// A candidate list that did not fit its buffer is never resolved (the overflow bit is set: the host fails the window).
// Launched cooperatively (the grid barrier needs every block resident).
__global__ void __launch_bounds__(1024) k_resolve(CandArrays C, const AgentArrays A, unsigned long long* best_t,
                                                  unsigned long long* best_gkey, uint32_t* claim, InfLog log, double T1,
                                                  uint32_t n_agents, long long* series)
{
    TlStamp tl_(A.ctr, TL_RESOLVE);
    const unsigned long long n_all = A.ctr->n_cand;
    if (n_all == 0 || n_all > (unsigned long long) C.cap)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            if (n_all) atomicOr(&A.ctr->overflow, 1u);
            close_window(A.ctr, series, 0ull);
        }
        return;
    }
    const uint32_t n = (uint32_t) n_all;
    const bool solo = n <= RESOLVE_SOLO;
    if (solo && blockIdx.x != 0) return;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const uint32_t first = solo ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = solo ? blockDim.x : gridDim.x * blockDim.x;
#define RESOLVE_BARRIER() do { if (solo) __syncthreads(); else grid.sync(); } while (0)
    for (uint32_t i = first; i < n; i += stride)
    {
        const uint32_t local = C.person[i] - A.person_base;
        if (local < n_agents) atomicMin(&best_t[local], (unsigned long long) f64_bits(C.t[i]));
    }
    RESOLVE_BARRIER();
    for (uint32_t i = first; i < n; i += stride)
    {
        const uint32_t local = C.person[i] - A.person_base;
        if (local < n_agents && f64_bits(C.t[i]) == __ldcg(&best_t[local]))
            atomicMin(&best_gkey[local], (unsigned long long) C.gkey[i]);
    }
    RESOLVE_BARRIER();
    for (uint32_t i = first; i < n; i += stride)
    {
        const uint32_t gid = C.person[i];
        const uint32_t person = gid - A.person_base;
        if (person >= n_agents) continue;
        if (f64_bits(C.t[i]) != __ldcg(&best_t[person]) || C.gkey[i] != __ldcg(&best_gkey[person])) continue;
        if (atomicExch(&claim[person], 1u) != 0u) continue;  // identical duplicate
        uint64_t v = A.view[person];
        if (!phase_is_susceptible(view_phase(v))) continue;
        const double t = C.t[i];
        double nt;
        expose_agent(A, gid, v, nt, t);
        double ill_end;
        run_transitions(A, gid, v, nt, T1, &ill_end);
        A.view[person] = v;
        A.next_time[person] = nt;
        const unsigned long long slot = atomicAdd(&A.ctr->n_inf_log, 1ull);
        if (slot < log.cap)
        {
            log.person[slot] = gid;
            log.gkey[slot] = C.gkey[i];
            log.t[slot] = t;
            log.p[slot] = C.p[i];
        }
        else
            atomicOr(&A.ctr->overflow, 8u);
        atomicAdd(&A.ctr->infections, 1ull);
    }
    RESOLVE_BARRIER();
    for (uint32_t i = first; i < n; i += stride)
    {
        const uint32_t person = C.person[i] - A.person_base;
        if (person >= n_agents) continue;
        best_t[person] = BEST_NONE;
        best_gkey[person] = BEST_NONE;
        claim[person] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) close_window(A.ctr, series, n_all);
#undef RESOLVE_BARRIER
}

// after a candidate buffer overflow the per-agent scratch of k_resolve may hold stale entries
__global__ void __launch_bounds__(TPB) k_reset_best(unsigned long long* best_t, unsigned long long* best_key, uint32_t* claim,
                                                    uint32_t n)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    best_t[a] = BEST_NONE; best_key[a] = BEST_NONE; claim[a] = 0u;
}

// ---- inputs ------------------------------------------------------------------------------------------------------------
struct SubmitIn { const int32_t* person; const int32_t* loc; const int16_t* sub; const double* tin; const double* tout; };

// Closed stays go to the pool; a closed stay also closes the matching open stay of its person (same t_in), an open
// stay (t_out = +inf) becomes the open stay of its person.  One pass: closing is a compare-and-swap of the slot's t_in
// to NaN ("nobody there"), opening overwrites key and t_in, so any interleaving of the close of the old stay and the
// opening of the next one (t_in differs) ends with the new stay in the slot.
__global__ void __launch_bounds__(TPB) k_submit(SubmitIn in, uint32_t n, PoolArrays P, uint32_t pool_cap,
                                                uint32_t* open_key, double* open_tin,
                                                const uint2* __restrict__ loc_key, uint32_t n_loc, uint32_t n_agents,
                                                uint32_t person_base, int32_t loc_origin, double t_now, uint32_t* count,
                                                Counters* ctr)
{
    TlStamp tl_(ctr, TL_SUBMIT);
    const uint32_t base = blockIdx.x * (TPB * IPT) + threadIdx.x;
    const double inf = bits_f64(0x7ff0000000000000ull);
    // the block's slots in the fresh pool: its cursor lives on the device (one atomic per block)
    __shared__ unsigned long long s_slot;
    if (threadIdx.x == 0)
        s_slot = atomicAdd(&ctr->n_fresh, (unsigned long long) min((uint32_t) (TPB * IPT), n - blockIdx.x * (TPB * IPT)));
    __syncthreads();
    const unsigned long long slot0 = s_slot;
    int32_t person[IPT], loc[IPT], sub[IPT];
    double tin[IPT], tout[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k)
    {
        const uint32_t i = base + k * TPB;
        person[k] = -1; loc[k] = -1; sub[k] = 0; tin[k] = 0.0; tout[k] = 0.0;
        if (i < n)
        {
            person[k] = (int32_t) ((uint32_t) in.person[i] - person_base);
            loc[k] = in.loc[i] - loc_origin;
            sub[k] = in.sub[i]; tin[k] = in.tin[i]; tout[k] = in.tout[i];
        }
    }
    bool ok[IPT];
    int nsub[IPT];
    uint32_t key[IPT];
    double cur_tin[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k)
    {
        ok[k] = person[k] >= 0 && (uint32_t) person[k] < n_agents && loc[k] >= 0 && (uint32_t) loc[k] < n_loc &&
                sub[k] >= 0 && tin[k] >= 0.0 && tin[k] < inf && tout[k] == tout[k];
        const uint2 lk = ok[k] ? loc_key[loc[k]] : make_uint2(0u, 0u);  // first key and number of sub-locations: one load
        nsub[k] = (int) lk.y;
        key[k] = lk.x;
        cur_tin[k] = ok[k] ? open_tin[person[k]] : 0.0;  // the open stay this one may close
    }
    unsigned late = 0, invalid = 0;
    uint32_t pk[IPT], ticket[IPT];
    bool good[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k)
    {
        const uint32_t i = base + k * TPB;
        pk[k] = KEY_NONE; good[k] = false;
        if (i >= n) continue;
        good[k] = ok[k] && sub[k] < max(nsub[k], 1);
        const bool open = good[k] && !(tout[k] < inf);
        const bool positive = good[k] && tout[k] > tin[k];
        if (open)
        {
            open_key[person[k]] = key[k] + (uint32_t) sub[k];
            open_tin[person[k]] = tin[k];
            late += tin[k] < t_now;
        }
        else if (positive)
        {
            pk[k] = key[k] + (uint32_t) sub[k];
            if (f64_bits(cur_tin[k]) == f64_bits(tin[k]))
                atomicCAS((unsigned long long*) &open_tin[person[k]], (unsigned long long) f64_bits(tin[k]), OPEN_CLOSED);
            late += tin[k] < t_now;
        }
        else if (!good[k])
            ++invalid;
    }
    // the tickets of the thread's stays are taken together: IPT atomics in flight instead of one after the other
#pragma unroll
    for (int k = 0; k < IPT; ++k) ticket[k] = pk[k] != KEY_NONE ? atomicAdd(&count[pk[k]], 1u) : KEY_NONE;
#pragma unroll
    for (int k = 0; k < IPT; ++k)
    {
        const uint32_t i = base + k * TPB;
        if (i >= n) continue;
        const unsigned long long at = slot0 + (unsigned long long) (k * TPB + threadIdx.x);
        if (at >= pool_cap) { atomicOr(&ctr->overflow, 32u); continue; }
        P.person[at] = good[k] ? (uint32_t) person[k] : 0u;
        P.key[at] = pk[k];
        P.tin[at] = tin[k];
        P.tout[at] = tout[k];
        P.rank[at] = ticket[k];  // its ticket for the next window
    }
    late = __reduce_add_sync(0xffffffffu, late);
    invalid = __reduce_add_sync(0xffffffffu, invalid);
    if ((threadIdx.x & 31) == 0)
    {
        if (late) atomicAdd(&ctr->late_visits, late);
        if (invalid) atomicAdd(&ctr->invalid_visits, invalid);
    }
}

__global__ void __launch_bounds__(TPB) k_expose(const AgentArrays A, uint32_t n, const int32_t* person,
                                                const double* at_time, uint32_t n_agents)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = (int32_t) ((uint32_t) person[i] - A.person_base);
    if (p < 0 || (uint32_t) p >= n_agents) return;
    const uint64_t v0 = A.view[p];
    if (!phase_is_susceptible(view_phase(v0))) return;  // ConstructHerosModel.java:1119
    // claim the agent first so that a person listed twice is exposed once
    const uint64_t claimed = view_set(v0, 0, 0xF, PH_E);
    if (atomicCAS((unsigned long long*) &A.view[p], (unsigned long long) v0, (unsigned long long) claimed) != v0) return;
    uint64_t v = v0;
    double nt;
    expose_agent(A, A.person_base + (uint32_t) p, v, nt, at_time[i]);
    A.view[p] = v;
    A.next_time[p] = nt;
}

__global__ void __launch_bounds__(TPB) k_set_agent_state(const AgentArrays A, uint32_t n, const int32_t* person,
                                                         const uint8_t* phase, const float* exposure, double now,
                                                         uint32_t n_agents)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = (int32_t) ((uint32_t) person[i] - A.person_base);
    if (p < 0 || (uint32_t) p >= n_agents) return;
    uint64_t v = A.view[p];
    const uint32_t ph = phase[i] & 7u;
    count_move(A, view_phase(v), ph);
    v = view_set(v, 0, 0xF, ph);
    v = view_set(v, 4, 0xF, ph);
    v = view_set(v, 13, 1, 0);
    v = view_set_exposure(v, exposure[i]);
    uint32_t next_phase;
    double nt;
    progression_next(*A.prog, A.seed, A.person_base + (uint32_t) p, view_age(v), view_female(v), ph, now, next_phase, nt);
    v = view_set(v, 8, 0xF, next_phase);
    A.view[p] = v;
    A.next_time[p] = nt;
}

__global__ void __launch_bounds__(TPB) k_init_agents(uint64_t* view, double* next_time, unsigned long long* best_t,
                                                     unsigned long long* best_key, uint32_t* claim, uint32_t* open_key,
                                                     double* open_tin, const int8_t* age, const uint8_t* female,
                                                     const uint8_t* role, uint32_t n)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    uint64_t v = 0;  // Susceptible, exposure time 0.0f (ConstructHerosModel.java:744-747)
    const int ag = age ? age[a] : 0;
    v = view_set(v, 16, 0xFF, (uint32_t) (ag < 0 ? 0 : ag));
    v = view_set(v, 12, 1, female ? (female[a] ? 1u : 0u) : 0u);
    v = view_set(v, 24, 0xFF, role ? role[a] : 0u);
    view[a] = v;
    next_time[a] = bits_f64(0x7ff0000000000000ull);
    best_t[a] = BEST_NONE;
    best_key[a] = BEST_NONE;
    claim[a] = 0u;
    open_key[a] = KEY_NONE;
    open_tin[a] = 0.0;
}

// 1:1 shim of infectPeople: a single thread walks the sets in the caller's order, exactly like the Java loops
struct ShimOut { int32_t calculated, n_infectious, n_infected, pad; double p; };

__global__ void k_infect_people(int model, AreaParams area, DistParams dist, double psi, double mu, uint64_t seed,
                                int infect_sub, int n_sub_locations, double total_area, double corr,
                                const uint64_t* view, uint32_t person_base, int n_sub, const int32_t* sub_ids, int n_all,
                                const int32_t* all_ids, double now, double duration, int32_t* infectious,
                                int32_t* infected, ShimOut* out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    ShimOut o = {0, 0, 0, 0, bits_f64(0x7ff8000000000000ull)};
    const double thr = model == HEROS_MODEL_AREA ? area.threshold : dist.threshold;
    if (duration < thr) { *out = o; return; }  // TArea:138 / TDist:120
    o.calculated = 1;
    const bool sub_branch = infect_sub || n_sub_locations < 2;  // TArea:149 / TDist:131
    const int32_t* ids = sub_branch ? sub_ids : all_ids;
    const int n = sub_branch ? n_sub : n_all;
    double a = total_area;
    if (sub_branch) a /= (double) n_sub_locations;  // TArea:155 / TDist:137
    double p = 0.0;
    bool have_p = false;
    if (model == HEROS_MODEL_AREA)
    {
        const double factor = sub_branch ? area_factor(area, duration, corr * a) : area_factor_total(area, duration, corr * a);
        if (factor != 0.0)
        {
            double sum = 0.0;
            for (int k = 0; k < n; ++k)
            {
                const uint64_t v = view[(uint32_t) ids[k] - person_base];
                if (phase_is_ill(view_phase(v)))
                {
                    const double te = now - (double) view_exposure(v);
                    sum += sub_branch ? area_contribution(area, te) : area_contribution_total(area, te);
                    infectious[o.n_infectious++] = ids[k];  // TArea:175: every ill occupant
                }
            }
            if (sum != 0.0) { p = 1.0 - hexp(factor * sum); have_p = true; }
        }
    }
    else
    {
        const double factor = dist_factor(dist, psi, mu, a, n_sub);  // TDist:138 / :196 (sub-location head count)
        if (factor != 0.0)
        {
            double sum = 0.0;
            for (int k = 0; k < n; ++k)
            {
                const uint64_t v = view[(uint32_t) ids[k] - person_base];
                if (phase_is_ill(view_phase(v)))
                {
                    const double v_t = dist_viral_load(dist, now - (double) view_exposure(v));
                    if (v_t > 0)
                    {
                        sum += dist_term(dist, factor, duration, v_t);
                        infectious[o.n_infectious++] = ids[k];  // TDist:165: only contagious occupants
                    }
                }
            }
            if (sum != 0.0) { p = 1.0 - hexp(-sum); have_p = true; }
        }
    }
    if (have_p)
    {
        o.p = p;
        for (int k = 0; k < n; ++k)
        {
            const uint64_t v = view[(uint32_t) ids[k] - person_base];
            if (phase_is_susceptible(view_phase(v)) && u01(seed, (uint32_t) ids[k], f64_bits(now), TAG_TRANSMIT) < p)
                infected[o.n_infected++] = ids[k];
        }
    }
    *out = o;
}

}  // namespace
}  // namespace heros

// ---------------------------------------------------------------------------------------------------------------------
// engine object
// ---------------------------------------------------------------------------------------------------------------------
using namespace heros;

namespace heros_impl {

int32_t fail(heros_handle h, int32_t code, const std::string& msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

AgentArrays agent_arrays(heros_handle h)
{
    AgentArrays A;
    A.view = h->d_view.p; A.next_time = h->d_next_time.p; A.ill_end = h->d_ill_end.p;
    A.prog = h->d_prog.p; A.seed = h->cfg.seed; A.person_base = h->person_base; A.ctr = h->d_ctr.p;
    A.tr_person = h->tr_person.p; A.tr_phase = h->tr_phase.p; A.tr_time = h->tr_time.p;
    A.tr_cap = h->tr_person.n;
    return A;
}

PoolArrays fresh_arrays(heros_handle h)
{
    return PoolArrays{h->fresh_person.p, h->fresh_key.p, h->fresh_tin.p, h->fresh_tout.p, h->fresh_rank.p};
}

PoolArrays carry_arrays(heros_handle h, int which)
{
    return PoolArrays{h->carry_person[which].p, h->carry_key[which].p, h->carry_tin[which].p, h->carry_tout[which].p,
                      h->carry_rank[which].p};
}

// Growing a pool that is in use: everything enqueued so far (it may read or write the old buffers) is waited for first.
static int32_t quiesce(heros_handle h)
{
    CU(h, cudaStreamSynchronize(h->sub_stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int32_t ensure_fresh(heros_handle h, size_t n)
{
    if (n <= h->fresh_person.n) return HEROS_OK;
    { const int32_t rc = quiesce(h); if (rc != HEROS_OK) return rc; }
    CU(h, h->fresh_person.ensure(n, true, h->stream)); CU(h, h->fresh_key.ensure(n, true, h->stream));
    CU(h, h->fresh_tin.ensure(n, true, h->stream)); CU(h, h->fresh_tout.ensure(n, true, h->stream));
    CU(h, h->fresh_rank.ensure(n, true, h->stream));
    return HEROS_OK;
}

static int32_t ensure_carry(heros_handle h, int which, size_t n)
{
    if (n <= h->carry_person[which].n) return HEROS_OK;
    // the target pool of a window holds nothing that is still needed
    CU(h, h->carry_person[which].ensure(n)); CU(h, h->carry_key[which].ensure(n)); CU(h, h->carry_tin[which].ensure(n));
    CU(h, h->carry_tout[which].ensure(n)); CU(h, h->carry_rank[which].ensure(n));
    return HEROS_OK;
}

// the engine stream waits for what was submitted on the side stream
int32_t join_submissions(heros_handle h)
{
    if (h->sub_used) CU(h, cudaStreamWaitEvent(h->stream, h->ev_submitted, 0));
    return HEROS_OK;
}

// Pinned staging of the getters.  Page-locking memory is slow (measured at N = 2: one regrow in the middle of a run held the
// process -- and, through the exchange flags, its peer -- for 15 ms), so it is sized once for what a day can bring
// (heros_set_persons) and grows by factors of four.
int32_t ensure_pinned(heros_handle h, size_t bytes)
{
    if (bytes <= h->pin_bytes) return HEROS_OK;
    size_t want = std::max<size_t>(h->pin_bytes ? 4 * h->pin_bytes : 0, std::max<size_t>(bytes, 1u << 20));
    if (h->pin_buf) cudaFreeHost(h->pin_buf);
    h->pin_buf = nullptr; h->pin_bytes = 0;
    CU(h, cudaMallocHost(&h->pin_buf, want));
    h->pin_bytes = want;
    return HEROS_OK;
}

int32_t check_idle(heros_handle h, const char* what)
{
    if (h->win_state != 0)
        return fail(h, HEROS_ERR_INVALID, std::string(what) + ": a window opened by heros_window_begin is still in progress");
    return HEROS_OK;
}

// stage times of the windows whose events have all been reached
static int32_t harvest(heros_handle h, WindowEvents& we)
{
    if (!we.pending) return HEROS_OK;
    CU(h, cudaEventSynchronize(we.resolved));
    cudaEvent_t marks[6] = {we.begin, we.opened, we.scattered, we.mid, we.transmitted, we.resolved};
    for (int k = 0; k < 5; ++k)
    {
        float ms = 0.f;
        CU(h, cudaEventElapsedTime(&ms, marks[k], marks[k + 1]));
        h->stage_ms[k] += ms;
    }
    float ms = 0.f;
    CU(h, cudaEventElapsedTime(&ms, we.opened, we.scanned));
    h->bucket_part_ms[1] += ms;
    CU(h, cudaEventElapsedTime(&ms, we.scanned, we.scattered));
    h->bucket_part_ms[2] += ms;
    for (int k = 0; k < N_AUX; ++k)
    {
        CU(h, cudaEventElapsedTime(&ms, we.fork, we.join[k]));
        h->aux_done_ms[k] += ms;
    }
    h->windows_timed++;
    we.pending = false;
    return HEROS_OK;
}

// Everything enqueued has finished; the host copy of the counters is current.  The bounds of the pools become exact,
// the compartment counts of the finished windows are fetched, the stage times are added up.
int32_t sync_counters(heros_handle h)
{
    { const int32_t rc_ = join_submissions(h); if (rc_ != HEROS_OK) return rc_; }
    CU(h, cudaMemcpyAsync(h->h_ctr, h->d_ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->ctr_current = true;
    for (auto& we : h->ring) { const int32_t rc = harvest(h, we); if (rc != HEROS_OK) return rc; }
    if (h->win_state == 0)
    {
        h->carry_ub = (size_t) h->h_ctr->n_carry;
        h->fresh_ub = (size_t) h->h_ctr->n_fresh;
        for (bool& p : h->keep_pending) p = false;  // nothing in flight can know better
    }
    // compartment counts at the end of every window finished since the last fetch (the ring holds SERIES_RING of them)
    const size_t done = (size_t) h->h_ctr->windows;
    if (done > h->series_fetched && h->d_series.p)
    {
        size_t first = h->series_fetched;
        if (done - first > (size_t) SERIES_RING) first = done - SERIES_RING;  // older rows were overwritten (never: see window_begin)
        h->series_counts.resize(done * 8, 0);
        // only the rows of the windows finished since the last fetch (one or two pieces of the ring)
        for (size_t w = first; w < done;)
        {
            const size_t slot = w % SERIES_RING, run = std::min(done - w, (size_t) SERIES_RING - slot);
            static_assert(sizeof(int64_t) == sizeof(long long), "series rows are copied as they are");
            CU(h, cudaMemcpy(h->series_counts.data() + w * 8, h->d_series.p + slot * 8, run * 64, cudaMemcpyDeviceToHost));
            w += run;
        }
        h->series_fetched = done;
    }
    if (h->h_ctr->overflow & 1u)
    {
        // a truncated candidate list was not resolved: the earliest-draw scratch may hold entries nobody will reset
        k_reset_best<<<blocks_for(h->n_agents), TPB, 0, h->stream>>>(h->d_best_t.p, h->d_best_key.p, h->d_claim.p, h->n_agents);
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return HEROS_OK;
}

// the host copy of the counters is refreshed by every call that ends with a synchronisation; getters only pay for a
// round trip when something was enqueued since
int32_t current_counters(heros_handle h) { return h->ctr_current ? HEROS_OK : sync_counters(h); }

// what a finished window may have to report: a device buffer that overflowed (the results of the call are not to be
// trusted: HEROS_ERR_CAPACITY) or rejected stays (the windows ran without them: HEROS_ERR_INVALID, once)
static int32_t report_window_errors(heros_handle h)
{
    if (h->h_ctr->overflow)
    {
        char buf[200];
        snprintf(buf, sizeof buf, "device buffer overflow (mask %u: 1 candidates, 2 scratch / work lists, 4 transition log, 8 infection log, 16 exchange block, 32 pool, 64 a peer did not answer)",
                 h->h_ctr->overflow);
        // the candidate buffer grows for the next attempt; the other buffers are sized from exact bounds
        if (h->h_ctr->overflow & 1u)
        {
            const size_t want = 2 * (size_t) h->cand_person.n;
            h->cand_person.ensure(want); h->cand_gkey.ensure(want); h->cand_t.ensure(want); h->cand_p.ensure(want);
        }
        unsigned int zero = 0;
        cudaMemcpyAsync(&h->d_ctr.p->overflow, &zero, sizeof zero, cudaMemcpyHostToDevice, h->stream);
        cudaStreamSynchronize(h->stream);
        h->h_ctr->overflow = 0;
        return fail(h, HEROS_ERR_CAPACITY, buf);
    }
    if (h->h_ctr->invalid_visits != h->invalid_reported)
    {
        // not silently: stays with ids outside the tables or NaN / negative times were left out of their window
        char buf[200];
        snprintf(buf, sizeof buf, "%u submitted stays were rejected (person, location or sub-location id out of range, or a NaN / negative time); the window ran without them",
                 h->h_ctr->invalid_visits - h->invalid_reported);
        h->invalid_reported = h->h_ctr->invalid_visits;
        return fail(h, HEROS_ERR_INVALID, buf);
    }
    return HEROS_OK;
}

int32_t check_ready(heros_handle h)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->n_agents == 0 || h->n_loc == 0 || h->n_types == 0)
        return fail(h, HEROS_ERR_INVALID, "location types, locations and persons must be set first");
    if (!h->have_prog) return fail(h, HEROS_ERR_INVALID, "heros_set_progression has not been called");
    if (h->cfg.model == HEROS_MODEL_AREA && !h->have_area)
        return fail(h, HEROS_ERR_INVALID, "heros_set_transmission_area has not been called");
    if (h->cfg.model == HEROS_MODEL_DISTANCE && !h->have_dist)
        return fail(h, HEROS_ERR_INVALID, "heros_set_transmission_distance has not been called");
    return HEROS_OK;
}

// A window [T0, T1) runs in three phases so that a multi-GPU host can exchange stays and candidate infections between
// them: begin (disease-phase state machine + tickets of the open and carried-over stays; agent words final for the
// window) -> [guests in] -> transmit (prefix sum, scatter + retire, every evaluation; candidates out / in) -> finish
// (resolve).  Nothing in them waits for the device: the fill counts of the pools, the work lists and the candidates
// stay in HBM, the host only keeps upper bounds (exact again after every synchronising call).
int32_t window_begin(heros_handle h, double T1)
{
    // a rate-based location that follows the reference group uses the exposures of the PREVIOUS simulated day, which must
    // be complete when the window starts: such a window lies inside one day
    if (h->rate_factor_mode && std::floor(h->t_now / 24.0) != std::floor(std::nextafter(T1, 0.0) / 24.0))
        return fail(h, HEROS_ERR_UNSUPPORTED, "with rate-based locations that follow the reference group a window must not span a midnight");
    cudaStream_t s = h->stream;
    const uint32_t N = h->n_agents;
    h->win_T0 = h->t_now;
    h->win_T1 = T1;
    h->ctr_current = false;
    h->win_launches = 0;
    h->n_guest = 0;
    { const int32_t rc_ = join_submissions(h); if (rc_ != HEROS_OK) return rc_; }
    // the fill of the carry pool after the latest window whose count has reached the host, plus what was submitted since
    {
        // A host that runs ahead is held back here until the count of the last window has arrived, i.e. until that window
        // has been bucketed: its whole transmission stage is still queued on the device while this window is enqueued
        // (which takes the host less time), so the device does not run dry -- and the bound is exact, so no buffer is
        // ever sized for stays that are not there.
        for (int k = 0; k < WIN_RING; ++k)
            if (h->keep_pending[k] && h->keep_seq[k] == h->win_seq) CU(h, cudaEventSynchronize(h->ev_keep[k]));
        int best = -1;
        for (int k = 0; k < WIN_RING; ++k)
            if (h->keep_pending[k] && (best < 0 || h->keep_seq[k] > h->keep_seq[best]) && cudaEventQuery(h->ev_keep[k]) == cudaSuccess)
                best = k;
        if (best >= 0)
        {
            const size_t bound = (size_t) h->h_keep[best] + (size_t) (h->fresh_cum - h->keep_cum[best]);
            if (bound < h->carry_ub) h->carry_ub = bound;
            for (int k = 0; k < WIN_RING; ++k)
                if (h->keep_pending[k] && h->keep_seq[k] <= h->keep_seq[best]) h->keep_pending[k] = false;
        }
        cudaGetLastError();  // cudaErrorNotReady of the queries is not an error
    }
    // bounds that were never made exact keep growing: refresh them before they get silly (a cheap, rare round trip)
    if (h->carry_ub > 4 * std::max<size_t>(h->fresh_seen, 1u << 20) || h->series_t.size() - h->series_fetched >= (size_t) SERIES_RING - 1)
    {
        const int32_t rc_ = sync_counters(h);
        if (rc_ != HEROS_OK) return rc_;
        h->ctr_current = false;
    }
    CU(h, h->bk_rank.ensure(N));
    { const int32_t rc_ = ensure_carry(h, h->carry_cur ^ 1, std::max<size_t>(h->carry_ub + h->fresh_ub, 1)); if (rc_ != HEROS_OK) return rc_; }
    h->win_ring = h->ring_next;
    h->ring_next = (h->ring_next + 1) % WIN_RING;
    WindowEvents& we = h->ring[h->win_ring];
    { const int32_t rc_ = harvest(h, we); if (rc_ != HEROS_OK) return rc_; }  // the host is a whole ring ahead: wait for that window
    CU(h, cudaEventRecord(we.begin, s));
    CU(h, cudaMemsetAsync(&h->d_ctr.p->n_cand, 0, sizeof(Counters) - offsetof(Counters, n_cand), s));
    // 1. disease-phase state machine, tickets of open / carried-over stays
    const int pool_blocks = h->carry_ub ? item_blocks(h->carry_ub) : 0;
    k_window_open<<<pool_blocks + item_blocks(N), TPB, 0, s>>>(agent_arrays(h), N, T1, carry_arrays(h, h->carry_cur),
                                                               (uint32_t) pool_blocks, h->d_open_key.p, h->d_open_tin.p,
                                                               h->bk_count.p, h->bk_rank.p);
    h->win_launches += 1;
    CU(h, cudaEventRecord(we.opened, s));
    CU(h, cudaGetLastError());
    h->win_state = 1;
    return HEROS_OK;
}

int32_t window_transmit(heros_handle h)
{
    cudaStream_t s = h->stream;
    const double T0 = h->win_T0, T1 = h->win_T1;
    const uint32_t N = h->n_agents;
    WindowEvents& we = h->ring[h->win_ring];
    const size_t pool_ub = h->carry_ub + h->fresh_ub;
    if (pool_ub + N + h->n_guest >= 0xFFFFFFF0ull) return fail(h, HEROS_ERR_CAPACITY, "more than 2^32 stays in one window");
    const uint32_t n_total = (uint32_t) (pool_ub + N + h->n_guest);
    int& launches = h->win_launches;

    // 2-3. bucket offsets and the scatter into bucket order
    const uint32_t n_keys = h->n_keys;
    const uint32_t n_tiles = (n_keys + SCAN_TILE - 1) / SCAN_TILE;
    CU(h, h->seg_start.ensure((size_t) n_keys + 2));
    if (h->scan_status.n < n_tiles)
    {
        CU(h, h->scan_status.ensure(n_tiles));
        CU(h, cudaMemsetAsync(h->scan_status.p, 0, h->scan_status.n * 8, s));
    }
    CU(h, h->rec.ensure(n_total));
    CU(h, h->rec_part.ensure(n_total));
    const GuestArrays G{h->g_gid.p, h->g_view.p, h->g_key.p, h->g_tin.p, h->g_tout.p, h->g_rank.p, h->n_guest};
    const size_t list_cap = (size_t) n_total / 32 + 16;
    const size_t occupied_cap = std::min<size_t>(n_keys, n_total) + 16;
    CU(h, h->sub_seg[0].ensure(occupied_cap));
    CU(h, h->sub_seg[1].ensure((size_t) n_total / 9 + 16));  // sub-locations of 9..32 stays
    for (auto& b : h->hot_seg) CU(h, b.ensure(list_cap));
    CU(h, h->huge_off.ensure(list_cap));
    // exact bounds: a sub-location has at most one evaluation per movement event (two per stay), and one work item per
    // 32 evaluations (rounded up per hot sub-location)
    const size_t tab_cap = 2 * (size_t) n_total + 64, item_cap = (size_t) n_total / 8 + 1024;
    CU(h, h->tab_now.ensure(tab_cap)); CU(h, h->tab_dur.ensure(tab_cap)); CU(h, h->tab_head.ensure(tab_cap));
    CU(h, h->item_key.ensure(item_cap)); CU(h, h->item_tab.ensure(item_cap)); CU(h, h->item_cnt.ensure(item_cap));
    CU(h, h->item_nill.ensure(item_cap)); CU(h, h->item_nsus.ensure(item_cap)); CU(h, h->item_part.ensure(item_cap));
    CU(h, h->part_table.ensure((size_t) n_total / 33 + 64));  // one per hot sub-location (> 32 stays each)
    // global scratch of the sub-locations too large for shared memory (> 8192 movement events): 42 B per stay + 82 KB
    // per such sub-location (each holds > 4096 stays), and 4 B per stay of a total-branch location
    CU(h, h->scratch.ensure(std::min<size_t>(72 * (size_t) n_total + (4u << 20), (size_t) 8 << 30)));

    BigLists big{};
    for (int k = 0; k < N_HOT; ++k) big.hot_seg[k] = h->hot_seg[k].p;
    big.hot_cap = (uint32_t) h->hot_seg[0].n;
    for (int k = 1; k < N_HOT; ++k) big.hot_cap = (uint32_t) std::min<size_t>(big.hot_cap, h->hot_seg[k].n);
    big.hot_cap = (uint32_t) std::min<size_t>(big.hot_cap, 0x7fffffffu);
    big.huge_off = h->huge_off.p; big.scratch_cap = h->scratch.n;
    big.key_total = (h->n_total_loc || h->n_rate_loc) ? h->d_key_total.p : nullptr;
    big.need_enters = h->cfg.trigger == HEROS_TRIGGER_ENTER_AND_EXIT || h->cfg.model == HEROS_MODEL_DISTANCE;
    k_scan_lookback<<<n_tiles, TPB, 0, s>>>(h->bk_count.p, n_keys, h->seg_start.p, h->scan_status.p, n_tiles, h->d_ctr.p,
                                            big);
    CU(h, cudaEventRecord(we.scanned, s));
    const size_t vb_ub = (size_t) item_blocks(h->carry_ub) + item_blocks(h->fresh_ub) + item_blocks(N) + item_blocks(h->n_guest);
    const int scatter_grid = (int) std::max<size_t>(1, std::min<size_t>(vb_ub, (size_t) h->n_sm * 32));
    k_bucket_scatter<<<scatter_grid, TPB, 0, s>>>(
        carry_arrays(h, h->carry_cur), fresh_arrays(h), carry_arrays(h, h->carry_cur ^ 1),
        (uint32_t) std::min<size_t>(h->carry_person[h->carry_cur ^ 1].n, 0xFFFFFFFFu), h->d_open_key.p, h->d_open_tin.p, N,
        h->bk_rank.p, G, h->person_base, T1, h->seg_start.p, h->d_view.p, h->rec.p, h->d_ctr.p,
        (uint32_t) std::min<size_t>(h->rec.n, 0xFFFFFFFFu), n_keys);
    k_window_roll<<<1, 32, 0, s>>>(h->d_ctr.p);
    CU(h, cudaEventRecord(h->ev_rolled, s));  // k_submit of the next window may run from here on, beside the transmission
    {
        // n_carry (= what outlived this window) to the host, beside the transmission stage.  Should the copy be late and
        // read the count of a later window, the bound computed from it is still an upper bound (it adds the fresh stays
        // of the windows in between once more).
        const int k = h->win_ring;
        h->fresh_cum += (uint64_t) h->fresh_ub;
        CU(h, cudaStreamWaitEvent(h->meta_stream, h->ev_rolled, 0));
        CU(h, cudaMemcpyAsync(h->h_keep + k, &h->d_ctr.p->n_carry, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->meta_stream));
        CU(h, cudaEventRecord(h->ev_keep[k], h->meta_stream));
        h->keep_pending[k] = true;
        h->keep_seq[k] = ++h->win_seq;
        h->keep_cum[k] = h->fresh_cum;
    }
    CU(h, cudaEventRecord(we.scattered, s));
    launches += 3;
    // the pools as the host sees them from now on: everything that may have outlived the window is in the other carry pool
    h->fresh_seen = std::max(h->fresh_seen, h->fresh_ub);
    h->carry_ub = pool_ub;
    h->fresh_ub = 0;
    h->carry_cur ^= 1;

    // 4. transmission
    WindowCtx c{};
    c.T0 = T0; c.T1 = T1; c.model = h->cfg.model; c.trigger = h->cfg.trigger; c.seed = h->cfg.seed;
    c.area = h->area; c.dist = h->dist; c.policy = h->d_policy.p; c.n_policy = (int) h->policy.size();
    c.key_loc = h->d_key_loc.p; c.loc_param = h->d_loc_param.p; c.last_calc = h->d_last_calc.p; c.n_keys = h->n_keys;
    c.key_total = (h->n_total_loc || h->n_rate_loc) ? h->d_key_total.p : nullptr; c.total_loc = h->d_total_loc.p; c.n_total_loc = h->n_total_loc;
    c.rate_loc = h->d_rate_loc.p; c.n_rate_loc = h->n_rate_loc;
    c.view = h->d_view.p; c.ill_end = h->d_ill_end.p; c.guest_ill_end = h->g_ill_end.p;
    c.person_base = h->person_base; c.n_agents = N;
    c.loc_first_key = h->d_loc_base.p; c.location_base = h->location_base;
    c.rec = h->rec.p; c.rec_part = h->rec_part.p; c.seg_start = h->seg_start.p;
    c.rec_cap = (uint32_t) std::min<size_t>(std::min(h->rec.n, h->rec_part.n), 0xFFFFFFFFu);
    c.cand_person = h->cand_person.p; c.cand_gkey = h->cand_gkey.p; c.cand_t = h->cand_t.p; c.cand_p = h->cand_p.p;
    c.cand_cap = (uint32_t) std::min<size_t>(std::min(std::min(h->cand_person.n, h->cand_gkey.n), std::min(h->cand_t.n, h->cand_p.n)), 0xFFFFFFFFu);
    c.best_t = h->d_best_t.p;
    for (int k = 0; k < 2; ++k) c.sub_seg[k] = h->sub_seg[k].p;
    for (int k = 0; k < 2; ++k) c.sub_cap[k] = (uint32_t) h->sub_seg[k].n;
    for (int k = 0; k < N_HOT; ++k) c.hot_seg[k] = h->hot_seg[k].p;
    c.hot_cap = big.hot_cap;
    c.tab_now = h->tab_now.p; c.tab_dur = h->tab_dur.p; c.tab_head = h->tab_head.p;
    c.tab_cap = std::min(std::min(h->tab_now.n, h->tab_dur.n), h->tab_head.n);
    c.item_key = h->item_key.p; c.item_tab = h->item_tab.p; c.item_cnt = h->item_cnt.p;
    c.item_nill = h->item_nill.p; c.item_nsus = h->item_nsus.p; c.item_part = h->item_part.p;
    c.part_table = h->part_table.p; c.part_cap = h->part_table.n;
    c.item_cap = std::min(std::min(std::min(h->item_key.n, h->item_tab.n), h->item_cnt.n), std::min(std::min(h->item_nill.n, h->item_nsus.n), h->item_part.n));
    c.huge_off = h->huge_off.p; c.scratch = h->scratch.p; c.scratch_cap = h->scratch.n;
    c.flags = h->cfg.flags;
    c.ctr = h->d_ctr.p;
    TransmitStreams ts{};
    ts.main = s;
    for (int k = 0; k < N_AUX; ++k) { ts.aux[k] = h->aux[k]; ts.join[k] = we.join[k]; }
    ts.fork = we.fork; ts.mid = we.mid;
    launch_transmit(c, h->n_sm, ts, launches);
    CU(h, cudaEventRecord(we.transmitted, s));
    CU(h, cudaGetLastError());
    h->win_state = 2;
    return HEROS_OK;
}

// 5. resolve: earliest successful draw per person wins.  Nothing is read back: the window is finished on the host as
// soon as it is enqueued.
int32_t window_finish(heros_handle h)
{
    cudaStream_t s = h->stream;
    double T1 = h->win_T1;
    uint32_t N = h->n_agents;
    WindowEvents& we = h->ring[h->win_ring];
    AgentArrays A = agent_arrays(h);
    CandArrays C{h->cand_person.p, h->cand_gkey.p, h->cand_t.p, h->cand_p.p,
                 (uint32_t) std::min<size_t>(std::min(std::min(h->cand_person.n, h->cand_gkey.n), std::min(h->cand_t.n, h->cand_p.n)), 0xFFFFFFFFu)};
    InfLog log{h->inf_person.p, h->inf_gkey.p, h->inf_t.p, h->inf_p.p, h->inf_person.n};
    unsigned long long* best_t = h->d_best_t.p;
    unsigned long long* best_key = h->d_best_key.p;
    uint32_t* claim = h->d_claim.p;
    long long* series = h->d_series.p;
    if (h->resolve_grid == 0)
    {
        int per_sm = 0;
        CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resolve, 1024, 0));
        if (per_sm < 1) return fail(h, HEROS_ERR_CUDA, "k_resolve does not fit on an SM");
        h->resolve_grid = h->n_sm;
    }
    void* args[] = {&C, &A, &best_t, &best_key, &claim, &log, &T1, &N, &series};
    CU(h, cudaLaunchCooperativeKernel((const void*) k_resolve, dim3(h->resolve_grid), dim3(1024), args, 0, s));
    h->win_launches += 1;
    CU(h, cudaEventRecord(we.resolved, s));
    if (h->d_timeline.p) k_timeline_fold<<<1, 32, 0, s>>>(h->d_timeline.p);
    CU(h, cudaGetLastError());
    we.pending = true;
    we.window_no = (int64_t) h->series_t.size();  // 0-based number of this window since the persons were set
    h->launches_total += h->win_launches;
    h->win_state = 0;
    h->n_guest = 0;
    h->t_now = T1;
    h->series_t.push_back(T1);
    return HEROS_OK;
}

int32_t check_window_length(heros_handle h, double T1)
{
    // batch == sequential only while an infection of the window cannot become contagious inside it;
    // the float exposure time moves the onset by < 1e-3 h
    const double latent = h->cfg.model == HEROS_MODEL_AREA ? h->area.t_e_min : h->dist.L;
    if (!(h->cfg.window_hours <= latent - 1e-3) || !(T1 - h->t_now <= latent - 1e-3))
        return fail(h, HEROS_ERR_UNSUPPORTED, "the window must be shorter than the latent period of the transmission model");

    return HEROS_OK;
}

}  // namespace heros_impl
using namespace heros_impl;

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

int32_t heros_abi_version(void) { return HEROS_ABI_VERSION; }

const char* heros_last_error(heros_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int32_t heros_create(const heros_config* cfg, heros_handle* out)
{
    if (!cfg || !out) return fail(nullptr, HEROS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != HEROS_ABI_VERSION) return fail(nullptr, HEROS_ERR_INVALID, "ABI version mismatch");
    if (cfg->model != HEROS_MODEL_AREA && cfg->model != HEROS_MODEL_DISTANCE)
        return fail(nullptr, HEROS_ERR_INVALID, "unknown transmission model");
    if (cfg->trigger != HEROS_TRIGGER_EXIT_ONLY && cfg->trigger != HEROS_TRIGGER_ENTER_AND_EXIT)
        return fail(nullptr, HEROS_ERR_INVALID, "unknown trigger");
    if (!(cfg->window_hours > 0)) return fail(nullptr, HEROS_ERR_INVALID, "window_hours must be positive");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, HEROS_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                                 " (the engine has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, HEROS_ERR_INVALID, "device ordinal out of range");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, HEROS_ERR_CUDA, cudaGetErrorString(e));
    heros_handle h = new (std::nothrow) heros_engine();
    if (!h) return fail(nullptr, HEROS_ERR_INVALID, "out of host memory");
    h->cfg = *cfg;
    auto bail = [&](const char* what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        heros_destroy(h);
        return (int32_t) HEROS_ERR_CUDA;
    };
    {
        // k_submit of the NEXT window (side stream) becomes ready in the middle of the current one; thread blocks are handed
        // out grid by grid in the order the grids became ready, so on equal terms its 2 000 blocks would be placed in front
        // of the transmission kernels of the window that is running (heros_debug_timeline: + 30 us).  It has a whole
        // window to finish: lower priority.
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // numerically: hi <= lo
        const bool flat = getenv("HEROS_TUNE_FLAT_PRIO") != nullptr;
        // three levels where the device has them: the hot path (below) above the engine stream above k_submit
        const int mid = hi < lo - 1 ? hi + 1 : hi;
        if ((e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, flat ? lo : mid)) != cudaSuccess) return bail("stream", e);
        if ((e = cudaStreamCreateWithPriority(&h->sub_stream, cudaStreamNonBlocking, lo)) != cudaSuccess) return bail("stream", e);
    }
    if ((e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&h->meta_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    for (auto& ev : h->ev_keep)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    if ((e = cudaMallocHost(&h->h_keep, sizeof(unsigned long long) * WIN_RING)) != cudaSuccess) return bail("pinned counters", e);
    {
        // The hot path (aux 0, 2) is the longer of the two branches of the transmission stage.  Thread blocks are handed
        // out grid by grid in the order the grids became ready, so behind the streaming kernel (and k_submit of the next
        // window) the chain kernels only start when those have no block left to place (heros_debug_timeline shows it);
        // on streams of higher priority (HEROS_TUNE_AUX_PRIO=1) their blocks go first whenever a slot frees up.
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        // HEROS_TUNE_AUX_PRIO=0: the hot path on the level of the engine stream (both branches then mostly follow each other)
        const char* knob = getenv("HEROS_TUNE_AUX_PRIO");
        const bool hot_first = !(knob && knob[0] == '0');
        const int mid = hi < lo - 1 ? hi + 1 : hi;
        for (int k = 0; k < N_AUX; ++k)
        {
            const int level = getenv("HEROS_TUNE_FLAT_PRIO") ? lo : ((hot_first && k != 1) ? hi : mid);
            if ((e = cudaStreamCreateWithPriority(&h->aux[k], cudaStreamNonBlocking, level)) != cudaSuccess) return bail("stream", e);
        }
    }
    for (auto& ev : h->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("event", e);
    for (cudaEvent_t* ev : {&h->ev_rolled, &h->ev_submitted, &h->ev_copied, &h->ev_stage_free[0], &h->ev_stage_free[1]})
        if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    for (auto& ev : h->ev_chunk)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    for (auto& we : h->ring)
    {
        for (cudaEvent_t* ev : {&we.begin, &we.opened, &we.scanned, &we.scattered, &we.mid, &we.transmitted, &we.resolved,
                                &we.fork, &we.join[0], &we.join[1], &we.join[2]})
            if ((e = cudaEventCreate(ev)) != cudaSuccess) return bail("event", e);
        if ((e = cudaEventCreateWithFlags(&we.filtered, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    }
    static_assert(N_AUX == 3, "the events of the auxiliary streams are listed by hand");
    if ((e = h->d_series.ensure((size_t) SERIES_RING * 9)) != cudaSuccess) return bail("series", e);
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
    if ((e = h->d_ctr.ensure(1)) != cudaSuccess) return bail("counters", e);
    if ((e = cudaMemsetAsync(h->d_ctr.p, 0, sizeof(Counters), h->stream)) != cudaSuccess) return bail("memset", e);
    if ((e = cudaMallocHost(&h->h_ctr, sizeof(Counters))) != cudaSuccess) return bail("pinned counters", e);
    if ((e = h->d_policy.ensure(16)) != cudaSuccess) return bail("policy", e);
    if ((e = h->d_prog.ensure(1)) != cudaSuccess) return bail("prog", e);
    *out = h;
    return HEROS_OK;
}

int32_t heros_destroy(heros_handle h)
{
    if (!h) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    if (h->sub_stream) cudaStreamSynchronize(h->sub_stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& a : h->aux) if (a) cudaStreamSynchronize(a);
    h->d_policy.release(); h->d_prog.release(); h->d_loc_base.release(); h->d_key_loc.release();
    h->d_loc_nsub.release(); h->d_loc_param.release(); h->d_last_calc.release(); h->d_view.release();
    h->d_key_total.release(); h->d_total_loc.release(); h->d_loc_key.release(); h->d_rate_loc.release();
    h->d_next_time.release(); h->d_ill_end.release(); h->d_open_tin.release(); h->d_best_t.release();
    h->d_best_key.release(); h->d_claim.release(); h->d_open_key.release();
    for (int i = 0; i < 2; ++i)
    {
        h->carry_person[i].release(); h->carry_key[i].release(); h->carry_tin[i].release(); h->carry_tout[i].release();
        h->carry_rank[i].release();
        h->st_person[i].release(); h->st_loc[i].release(); h->st_sub[i].release(); h->st_tin[i].release(); h->st_tout[i].release();
    }
    h->fresh_person.release(); h->fresh_key.release(); h->fresh_tin.release(); h->fresh_tout.release(); h->fresh_rank.release();
    h->d_series.release();
    h->bk_count.release(); h->bk_rank.release(); h->scan_status.release(); h->rec.release();
    h->seg_start.release(); h->g_rank.release(); h->cand_person.release();
    h->cand_gkey.release(); h->cand_t.release();
    h->cand_p.release(); h->inf_person.release(); h->inf_gkey.release(); h->tr_person.release(); h->inf_t.release();
    h->g_gid.release(); h->g_key.release(); h->g_view.release(); h->g_ill_end.release(); h->g_tin.release();
    h->g_tout.release(); h->d_nguest_cand.release();
    h->inf_p.release(); h->tr_time.release(); h->tr_phase.release(); h->sub_seg[0].release(); h->sub_seg[1].release();
    h->item_key.release(); h->item_tab.release(); h->item_cnt.release();
    h->tab_now.release(); h->tab_dur.release(); h->tab_head.release();
    for (auto& b : h->hot_seg) b.release();
    h->rec_part.release(); h->item_nill.release(); h->item_nsus.release(); h->item_part.release(); h->part_table.release();
    h->d_timeline.release();
    h->huge_off.release(); h->scratch.release();
    h->s_closure.release(); h->s_loc_type.release(); h->m_times.release(); h->m_sub.release(); h->m_loc.release();
    h->s_order.release(); h->s_acts.release(); h->s_pat_start.release(); h->s_pat_count.release(); h->s_choice_start.release();
    h->s_choice_type.release(); h->s_nearest.release(); h->s_ncand.release(); h->s_cand.release(); h->d_home_loc.release();
    h->d_work_loc.release(); h->s_choice_cum.release(); h->d_home_sub.release(); h->d_work_sub.release();
    h->d_ctr.release(); h->x_person.release(); h->x_time.release();
    for (int d = 0; d < 64; ++d)
        if (h->xc.opened[d]) cudaIpcCloseMemHandle(h->xc.peer[d]);
    if (h->xc.region) cudaFree(h->xc.region);
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    if (h->h_keep) cudaFreeHost(h->h_keep);
    if (h->pin_buf) cudaFreeHost(h->pin_buf);
    for (auto& ev : h->ev_keep) if (ev) cudaEventDestroy(ev);
    if (h->meta_stream) cudaStreamDestroy(h->meta_stream);
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : {h->ev_rolled, h->ev_submitted, h->ev_copied, h->ev_stage_free[0], h->ev_stage_free[1]})
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->ev_chunk) if (ev) cudaEventDestroy(ev);
    for (auto& we : h->ring)
        for (cudaEvent_t ev : {we.begin, we.opened, we.scanned, we.scattered, we.mid, we.transmitted, we.resolved, we.fork,
                               we.filtered, we.join[0], we.join[1], we.join[2]})
            if (ev) cudaEventDestroy(ev);
    for (auto& a : h->aux) if (a) cudaStreamDestroy(a);
    if (h->sub_stream) cudaStreamDestroy(h->sub_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return HEROS_OK;
}

int32_t heros_set_location_types(heros_handle h, int32_t n, const uint8_t* infect_sub, const double* corr,
                                 const uint8_t* cap_constrained, const double* cap_per_m2, const double*)
{
    if (!h) return HEROS_ERR_INVALID;
    if (n <= 0 || n > 256 || !infect_sub || !corr) return fail(h, HEROS_ERR_INVALID, "bad location type table");
    h->n_types = n;
    h->type_infect_sub.assign(infect_sub, infect_sub + n);
    h->type_corr.assign(corr, corr + n);
    // capConstrained / capPersonsPerM2 (locationtypes.csv columns 8-9): read by the capacity-constrained locators of the
    // device-side activity schedule
    h->type_cap_constrained.assign((size_t) n, 0);
    h->type_cap_per_m2.assign((size_t) n, 0.0);
    if (cap_constrained) h->type_cap_constrained.assign(cap_constrained, cap_constrained + n);
    if (cap_per_m2) h->type_cap_per_m2.assign(cap_per_m2, cap_per_m2 + n);
    return HEROS_OK;
}

int32_t heros_set_locations(heros_handle h, int32_t n, const uint8_t* type, const int16_t* n_sub, const float* area,
                            const float*, const float*)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->n_types == 0) return fail(h, HEROS_ERR_INVALID, "heros_set_location_types must be called first");
    if (n <= 0 || !type || !n_sub || !area) return fail(h, HEROS_ERR_INVALID, "bad location table");
    cudaSetDevice(h->cfg.device);
    std::vector<uint32_t> base((size_t) n + 1);
    std::vector<LocParam> lp(n);
    std::vector<TotalLoc> total;
    uint64_t b = 0;
    for (int i = 0; i < n; ++i)
    {
        if (type[i] >= h->n_types) return fail(h, HEROS_ERR_INVALID, "location type id out of range");
        if (n_sub[i] < 0) return fail(h, HEROS_ERR_INVALID, "negative number of sub-locations");
        if (!h->type_infect_sub[type[i]] && n_sub[i] >= 2)  // total-location branch (TArea:149 / TDist:131)
        {
            TotalLoc t{};
            t.first_key = (uint32_t) b; t.end_key = (uint32_t) (b + (uint64_t) n_sub[i]);
            t.area_total = (double) area[i];
            t.denom_total = h->type_corr[type[i]] * t.area_total;
            total.push_back(t);
        }
        base[i] = (uint32_t) b;
        b += (uint64_t) std::max<int>(n_sub[i], 1);  // a location with 0 sub-locations still owns one key
        double a = (double) area[i];                 // float getTotalSurfaceM2() widened (TArea:144)
        a /= (double) n_sub[i];                      // TArea:155 / TDist:137 (inf or NaN for n_sub == 0, as in Java)
        lp[i].area_sub = a;
        lp[i].denom = h->type_corr[type[i]] * a;     // TArea:156
    }
    if (b >= 0xFFFFFFF0ull) return fail(h, HEROS_ERR_UNSUPPORTED, "more than 2^32 sub-locations");
    base[n] = (uint32_t) b;
    h->n_loc = (uint32_t) n;
    h->n_keys = (uint32_t) b;
    h->h_loc_base = base;
    h->h_loc_type.assign(type, type + n);
    h->h_loc_nsub.assign(n_sub, n_sub + n);
    h->h_loc_area.assign(area, area + n);
    std::vector<uint32_t> key_loc(h->n_keys);
    for (int i = 0; i < n; ++i)
        for (uint32_t k = base[i]; k < base[i + 1]; ++k) key_loc[k] = (uint32_t) i;
    CU(h, h->d_loc_base.ensure((size_t) n + 1)); CU(h, h->d_key_loc.ensure(h->n_keys));
    CU(h, h->d_loc_nsub.ensure(n)); CU(h, h->d_loc_param.ensure(n)); CU(h, h->d_last_calc.ensure(h->n_keys));
    CU(h, cudaMemcpyAsync(h->d_loc_base.p, base.data(), ((size_t) n + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_key_loc.p, key_loc.data(), (size_t) h->n_keys * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_loc_nsub.p, n_sub, (size_t) n * 2, cudaMemcpyHostToDevice, h->stream));
    {
        std::vector<uint2> lk(n);
        for (int i = 0; i < n; ++i) lk[i] = make_uint2(base[i], (uint32_t) (int32_t) n_sub[i]);
        CU(h, h->d_loc_key.ensure(n));
        CU(h, cudaMemcpyAsync(h->d_loc_key.p, lk.data(), (size_t) n * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    CU(h, cudaMemcpyAsync(h->d_loc_param.p, lp.data(), (size_t) n * sizeof(LocParam), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemsetAsync(h->d_last_calc.p, 0, (size_t) h->n_keys * 8, h->stream));
    h->n_total_loc = (uint32_t) total.size();
    h->n_rate_loc = 0; h->rate_factor_mode = false;
    h->d_key_total.release(); h->d_total_loc.release(); h->d_rate_loc.release();
    h->h_key_flag.assign(h->n_keys, 0);
    if (!total.empty())
    {
        std::vector<uint8_t>& flag = h->h_key_flag;
        for (const TotalLoc& t : total)
            for (uint32_t k = t.first_key; k < t.end_key; ++k) flag[k] = KEYFLAG_TOTAL;
        CU(h, h->d_key_total.ensure(h->n_keys)); CU(h, h->d_total_loc.ensure(total.size()));
        CU(h, cudaMemcpyAsync(h->d_key_total.p, flag.data(), h->n_keys, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(h->d_total_loc.p, total.data(), total.size() * sizeof(TotalLoc), cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    CU(h, h->bk_count.ensure((size_t) h->n_keys + 1));
    CU(h, cudaMemsetAsync(h->bk_count.p, 0, h->bk_count.n * 4, h->stream));  // k_scan_lookback leaves it zeroed from then on
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int32_t heros_set_persons(heros_handle h, int32_t n, const int8_t* age, const uint8_t* female, const uint8_t* role,
                          const int32_t* home_loc, const int16_t* home_sub, const int32_t* work_loc)
{
    if (!h) return HEROS_ERR_INVALID;
    if (n <= 0) return fail(h, HEROS_ERR_INVALID, "bad person table");
    cudaSetDevice(h->cfg.device);
    CU(h, cudaStreamSynchronize(h->sub_stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->sub_used = false;
    const size_t N = (size_t) n;
    h->n_agents = (uint32_t) n;
    h->have_schedule = false;
    h->h_role.assign((size_t) n, 0);
    if (role) h->h_role.assign(role, role + n);
    h->d_home_loc.release(); h->d_home_sub.release(); h->d_work_loc.release();
    h->h_home_loc.clear(); h->h_home_sub.clear(); h->h_work_loc.clear();
    if (home_loc && home_sub && work_loc)  // only read by the device-side activity schedule
    {
        // host copies: heros_set_schedule checks them against the location table before a kernel indexes with them
        h->h_home_loc.assign(home_loc, home_loc + n); h->h_home_sub.assign(home_sub, home_sub + n);
        h->h_work_loc.assign(work_loc, work_loc + n);
        CU(h, h->d_home_loc.ensure(N)); CU(h, h->d_home_sub.ensure(N)); CU(h, h->d_work_loc.ensure(N));
        CU(h, cudaMemcpyAsync(h->d_home_loc.p, home_loc, N * 4, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(h->d_home_sub.p, home_sub, N * 2, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(h->d_work_loc.p, work_loc, N * 4, cudaMemcpyHostToDevice, h->stream));
    }
    CU(h, h->d_view.ensure(N)); CU(h, h->d_next_time.ensure(N)); CU(h, h->d_ill_end.ensure(N));
    CU(h, h->d_open_tin.ensure(N)); CU(h, h->d_best_t.ensure(N)); CU(h, h->d_best_key.ensure(N));
    CU(h, h->d_claim.ensure(N)); CU(h, h->d_open_key.ensure(N));
    CU(h, h->cand_person.ensure(N + 65536)); CU(h, h->cand_gkey.ensure(N + 65536));
    CU(h, h->cand_t.ensure(N + 65536)); CU(h, h->cand_p.ensure(N + 65536));
    CU(h, h->inf_person.ensure(N)); CU(h, h->inf_gkey.ensure(N)); CU(h, h->inf_t.ensure(N)); CU(h, h->inf_p.ensure(N));
    CU(h, h->tr_person.ensure(5 * N)); CU(h, h->tr_phase.ensure(5 * N)); CU(h, h->tr_time.ensure(5 * N));
    DevBuf<int8_t> d_age; DevBuf<uint8_t> d_female, d_role;
    if (age) { CU(h, d_age.ensure(N)); CU(h, cudaMemcpyAsync(d_age.p, age, N, cudaMemcpyHostToDevice, h->stream)); }
    if (female) { CU(h, d_female.ensure(N)); CU(h, cudaMemcpyAsync(d_female.p, female, N, cudaMemcpyHostToDevice, h->stream)); }
    if (role) { CU(h, d_role.ensure(N)); CU(h, cudaMemcpyAsync(d_role.p, role, N, cudaMemcpyHostToDevice, h->stream)); }
    k_init_agents<<<blocks_for(N), TPB, 0, h->stream>>>(h->d_view.p, h->d_next_time.p, h->d_best_t.p, h->d_best_key.p,
                                                        h->d_claim.p, h->d_open_key.p, h->d_open_tin.p, d_age.p,
                                                        d_female.p, d_role.p, (uint32_t) n);
    Counters zero{};
    zero.phase_counts[PH_S] = n;
    if (role)
        for (int32_t i = 0; i < n; ++i) zero.n_ref += role[i] == (uint8_t) ROLE_REFERENCE ? 1ull : 0ull;
    { Counters live{}; cudaMemcpy(&live, h->d_ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost); zero.timeline = live.timeline; }
    CU(h, cudaMemcpyAsync(h->d_ctr.p, &zero, sizeof(Counters), cudaMemcpyHostToDevice, h->stream));
    { const int32_t rc_ = sync_counters(h); if (rc_ != HEROS_OK) return rc_; }
    d_age.release(); d_female.release(); d_role.release();
    h->fresh_ub = h->carry_ub = 0; h->t_now = 0.0; h->win_state = 0;
    cudaStreamSynchronize(h->meta_stream);
    for (bool& p : h->keep_pending) p = false;
    if (h->bk_count.p) CU(h, cudaMemsetAsync(h->bk_count.p, 0, h->bk_count.n * 4, h->stream));  // tickets of dropped stays
    h->series_t.clear(); h->series_counts.clear(); h->series_fetched = 0;
    h->o_inf_n = h->o_tr_n = ~0ull;  // the event logs start again: the sorted output caches are stale
    h->invalid_reported = 0;
    { const int32_t rc_ = ensure_pinned(h, 32 * std::min<size_t>(N, (size_t) 1 << 18) + 4096); if (rc_ != HEROS_OK) return rc_; }
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int32_t heros_set_rate_locations(heros_handle h, int32_t n, const int32_t* location, const double* rate_factor,
                                 const double* rate)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->n_loc == 0) return fail(h, HEROS_ERR_INVALID, "heros_set_locations must be called first");
    if (n < 0 || (n > 0 && (!location || !rate_factor || !rate))) return fail(h, HEROS_ERR_INVALID, "null argument");
    cudaSetDevice(h->cfg.device);
    int32_t rc = check_idle(h, "heros_set_rate_locations");
    if (rc != HEROS_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    std::vector<RateLoc> table;
    std::vector<uint8_t> flag = h->h_key_flag;
    for (uint8_t& f : flag) f &= (uint8_t) ~KEYFLAG_RATE;
    bool factor_mode = false;
    for (int i = 0; i < n; ++i)
    {
        const int64_t l = (int64_t) location[i] - h->location_base;
        if (l < 0 || l >= (int64_t) h->n_loc) return fail(h, HEROS_ERR_INVALID, "rate-based location outside the location table");
        if (!(rate[i] >= 0.0) && !(rate_factor[i] >= 0.0))
            return fail(h, HEROS_ERR_INVALID, "a rate-based location needs an infection rate or an infection-rate factor (>= 0)");
        RateLoc r{h->h_loc_base[l], h->h_loc_base[l + 1], rate_factor[i], rate[i] >= 0.0 ? rate[i] : -1.0};
        for (uint32_t k = r.first_key; k < r.end_key; ++k)
        {
            if (flag[k] & KEYFLAG_RATE) return fail(h, HEROS_ERR_INVALID, "a rate-based location is listed twice");
            flag[k] |= KEYFLAG_RATE;
        }
        factor_mode = factor_mode || !(rate[i] >= 0.0);
        table.push_back(r);
    }
    h->h_key_flag = flag;
    h->n_rate_loc = (uint32_t) table.size();
    h->rate_factor_mode = factor_mode;
    if (h->n_rate_loc || h->n_total_loc)
    {
        CU(h, h->d_key_total.ensure(h->n_keys));
        CU(h, cudaMemcpy(h->d_key_total.p, flag.data(), h->n_keys, cudaMemcpyHostToDevice));
    }
    if (h->n_rate_loc)
    {
        CU(h, h->d_rate_loc.ensure(table.size()));
        CU(h, cudaMemcpy(h->d_rate_loc.p, table.data(), table.size() * sizeof(RateLoc), cudaMemcpyHostToDevice));
    }
    return HEROS_OK;
}

int32_t heros_set_transmission_area(heros_handle h, const heros_area_params* p)
{
    if (!h || !p) return HEROS_ERR_INVALID;
    h->area.contagiousness = p->contagiousness;                  // TArea:63
    h->area.beta = p->beta;                                      // TArea:64
    h->area.t_e_min = p->t_e_min * 24.0;                         // TArea:65
    h->area.t_e_mode = p->t_e_mode * 24.0;                       // TArea:66
    h->area.t_e_max = p->t_e_max * 24.0;                         // TArea:67
    h->area.threshold = p->calculation_threshold / 3600.0;       // TArea:68
    h->have_area = true;
    return HEROS_OK;
}

int32_t heros_set_transmission_distance(heros_handle h, const heros_dist_params* p)
{
    if (!h || !p) return HEROS_ERR_INVALID;
    h->dist.L = p->L * 24.0;                                     // TDist:59
    h->dist.I = p->I * 24.0;                                     // TDist:60
    h->dist.C = p->C * 24.0;                                     // TDist:61
    h->dist.v_max = p->v_max; h->dist.v_0 = p->v_0; h->dist.r = p->r;
    h->dist.psi = p->psi; h->dist.alpha = p->alpha; h->dist.mu = p->mu;
    h->dist.threshold = p->calculation_threshold / 3600.0;       // TDist:68
    h->have_dist = true;
    return HEROS_OK;
}

int32_t heros_set_progression(heros_handle h, const heros_prog_params* p)
{
    if (!h || !p) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    ProgParams pp;
    static_assert(sizeof(pp.fraction) == sizeof(p->fraction), "fraction table layout");
    memcpy(pp.fraction, p->fraction, sizeof(pp.fraction));
    for (int i = 0; i < N_PERIOD; ++i)
    {
        const heros_duration& d = p->period[i];
        if (d.kind < HEROS_DIST_CONSTANT || d.kind > HEROS_DIST_TRIANGULAR)
            return fail(h, HEROS_ERR_UNSUPPORTED, "unsupported duration distribution kind");
        pp.period[i] = Duration{d.kind, 0, d.a, d.b, d.c};
    }
    CU(h, cudaMemcpyAsync(h->d_prog.p, &pp, sizeof(pp), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_prog = true;
    return HEROS_OK;
}

int32_t heros_set_parameter(heros_handle h, const char* name, double value, double at_time)
{
    if (!h || !name) return HEROS_ERR_INVALID;
    int which;
    if (strcmp(name, "psi") == 0) which = 0;       // TDist:254
    else if (strcmp(name, "mu") == 0) which = 1;   // TDist:258
    else return fail(h, HEROS_ERR_UNKNOWN_PARAMETER, std::string("Unrecognized variable name ") + name);
    if (h->cfg.model != HEROS_MODEL_DISTANCE)      // TArea:251-255 rejects every name
        return fail(h, HEROS_ERR_UNKNOWN_PARAMETER, std::string("Unrecognized variable name ") + name);
    cudaSetDevice(h->cfg.device);
    PolicyChange pc{at_time, which, 0, value};
    auto it = std::upper_bound(h->policy.begin(), h->policy.end(), pc,
                               [](const PolicyChange& a, const PolicyChange& b) { return a.t < b.t; });
    h->policy.insert(it, pc);
    CU(h, h->d_policy.ensure(h->policy.size()));
    CU(h, cudaMemcpyAsync(h->d_policy.p, h->policy.data(), h->policy.size() * sizeof(PolicyChange),
                          cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int32_t heros_expose(heros_handle h, int32_t n, const int32_t* person, const double* at_time)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (n < 0 || (n > 0 && (!person || !at_time))) return fail(h, HEROS_ERR_INVALID, "null argument");
    if (n == 0) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    CU(h, h->x_person.ensure(n)); CU(h, h->x_time.ensure(n));
    CU(h, cudaMemcpyAsync(h->x_person.p, person, (size_t) n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->x_time.p, at_time, (size_t) n * 8, cudaMemcpyHostToDevice, h->stream));
    k_expose<<<blocks_for(n), TPB, 0, h->stream>>>(agent_arrays(h), (uint32_t) n, h->x_person.p, h->x_time.p,
                                                   h->n_agents);
    CU(h, cudaGetLastError());
    return sync_counters(h);  // also refreshes the host copy of the counters (transitions, phase counts)
}

int32_t heros_seed_infections(heros_handle h, int32_t n_infected, int32_t min_age, int32_t max_age,
                              int32_t* selected_out)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (n_infected < 0) return fail(h, HEROS_ERR_INVALID, "negative number of infections");
    cudaSetDevice(h->cfg.device);
    std::vector<uint64_t> view(h->n_agents);
    CU(h, cudaMemcpyAsync(view.data(), h->d_view.p, (size_t) h->n_agents * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    // the n eligible persons with the smallest keyed hash: independent of table order, sharding and thread count
    std::vector<std::pair<uint64_t, int32_t>> eligible;
    for (uint32_t a = 0; a < h->n_agents; ++a)
    {
        const int age = (int) view_age(view[a]);
        if (!phase_is_susceptible(view_phase(view[a])) || age < min_age || age > max_age) continue;
        const uint32_t gid = h->person_base + a;
        const Philox4 r = philox4x32_10(gid, 0, 0, TAG_SEEDING, (uint32_t) h->cfg.seed, (uint32_t) (h->cfg.seed >> 32));
        eligible.emplace_back(((uint64_t) r.x << 32) | r.y, (int32_t) gid);
    }
    if ((size_t) n_infected > eligible.size())
        return fail(h, HEROS_ERR_INVALID, "fewer eligible persons than infections requested");
    std::partial_sort(eligible.begin(), eligible.begin() + n_infected, eligible.end());
    std::vector<int32_t> chosen(n_infected);
    for (int i = 0; i < n_infected; ++i) chosen[i] = eligible[i].second;
    std::sort(chosen.begin(), chosen.end());
    std::vector<double> t0(n_infected, 0.0);
    if (selected_out) memcpy(selected_out, chosen.data(), (size_t) n_infected * 4);
    return heros_expose(h, n_infected, chosen.data(), t0.data());
}

// k_submit of a batch of stays: on the side stream, so that it can run beside the transmission stage of a window that
// is still in progress (it only needs that window's pools to have been consumed: ev_rolled).  ready: an event the
// input columns depend on (the host -> device copies of heros_submit_visits), or nullptr.
static int32_t submit_from_device(heros_handle h, int64_t n, const int32_t* person, const int32_t* loc,
                                  const int16_t* sub, const double* tin, const double* tout, cudaEvent_t ready, int stage)
{
    if (h->win_state != 0)
        return fail(h, HEROS_ERR_INVALID, "stays cannot be submitted while a window opened by heros_window_begin is in progress");
    if ((uint64_t) h->carry_ub + (uint64_t) h->fresh_ub + (uint64_t) n + h->n_agents >= 0xFFFFFFF0ull)
        return fail(h, HEROS_ERR_CAPACITY, "more than 2^32 stays in one window");
    int32_t rc = ensure_fresh(h, h->fresh_ub + (size_t) n);
    if (rc != HEROS_OK) return rc;
    cudaStream_t ss = h->sub_stream;
    if (ready) CU(h, cudaStreamWaitEvent(ss, ready, 0));
    if (!(h->cfg.flags & HEROS_FLAG_RESIDENT_INPUTS) && !ready)
    {
        // the caller's device buffers may have been produced on the engine stream: order the kernel behind it
        CU(h, cudaEventRecord(h->ev_copied, h->stream));
        CU(h, cudaStreamWaitEvent(ss, h->ev_copied, 0));
    }
    else
        CU(h, cudaStreamWaitEvent(ss, h->ev_rolled, 0));  // the pools of the last window have been consumed
    const SubmitIn in{person, loc, sub, tin, tout};
    h->ctr_current = false;
    k_submit<<<item_blocks(n), TPB, 0, ss>>>(in, (uint32_t) n, fresh_arrays(h), (uint32_t) std::min<size_t>(h->fresh_person.n, 0xFFFFFFFFu),
                                             h->d_open_key.p, h->d_open_tin.p, h->d_loc_key.p, h->n_loc, h->n_agents,
                                             h->person_base, h->location_base, h->t_now, h->bk_count.p, h->d_ctr.p);
    CU(h, cudaGetLastError());
    CU(h, cudaEventRecord(h->ev_submitted, ss));
    if (stage >= 0) CU(h, cudaEventRecord(h->ev_stage_free[stage], ss));
    h->sub_used = true;
    h->fresh_ub += (size_t) n;
    h->launches_total += 1;
    return HEROS_OK;
}

// The two halves of heros_submit_visits.  begin: the columns start travelling to one of two staging sets on the copy
// stream, in ST_CHUNKS pieces -- beside whatever the engine stream is doing -- and the call returns at once.  end: k_submit
// of every piece is enqueued behind its copy (and behind the window that consumes the pools, which must have been enqueued
// by now), so that all but the last piece are taken in while the rest still crosses PCIe; the call returns when the
// caller's buffers have been read.
static inline int64_t chunk_begin(int64_t n, int c) { return (n * c / heros_engine::ST_CHUNKS) & ~(int64_t) 63; }

int32_t heros_submit_visits_begin(heros_handle h, int64_t n, const int32_t* person, const int32_t* loc, const int16_t* sub,
                                  const double* tin, const double* tout)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (n < 0 || (n > 0 && (!person || !loc || !sub || !tin || !tout))) return fail(h, HEROS_ERR_INVALID, "null argument");
    if (h->st_pending_n >= 0) return fail(h, HEROS_ERR_INVALID, "heros_submit_visits_begin: the previous submission was not ended");
    cudaSetDevice(h->cfg.device);
    const int k = h->st_cur;
    h->st_cur ^= 1;
    h->st_pending_n = n;
    h->st_pending_k = k;
    if (n == 0) return HEROS_OK;
    cudaStream_t cs = h->copy_stream;
    CU(h, cudaStreamWaitEvent(cs, h->ev_stage_free[k], 0));  // the kernel that read this staging set has finished
    if ((size_t) n > h->st_person[k].n)
    {
        CU(h, cudaStreamSynchronize(h->sub_stream));
        CU(h, h->st_person[k].ensure(n)); CU(h, h->st_loc[k].ensure(n)); CU(h, h->st_sub[k].ensure(n));
        CU(h, h->st_tin[k].ensure(n)); CU(h, h->st_tout[k].ensure(n));
    }
    for (int c = 0; c < heros_engine::ST_CHUNKS; ++c)
    {
        const int64_t a = c == 0 ? 0 : chunk_begin(n, c), b = c + 1 == heros_engine::ST_CHUNKS ? n : chunk_begin(n, c + 1);
        if (b > a)
        {
            const size_t m = (size_t) (b - a);
            CU(h, cudaMemcpyAsync(h->st_person[k].p + a, person + a, m * 4, cudaMemcpyHostToDevice, cs));
            CU(h, cudaMemcpyAsync(h->st_loc[k].p + a, loc + a, m * 4, cudaMemcpyHostToDevice, cs));
            CU(h, cudaMemcpyAsync(h->st_sub[k].p + a, sub + a, m * 2, cudaMemcpyHostToDevice, cs));
            CU(h, cudaMemcpyAsync(h->st_tin[k].p + a, tin + a, m * 8, cudaMemcpyHostToDevice, cs));
            CU(h, cudaMemcpyAsync(h->st_tout[k].p + a, tout + a, m * 8, cudaMemcpyHostToDevice, cs));
        }
        CU(h, cudaEventRecord(h->ev_chunk[c], cs));
    }
    return HEROS_OK;
}

int32_t heros_submit_visits_end(heros_handle h)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->st_pending_n < 0) return fail(h, HEROS_ERR_INVALID, "heros_submit_visits_end without heros_submit_visits_begin");
    cudaSetDevice(h->cfg.device);
    const int64_t n = h->st_pending_n;
    const int k = h->st_pending_k;
    h->st_pending_n = -1;
    if (n == 0) return HEROS_OK;
    int last = 0;
    for (int c = 0; c < heros_engine::ST_CHUNKS; ++c)
        if ((c + 1 == heros_engine::ST_CHUNKS ? n : chunk_begin(n, c + 1)) > (c == 0 ? 0 : chunk_begin(n, c))) last = c;
    int32_t rc = HEROS_OK;
    for (int c = 0; c < heros_engine::ST_CHUNKS && rc == HEROS_OK; ++c)
    {
        const int64_t a = c == 0 ? 0 : chunk_begin(n, c), b = c + 1 == heros_engine::ST_CHUNKS ? n : chunk_begin(n, c + 1);
        if (b > a)
            rc = submit_from_device(h, b - a, h->st_person[k].p + a, h->st_loc[k].p + a, h->st_sub[k].p + a, h->st_tin[k].p + a,
                                    h->st_tout[k].p + a, h->ev_chunk[c], c == last ? k : -1);
    }
    if (rc != HEROS_OK)
    {
        // nothing may be left reading the caller's buffers or writing the staging set when the error is reported
        cudaStreamSynchronize(h->copy_stream);
        cudaEventRecord(h->ev_stage_free[k], h->sub_stream);
        return rc;
    }
    CU(h, cudaStreamSynchronize(h->copy_stream));  // the caller may reuse its buffers
    return HEROS_OK;
}

int32_t heros_submit_visits(heros_handle h, int64_t n, const int32_t* person, const int32_t* loc, const int16_t* sub,
                            const double* tin, const double* tout)
{
    if (h && h->win_state != 0)
        return fail(h, HEROS_ERR_INVALID, "stays cannot be submitted while a window opened by heros_window_begin is in progress");
    const int32_t rc = heros_submit_visits_begin(h, n, person, loc, sub, tin, tout);
    if (rc != HEROS_OK) return rc;
    return heros_submit_visits_end(h);
}

int32_t heros_submit_visits_device(heros_handle h, int64_t n, const int32_t* person, const int32_t* loc,
                                   const int16_t* sub, const double* tin, const double* tout)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (n < 0 || (n > 0 && (!person || !loc || !sub || !tin || !tout))) return fail(h, HEROS_ERR_INVALID, "null argument");
    if (n == 0) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    return submit_from_device(h, n, person, loc, sub, tin, tout, nullptr, -1);  // asynchronous
}

// Statistics of a call (or of the windows between two heros_wait calls) = differences of the device counters, which
// count over the whole life of the engine, plus the stage times of the windows harvested in between.
static void begin_stats(heros_handle h)
{
    h->win_before = *h->h_ctr;
    h->launches_before = h->launches_total;
    for (double& m : h->stage_ms) m = 0.0;
    for (double& m : h->bucket_part_ms) m = 0.0;
    for (double& m : h->aux_done_ms) m = 0.0;
}

static void fill_stats(heros_handle h, heros_step_stats& st)
{
    const Counters& now = *h->h_ctr;
    const Counters& before = h->win_before;
    st.windows = (int64_t) (now.windows - before.windows);
    st.visits = (int64_t) (now.visits_total - before.visits_total);
    st.sublocations = (int64_t) (now.seg_total - before.seg_total);
    st.big_sublocations = (int64_t) (now.hot_total - before.hot_total);
    st.gpu_launches = h->launches_total - h->launches_before;
    st.evaluations = (int64_t) (now.evaluations - before.evaluations);
    st.draws = (int64_t) (now.draws - before.draws);
    st.infections = (int64_t) (now.infections - before.infections);
    st.transitions = (int64_t) (now.transitions - before.transitions);
    st.near_ties = (int64_t) (now.near_ties - before.near_ties);
    st.big_stays = (int64_t) (now.big_stays - before.big_stays);
    st.progress_ms = h->stage_ms[0]; st.bucket_ms = h->stage_ms[1];
    st.transmit_small_ms = h->stage_ms[2]; st.transmit_big_ms = h->stage_ms[3];
    st.resolve_ms = h->stage_ms[4]; st.retire_ms = h->stage_ms[5];
    st.transmit_ms = h->stage_ms[2] + h->stage_ms[3];
    st.bucket_count_ms = h->bucket_part_ms[0]; st.bucket_scan_ms = h->bucket_part_ms[1];
    st.bucket_scatter_ms = h->bucket_part_ms[2];
    // auxiliary streams, time from the fork (after the scatter): 0 the hot path (filter, chain, evaluation), 2 the chain
    // kernel of the small class, 1 the 32-lane sub-warp kernel
    st.group_done_ms[0] = h->aux_done_ms[0]; st.group_done_ms[1] = h->aux_done_ms[2];
    st.group_done_ms[2] = 0.0; st.group_done_ms[3] = 0.0;
    st.sub32_done_ms = h->aux_done_ms[1];
}

int32_t heros_debug_get_stays(heros_handle h, int64_t capacity, int32_t* person, int32_t* location, int16_t* sublocation,
                              double* t_in, double* t_out, int64_t* n_out)
{
    if (!h || !n_out || capacity < 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    int32_t rc = check_idle(h, "heros_debug_get_stays");
    if (rc != HEROS_OK) return rc;
    if ((rc = sync_counters(h)) != HEROS_OK) return rc;  // the fill counts of the pools
    const size_t nf = (size_t) h->h_ctr->n_fresh, nc = (size_t) h->h_ctr->n_carry, np = nf + nc, N = h->n_agents;
    std::vector<uint32_t> pp(np), pk(np), ok(N);
    std::vector<double> pi(np), po(np), oi(N);
    const PoolArrays F = fresh_arrays(h), Cy = carry_arrays(h, h->carry_cur);
    auto fetch = [&](const PoolArrays& P, size_t n, size_t at) -> cudaError_t {
        if (!n) return cudaSuccess;
        cudaError_t e;
        if ((e = cudaMemcpyAsync(pp.data() + at, P.person, n * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(pk.data() + at, P.key, n * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(pi.data() + at, P.tin, n * 8, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return e;
        return cudaMemcpyAsync(po.data() + at, P.tout, n * 8, cudaMemcpyDeviceToHost, h->stream);
    };
    CU(h, fetch(Cy, nc, 0));
    CU(h, fetch(F, nf, nc));
    CU(h, cudaMemcpyAsync(ok.data(), h->d_open_key.p, N * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(oi.data(), h->d_open_tin.p, N * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    int64_t n = 0;
    auto put = [&](uint32_t per, uint32_t key, double a, double b) {
        if (n < capacity)
        {
            const uint32_t loc = (uint32_t) (std::upper_bound(h->h_loc_base.begin(), h->h_loc_base.end(), key) - h->h_loc_base.begin()) - 1;
            if (person) person[n] = (int32_t) (h->person_base + per);
            if (location) location[n] = h->location_base + (int32_t) loc;
            if (sublocation) sublocation[n] = (int16_t) (key - h->h_loc_base[loc]);
            if (t_in) t_in[n] = a;
            if (t_out) t_out[n] = b;
        }
        ++n;
    };
    for (size_t i = 0; i < np; ++i)
        if (pk[i] != KEY_NONE) put(pp[i], pk[i], pi[i], po[i]);
    for (size_t a = 0; a < N; ++a)
        if (ok[a] != KEY_NONE && oi[a] == oi[a]) put((uint32_t) a, ok[a], oi[a], std::numeric_limits<double>::infinity());
    *n_out = n;
    return n > capacity && capacity > 0 ? HEROS_ERR_CAPACITY : HEROS_OK;
}

// the windows up to t_until, enqueued without waiting for any of them
static int32_t enqueue_windows(heros_handle h, double t_until)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if ((rc = check_idle(h, "heros_step")) != HEROS_OK) return rc;
    if (!(t_until >= h->t_now)) return fail(h, HEROS_ERR_INVALID, "t_until lies before the engine time");
    const double latent = h->cfg.model == HEROS_MODEL_AREA ? h->area.t_e_min : h->dist.L;
    if (!(h->cfg.window_hours <= latent - 1e-3))
        return fail(h, HEROS_ERR_UNSUPPORTED, "window_hours must be below the latent period of the transmission model");
    cudaSetDevice(h->cfg.device);
    if (!h->async_open)
    {
        begin_stats(h);
        CU(h, cudaEventRecord(h->ev[0], h->stream));
        h->async_open = true;
    }
    while (h->t_now < t_until)
    {
        const double T1 = std::min(h->t_now + h->cfg.window_hours, t_until);
        static const bool host_times = getenv("HEROS_TUNE_HOSTTIME") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        if ((rc = window_begin(h, T1)) != HEROS_OK) return rc;
        const auto t1 = std::chrono::steady_clock::now();
        if ((rc = window_transmit(h)) != HEROS_OK) return rc;
        const auto t2 = std::chrono::steady_clock::now();
        if ((rc = window_finish(h)) != HEROS_OK) return rc;
        if (host_times)
        {
            const auto t3 = std::chrono::steady_clock::now();
            auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
            fprintf(stderr, "[heros] window to %.0f h: host us begin %.0f transmit %.0f finish %.0f | carry_ub %zu fresh_seen %zu\n", T1,
                    us(t0, t1), us(t1, t2), us(t2, t3), h->carry_ub, h->fresh_seen);
        }
    }
    return HEROS_OK;
}

// waits for everything enqueued, fills the statistics since the last wait and reports what the windows have to report
static int32_t wait_and_report(heros_handle h, heros_step_stats* stats)
{
    cudaSetDevice(h->cfg.device);
    float ms = 0.f;
    if (h->async_open) CU(h, cudaEventRecord(h->ev[1], h->stream));
    int32_t rc = sync_counters(h);
    if (rc != HEROS_OK) return rc;
    if (h->async_open) CU(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    heros_step_stats st{};
    if (h->async_open) fill_stats(h, st);
    st.gpu_ms = ms;
    h->async_open = false;
    if (stats) *stats = st;
    return report_window_errors(h);
}

int32_t heros_step(heros_handle h, double t_until, heros_step_stats* stats)
{
    const int32_t rc = enqueue_windows(h, t_until);
    if (rc != HEROS_OK) return rc;
    return wait_and_report(h, stats);
}

int32_t heros_step_async(heros_handle h, double t_until) { return enqueue_windows(h, t_until); }

int32_t heros_wait(heros_handle h, heros_step_stats* stats)
{
    if (!h) return HEROS_ERR_INVALID;
    const int32_t rc = check_idle(h, "heros_wait");
    if (rc != HEROS_OK) return rc;
    return wait_and_report(h, stats);
}

int32_t heros_set_id_bases(heros_handle h, int32_t person_base, int32_t location_base)
{
    if (!h || person_base < 0 || location_base < 0) return HEROS_ERR_INVALID;
    if (h->fresh_ub != 0 || h->carry_ub != 0 || h->t_now != 0.0) return fail(h, HEROS_ERR_INVALID, "id bases must be set before any stay is submitted");
    h->person_base = (uint32_t) person_base;
    h->location_base = location_base;
    return HEROS_OK;
}

int32_t heros_window_begin(heros_handle h, double t_until)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (h->win_state != 0) return fail(h, HEROS_ERR_INVALID, "a window is already in progress");
    if (!(t_until > h->t_now)) return fail(h, HEROS_ERR_INVALID, "t_until must lie after the engine time");
    if ((rc = check_window_length(h, t_until)) != HEROS_OK) return rc;
    cudaSetDevice(h->cfg.device);
    if (!h->async_open)
    {
        begin_stats(h);
        CU(h, cudaEventRecord(h->ev[0], h->stream));
        h->async_open = true;
    }
    return window_begin(h, t_until);
}

int32_t heros_window_transmit(heros_handle h)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->win_state != 1) return fail(h, HEROS_ERR_INVALID, "heros_window_begin has not been called");
    cudaSetDevice(h->cfg.device);
    return window_transmit(h);
}

// A window that cannot be completed (an exchange call failed between its phases) is carried out with what it has: no
// further guests, no further candidates.  The engine is idle again afterwards; peers that wait for this shard's flags
// are the caller's to release (heros_window_send_guests / _candidates with no rows raise them).
int32_t heros_window_abort(heros_handle h)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->win_state == 0) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    int32_t rc = HEROS_OK;
    if (h->win_state == 1) rc = window_transmit(h);
    if (rc == HEROS_OK && h->win_state == 2) rc = window_finish(h);
    if (rc != HEROS_OK)
    {
        // not even that: the pipeline state is reset by hand (the window's stays are lost)
        h->win_state = 0;
        h->n_guest = 0;
        h->t_now = h->win_T1;
    }
    h->async_open = false;
    sync_counters(h);
    return rc;
}

int32_t heros_window_finish(heros_handle h, heros_step_stats* stats)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "heros_window_transmit has not been called");
    cudaSetDevice(h->cfg.device);
    const int32_t rc = window_finish(h);
    if (rc != HEROS_OK) return rc;
    return wait_and_report(h, stats);
}

// the same without waiting: the window is finished on the device some time later (heros_wait)
int32_t heros_window_finish_async(heros_handle h)
{
    if (!h) return HEROS_ERR_INVALID;
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "heros_window_transmit has not been called");
    cudaSetDevice(h->cfg.device);
    return window_finish(h);
}

int32_t heros_get_infections(heros_handle h, int64_t first, int64_t capacity, int32_t* person, int32_t* location,
                             int16_t* sublocation, double* time, double* p, int64_t* n_out)
{
    if (!h || !n_out || first < 0 || capacity < 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    int32_t rc = current_counters(h);
    if (rc != HEROS_OK) return rc;
    const unsigned long long n = std::min<unsigned long long>(h->h_ctr->n_inf_log, h->inf_person.n);
    if (n < h->o_inf_n || h->o_inf_n == ~0ull) { h->o_inf.clear(); h->o_inf_n = 0; }  // the log was reset
    if (n != h->o_inf_n)
    {
        // only the events logged since the last call are fetched; they are sorted and merged behind the older ones
        // (the events of a later window are later, so the merge normally has nothing to move)
        const size_t old = (size_t) h->o_inf_n, add = (size_t) (n - h->o_inf_n);
        // into pinned memory: four asynchronous copies, one wait (pageable targets would be staged one after the other)
        const size_t need = add * 32 + 64;
        { const int32_t rc_ = ensure_pinned(h, need); if (rc_ != HEROS_OK) return rc_; }
        unsigned char* pb = static_cast<unsigned char*>(h->pin_buf);
        double* t = reinterpret_cast<double*>(pb);
        double* pp = t + add;
        uint64_t* key = reinterpret_cast<uint64_t*>(pp + add);
        uint32_t* per = reinterpret_cast<uint32_t*>(key + add);
        CU(h, cudaMemcpyAsync(per, h->inf_person.p + old, add * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(key, h->inf_gkey.p + old, add * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(t, h->inf_t.p + old, add * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(pp, h->inf_p.p + old, add * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        if (h->o_inf.capacity() < old + add) h->o_inf.reserve(std::max<size_t>(2 * (old + add), 1u << 16));
        h->o_inf.resize(old + add);
        auto before = [](const heros_engine::InfEvent& x, const heros_engine::InfEvent& y) {
            return x.t < y.t || (x.t == y.t && x.person < y.person);
        };
        {
            // order of the new events by (time, person): times are non-negative, so their bit patterns order as integers.
            // LSD radix sort of 16-byte keys by the time bits (digits that are the same in every key -- the high ones of a
            // 24 h window -- are skipped), stable; runs of equal times are then ordered by person; one gather.
            using Key = heros_engine::SortKey;
            std::vector<Key>& keys = h->sort_keys;
            std::vector<Key>& tmp = h->sort_tmp;
            keys.resize(add);
            tmp.resize(add);
            for (size_t i = 0; i < add; ++i)
            {
                uint64_t b;
                memcpy(&b, &t[i], 8);
                keys[i] = Key{b, per[i], (uint32_t) i};
            }
            for (int shift = 0; shift < 64; shift += 11)
            {
                uint32_t hist[2049] = {};
                for (const Key& k : keys) ++hist[((k.t >> shift) & 0x7FF) + 1];
                bool trivial = false;
                for (int b = 1; b <= 2048; ++b) trivial = trivial || hist[b] == add;
                if (trivial) continue;
                for (int b = 0; b < 2048; ++b) hist[b + 1] += hist[b];
                for (const Key& k : keys) tmp[hist[(k.t >> shift) & 0x7FF]++] = k;
                keys.swap(tmp);
            }
            for (size_t i = 0; i < add;)
            {
                size_t j = i + 1;
                while (j < add && keys[j].t == keys[i].t) ++j;
                if (j - i > 1) std::sort(keys.begin() + i, keys.begin() + j, [](const Key& x, const Key& y) { return x.person < y.person; });
                i = j;
            }
            for (size_t i = 0; i < add; ++i)
            {
                const uint32_t k = keys[i].idx;
                h->o_inf[old + i] = heros_engine::InfEvent{t[k], pp[k], key[k], per[k]};
            }
        }
        if (old && add && before(h->o_inf[old], h->o_inf[old - 1]))
            std::inplace_merge(h->o_inf.begin(), h->o_inf.begin() + old, h->o_inf.end(), before);
        h->o_inf_n = n;
    }
    const int64_t avail = std::max<int64_t>(0, (int64_t) n - first);
    const int64_t m = std::min(avail, capacity);
    for (int64_t i = 0; i < m; ++i)
    {
        const size_t k = (size_t) (first + i);
        const uint64_t gkey = h->o_inf[k].gkey;  // (global location id << 16) | sub-location
        if (person) person[i] = (int32_t) h->o_inf[k].person;
        if (location) location[i] = (int32_t) (gkey >> 16);
        if (sublocation) sublocation[i] = (int16_t) (gkey & 0xFFFFu);
        if (time) time[i] = h->o_inf[k].t;
        if (p) p[i] = h->o_inf[k].p;
    }
    *n_out = avail;
    return m < avail && capacity > 0 ? HEROS_ERR_CAPACITY : HEROS_OK;
}

int32_t heros_get_infections_unordered(heros_handle h, int64_t first, int64_t capacity, int32_t* person, int32_t* location,
                                       int16_t* sublocation, double* time, double* p, int64_t* n_out)
{
    if (!h || !n_out || first < 0 || capacity < 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    int32_t rc = current_counters(h);
    if (rc != HEROS_OK) return rc;
    const int64_t n = (int64_t) std::min<unsigned long long>(h->h_ctr->n_inf_log, h->inf_person.n);
    const int64_t avail = std::max<int64_t>(0, n - first);
    const int64_t m = std::min(avail, capacity);
    *n_out = avail;
    if (m > 0)
    {
        const size_t add = (size_t) m, need = add * 28 + 64;
        { const int32_t rc_ = ensure_pinned(h, need); if (rc_ != HEROS_OK) return rc_; }
        double* t = reinterpret_cast<double*>(h->pin_buf);
        double* pp = t + add;
        uint64_t* key = reinterpret_cast<uint64_t*>(pp + add);
        uint32_t* per = reinterpret_cast<uint32_t*>(key + add);
        if (person) CU(h, cudaMemcpyAsync(per, h->inf_person.p + first, add * 4, cudaMemcpyDeviceToHost, h->stream));
        if (location || sublocation) CU(h, cudaMemcpyAsync(key, h->inf_gkey.p + first, add * 8, cudaMemcpyDeviceToHost, h->stream));
        if (time) CU(h, cudaMemcpyAsync(t, h->inf_t.p + first, add * 8, cudaMemcpyDeviceToHost, h->stream));
        if (p) CU(h, cudaMemcpyAsync(pp, h->inf_p.p + first, add * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        if (person) memcpy(person, per, add * 4);
        if (time) memcpy(time, t, add * 8);
        if (p) memcpy(p, pp, add * 8);
        for (size_t i = 0; i < add && (location || sublocation); ++i)
        {
            if (location) location[i] = (int32_t) (key[i] >> 16);  // (global location id << 16) | sub-location
            if (sublocation) sublocation[i] = (int16_t) (key[i] & 0xFFFFu);
        }
    }
    return m < avail && capacity > 0 ? HEROS_ERR_CAPACITY : HEROS_OK;
}

// Results of ONE finished window without waiting for the windows enqueued behind it: the call waits for that window's
// last kernel only and reads on a side stream (the compartment counts were left in the series ring by k_resolve, the
// infection log is append-only and the rows of a finished window never change).
int32_t heros_window_results(heros_handle h, int64_t window, int64_t counts[8], int64_t* infections_logged)
{
    if (!h || window < 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    if (window >= (int64_t) h->series_t.size()) return fail(h, HEROS_ERR_INVALID, "heros_window_results: that window has not been enqueued");
    if ((int64_t) h->series_t.size() - window > SERIES_RING - 1) return fail(h, HEROS_ERR_INVALID, "heros_window_results: that window is too long ago");
    bool found = false;
    for (auto& we : h->ring)
        if (we.window_no == window) { CU(h, cudaEventSynchronize(we.resolved)); found = true; }
    if (!found) CU(h, cudaStreamSynchronize(h->stream));  // its events have been recycled: it finished long ago, or nothing is behind it
    { const int32_t rc_ = ensure_pinned(h, 128); if (rc_ != HEROS_OK) return rc_; }
    long long* pin = static_cast<long long*>(h->pin_buf);
    const size_t slot = (size_t) (window % SERIES_RING);
    CU(h, cudaMemcpyAsync(pin, h->d_series.p + slot * 8, 64, cudaMemcpyDeviceToHost, h->meta_stream));
    CU(h, cudaMemcpyAsync(pin + 8, h->d_series.p + (size_t) SERIES_RING * 8 + slot, 8, cudaMemcpyDeviceToHost, h->meta_stream));
    CU(h, cudaStreamSynchronize(h->meta_stream));
    if (counts) for (int k = 0; k < 8; ++k) counts[k] = (int64_t) pin[k];
    if (infections_logged) *infections_logged = (int64_t) pin[8];
    return HEROS_OK;
}

// Rows [first, first + n) of the infection log as the device wrote them, read on a side stream: for rows of windows that
// have finished (heros_window_results tells how many the log held at the end of a window).  Nothing is waited for.
int32_t heros_get_infections_range(heros_handle h, int64_t first, int64_t n, int32_t* person, int32_t* location,
                                   int16_t* sublocation, double* time, double* p)
{
    if (!h || first < 0 || n < 0) return HEROS_ERR_INVALID;
    if (n == 0) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    if ((size_t) (first + n) > h->inf_person.n) return fail(h, HEROS_ERR_INVALID, "heros_get_infections_range: beyond the log");
    const size_t add = (size_t) n, need = add * 28 + 64;
    { const int32_t rc_ = ensure_pinned(h, need); if (rc_ != HEROS_OK) return rc_; }
    double* t = reinterpret_cast<double*>(h->pin_buf);
    double* pp = t + add;
    uint64_t* key = reinterpret_cast<uint64_t*>(pp + add);
    uint32_t* per = reinterpret_cast<uint32_t*>(key + add);
    cudaStream_t ms = h->meta_stream;
    if (person) CU(h, cudaMemcpyAsync(per, h->inf_person.p + first, add * 4, cudaMemcpyDeviceToHost, ms));
    if (location || sublocation) CU(h, cudaMemcpyAsync(key, h->inf_gkey.p + first, add * 8, cudaMemcpyDeviceToHost, ms));
    if (time) CU(h, cudaMemcpyAsync(t, h->inf_t.p + first, add * 8, cudaMemcpyDeviceToHost, ms));
    if (p) CU(h, cudaMemcpyAsync(pp, h->inf_p.p + first, add * 8, cudaMemcpyDeviceToHost, ms));
    CU(h, cudaStreamSynchronize(ms));
    if (person) memcpy(person, per, add * 4);
    if (time) memcpy(time, t, add * 8);
    if (p) memcpy(p, pp, add * 8);
    for (size_t i = 0; i < add && (location || sublocation); ++i)
    {
        if (location) location[i] = (int32_t) (key[i] >> 16);
        if (sublocation) sublocation[i] = (int16_t) (key[i] & 0xFFFFu);
    }
    return HEROS_OK;
}

int32_t heros_get_transitions(heros_handle h, int64_t first, int64_t capacity, int32_t* person, uint8_t* phase,
                              double* time, int64_t* n_out)
{
    if (!h || !n_out || first < 0 || capacity < 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    int32_t rc = current_counters(h);
    if (rc != HEROS_OK) return rc;
    const unsigned long long n = std::min<unsigned long long>(h->h_ctr->n_tr_log, h->tr_person.n);
    if (n < h->o_tr_n || h->o_tr_n == ~0ull) { h->o_tr.clear(); h->o_tr_n = 0; }  // the log was reset
    if (n != h->o_tr_n)
    {
        // only what was logged since the last call is fetched, sorted and merged behind the older entries (the
        // transitions of a later window are later, so the merge normally has nothing to move)
        const size_t old = (size_t) h->o_tr_n, add = (size_t) (n - h->o_tr_n);
        std::vector<uint32_t> per(add);
        std::vector<uint8_t> ph(add);
        std::vector<double> t(add);
        CU(h, cudaMemcpyAsync(per.data(), h->tr_person.p + old, add * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(ph.data(), h->tr_phase.p + old, add, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(t.data(), h->tr_time.p + old, add * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->o_tr.resize(old + add);
        for (size_t i = 0; i < add; ++i) h->o_tr[old + i] = heros_engine::TrEvent{t[i], per[i], ph[i]};
        auto before = [](const heros_engine::TrEvent& x, const heros_engine::TrEvent& y) {
            return x.t < y.t || (x.t == y.t && x.person < y.person);
        };
        std::sort(h->o_tr.begin() + old, h->o_tr.end(), before);
        if (old && add && before(h->o_tr[old], h->o_tr[old - 1]))
            std::inplace_merge(h->o_tr.begin(), h->o_tr.begin() + old, h->o_tr.end(), before);
        h->o_tr_n = n;
    }
    const int64_t avail = std::max<int64_t>(0, (int64_t) n - first);
    const int64_t m = std::min(avail, capacity);
    for (int64_t i = 0; i < m; ++i)
    {
        const size_t k = (size_t) (first + i);
        if (person) person[i] = (int32_t) h->o_tr[k].person;
        if (phase) phase[i] = h->o_tr[k].phase;
        if (time) time[i] = h->o_tr[k].t;
    }
    *n_out = avail;
    return m < avail && capacity > 0 ? HEROS_ERR_CAPACITY : HEROS_OK;
}

int32_t heros_get_phase_counts(heros_handle h, int64_t counts[8])
{
    if (!h || !counts) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    int32_t rc = current_counters(h);
    if (rc != HEROS_OK) return rc;
    for (int k = 0; k < 8; ++k) counts[k] = (int64_t) h->h_ctr->phase_counts[k];
    return HEROS_OK;
}

int32_t heros_get_count_series(heros_handle h, int64_t capacity, double* times, int64_t* counts, int64_t* n_out)
{
    if (!h || !n_out || capacity < 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    { const int32_t rc_ = current_counters(h); if (rc_ != HEROS_OK) return rc_; }  // fetches the rows of the finished windows
    const int64_t n = (int64_t) std::min(h->series_t.size(), h->series_counts.size() / 8);
    const int64_t m = std::min(n, capacity);
    if (times) memcpy(times, h->series_t.data(), (size_t) m * 8);
    if (counts) memcpy(counts, h->series_counts.data(), (size_t) m * 64);
    *n_out = n;
    return m < n && capacity > 0 ? HEROS_ERR_CAPACITY : HEROS_OK;
}

int32_t heros_get_agent_state(heros_handle h, uint8_t* phase, float* exposure, double* next_time, uint8_t* next_phase)
{
    if (!h || h->n_agents == 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    std::vector<uint64_t> view(h->n_agents);
    CU(h, cudaMemcpyAsync(view.data(), h->d_view.p, (size_t) h->n_agents * 8, cudaMemcpyDeviceToHost, h->stream));
    if (next_time)
        CU(h, cudaMemcpyAsync(next_time, h->d_next_time.p, (size_t) h->n_agents * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    for (uint32_t a = 0; a < h->n_agents; ++a)
    {
        if (phase) phase[a] = (uint8_t) view_phase(view[a]);
        if (exposure) exposure[a] = view_exposure(view[a]);
        if (next_phase) next_phase[a] = (uint8_t) view_next_phase(view[a]);
    }
    return HEROS_OK;
}

int32_t heros_set_agent_state(heros_handle h, int32_t n, const int32_t* person, const uint8_t* phase,
                              const float* exposure, double now)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (n < 0 || (n > 0 && (!person || !phase || !exposure))) return fail(h, HEROS_ERR_INVALID, "null argument");
    if (n == 0) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    DevBuf<uint8_t> d_phase; DevBuf<float> d_exp;
    CU(h, h->x_person.ensure(n)); CU(h, d_phase.ensure(n)); CU(h, d_exp.ensure(n));
    CU(h, cudaMemcpyAsync(h->x_person.p, person, (size_t) n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(d_phase.p, phase, (size_t) n, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(d_exp.p, exposure, (size_t) n * 4, cudaMemcpyHostToDevice, h->stream));
    k_set_agent_state<<<blocks_for(n), TPB, 0, h->stream>>>(agent_arrays(h), (uint32_t) n, h->x_person.p, d_phase.p,
                                                            d_exp.p, now, h->n_agents);
    CU(h, cudaGetLastError());
    rc = sync_counters(h);
    d_phase.release(); d_exp.release();
    return rc;
}

int32_t heros_get_last_calculation(heros_handle h, int32_t location, int16_t sublocation, double* out)
{
    if (!h || !out) return HEROS_ERR_INVALID;
    location -= h->location_base;
    if (location < 0 || (uint32_t) location >= h->n_loc || sublocation < 0) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    const uint32_t key = h->h_loc_base[location] + (uint32_t) sublocation;
    if (key >= h->h_loc_base[location + 1]) return fail(h, HEROS_ERR_INVALID, "sub-location out of range");
    CU(h, cudaMemcpyAsync(out, h->d_last_calc.p + key, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int32_t heros_get_phase_clocks(heros_handle h, uint64_t* clocks)
{
    // development aid of the former group kernels; kept in the ABI, reports nothing
    if (!h || !clocks) return HEROS_ERR_INVALID;
    memset(clocks, 0, 64 * sizeof(uint64_t));
    return HEROS_OK;
}

int32_t heros_debug_timeline(heros_handle h, int32_t enable, double* start_us, double* end_us, int64_t* n_windows)
{
    if (!h) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    int32_t rc = check_idle(h, "heros_debug_timeline");
    if (rc != HEROS_OK) return rc;
    if ((rc = sync_counters(h)) != HEROS_OK) return rc;
    constexpr size_t WORDS = 4 * TL_N + 1 + TL_N;
    std::vector<unsigned long long> host(WORDS, 0ull);
    if (h->d_timeline.p && (start_us || end_us || n_windows))
    {
        CU(h, cudaMemcpy(host.data(), h->d_timeline.p, WORDS * 8, cudaMemcpyDeviceToHost));
        const long long* acc = reinterpret_cast<const long long*>(host.data() + 2 * TL_N);
        for (int k = 0; k < TL_N; ++k)
        {
            const double runs = (double) host[4 * TL_N + 1 + k];
            if (start_us) start_us[k] = runs > 0 ? 1e-3 * (double) acc[2 * k] / runs : -1.0;
            if (end_us) end_us[k] = runs > 0 ? 1e-3 * (double) acc[2 * k + 1] / runs : -1.0;
        }
        if (n_windows) *n_windows = (int64_t) host[4 * TL_N];
    }
    else if (n_windows)
        *n_windows = 0;
    unsigned long long* dev = nullptr;
    if (enable)
    {
        CU(h, h->d_timeline.ensure(WORDS));
        for (int k = 0; k < TL_N; ++k) { host[2 * k] = ~0ull; host[2 * k + 1] = 0ull; }
        for (size_t k = 2 * TL_N; k < WORDS; ++k) host[k] = 0ull;
        CU(h, cudaMemcpy(h->d_timeline.p, host.data(), WORDS * 8, cudaMemcpyHostToDevice));
        dev = h->d_timeline.p;
    }
    else
        h->d_timeline.release();
    CU(h, cudaMemcpy(&h->d_ctr.p->timeline, &dev, sizeof dev, cudaMemcpyHostToDevice));
    return HEROS_OK;
}

int32_t heros_get_stream(heros_handle h, void** stream_out)
{
    if (!h || !stream_out) return HEROS_ERR_INVALID;
    *stream_out = (void*) h->stream;
    return HEROS_OK;
}

int32_t heros_get_time(heros_handle h, double* now)
{
    if (!h || !now) return HEROS_ERR_INVALID;
    *now = h->t_now;
    return HEROS_OK;
}

int32_t heros_infect_people(heros_handle h, int32_t location, int32_t n_persons, const int32_t* persons, int32_t n_all,
                            const int32_t* all_persons, double now, double duration, int32_t* calculated,
                            int32_t* infectious_out, int32_t* n_infectious, int32_t* infected_out, int32_t* n_infected,
                            double* p_infection)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    location -= h->location_base;
    if (location < 0 || (uint32_t) location >= h->n_loc || n_persons < 0 || n_all < 0 || (n_persons && !persons))
        return fail(h, HEROS_ERR_INVALID, "bad argument");
    const int type = h->h_loc_type[location];
    const int infect_sub = h->type_infect_sub[type];
    const int n_sub_locations = h->h_loc_nsub[location];
    const bool total_branch = !(infect_sub || n_sub_locations < 2);
    if (total_branch && n_all > 0 && !all_persons) return fail(h, HEROS_ERR_INVALID, "all_persons required for the total-location branch");
    for (int i = 0; i < n_persons; ++i)
        if ((uint32_t) persons[i] - h->person_base >= h->n_agents) return fail(h, HEROS_ERR_INVALID, "person id out of range");
    for (int i = 0; i < (total_branch ? n_all : 0); ++i)
        if ((uint32_t) all_persons[i] - h->person_base >= h->n_agents) return fail(h, HEROS_ERR_INVALID, "person id out of range");
    cudaSetDevice(h->cfg.device);
    const int cap = std::max(std::max(n_persons, n_all), 1);
    DevBuf<int32_t> d_sub, d_all, d_infectious, d_infected;
    DevBuf<ShimOut> d_out;
    CU(h, d_sub.ensure(cap)); CU(h, d_all.ensure(cap)); CU(h, d_infectious.ensure(cap)); CU(h, d_infected.ensure(cap));
    CU(h, d_out.ensure(1));
    if (n_persons) CU(h, cudaMemcpyAsync(d_sub.p, persons, (size_t) n_persons * 4, cudaMemcpyHostToDevice, h->stream));
    if (total_branch && n_all) CU(h, cudaMemcpyAsync(d_all.p, all_persons, (size_t) n_all * 4, cudaMemcpyHostToDevice, h->stream));
    double psi = h->dist.psi, mu = h->dist.mu;
    for (const PolicyChange& pc : h->policy)
    {
        if (!(pc.t <= now)) break;
        if (pc.which == 0) psi = pc.value; else mu = pc.value;
    }
    k_infect_people<<<1, 32, 0, h->stream>>>(h->cfg.model, h->area, h->dist, psi, mu, h->cfg.seed, infect_sub,
                                             n_sub_locations, (double) h->h_loc_area[location], h->type_corr[type],
                                             h->d_view.p, h->person_base, n_persons, d_sub.p, total_branch ? n_all : 0,
                                             d_all.p, now,
                                             duration, d_infectious.p, d_infected.p, d_out.p);
    CU(h, cudaGetLastError());
    ShimOut o{};
    CU(h, cudaMemcpyAsync(&o, d_out.p, sizeof(o), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (infectious_out && o.n_infectious)
        CU(h, cudaMemcpy(infectious_out, d_infectious.p, (size_t) o.n_infectious * 4, cudaMemcpyDeviceToHost));
    if (infected_out && o.n_infected)
        CU(h, cudaMemcpy(infected_out, d_infected.p, (size_t) o.n_infected * 4, cudaMemcpyDeviceToHost));
    if (calculated) *calculated = o.calculated;
    if (n_infectious) *n_infectious = o.n_infectious;
    if (n_infected) *n_infected = o.n_infected;
    if (p_infection) *p_infection = o.p;
    d_sub.release(); d_all.release(); d_infectious.release(); d_infected.release(); d_out.release();
    return HEROS_OK;
}

}  // extern "C"
