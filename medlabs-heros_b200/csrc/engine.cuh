// engine.cuh -- engine state shared by the translation units of libheros_gpu.so
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/heros_gpu.h"
#include "heros_math.cuh"

namespace heros {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr uint64_t BEST_NONE = ~0ull;

// per-location constants of the transmission formulas, precomputed once (heros_set_locations)
struct LocParam
{
    double area_sub;  // (double) totalSurfaceM2 / nSub         (TArea:155, TDist:137)
    double denom;     // correctionFactorArea * area_sub        (TArea:156)
};

// a location evaluated by the total-location branch (infectInSublocation = false and >= 2 sub-locations,
// TArea:198-246 / TDist:189-246): every evaluation of one of its sub-locations scans the persons of ALL of them
struct TotalLoc
{
    uint32_t first_key, end_key;  // its sub-location keys
    double area_total;            // (double) totalSurfaceM2, not divided  (TArea:144, TDist:126)
    double denom_total;           // correctionFactorArea * area_total     (TArea:205)
};

// one stay of the window in bucket order: everything transmission needs from it in one 32-byte sector
struct alignas(32) StayRec
{
    double tin, tout;
    uint64_t view;    // packed agent word of the person (phase at window start, exposure time, ...)
    uint32_t person;  // GLOBAL person id (person_base + local index for the engine's own agents)
    uint32_t pad;     // bit 31: guest (a person of another shard, submitted for this window only); bits 0..30: its index
};
constexpr uint32_t REC_GUEST = 0x80000000u;

#ifdef __CUDACC__
// a record is one 32-byte sector: moved with ONE 256-bit access (sm_100: LDG.E.256 / STG.E.256)
__device__ __forceinline__ StayRec ld_rec(const StayRec* p)
{
#ifdef HEROS_LD256
    uint64_t a, b, v, w;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(v), "=l"(w) : "l"(p));
    StayRec r;
    r.tin = bits_f64(a); r.tout = bits_f64(b); r.view = v; r.person = (uint32_t) w; r.pad = (uint32_t) (w >> 32);
    return r;
#else
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 h0 = __ldg(q), h1 = __ldg(q + 1);  // {t_in, t_out}, {view, person, pad}
    StayRec r;
    r.tin = bits_f64(((uint64_t) h0.y << 32) | h0.x); r.tout = bits_f64(((uint64_t) h0.w << 32) | h0.z);
    r.view = ((uint64_t) h1.y << 32) | h1.x; r.person = h1.z; r.pad = h1.w;
    return r;
#endif
}
__device__ __forceinline__ void st_rec(StayRec* p, const StayRec& r)
{
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(f64_bits(r.tin)), "l"(f64_bits(r.tout)), "l"(r.view),
                 "l"((uint64_t) r.person | ((uint64_t) r.pad << 32))
                 : "memory");
}
#endif

// ---- heros_debug_timeline: when and for how long every kernel of a window actually ran (the kernels of the transmission
// stage run side by side on several streams; CUDA events only give the end of each stream's work) ----
enum TlId { TL_OPEN = 0, TL_SCAN, TL_SCATTER, TL_ROLL, TL_STREAM, TL_SUB8, TL_SUB32, TL_CHAIN0, TL_CHAIN1, TL_CHAIN2, TL_EVAL,
            TL_TOTAL, TL_RESOLVE, TL_SUBMIT, TL_ADVANCE, TL_RATE, TL_N = 32 };
#ifdef __CUDACC__
struct Counters;
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// a rate-based location (medlabs LocationProbBased, factory/ConstructHerosModel.java:382-388): no occupancy-based
// transmission; a susceptible person leaving it after a stay of d hours is exposed with p = 1 - exp(-lambda d / 24),
// lambda = rate (per day) if rate >= 0, else factor x the share of the reference group exposed the day before
struct RateLoc { uint32_t first_key, end_key; double factor, rate; };
constexpr uint32_t TAG_RATE = 0x500u;     // draw of a rate-based location, keyed by (person, exit time)
constexpr uint32_t ROLE_REFERENCE = 7u;   // Worker: the reference group of the satellite person types (ConstructHerosModel.java:303-308)
constexpr uint8_t KEYFLAG_TOTAL = 1, KEYFLAG_RATE = 2;

struct PolicyChange { double t; int32_t which; int32_t pad; double value; };  // which: 0 psi, 1 mu

// device-side counters of one engine.  The block from `n_cand` on is zeroed at the start of every window.
constexpr int N_HOT = 3;     // classes of sub-locations of > 32 stays (the hot path): <= 1024 / <= 8192 movement events in
                             // shared memory, anything larger in global scratch
constexpr int HOT_HUGE = 2;
constexpr int SERIES_RING = 4096;  // windows whose compartment counts are kept on the device between two host reads
struct Counters
{
    unsigned long long evaluations, draws, near_ties, infections, transitions;
    unsigned long long big_stays;    // stays of the sub-locations handed to the hot path
    unsigned long long n_inf_log;    // cursor of the infection log
    unsigned long long n_tr_log;     // cursor of the transition log
    unsigned long long n_sched;      // closed stays produced by the last heros_advance_schedule
    unsigned int overflow;           // bit 0 cand, 1 scratch, 2 transition log, 3 infection log, 4 exchange block, 5 pool
    unsigned int invalid_visits;     // rejected by heros_submit_visits
    unsigned int late_visits;        // t_in before the engine time
    unsigned int pad;
    long long phase_counts[8];
    // the pools of stays live on the device: nothing of a window is read back by the host
    unsigned long long n_fresh;      // stays submitted since the last window (fresh pool)
    unsigned long long n_carry;      // stays carried over from earlier windows (carry pool in use)
    unsigned long long windows;      // windows finished
    unsigned long long visits_total; // stays bucketed, summed over the windows
    unsigned long long seg_total;    // occupied sub-locations, summed over the windows
    unsigned long long hot_total;    // sub-locations handed to the hot path, summed over the windows
    unsigned long long cand_total;   // candidate infections kept, summed over the windows
    unsigned long long rate_draws;   // draws of the rate-based locations (LocationProbBased)
    unsigned long long* timeline;    // heros_debug_timeline: per kernel {first block start, last block end} (globaltimer ns), or NULL
    // rate-based locations (LocationProbBased): persons of the reference group (role Worker) exposed per simulated day,
    // a ring over day % 4, and the size of the group
    unsigned long long ref_exposed[4];
    unsigned long long n_ref;
    unsigned long long reserved_[2];
    // ---- per window ----
    unsigned long long n_cand;       // candidate infections of the current window (first field of the window block)
    unsigned long long n_sub[2];     // segments queued for the sub-warp kernels: [0] <= 8 stays, [1] 9..32 stays
    unsigned long long n_hot[N_HOT]; // sub-locations of > 32 stays, by class (queued by the prefix-sum kernel)
    unsigned long long reserved_run[N_HOT];
    unsigned long long scratch_top;  // bump allocator for the global scratch of huge segments (bytes)
    unsigned long long n_keep;       // stays kept in the pool after the window
    unsigned long long n_seg;        // occupied sub-locations of the current window
    unsigned long long n_active;     // active stays of the current window
    unsigned long long n_items;      // work items (sub-location, 32 evaluations) of the hot evaluation kernel
    unsigned long long tab_top;      // entries of the table of evaluations of the hot sub-locations
    unsigned int scan_tile, scan_done;  // single-pass prefix sum: next tile to hand out, tiles finished
    // work of very different size (a corner shop, a supermarket at noon) is handed out block by block, not by a fixed stride
    unsigned long long eval_next;      // next work item of k_hot_eval
    unsigned long long n_part;         // PartTables handed out
    unsigned long long chain_next[N_HOT];  // next entry of hot_seg[class] for k_hot_chain
};
// Debug build (-DHEROS_DEBUG, csrc/Makefile `make debug`): every index a window kernel computes is checked before it is used;
// a failed check raises bit 7 of the overflow mask (the window then fails with HEROS_ERR_CAPACITY naming "128") and the
// access is skipped.  In the normal build the macro is the bare condition-free statement.
#ifdef HEROS_DEBUG
#define HEROS_CHECK(ctr, cond) ((cond) ? true : (atomicOr(&(ctr)->overflow, 128u), false))
#else
#define HEROS_CHECK(ctr, cond) (true)
#endif

#ifdef __CUDACC__
// thread 0 of every block stamps the kernel's entry when it starts and when it leaves the kernel (any return path)
struct TlStamp
{
    unsigned long long* tl;
    __device__ __forceinline__ TlStamp(const Counters* ctr, int id) : tl(nullptr)
    {
        if (threadIdx.x == 0)
        {
            unsigned long long* p = ctr->timeline;
            if (p) { tl = p + 2 * id; atomicMin(tl, global_ns()); }
        }
    }
    __device__ __forceinline__ ~TlStamp() { if (tl) atomicMax(tl + 1, global_ns()); }
};
#endif

template <typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t want, bool keep = false, cudaStream_t s = 0)
    {
        if (want <= n) return cudaSuccess;
        size_t cap = want + want / 4 + 1024;
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, cap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p && n) e = cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, s);
        if (p) { cudaStreamSynchronize(s); cudaFree(p); }
        p = q; n = cap;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// The partitioned copy of a hot segment is ordered by the time slot (PART_SLOTS over the window) in which a stay BEGAN, the
// ill stays and the susceptible stays each; per slot the table keeps where its stays start and the latest t_out among them.
// A work item of the evaluation kernel covers a short time span: it reads the stays from the first slot that still holds
// somebody present at its first evaluation to the slot of its last evaluation -- a few slots instead of the whole day.
constexpr int PART_SLOTS = 32;
struct PartTable
{
    uint32_t start[2][PART_SLOTS + 1];   // [ill, susceptible]: first stay of every slot (relative to the list's first stay)
    uint32_t pad_[2];
    double max_tout[2][PART_SLOTS];      // latest t_out of the stays of the slot (-inf: none)
};

// everything the window kernels need, passed by value
struct WindowCtx
{
    // window
    double T0, T1;
    int model, trigger;
    uint64_t seed;
    AreaParams area;
    DistParams dist;
    const PolicyChange* policy; int n_policy;
    // static tables
    const uint32_t* key_loc; const LocParam* loc_param; double* last_calc; uint32_t n_keys;
    const uint8_t* key_total;  // per key: KEYFLAG_TOTAL = its location takes the total-location branch, KEYFLAG_RATE = rate-based
                               // (neither is touched by the occupancy kernels; nullptr: no such location)
    const RateLoc* rate_loc; uint32_t n_rate_loc;
    const TotalLoc* total_loc; uint32_t n_total_loc;
    // agents (own: indexed by global id - person_base; guests: by the index in StayRec::pad)
    const uint64_t* view; const double* ill_end; const double* guest_ill_end;
    uint32_t person_base, n_agents;
    const uint32_t* loc_first_key; int32_t location_base;  // for the global sub-location id of a candidate
    // the window's stays bucketed by sub-location: stays of key k are rec[seg_start[k] .. seg_start[k + 1])
    StayRec* rec;
    uint32_t rec_cap;           // records allocated (debug build: every segment must end inside)
    const uint32_t* seg_start;  // n_keys + 1 entries
    // candidates; best_t: earliest successful draw so far of every own agent (time bits), which lets later and therefore
    // hopeless candidates of the same person be dropped where they arise
    uint32_t* cand_person; uint64_t* cand_gkey; double* cand_t; double* cand_p; uint32_t cand_cap;
    unsigned long long* best_t;
    // work lists filled by k_transmit_stream
    uint32_t* sub_seg[2]; uint32_t sub_cap[2];          // sub-location keys queued for the sub-warp kernels (8 / 32 lanes)
    // hot sub-locations (> 32 stays): queued by the prefix-sum kernel
    uint32_t* hot_seg[N_HOT]; uint32_t hot_cap;
    // partitioned copy of the hot segments that need evaluations (same offsets as rec): ill stays from the front of a
    // segment, susceptible ones from its back -- all the evaluation kernel reads
    StayRec* rec_part;
    unsigned long long* huge_off;  // byte offset into scratch of every entry of hot_seg[HOT_HUGE]
    unsigned char* scratch; unsigned long long scratch_cap;
    // the chain kernel leaves one row per calculated evaluation of a hot sub-location (time, duration, head count) and
    // one work item per 32 of them; the evaluation kernel spreads the items over the whole GPU
    double* tab_now; double* tab_dur; uint32_t* tab_head; unsigned long long tab_cap;
    uint32_t* item_key; unsigned long long* item_tab; uint32_t* item_cnt; uint32_t* item_nill; uint32_t* item_nsus;
    uint32_t* item_part;             // index of the sub-location's PartTable
    unsigned long long item_cap;
    PartTable* part_table; unsigned long long part_cap;
    uint32_t flags;
    Counters* ctr;
};

// Sub-locations of > 32 stays take the hot path.  The prefix-sum kernel (it sees every bucket size) files them by the
// number of movement events the chain kernel has to order: <= 1024 and <= 8192 events are ordered in shared memory
// (26 KB / 205 KB per thread block), anything larger in global scratch.
constexpr int HOT_M_CAP0 = 1024, HOT_M_CAP1 = 8192;
constexpr int HOT_NB_CAP0 = 512, HOT_NB_CAP = 4096;  // time buckets of the event ordering
__host__ __device__ inline unsigned long long hot_scratch_bytes(unsigned long long m)
{
    // ev_t f64[m] | 5 bucket arrays u32[NB + 2] | 3 index arrays u32[m + 1] | ev_k u8[m]
    return (8ull * m + 5ull * 4ull * (HOT_NB_CAP + 2) + 3ull * 4ull * (m + 1) + m + 256ull + 15ull) & ~15ull;
}
struct BigLists
{
    uint32_t* hot_seg[N_HOT]; uint32_t hot_cap;
    unsigned long long* huge_off; unsigned long long scratch_cap;
    const uint8_t* key_total;  // sub-locations of total-branch locations are not queued here
    int need_enters;           // enter events are needed (distance model head counts / ENTER_AND_EXIT trigger)
};

#ifdef __CUDACC__
__device__ __forceinline__ void queue_big(const BigLists& b, Counters* ctr, uint32_t key, uint32_t n)
{
    const uint32_t m = b.need_enters ? 2 * n : n;
    const int cls = m <= (uint32_t) HOT_M_CAP0 ? 0 : (m <= (uint32_t) HOT_M_CAP1 ? 1 : HOT_HUGE);
    const unsigned long long slot = atomicAdd(&ctr->n_hot[cls], 1ull);
    if (slot >= b.hot_cap) { atomicOr(&ctr->overflow, 2u); return; }
    b.hot_seg[cls][slot] = key;
    if (cls == HOT_HUGE)
    {
        const unsigned long long bytes = hot_scratch_bytes(m);
        const unsigned long long off = atomicAdd(&ctr->scratch_top, bytes);
        if (off + bytes <= b.scratch_cap) b.huge_off[slot] = off;
        else { b.hot_seg[cls][slot] = KEY_NONE; atomicOr(&ctr->overflow, 2u); }
    }
    atomicAdd(&ctr->big_stays, (unsigned long long) n);
}
#endif

// the engine stream, the auxiliary streams of the kernels that run side by side, and the fork / join events
constexpr int N_AUX = 3;  // 0: the hot path (chain kernels of the large classes, evaluation), 1: the 32-lane
                          // sub-warp kernel, 2: the chain kernel of the small class
struct TransmitStreams { cudaStream_t main; cudaStream_t aux[N_AUX]; cudaEvent_t fork, mid; cudaEvent_t join[N_AUX]; };

}  // namespace heros
