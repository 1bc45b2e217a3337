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
struct alignas(16) StayRec
{
    double tin, tout;
    uint64_t view;    // packed agent word of the person (phase at window start, exposure time, ...)
    uint32_t person;  // GLOBAL person id (person_base + local index for the engine's own agents)
    uint32_t pad;     // bit 31: guest (a person of another shard, submitted for this window only); bits 0..30: its index
};
constexpr uint32_t REC_GUEST = 0x80000000u;

struct PolicyChange { double t; int32_t which; int32_t pad; double value; };  // which: 0 psi, 1 mu

constexpr int N_CLS = 5;     // classes of sub-locations of > 32 stays (the group kernels)
constexpr int CLS_HUGE = 4;  // ... the last one works in global scratch

// device-side counters of one engine.  The block from `n_cand` on is zeroed at the start of every window.
struct Counters
{
    unsigned long long evaluations, draws, near_ties, infections, transitions;
    unsigned long long big_stays;    // stays of the sub-locations handed to the group kernels
    unsigned long long n_inf_log;    // cursor of the infection log
    unsigned long long n_tr_log;     // cursor of the transition log
    unsigned long long n_sched;      // closed stays produced by the last heros_advance_schedule
    unsigned int overflow;           // bit 0 cand, 1 scratch, 2 transition log, 3 infection log, 4 exchange block, 5 pool
    unsigned int invalid_visits;     // rejected by heros_submit_visits
    unsigned int late_visits;        // t_in before the engine time
    unsigned int pad;
    long long phase_counts[8];
    // ---- per window ----
    unsigned long long n_cand;       // candidate infections of the current window (first field of the window block)
    unsigned long long n_sub[2];     // segments queued for the sub-warp kernels: [0] <= 8 stays, [1] 9..32 stays
    unsigned long long n_blk[N_CLS]; // segments handed to the group kernels (<= 256 / 1024 / 4096 / 8192 events in shared
                                     // memory, larger ones in global scratch)
    unsigned long long n_run[N_CLS]; // ... of which the filter kernel passed on to the group kernels
    unsigned long long scratch_top;  // bump allocator for the global scratch of huge segments (bytes)
    unsigned long long n_keep;       // stays kept in the pool after the window
    unsigned long long n_seg;        // occupied sub-locations of the current window
    unsigned long long n_active;     // active stays of the current window
    unsigned long long n_draw;       // draw items << 40 | (stay, evaluation) pairs they stand for
    unsigned long long tab_top;      // entries of the table of evaluations with p > 0
    unsigned int scan_tile, scan_done;  // single-pass prefix sum: next tile to hand out, tiles finished
    unsigned long long dbg_clk[N_CLS][16];  // HEROS_FLAG_PHASE_CLOCKS: max cycles per phase of k_transmit_group, per class
};
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
    const uint8_t* key_total;  // per key: 1 = its location takes the total-location branch (nullptr: no such location)
    const TotalLoc* total_loc; uint32_t n_total_loc;
    // agents (own: indexed by global id - person_base; guests: by the index in StayRec::pad)
    const uint64_t* view; const double* ill_end; const double* guest_ill_end;
    uint32_t person_base, n_agents;
    const uint32_t* loc_first_key; int32_t location_base;  // for the global sub-location id of a candidate
    // the window's stays bucketed by sub-location: stays of key k are rec[seg_start[k] .. seg_start[k + 1])
    StayRec* rec;
    const uint32_t* seg_start;  // n_keys + 1 entries
    // candidates
    uint32_t* cand_person; uint64_t* cand_gkey; double* cand_t; double* cand_p; uint32_t cand_cap;
    // work lists filled by k_transmit_thread
    uint32_t* sub_seg[2]; uint32_t sub_cap[2];          // sub-location keys queued for the sub-warp kernels (8 / 32 lanes)
    uint32_t* blk_seg[N_CLS]; uint32_t blk_cap;      // ... and for the group kernels
    // draws queued by the group kernels for k_draw: the evaluations with p > 0 of every sub-location (time, p), and one
    // item per susceptible stay = a run of table entries, numbered by its first (stay, evaluation) pair
    double* tab_t; double* tab_p; unsigned long long tab_cap;
    uint32_t* item_person; uint32_t* item_key; unsigned long long* item_tab; unsigned long long* item_pair;
    unsigned long long item_cap;
    unsigned long long* huge_off;  // byte offset into scratch of every entry of blk_seg[CLS_HUGE]
    uint32_t* run_seg[N_CLS];      // the queued sub-locations of classes 0 and 1 that need more than lastCalculation
                                   // advanced (k_group_filter)
    unsigned char* scratch; unsigned long long scratch_cap;
    uint32_t flags;
    Counters* ctr;
};

// Sub-locations of > 32 stays go to the group kernels, by the number of events of the sub-location and the bound on
// its evaluations: classes 0 - 3 keep <= 256 / 1024 / 4096 / 8192 events in shared memory (17 / 66 / 126 / 196 KB per
// group; groups of classes 0 - 2 fit on an SM side by side), class 4 works in global scratch.
// The lists are filled by the prefix-sum kernel (it sees every bucket size), so the group kernels can start as soon as
// the stays are in bucket order, side by side with the kernel that streams the small sub-locations.
constexpr int BLOCK_P_CAP = 2048;  // evaluations per segment the shared-memory group kernels keep probabilities for
constexpr int BLOCK_S_CAP = 4096;  // events per segment of class 2
constexpr int BLOCK_S_CAP3 = 8192; // ... and of class 3, the largest shared-memory class
struct BigLists
{
    uint32_t* blk_seg[N_CLS]; uint32_t blk_cap;
    unsigned long long* huge_off; unsigned long long scratch_cap;
    const uint8_t* key_total;  // sub-locations of total-branch locations are not queued here
    int need_enters, both;  // enter events are needed (distance model head counts / ENTER_AND_EXIT trigger)
    double r_window;        // (T1 - T0) / threshold + 2: no window holds more evaluations of one sub-location
};

#ifdef __CUDACC__
__device__ __forceinline__ void queue_big(const BigLists& b, Counters* ctr, uint32_t key, uint32_t n)
{
    const uint32_t slots = b.need_enters ? 2 * n : n;
    double r_bound = b.both ? 2.0 * n : (double) n;
    if (b.r_window < r_bound) r_bound = b.r_window;
    const bool r_fits = r_bound <= (double) BLOCK_P_CAP;
    const int cls = slots <= 256 ? 0 : (slots <= 1024 ? 1 : ((slots <= (uint32_t) BLOCK_S_CAP && r_fits) ? 2 :
                    ((slots <= (uint32_t) BLOCK_S_CAP3 && r_fits) ? 3 : CLS_HUGE)));
    const unsigned long long slot = atomicAdd(&ctr->n_blk[cls], 1ull);
    if (slot >= b.blk_cap) { atomicOr(&ctr->overflow, 2u); return; }
    b.blk_seg[cls][slot] = key;
    if (cls == CLS_HUGE)
    {
        unsigned long long mp = 64;
        while (mp < slots) mp <<= 1;
        const unsigned long long bytes = (unsigned long long) slots * 72 + 8192 + 20ull * (mp < 65536 ? mp : 65536);
        const unsigned long long off = atomicAdd(&ctr->scratch_top, bytes);
        if (off + bytes <= b.scratch_cap) b.huge_off[slot] = off;
        else { b.blk_seg[cls][slot] = KEY_NONE; atomicOr(&ctr->overflow, 2u); }
    }
    atomicAdd(&ctr->big_stays, (unsigned long long) n);
}
#endif

// the engine stream, four auxiliary streams for the kernels that run side by side, and the fork / join events
constexpr int N_AUX = 5;
struct TransmitStreams { cudaStream_t main; cudaStream_t aux[N_AUX]; cudaEvent_t fork, mid, filtered; cudaEvent_t join[N_AUX]; };
constexpr int DRAW_ITEM_SHIFT = 40;
constexpr unsigned long long DRAW_PAIR_MASK = (1ull << DRAW_ITEM_SHIFT) - 1ull;

}  // namespace heros
