// engine_state.cuh -- the engine object and the small shared definitions of the translation units that implement the
// C ABI (engine.cu: lifecycle, inputs, the window pipeline, outputs; exchange.cu: multi-GPU exchange; schedule.cu: the
// device-side activity schedule; monitor.cu: occupancy sampling).  transmit.cu only needs engine.cuh.
#pragma once

#include "engine.cuh"

#include <algorithm>
#include <limits>
#include <string>
#include <vector>

namespace heros {

constexpr int TPB = 256;
constexpr unsigned long long OPEN_CLOSED = 0x7ff8000000000000ull;  // t_in of an open-stay slot nobody is in (NaN)
inline int blocks_for(size_t n, int per = TPB) { return (int) std::max<size_t>(1, (n + per - 1) / per); }

// per-agent state and the transition log, as the kernels that change disease phases see them
struct AgentArrays
{
    uint64_t* view; double* next_time; double* ill_end;
    const ProgParams* prog; uint64_t seed;
    uint32_t person_base;  // global id of agent 0: every draw and every log entry uses global person ids
    Counters* ctr;
    uint32_t* tr_person; uint8_t* tr_phase; double* tr_time; unsigned long long tr_cap;
};

// the pool of submitted stays (rank = the ticket of the stay at its sub-location)
struct PoolArrays { uint32_t* person; uint32_t* key; double* tin; double* tout; uint32_t* rank; };

// stays of other shards' persons in this shard's locations, valid for the window in progress
struct GuestArrays
{
    const uint32_t* gid; const uint64_t* view; const uint32_t* key; const double* tin; const double* tout;
    const uint32_t* rank; uint32_t n;
};

// the streaming kernels handle IPT items per thread (their loads and atomics are in flight together)
constexpr int IPT = 4;
inline int item_blocks(size_t n) { return (int) ((n + (size_t) TPB * IPT - 1) / ((size_t) TPB * IPT)); }

// candidate infections of the window / the infection log
struct CandArrays { const uint32_t* person; const uint64_t* gkey; const double* t; const double* p; uint32_t cap; };
struct InfLog { uint32_t* person; uint64_t* gkey; double* t; double* p; unsigned long long cap; };

struct Activity { int32_t kind, locator, choice, dist; double mn, mode, mx, until; };  // = heros_activity

}  // namespace heros

using namespace heros;

// events of one window, for the per-stage device times (a small ring: the host may be several windows ahead of the device)
constexpr int WIN_RING = 8;
struct WindowEvents
{
    cudaEvent_t begin, opened, scanned, scattered, mid, transmitted, resolved;  // on the engine stream, in this order
    cudaEvent_t fork, filtered, join[N_AUX];                                     // the auxiliary streams of the stage
    bool pending = false;  // recorded, not yet added to the stage times
    int64_t window_no = -1; // which window of the run these events belong to
};

struct heros_engine
{
    heros_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;       // the engine stream: every window kernel is ordered on it
    cudaStream_t sub_stream = nullptr;   // k_submit of the next window runs here, beside the transmission of the current one
    cudaStream_t copy_stream = nullptr;  // host -> device copies of heros_submit_visits, beside everything else
    cudaStream_t aux[N_AUX]{};
    cudaEvent_t ev[2]{};                 // begin / end of a heros_step (or of the windows since the last heros_wait)
    cudaEvent_t ev_rolled = nullptr;     // the pools of the last window have been consumed (k_submit may write again)
    cudaEvent_t ev_submitted = nullptr;  // everything submitted so far is in the pool (the next window may begin)
    cudaEvent_t ev_stage_free[2]{};      // the staging buffers of heros_submit_visits have been read
    cudaEvent_t ev_copied = nullptr;
    bool sub_used = false, async_open = false;
    WindowEvents ring[WIN_RING];
    int ring_next = 0;
    double stage_ms[6]{};  // window open (state machine + tickets), prefix sum + scatter, streaming kernel, rest of the
                           // transmission stage, resolve, (unused)
    double bucket_part_ms[3]{};  // (unused), prefix sum, scatter
    double aux_done_ms[N_AUX]{}; // fork -> end of the work of every auxiliary stream
    int64_t windows_timed = 0;
    int n_sm = 148;
    double t_now = 0.0;

    bool have_area = false, have_dist = false, have_prog = false;
    AreaParams area{};
    DistParams dist{};
    std::vector<PolicyChange> policy;
    DevBuf<PolicyChange> d_policy;
    DevBuf<ProgParams> d_prog;

    int n_types = 0;
    std::vector<uint8_t> type_infect_sub;
    std::vector<double> type_corr;
    std::vector<uint8_t> type_cap_constrained;   // locationtypes.csv capConstrained / capPersonsPerM2 (schedule.cu)
    std::vector<double> type_cap_per_m2;

    uint32_t n_loc = 0, n_keys = 0;
    std::vector<uint32_t> h_loc_base;
    std::vector<uint8_t> h_loc_type;
    std::vector<int16_t> h_loc_nsub;
    std::vector<float> h_loc_area;
    DevBuf<uint32_t> d_loc_base, d_key_loc;
    DevBuf<int16_t> d_loc_nsub;
    DevBuf<uint2> d_loc_key;  // {first key, number of sub-locations} of every location, for k_submit
    DevBuf<LocParam> d_loc_param;
    DevBuf<double> d_last_calc;
    DevBuf<uint8_t> d_key_total;     // per key: KEYFLAG_TOTAL / KEYFLAG_RATE (allocated only if there is such a location)
    std::vector<uint8_t> h_key_flag;
    DevBuf<RateLoc> d_rate_loc;      // rate-based locations (heros_set_rate_locations)
    uint32_t n_rate_loc = 0;
    bool rate_factor_mode = false;   // some location follows the reference group: windows must not span a midnight
    DevBuf<TotalLoc> d_total_loc;
    uint32_t n_total_loc = 0;

    uint32_t n_agents = 0;
    DevBuf<uint64_t> d_view;
    DevBuf<double> d_next_time, d_ill_end, d_open_tin;
    DevBuf<unsigned long long> d_best_t, d_best_key;
    DevBuf<uint32_t> d_claim, d_open_key;
    uint32_t person_base = 0;   // global id of agent 0
    int32_t location_base = 0;  // global id of location 0

    // guests: stays of other shards' persons in this shard's locations, valid for the window in progress
    DevBuf<uint32_t> g_gid, g_key, g_rank;
    DevBuf<uint64_t> g_view;
    DevBuf<double> g_ill_end, g_tin, g_tout;
    uint32_t n_guest = 0;
    // peer-to-peer exchange region (heros_exchange_create / _connect): this shard's receive region and the peers'
    struct PeerExchange
    {
        unsigned char* region = nullptr; size_t bytes = 0; int rank = 0, n = 0; size_t gc = 0, cc = 0;
        unsigned char* peer[64] = {}; bool opened[64] = {}; bool connected = false; unsigned long long seq = 0;
        unsigned long long sent_to = 0;  // peers that got guests from this shard in the window in progress
        unsigned long long sources = ~0ull;  // peers that may send guests to this shard (heros_exchange_set_sources)
    } xc;
    int win_state = 0;          // 0 idle, 1 after heros_window_begin, 2 after heros_window_transmit
    double win_T0 = 0.0, win_T1 = 0.0;
    int win_launches = 0;
    int64_t launches_total = 0;  // kernels launched since the engine was created
    int win_ring = 0;            // ring slot of the window in progress
    Counters win_before{};
    int64_t launches_before = 0;

    // The pools of stays.  fresh: submitted since the last window (k_submit / k_advance_day append, taking every stay's
    // ticket on the way); carry[2]: stays that outlive a window move from one to the other.  Their fill counts live on
    // the DEVICE (Counters::n_fresh / n_carry); the host only keeps upper bounds, for allocation and grid sizes.
    DevBuf<uint32_t> fresh_person, fresh_key, fresh_rank, carry_person[2], carry_key[2], carry_rank[2];
    DevBuf<double> fresh_tin, fresh_tout, carry_tin[2], carry_tout[2];
    int carry_cur = 0;
    size_t fresh_ub = 0, carry_ub = 0;   // upper bounds of n_fresh / n_carry
    size_t fresh_seen = 0;               // largest fresh_ub of a window so far
    // The exact fill of the carry pool after every window travels to the host on its own stream (8 bytes, pinned);
    // whatever has arrived by the time the next window is enqueued tightens carry_ub -- without it a host that never
    // waits (heros_step_async back to back) would see the bound, and every buffer sized from it, grow window by window.
    cudaStream_t meta_stream = nullptr;
    unsigned long long* h_keep = nullptr;        // pinned, WIN_RING entries
    cudaEvent_t ev_keep[WIN_RING]{};
    bool keep_pending[WIN_RING]{};
    uint64_t keep_seq[WIN_RING]{}, win_seq = 0;
    uint64_t keep_cum[WIN_RING]{}, fresh_cum = 0;  // fresh stays consumed by the windows enqueued so far (running sum)

    // window buffers
    DevBuf<uint32_t> bk_count, bk_rank, seg_start;
    DevBuf<unsigned long long> scan_status;
    DevBuf<StayRec> rec, rec_part;
    DevBuf<unsigned long long> d_timeline;  // heros_debug_timeline
    int resolve_grid = 0;  // blocks of the cooperative k_resolve launch

    // candidates, logs
    DevBuf<uint32_t> cand_person;
    DevBuf<uint64_t> cand_gkey, inf_gkey;
    DevBuf<double> cand_t, cand_p;
    DevBuf<uint32_t> inf_person, tr_person;
    DevBuf<unsigned long long> d_nguest_cand;
    DevBuf<double> inf_t, inf_p, tr_time;
    DevBuf<uint8_t> tr_phase;
    DevBuf<long long> d_series;   // compartment counts at the end of the last SERIES_RING windows
    size_t series_fetched = 0;    // windows whose counts have been copied into series_counts

    // work lists of the transmission kernels + global scratch of huge segments
    DevBuf<uint32_t> sub_seg[2], hot_seg[N_HOT], item_key, item_cnt, item_nill, item_nsus, item_part, tab_head;
    DevBuf<PartTable> part_table;
    DevBuf<unsigned long long> item_tab;
    DevBuf<double> tab_now, tab_dur;
    DevBuf<unsigned long long> huge_off;
    DevBuf<unsigned char> scratch;

    // activity schedule (heros_set_schedule)
    bool have_schedule = false;
    unsigned int invalid_reported = 0;  // of Counters::invalid_visits, already returned as an error
    DevBuf<double> m_times; DevBuf<int32_t> m_sub, m_loc;  // heros_sample_occupancy
    DevBuf<double> s_divert;  // heros_set_capacity_diversion (nullptr: no *Cap locator sheds anything)
    DevBuf<double2> s_closure; DevBuf<uint8_t> s_loc_type; std::vector<double> h_closure; bool closure_active = false;
    int sched_max_acts = 0, sched_n_types = 0, sched_n_cand = 0, sched_accommodation = 0;
    uint64_t sched_seed = 0;
    size_t sched_room = 0;  // stays one simulated day can end, at most (heros_set_schedule)
    DevBuf<Activity> s_acts;
    DevBuf<int32_t> s_pat_start, s_pat_count, s_choice_start, s_choice_type, s_nearest, s_ncand, s_cand, d_home_loc,
        d_work_loc;
    DevBuf<double> s_choice_cum;
    DevBuf<int16_t> d_home_sub, d_work_sub;
    DevBuf<uint32_t> s_order;      // persons grouped by social role
    std::vector<uint8_t> h_role;   // host copy of the roles (heros_set_persons)
    std::vector<int32_t> h_home_loc, h_work_loc;  // host copies for the range checks of heros_set_schedule
    std::vector<int16_t> h_home_sub;

    DevBuf<Counters> d_ctr;
    Counters* h_ctr = nullptr;  // pinned
    void* pin_buf = nullptr; size_t pin_bytes = 0;  // pinned staging of the getters
    bool ctr_current = false;   // nothing that changes the counters was enqueued since h_ctr was refreshed

    // staging for host submissions (double buffered: the copy of the next batch runs beside the kernel that reads this one)
    DevBuf<int32_t> st_person[2], st_loc[2];
    DevBuf<int16_t> st_sub[2];
    DevBuf<double> st_tin[2], st_tout[2];
    int st_cur = 0;
    // a submission begun by heros_submit_visits_begin: the columns travel in ST_CHUNKS pieces, k_submit follows piece by piece
    static constexpr int ST_CHUNKS = 4;
    cudaEvent_t ev_chunk[ST_CHUNKS]{};
    int64_t st_pending_n = -1;   // -1: nothing begun
    int st_pending_k = 0;
    DevBuf<int32_t> x_person; DevBuf<double> x_time;  // heros_expose / heros_set_agent_state

    std::vector<double> series_t;
    std::vector<int64_t> series_counts;

    // sorted output caches
    struct InfEvent { double t, p; uint64_t gkey; uint32_t person; };
    std::vector<InfEvent> o_inf;
    struct SortKey { uint64_t t; uint32_t person, idx; };
    std::vector<SortKey> sort_keys, sort_tmp;  // scratch of the (time, person) ordering of newly fetched events
    struct TrEvent { double t; uint32_t person; uint8_t phase; };
    std::vector<TrEvent> o_tr;
    unsigned long long o_inf_n = ~0ull, o_tr_n = ~0ull;
};

namespace heros_impl {
using namespace heros;

#define CU(h, call)                                                                                       \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
            return HEROS_ERR_CUDA;                                                                        \
        }                                                                                                 \
    } while (0)

int32_t fail(heros_handle h, int32_t code, const std::string& msg);
heros::AgentArrays agent_arrays(heros_handle h);
heros::PoolArrays fresh_arrays(heros_handle h);
heros::PoolArrays carry_arrays(heros_handle h, int which);
int32_t ensure_fresh(heros_handle h, size_t n);  // room for n stays in the fresh pool (contents kept)
int32_t sync_counters(heros_handle h);     // D2H of the counters + synchronisation of every stream of the engine
int32_t current_counters(heros_handle h);  // ... only if something was enqueued since the last one
int32_t check_ready(heros_handle h);
int32_t join_submissions(heros_handle h);  // the engine stream waits for what was submitted on the side stream
int32_t check_idle(heros_handle h, const char* what);  // no window in progress

}  // namespace heros_impl
