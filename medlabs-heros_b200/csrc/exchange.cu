// exchange.cu -- multi-GPU exchange of a sharded run: stays of a shard's persons in other shards'
// locations go out with the packed agent word of the person, successful draws on those guests come back.  Two forms:
// variable-size rows (the host routes and negotiates the sizes) and fixed-capacity blocks (kernels pack and parse one
// block per peer, the two all-to-alls of a window are equal-split collectives on the engine stream).
#include "engine_state.cuh"

namespace heros {
namespace {

// ---- multi-GPU: stays of this shard's persons in other shards' locations, and the way back -----------------------------
// the packed agent words (and the end of illness, if inside the window) of the listed persons, after k_progress
__global__ void __launch_bounds__(TPB) k_export_words(uint32_t n, const int32_t* person, uint32_t person_base,
                                                      uint32_t n_agents, const uint64_t* view, const double* ill_end,
                                                      uint64_t* view_out, double* ill_end_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t local = (uint32_t) person[i] - person_base;
    uint64_t v = 0;
    double e = bits_f64(0x7ff0000000000000ull);
    if (local < n_agents)
    {
        v = view[local];
        if (view_ends(v)) e = ill_end[local];
    }
    view_out[i] = v;
    ill_end_out[i] = e;
}

__global__ void __launch_bounds__(TPB) k_guest_keys(uint32_t n, const int32_t* loc, const int16_t* sub, int32_t loc_base,
                                                    const uint32_t* loc_first_key, const int16_t* loc_nsub, uint32_t n_loc,
                                                    const double* tin, const double* tout, double T1,
                                                    uint32_t* key_out, uint32_t* count, uint32_t* rank_out,
                                                    Counters* ctr)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t l = loc[i] - loc_base;
    const int32_t sb = sub[i];
    bool ok = l >= 0 && (uint32_t) l < n_loc && sb >= 0 && tin[i] >= 0.0 && tout[i] > tin[i];
    if (ok) ok = sb < max((int) loc_nsub[l], 1);
    const uint32_t key = ok ? loc_first_key[l] + (uint32_t) sb : KEY_NONE;
    key_out[i] = key;
    // guests take their tickets at the sub-location when they are submitted (the window's own stays did in k_window_open)
    rank_out[i] = (ok && tin[i] < T1) ? atomicAdd(&count[key], 1u) : KEY_NONE;
    if (!ok) atomicAdd(&ctr->invalid_visits, 1u);
}

// candidates whose person lives on another shard, compacted for the exchange
__global__ void __launch_bounds__(TPB) k_export_guest_candidates(CandArrays C, const Counters* ctr, uint32_t person_base,
                                                                 uint32_t n_agents, uint32_t cap, uint32_t* gid,
                                                                 int32_t* loc, int16_t* sub, double* t, double* p,
                                                                 unsigned long long* n_out)
{
    const uint32_t n = (uint32_t) min((unsigned long long) C.cap, ctr->n_cand);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        if (C.person[i] - person_base < n_agents) continue;
        const unsigned long long slot = atomicAdd(n_out, 1ull);
        if (slot >= cap) continue;
        gid[slot] = C.person[i];
        loc[slot] = (int32_t) (C.gkey[i] >> 16);
        sub[slot] = (int16_t) (C.gkey[i] & 0xFFFFu);
        t[slot] = C.t[i];
        p[slot] = C.p[i];
    }
}

__global__ void __launch_bounds__(TPB) k_import_candidates(uint32_t n, const uint32_t* gid, const int32_t* loc,
                                                           const int16_t* sub, const double* t, const double* p,
                                                           uint32_t* cand_person, uint64_t* cand_gkey, double* cand_t,
                                                           double* cand_p, uint32_t cand_cap, Counters* ctr)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long slot = atomicAdd(&ctr->n_cand, 1ull);
    if (slot >= cand_cap) { atomicOr(&ctr->overflow, 1u); return; }
    cand_person[slot] = gid[i];
    cand_gkey[slot] = ((uint64_t) (uint32_t) loc[i] << 16) | (uint64_t) (uint16_t) sub[i];
    cand_t[slot] = t[i];
    cand_p[slot] = p[i];
}

// ---- multi-GPU, block exchange: fixed-capacity blocks, one per peer, so that a window's two all-to-alls are
//      equal-split collectives with nothing to negotiate and no host synchronisation in between --------------------------
// guest block   : u64 count | pad | t_in f64[C] | t_out f64[C] | word u64[C] | ill_end f64[C] | person i32[C] |
//                 loc i32[C] | sub i16[C]                                  (C a multiple of 4: every column 8-byte aligned)
// candidate block: u64 count | pad | time f64[C] | p f64[C] | person i32[C] | loc i32[C] | sub i16[C]
constexpr int MAX_PEERS = 64;
__host__ __device__ inline size_t guest_block_bytes(size_t C) { return 16 + 42 * C; }
__host__ __device__ inline size_t candidate_block_bytes(size_t C) { return 16 + 26 * C; }
struct PeerOffsets { long long v[MAX_PEERS + 1]; };
struct PeerBases { int v[MAX_PEERS + 1]; };

__global__ void __launch_bounds__(TPB) k_pack_guest_blocks(uint32_t n_rows, PeerOffsets off, int n_blocks,
                                                           const int32_t* person, const int32_t* loc, const int16_t* sub,
                                                           const double* tin, const double* tout, uint32_t person_base,
                                                           uint32_t n_agents, const uint64_t* view, const double* ill_end,
                                                           uint32_t C, unsigned char* out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t bb = guest_block_bytes(C);
    if (i < (uint32_t) n_blocks)
        *reinterpret_cast<unsigned long long*>(out + (size_t) i * bb) = (unsigned long long) (off.v[i + 1] - off.v[i]);
    if (i >= n_rows) return;
    int d = 0;
    while (d + 1 < n_blocks && (long long) i >= off.v[d + 1]) ++d;
    const uint32_t j = i - (uint32_t) off.v[d];
    if (j >= C) return;
    unsigned char* b = out + (size_t) d * bb + 16;
    const uint32_t local = (uint32_t) person[i] - person_base;
    uint64_t v = 0;
    double e = bits_f64(0x7ff0000000000000ull);
    if (local < n_agents)
    {
        v = view[local];
        if (view_ends(v)) e = ill_end[local];
    }
    reinterpret_cast<double*>(b)[j] = tin[i];
    reinterpret_cast<double*>(b + 8 * (size_t) C)[j] = tout[i];
    reinterpret_cast<uint64_t*>(b + 16 * (size_t) C)[j] = v;
    reinterpret_cast<double*>(b + 24 * (size_t) C)[j] = e;
    reinterpret_cast<int32_t*>(b + 32 * (size_t) C)[j] = person[i];
    reinterpret_cast<int32_t*>(b + 36 * (size_t) C)[j] = loc[i];
    reinterpret_cast<int16_t*>(b + 40 * (size_t) C)[j] = sub[i];
}

__global__ void __launch_bounds__(TPB) k_submit_guest_blocks(int n_blocks, uint32_t C, const unsigned char* blocks,
                                                             double T1, int32_t loc_base, const uint32_t* loc_first_key,
                                                             const int16_t* loc_nsub, uint32_t n_loc, uint32_t* gid,
                                                             uint64_t* gview, double* gill_end, uint32_t* gkey,
                                                             double* gtin, double* gtout, uint32_t* grank, uint32_t* count,
                                                             Counters* ctr)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint32_t) n_blocks * C) return;
    const uint32_t bk = idx / C, j = idx % C;
    const size_t bb = guest_block_bytes(C);
    const unsigned char* b = blocks + (size_t) bk * bb;
    const unsigned long long cnt = *reinterpret_cast<const unsigned long long*>(b);
    if (j == 0 && cnt > C) atomicOr(&ctr->overflow, 16u);
    uint32_t key = KEY_NONE, rank = KEY_NONE;
    if (j < cnt)
    {
        b += 16;
        const double tin = reinterpret_cast<const double*>(b)[j], tout = reinterpret_cast<const double*>(b + 8 * (size_t) C)[j];
        const int32_t l = reinterpret_cast<const int32_t*>(b + 36 * (size_t) C)[j] - loc_base;
        const int32_t sb = reinterpret_cast<const int16_t*>(b + 40 * (size_t) C)[j];
        bool ok = l >= 0 && (uint32_t) l < n_loc && sb >= 0 && tin >= 0.0 && tout > tin;
        if (ok) ok = sb < max((int) loc_nsub[l], 1);
        if (ok)
        {
            key = loc_first_key[l] + (uint32_t) sb;
            if (tin < T1) rank = atomicAdd(&count[key], 1u);
        }
        else
            atomicAdd(&ctr->invalid_visits, 1u);
        gid[idx] = (uint32_t) reinterpret_cast<const int32_t*>(b + 32 * (size_t) C)[j];
        gview[idx] = reinterpret_cast<const uint64_t*>(b + 16 * (size_t) C)[j];
        gill_end[idx] = reinterpret_cast<const double*>(b + 24 * (size_t) C)[j];
        gtin[idx] = tin;
        gtout[idx] = tout;
    }
    gkey[idx] = key;
    grank[idx] = rank;
}

__global__ void k_zero_block_headers(int n_blocks, size_t block_bytes, unsigned char* out)
{
    if ((int) threadIdx.x < n_blocks) *reinterpret_cast<unsigned long long*>(out + threadIdx.x * block_bytes) = 0ull;
}

// successful draws on guests, routed to the block of the shard that owns the person
__global__ void __launch_bounds__(TPB) k_export_candidate_blocks(CandArrays Cd, Counters* ctr, PeerBases bases, int n_blocks,
                                                                 uint32_t person_base, uint32_t n_agents, uint32_t C,
                                                                 unsigned char* out)
{
    const uint32_t n = (uint32_t) min((unsigned long long) Cd.cap, ctr->n_cand);
    const size_t bb = candidate_block_bytes(C);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t gid = Cd.person[i];
        if (gid - person_base < n_agents) continue;
        int d = 0;
        while (d + 1 < n_blocks && (int) gid >= bases.v[d + 1]) ++d;
        unsigned char* b = out + (size_t) d * bb;
        const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(b), 1ull);
        if (slot >= C) { atomicOr(&ctr->overflow, 16u); continue; }
        b += 16;
        reinterpret_cast<double*>(b)[slot] = Cd.t[i];
        reinterpret_cast<double*>(b + 8 * (size_t) C)[slot] = Cd.p[i];
        reinterpret_cast<int32_t*>(b + 16 * (size_t) C)[slot] = (int32_t) gid;
        reinterpret_cast<int32_t*>(b + 20 * (size_t) C)[slot] = (int32_t) (Cd.gkey[i] >> 16);
        reinterpret_cast<int16_t*>(b + 24 * (size_t) C)[slot] = (int16_t) (Cd.gkey[i] & 0xFFFFu);
    }
}

__global__ void __launch_bounds__(TPB) k_import_candidate_blocks(int n_blocks, uint32_t C, const unsigned char* blocks,
                                                                 uint32_t* cand_person, uint64_t* cand_gkey, double* cand_t,
                                                                 double* cand_p, uint32_t cand_cap, Counters* ctr)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint32_t) n_blocks * C) return;
    const uint32_t bk = idx / C, j = idx % C;
    const unsigned char* b = blocks + (size_t) bk * candidate_block_bytes(C);
    const unsigned long long cnt = *reinterpret_cast<const unsigned long long*>(b);
    if (j >= cnt) return;
    b += 16;
    const unsigned long long slot = atomicAdd(&ctr->n_cand, 1ull);
    if (slot >= cand_cap) { atomicOr(&ctr->overflow, 1u); return; }
    cand_person[slot] = (uint32_t) reinterpret_cast<const int32_t*>(b + 16 * (size_t) C)[j];
    cand_gkey[slot] = ((uint64_t) (uint32_t) reinterpret_cast<const int32_t*>(b + 20 * (size_t) C)[j] << 16) |
                      (uint64_t) (uint16_t) reinterpret_cast<const int16_t*>(b + 24 * (size_t) C)[j];
    cand_t[slot] = reinterpret_cast<const double*>(b)[j];
    cand_p[slot] = reinterpret_cast<const double*>(b + 8 * (size_t) C)[j];
}

// ---- the same blocks written STRAIGHT INTO THE PEERS' MEMORY over NVLink (peer pointers from CUDA IPC): the pack /
//      export kernel stores every row in the receive region of the shard it is meant for, the block that finishes
//      last publishes the counts and raises this shard's flag at every peer, and the receiver waits for the flags of
//      all its peers with a one-block kernel before it parses its region.  No collective call, no host in the loop.
//      Region of a shard: guest flags u64[64] | candidate flags u64[64] | local counters u64[64] | done counter |
//      (pad to 4096) | guest blocks [2][n_peers] | candidate blocks [n_peers]  (block s is written by shard s).
//      A flag carries the window number, so nothing is ever reset.  Guests: every shard waits for the flags of ALL
//      peers (anybody may send guests), and the blocks alternate between two buffers with the window parity: a peer
//      can be at most one window ahead (its next-but-one window needs this shard's flag of the next one, raised after
//      this window was parsed).  Candidates can only come back from the shards this one sent guests to, so only their
//      flags are awaited -- a shard does not wait for the slowest shard of the node, only for its neighbours. ----
constexpr size_t REGION_HEADER = 4096;
struct PeerPtrs { unsigned char* p[MAX_PEERS]; };
struct PeerFlags { unsigned long long* p[MAX_PEERS]; };

__device__ __forceinline__ void publish_when_last(unsigned int* done, int n_peers, const PeerFlags& flags, unsigned long long seq)
{
    __threadfence_system();  // this thread's rows are visible to the peers ...
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned int prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1)  // ... and so are everybody's by the time the last block gets here
        {
            *done = 0u;
            __threadfence_system();
            for (int d = 0; d < n_peers; ++d) *reinterpret_cast<volatile unsigned long long*>(flags.p[d]) = seq;
        }
    }
}

__global__ void __launch_bounds__(TPB) k_p2p_pack_guests(uint32_t n_rows, PeerOffsets off, int n_peers, const int32_t* person,
                                                         const int32_t* loc, const int16_t* sub, const double* tin,
                                                         const double* tout, uint32_t person_base, uint32_t n_agents,
                                                         const uint64_t* view, const double* ill_end, uint32_t C,
                                                         PeerPtrs dst, PeerFlags flags, unsigned long long seq,
                                                         unsigned int* done)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (uint32_t) n_peers)
        *reinterpret_cast<unsigned long long*>(dst.p[i]) = (unsigned long long) (off.v[i + 1] - off.v[i]);
    if (i < n_rows)
    {
        int d = 0;
        while (d + 1 < n_peers && (long long) i >= off.v[d + 1]) ++d;
        const uint32_t j = i - (uint32_t) off.v[d];
        if (j < C)
        {
            unsigned char* b = dst.p[d] + 16;
            const uint32_t local = (uint32_t) person[i] - person_base;
            uint64_t v = 0;
            double e = bits_f64(0x7ff0000000000000ull);
            if (local < n_agents)
            {
                v = view[local];
                if (view_ends(v)) e = ill_end[local];
            }
            reinterpret_cast<double*>(b)[j] = tin[i];
            reinterpret_cast<double*>(b + 8 * (size_t) C)[j] = tout[i];
            reinterpret_cast<uint64_t*>(b + 16 * (size_t) C)[j] = v;
            reinterpret_cast<double*>(b + 24 * (size_t) C)[j] = e;
            reinterpret_cast<int32_t*>(b + 32 * (size_t) C)[j] = person[i];
            reinterpret_cast<int32_t*>(b + 36 * (size_t) C)[j] = loc[i];
            reinterpret_cast<int16_t*>(b + 40 * (size_t) C)[j] = sub[i];
        }
    }
    publish_when_last(done, n_peers, flags, seq);
}

__global__ void __launch_bounds__(TPB) k_p2p_export_candidates(CandArrays Cd, Counters* ctr, PeerBases bases, int n_peers,
                                                               uint32_t person_base, uint32_t n_agents, uint32_t C,
                                                               PeerPtrs dst, PeerFlags flags, unsigned long long seq,
                                                               unsigned long long* local_cnt, unsigned int* done)
{
    const uint32_t n = (uint32_t) min((unsigned long long) Cd.cap, ctr->n_cand);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t gid = Cd.person[i];
        if (gid - person_base < n_agents) continue;
        int d = 0;
        while (d + 1 < n_peers && (int) gid >= bases.v[d + 1]) ++d;
        const unsigned long long slot = atomicAdd(&local_cnt[d], 1ull);
        if (slot >= C) { atomicOr(&ctr->overflow, 16u); continue; }
        unsigned char* b = dst.p[d] + 16;
        reinterpret_cast<double*>(b)[slot] = Cd.t[i];
        reinterpret_cast<double*>(b + 8 * (size_t) C)[slot] = Cd.p[i];
        reinterpret_cast<int32_t*>(b + 16 * (size_t) C)[slot] = (int32_t) gid;
        reinterpret_cast<int32_t*>(b + 20 * (size_t) C)[slot] = (int32_t) (Cd.gkey[i] >> 16);
        reinterpret_cast<int16_t*>(b + 24 * (size_t) C)[slot] = (int16_t) (Cd.gkey[i] & 0xFFFFu);
    }
    // the last block writes the counts into the headers at the peers, then raises the flags
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if ((int) threadIdx.x < n_peers)
    {
        const unsigned long long cnt = *reinterpret_cast<volatile unsigned long long*>(&local_cnt[threadIdx.x]);
        *reinterpret_cast<unsigned long long*>(dst.p[threadIdx.x]) = cnt < C ? cnt : C;
        local_cnt[threadIdx.x] = 0ull;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        *done = 0u;
        for (int d = 0; d < n_peers; ++d) *reinterpret_cast<volatile unsigned long long*>(flags.p[d]) = seq;
    }
}

// one block waits until every peer has raised its flag for this window (2^33 cycles ~ 4 s, then the window fails)
__global__ void k_p2p_wait(const unsigned long long* my_flags, int n_peers, unsigned long long mask, unsigned long long seq,
                           Counters* ctr)
{
    if ((int) threadIdx.x >= n_peers || !((mask >> threadIdx.x) & 1ull)) return;
    const volatile unsigned long long* f = my_flags + threadIdx.x;
    const long long t0 = clock64();
    while (*f < seq)
    {
        if (clock64() - t0 > (1ll << 33)) { atomicOr(&ctr->overflow, 64u); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace
}  // namespace heros

using namespace heros;
using namespace heros_impl;

extern "C" {

int32_t heros_export_agent_words(heros_handle h, int64_t n, const int32_t* person, uint64_t* view_out, double* ill_end_out)
{
    if (!h || n < 0 || (n > 0 && (!person || !view_out || !ill_end_out))) return HEROS_ERR_INVALID;
    if (h->win_state != 1) return fail(h, HEROS_ERR_INVALID, "agent words are exported between heros_window_begin and heros_window_transmit");
    if (n == 0) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    k_export_words<<<blocks_for(n), TPB, 0, h->stream>>>((uint32_t) n, person, h->person_base, h->n_agents, h->d_view.p,
                                                         h->d_ill_end.p, view_out, ill_end_out);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int32_t heros_submit_guest_visits_device(heros_handle h, int64_t n, const int32_t* person, const uint64_t* view,
                                         const double* ill_end, const int32_t* location, const int16_t* sublocation,
                                         const double* t_in, const double* t_out)
{
    if (!h || n < 0 || (n > 0 && (!person || !view || !ill_end || !location || !sublocation || !t_in || !t_out)))
        return HEROS_ERR_INVALID;
    if (h->win_state != 1) return fail(h, HEROS_ERR_INVALID, "guest stays are submitted between heros_window_begin and heros_window_transmit");
    if (n == 0) return HEROS_OK;
    if ((uint64_t) h->n_guest + (uint64_t) n >= 0x7fffffffull) return fail(h, HEROS_ERR_CAPACITY, "too many guest stays");
    cudaSetDevice(h->cfg.device);
    const size_t tot = (size_t) h->n_guest + (size_t) n, off = h->n_guest;
    CU(h, h->g_gid.ensure(tot, true, h->stream)); CU(h, h->g_key.ensure(tot, true, h->stream));
    CU(h, h->g_rank.ensure(tot, true, h->stream));
    CU(h, h->g_view.ensure(tot, true, h->stream)); CU(h, h->g_ill_end.ensure(tot, true, h->stream));
    CU(h, h->g_tin.ensure(tot, true, h->stream)); CU(h, h->g_tout.ensure(tot, true, h->stream));
    cudaStream_t s = h->stream;
    CU(h, cudaMemcpyAsync(h->g_gid.p + off, person, (size_t) n * 4, cudaMemcpyDeviceToDevice, s));
    CU(h, cudaMemcpyAsync(h->g_view.p + off, view, (size_t) n * 8, cudaMemcpyDeviceToDevice, s));
    CU(h, cudaMemcpyAsync(h->g_ill_end.p + off, ill_end, (size_t) n * 8, cudaMemcpyDeviceToDevice, s));
    CU(h, cudaMemcpyAsync(h->g_tin.p + off, t_in, (size_t) n * 8, cudaMemcpyDeviceToDevice, s));
    CU(h, cudaMemcpyAsync(h->g_tout.p + off, t_out, (size_t) n * 8, cudaMemcpyDeviceToDevice, s));
    k_guest_keys<<<blocks_for(n), TPB, 0, s>>>((uint32_t) n, location, sublocation, h->location_base, h->d_loc_base.p,
                                               h->d_loc_nsub.p, h->n_loc, t_in, t_out, h->win_T1, h->g_key.p + off,
                                               h->bk_count.p, h->g_rank.p + off, h->d_ctr.p);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(s));
    h->n_guest += (uint32_t) n;
    return HEROS_OK;
}

int32_t heros_export_guest_candidates(heros_handle h, int64_t capacity, int32_t* person, int32_t* location,
                                      int16_t* sublocation, double* time, double* p, int64_t* n_out)
{
    if (!h || !n_out || capacity < 0) return HEROS_ERR_INVALID;
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "candidates are exported between heros_window_transmit and heros_window_finish");
    cudaSetDevice(h->cfg.device);
    CU(h, h->d_nguest_cand.ensure(1));
    CU(h, cudaMemsetAsync(h->d_nguest_cand.p, 0, 8, h->stream));
    const CandArrays C{h->cand_person.p, h->cand_gkey.p, h->cand_t.p, h->cand_p.p, (uint32_t) h->cand_person.n};
    k_export_guest_candidates<<<h->n_sm * 2, TPB, 0, h->stream>>>(C, h->d_ctr.p, h->person_base, h->n_agents,
                                                                 (uint32_t) std::min<int64_t>(capacity, 0x7fffffff),
                                                                 (uint32_t*) person, location, sublocation, time, p,
                                                                 h->d_nguest_cand.p);
    CU(h, cudaGetLastError());
    unsigned long long n = 0;
    CU(h, cudaMemcpyAsync(&n, h->d_nguest_cand.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    *n_out = (int64_t) n;
    return (int64_t) n > capacity ? HEROS_ERR_CAPACITY : HEROS_OK;
}

int32_t heros_import_candidates_device(heros_handle h, int64_t n, const int32_t* person, const int32_t* location,
                                       const int16_t* sublocation, const double* time, const double* p)
{
    if (!h || n < 0 || (n > 0 && (!person || !location || !sublocation || !time || !p))) return HEROS_ERR_INVALID;
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "candidates are imported between heros_window_transmit and heros_window_finish");
    if (n == 0) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    k_import_candidates<<<blocks_for(n), TPB, 0, h->stream>>>((uint32_t) n, (const uint32_t*) person, location, sublocation,
                                                              time, p, h->cand_person.p, h->cand_gkey.p, h->cand_t.p,
                                                              h->cand_p.p, (uint32_t) h->cand_person.n, h->d_ctr.p);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int64_t heros_guest_block_bytes(int64_t capacity) { return capacity > 0 ? (int64_t) guest_block_bytes((size_t) capacity) : 0; }
int64_t heros_candidate_block_bytes(int64_t capacity) { return capacity > 0 ? (int64_t) candidate_block_bytes((size_t) capacity) : 0; }

static int32_t check_blocks(heros_handle h, int32_t n_blocks, int64_t capacity, const void* blocks)
{
    if (!h || !blocks) return HEROS_ERR_INVALID;
    if (n_blocks < 1 || n_blocks > MAX_PEERS) return fail(h, HEROS_ERR_INVALID, "between 1 and 64 peer blocks");
    if (capacity < 4 || (capacity & 3) || capacity > (1 << 28)) return fail(h, HEROS_ERR_INVALID, "block capacity must be a multiple of 4");
    return HEROS_OK;
}

int32_t heros_pack_guest_blocks_device(heros_handle h, int32_t n_blocks, const int64_t* row_offsets, const int32_t* person,
                                       const int32_t* location, const int16_t* sublocation, const double* t_in,
                                       const double* t_out, int64_t capacity, void* blocks_out)
{
    int32_t rc = check_blocks(h, n_blocks, capacity, blocks_out);
    if (rc != HEROS_OK) return rc;
    if (!row_offsets) return HEROS_ERR_INVALID;
    if (h->win_state != 1) return fail(h, HEROS_ERR_INVALID, "guest blocks are packed between heros_window_begin and heros_window_transmit");
    PeerOffsets off{};
    for (int d = 0; d <= n_blocks; ++d) off.v[d] = row_offsets[d];
    for (int d = 0; d < n_blocks; ++d)
    {
        if (off.v[d + 1] < off.v[d] || off.v[0] != 0) return fail(h, HEROS_ERR_INVALID, "row offsets must ascend from 0");
        if (off.v[d + 1] - off.v[d] > capacity) return fail(h, HEROS_ERR_CAPACITY, "more rows for a peer than the block capacity");
    }
    const int64_t n = off.v[n_blocks];
    if (n > 0 && (!person || !location || !sublocation || !t_in || !t_out)) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    k_pack_guest_blocks<<<blocks_for((size_t) std::max<int64_t>(n, n_blocks)), TPB, 0, h->stream>>>(
        (uint32_t) n, off, n_blocks, person, location, sublocation, t_in, t_out, h->person_base, h->n_agents, h->d_view.p,
        h->d_ill_end.p, (uint32_t) capacity, (unsigned char*) blocks_out);
    CU(h, cudaGetLastError());
    return HEROS_OK;
}

static int32_t submit_guest_blocks_append(heros_handle h, int32_t n_blocks, int64_t capacity, const void* blocks)
{
    int32_t rc = check_blocks(h, n_blocks, capacity, blocks);
    if (rc != HEROS_OK) return rc;
    if (h->win_state != 1) return fail(h, HEROS_ERR_INVALID, "guest stays are submitted between heros_window_begin and heros_window_transmit");
    cudaSetDevice(h->cfg.device);
    const size_t off = h->n_guest, add = (size_t) n_blocks * (size_t) capacity, tot = off + add;
    if (tot >= 0x7fffffffull) return fail(h, HEROS_ERR_CAPACITY, "too many guest stays");
    CU(h, h->g_gid.ensure(tot, true, h->stream)); CU(h, h->g_key.ensure(tot, true, h->stream));
    CU(h, h->g_rank.ensure(tot, true, h->stream)); CU(h, h->g_view.ensure(tot, true, h->stream));
    CU(h, h->g_ill_end.ensure(tot, true, h->stream)); CU(h, h->g_tin.ensure(tot, true, h->stream));
    CU(h, h->g_tout.ensure(tot, true, h->stream));
    k_submit_guest_blocks<<<blocks_for(add), TPB, 0, h->stream>>>(
        n_blocks, (uint32_t) capacity, (const unsigned char*) blocks, h->win_T1, h->location_base, h->d_loc_base.p,
        h->d_loc_nsub.p, h->n_loc, h->g_gid.p + off, h->g_view.p + off, h->g_ill_end.p + off, h->g_key.p + off,
        h->g_tin.p + off, h->g_tout.p + off, h->g_rank.p + off, h->bk_count.p, h->d_ctr.p);
    CU(h, cudaGetLastError());
    h->n_guest = (uint32_t) tot;
    return HEROS_OK;
}

int32_t heros_submit_guest_blocks_device(heros_handle h, int32_t n_blocks, int64_t capacity, const void* blocks)
{
    if (h && h->n_guest != 0) return fail(h, HEROS_ERR_INVALID, "guest blocks cannot be combined with other guest submissions of the window");
    return submit_guest_blocks_append(h, n_blocks, capacity, blocks);
}

int32_t heros_export_candidate_blocks_device(heros_handle h, int32_t n_blocks, const int32_t* person_bases, int64_t capacity,
                                             void* blocks_out)
{
    int32_t rc = check_blocks(h, n_blocks, capacity, blocks_out);
    if (rc != HEROS_OK) return rc;
    if (!person_bases) return HEROS_ERR_INVALID;
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "candidates are exported between heros_window_transmit and heros_window_finish");
    PeerBases bases{};
    for (int d = 0; d <= n_blocks; ++d) bases.v[d] = person_bases[d];
    cudaSetDevice(h->cfg.device);
    k_zero_block_headers<<<1, MAX_PEERS, 0, h->stream>>>(n_blocks, candidate_block_bytes((size_t) capacity), (unsigned char*) blocks_out);
    const CandArrays C{h->cand_person.p, h->cand_gkey.p, h->cand_t.p, h->cand_p.p, (uint32_t) h->cand_person.n};
    k_export_candidate_blocks<<<h->n_sm * 2, TPB, 0, h->stream>>>(C, h->d_ctr.p, bases, n_blocks, h->person_base, h->n_agents,
                                                                 (uint32_t) capacity, (unsigned char*) blocks_out);
    CU(h, cudaGetLastError());
    return HEROS_OK;
}

int32_t heros_import_candidate_blocks_device(heros_handle h, int32_t n_blocks, int64_t capacity, const void* blocks)
{
    int32_t rc = check_blocks(h, n_blocks, capacity, blocks);
    if (rc != HEROS_OK) return rc;
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "candidates are imported between heros_window_transmit and heros_window_finish");
    cudaSetDevice(h->cfg.device);
    k_import_candidate_blocks<<<blocks_for((size_t) n_blocks * (size_t) capacity), TPB, 0, h->stream>>>(
        n_blocks, (uint32_t) capacity, (const unsigned char*) blocks, h->cand_person.p, h->cand_gkey.p, h->cand_t.p,
        h->cand_p.p, (uint32_t) h->cand_person.n, h->d_ctr.p);
    CU(h, cudaGetLastError());
    return HEROS_OK;
}

// ---- peer-to-peer exchange over NVLink (CUDA IPC) ----------------------------------------------------------------------
static int32_t heros_import_candidate_blocks_from(heros_handle h, unsigned long long mask);
static int32_t submit_guest_blocks_from(heros_handle h, unsigned long long mask);

static size_t region_bytes(int n, size_t gc, size_t cc) { return REGION_HEADER + (size_t) n * (2 * guest_block_bytes(gc) + candidate_block_bytes(cc)); }

int32_t heros_exchange_create(heros_handle h, int32_t rank, int32_t n_peers, int64_t guest_capacity, int64_t candidate_capacity,
                              void* ipc_handle_out, void** region_out)
{
    if (!h || !ipc_handle_out) return HEROS_ERR_INVALID;
    if (n_peers < 1 || n_peers > MAX_PEERS || rank < 0 || rank >= n_peers) return fail(h, HEROS_ERR_INVALID, "bad rank / number of peers");
    if (guest_capacity < 4 || (guest_capacity & 3) || candidate_capacity < 4 || (candidate_capacity & 3))
        return fail(h, HEROS_ERR_INVALID, "block capacities must be multiples of 4");
    if (h->xc.region) return fail(h, HEROS_ERR_INVALID, "the exchange region exists already");
    cudaSetDevice(h->cfg.device);
    h->xc.rank = rank; h->xc.n = n_peers; h->xc.gc = (size_t) guest_capacity; h->xc.cc = (size_t) candidate_capacity;
    h->xc.bytes = region_bytes(n_peers, h->xc.gc, h->xc.cc);
    CU(h, cudaMalloc(&h->xc.region, h->xc.bytes));
    CU(h, cudaMemset(h->xc.region, 0, h->xc.bytes));
    cudaIpcMemHandle_t ipc;
    CU(h, cudaIpcGetMemHandle(&ipc, h->xc.region));
    static_assert(sizeof(ipc) == 64, "CUDA IPC handle size");
    memcpy(ipc_handle_out, &ipc, sizeof(ipc));
    if (region_out) *region_out = h->xc.region;
    return HEROS_OK;
}

int32_t heros_exchange_connect(heros_handle h, const void* ipc_handles, void* const* local_regions)
{
    if (!h || (!ipc_handles && !local_regions)) return HEROS_ERR_INVALID;
    if (!h->xc.region) return fail(h, HEROS_ERR_INVALID, "heros_exchange_create has not been called");
    cudaSetDevice(h->cfg.device);
    for (int d = 0; d < h->xc.n; ++d)
    {
        if (d == h->xc.rank) { h->xc.peer[d] = h->xc.region; continue; }
        if (local_regions && local_regions[d]) { h->xc.peer[d] = (unsigned char*) local_regions[d]; continue; }
        if (!ipc_handles) return fail(h, HEROS_ERR_INVALID, "no handle for a peer");
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, (const unsigned char*) ipc_handles + (size_t) d * sizeof(ipc), sizeof(ipc));
        void* ptr = nullptr;
        CU(h, cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
        h->xc.peer[d] = (unsigned char*) ptr;
        h->xc.opened[d] = true;
    }
    h->xc.connected = true;
    return HEROS_OK;
}

static void peer_targets(heros_handle h, bool candidates, PeerPtrs& dst, PeerFlags& flags)
{
    const size_t gb = guest_block_bytes(h->xc.gc), cb = candidate_block_bytes(h->xc.cc);
    const size_t parity = (size_t) (h->xc.seq & 1ull);  // guest blocks alternate between two buffers
    for (int d = 0; d < h->xc.n; ++d)
    {
        unsigned char* base = h->xc.peer[d];
        dst.p[d] = candidates ? base + REGION_HEADER + 2 * (size_t) h->xc.n * gb + (size_t) h->xc.rank * cb
                              : base + REGION_HEADER + (parity * (size_t) h->xc.n + (size_t) h->xc.rank) * gb;
        flags.p[d] = reinterpret_cast<unsigned long long*>(base) + (candidates ? 64 : 0) + h->xc.rank;
    }
}

int32_t heros_exchange_set_sources(heros_handle h, uint64_t source_mask)
{
    if (!h) return HEROS_ERR_INVALID;
    if (!h->xc.region) return fail(h, HEROS_ERR_INVALID, "heros_exchange_create has not been called");
    h->xc.sources = source_mask | (1ull << h->xc.rank);
    return HEROS_OK;
}

int32_t heros_window_run_peers(heros_handle h, double t_until, const int64_t* row_offsets, const int32_t* person,
                               const int32_t* location, const int16_t* sublocation, const double* t_in, const double* t_out,
                               const int32_t* person_bases, heros_step_stats* stats)
{
    int32_t rc;
    if ((rc = heros_window_begin(h, t_until)) != HEROS_OK) return rc;
    if ((rc = heros_window_send_guests(h, row_offsets, person, location, sublocation, t_in, t_out)) != HEROS_OK) return rc;
    if ((rc = heros_window_recv_guests(h)) != HEROS_OK) return rc;
    if ((rc = heros_window_transmit(h)) != HEROS_OK) return rc;
    if ((rc = heros_window_send_candidates(h, person_bases)) != HEROS_OK) return rc;
    if ((rc = heros_window_recv_candidates(h)) != HEROS_OK) return rc;
    return heros_window_finish(h, stats);
}

int32_t heros_window_run_peers_async(heros_handle h, double t_until, const int64_t* row_offsets, const int32_t* person,
                                     const int32_t* location, const int16_t* sublocation, const double* t_in, const double* t_out,
                                     const int32_t* person_bases)
{
    int32_t rc;
    if ((rc = heros_window_begin(h, t_until)) != HEROS_OK) return rc;
    if ((rc = heros_window_send_guests(h, row_offsets, person, location, sublocation, t_in, t_out)) != HEROS_OK) return rc;
    if ((rc = heros_window_recv_guests(h)) != HEROS_OK) return rc;
    if ((rc = heros_window_transmit(h)) != HEROS_OK) return rc;
    if ((rc = heros_window_send_candidates(h, person_bases)) != HEROS_OK) return rc;
    if ((rc = heros_window_recv_candidates(h)) != HEROS_OK) return rc;
    return heros_window_finish_async(h);
}

int32_t heros_window_send_guests(heros_handle h, const int64_t* row_offsets, const int32_t* person, const int32_t* location,
                                 const int16_t* sublocation, const double* t_in, const double* t_out)
{
    if (!h || !row_offsets) return HEROS_ERR_INVALID;
    if (!h->xc.connected) return fail(h, HEROS_ERR_INVALID, "heros_exchange_connect has not been called");
    if (h->win_state != 1) return fail(h, HEROS_ERR_INVALID, "guest stays are sent between heros_window_begin and heros_window_transmit");
    const int n = h->xc.n;
    PeerOffsets off{};
    for (int d = 0; d <= n; ++d) off.v[d] = row_offsets[d];
    for (int d = 0; d < n; ++d)
    {
        if (off.v[d + 1] < off.v[d] || off.v[0] != 0) return fail(h, HEROS_ERR_INVALID, "row offsets must ascend from 0");
        if ((size_t) (off.v[d + 1] - off.v[d]) > h->xc.gc) return fail(h, HEROS_ERR_CAPACITY, "more rows for a peer than the block capacity");
    }
    const int64_t rows = off.v[n];
    if (rows > 0 && (!person || !location || !sublocation || !t_in || !t_out)) return HEROS_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    PeerPtrs dst{};
    PeerFlags flags{};
    ++h->xc.seq;
    peer_targets(h, false, dst, flags);
    h->xc.sent_to = 0;  // candidates can only come back from the shards that get guests from this one
    for (int d = 0; d < n; ++d)
        if (off.v[d + 1] > off.v[d] && d != h->xc.rank) h->xc.sent_to |= 1ull << d;
    unsigned int* done = reinterpret_cast<unsigned int*>(h->xc.region + 1536);
    k_p2p_pack_guests<<<blocks_for((size_t) std::max<int64_t>(rows, n)), TPB, 0, h->stream>>>(
        (uint32_t) rows, off, n, person, location, sublocation, t_in, t_out, h->person_base, h->n_agents, h->d_view.p,
        h->d_ill_end.p, (uint32_t) h->xc.gc, dst, flags, h->xc.seq, done);
    CU(h, cudaGetLastError());
    return HEROS_OK;
}

int32_t heros_window_recv_guests(heros_handle h)
{
    if (!h) return HEROS_ERR_INVALID;
    if (!h->xc.connected) return fail(h, HEROS_ERR_INVALID, "heros_exchange_connect has not been called");
    if (h->win_state != 1) return fail(h, HEROS_ERR_INVALID, "guest stays are received between heros_window_begin and heros_window_transmit");
    cudaSetDevice(h->cfg.device);
    k_p2p_wait<<<1, MAX_PEERS, 0, h->stream>>>(reinterpret_cast<const unsigned long long*>(h->xc.region), h->xc.n,
                                               h->xc.sources, h->xc.seq, h->d_ctr.p);
    CU(h, cudaGetLastError());
    return submit_guest_blocks_from(h, h->xc.sources);
}

int32_t heros_window_send_candidates(heros_handle h, const int32_t* person_bases)
{
    if (!h || !person_bases) return HEROS_ERR_INVALID;
    if (!h->xc.connected) return fail(h, HEROS_ERR_INVALID, "heros_exchange_connect has not been called");
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "candidates are sent between heros_window_transmit and heros_window_finish");
    cudaSetDevice(h->cfg.device);
    PeerBases bases{};
    for (int d = 0; d <= h->xc.n; ++d) bases.v[d] = person_bases[d];
    PeerPtrs dst{};
    PeerFlags flags{};
    peer_targets(h, true, dst, flags);
    const CandArrays C{h->cand_person.p, h->cand_gkey.p, h->cand_t.p, h->cand_p.p, (uint32_t) h->cand_person.n};
    k_p2p_export_candidates<<<h->n_sm * 2, TPB, 0, h->stream>>>(
        C, h->d_ctr.p, bases, h->xc.n, h->person_base, h->n_agents, (uint32_t) h->xc.cc, dst, flags, h->xc.seq,
        reinterpret_cast<unsigned long long*>(h->xc.region + 1024), reinterpret_cast<unsigned int*>(h->xc.region + 1536));
    CU(h, cudaGetLastError());
    return HEROS_OK;
}

int32_t heros_window_recv_candidates(heros_handle h)
{
    if (!h) return HEROS_ERR_INVALID;
    if (!h->xc.connected) return fail(h, HEROS_ERR_INVALID, "heros_exchange_connect has not been called");
    if (h->win_state != 2) return fail(h, HEROS_ERR_INVALID, "candidates are received between heros_window_transmit and heros_window_finish");
    cudaSetDevice(h->cfg.device);
    // Candidates can only come back from the shards this one sent guests to, so only their blocks are parsed.  Which
    // flags are awaited is a matter of flow control: with every shard listening to every shard for guests, a peer can be
    // at most one window ahead and waiting for the senders' targets is enough.  With a source mask
    // (heros_exchange_set_sources) the guest wait no longer couples all shards, and a shard that sends nothing for two
    // windows could overwrite a guest block that has not been parsed yet: then the candidate flags of ALL peers are
    // awaited (every shard raises them every window, after it has parsed its guests).
    const unsigned long long all = h->xc.n >= 64 ? ~0ull : ((1ull << h->xc.n) - 1ull);
    const unsigned long long others = all & ~(1ull << h->xc.rank);
    const bool masked = (h->xc.sources & others) != others;
    k_p2p_wait<<<1, MAX_PEERS, 0, h->stream>>>(reinterpret_cast<const unsigned long long*>(h->xc.region) + 64, h->xc.n,
                                               masked ? others : h->xc.sent_to, h->xc.seq, h->d_ctr.p);
    CU(h, cudaGetLastError());
    return heros_import_candidate_blocks_from(h, h->xc.sent_to);
}

}  // extern "C"

// candidate blocks of the peers in `mask` only (the others were not awaited: their blocks may still be written)
static int32_t heros_import_candidate_blocks_from(heros_handle h, unsigned long long mask)
{
    const size_t cb = candidate_block_bytes(h->xc.cc);
    const unsigned char* first = h->xc.region + REGION_HEADER + 2 * (size_t) h->xc.n * guest_block_bytes(h->xc.gc);
    for (int d = 0; d < h->xc.n; ++d)
    {
        if (!((mask >> d) & 1ull)) continue;
        const int32_t rc = heros_import_candidate_blocks_device(h, 1, (int64_t) h->xc.cc, first + (size_t) d * cb);
        if (rc != HEROS_OK) return rc;
    }
    return HEROS_OK;
}

// guest blocks of the peers in `mask` only: the others were declared not to send guests (heros_exchange_set_sources)
// and were not awaited.  The guest arrays are indexed by block * capacity + row, so every peer keeps its slot.
static int32_t submit_guest_blocks_from(heros_handle h, unsigned long long mask)
{
    const unsigned char* first = h->xc.region + REGION_HEADER +
                                 (size_t) (h->xc.seq & 1ull) * (size_t) h->xc.n * guest_block_bytes(h->xc.gc);
    bool all = true;
    for (int d = 0; d < h->xc.n; ++d) all = all && ((mask >> d) & 1ull);
    if (all) return heros_submit_guest_blocks_device(h, h->xc.n, (int64_t) h->xc.gc, first);
    // compact the awaited blocks' indices: parse them one by one into consecutive guest slots
    int32_t rc = HEROS_OK;
    for (int d = 0; d < h->xc.n && rc == HEROS_OK; ++d)
        if ((mask >> d) & 1ull) rc = submit_guest_blocks_append(h, 1, (int64_t) h->xc.gc, first + (size_t) d * guest_block_bytes(h->xc.gc));
    return rc;
}
