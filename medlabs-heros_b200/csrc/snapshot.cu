// snapshot.cu -- checkpoint / resume of the dynamic state of an engine: heros_snapshot_size,
// heros_snapshot_save, heros_snapshot_load.  Host code only: the state is copied between HBM and one caller-owned blob.
#include "engine_state.cuh"

#include <cstring>

namespace {

constexpr uint64_t SNAPSHOT_MAGIC = 0x31544f4853534f52ull;  // "ROSSHOT1"

struct SnapshotHeader
{
    uint64_t magic;
    uint32_t version, model;
    uint64_t seed;
    uint32_t n_agents, n_keys, n_fresh, n_carry;
    uint32_t n_policy, n_series, n_closure, closure_active;
    uint32_t invalid_reported, pad;
    uint64_t n_inf, n_tr;
    double t_now;
    Counters ctr;
};

struct Piece { void* dev; void* host; size_t bytes; };  // one of dev / host is set

// the dynamic state of an idle engine, in the order it is laid out behind the header
std::vector<Piece> pieces(heros_handle h, const SnapshotHeader& s)
{
    const size_t N = s.n_agents, K = s.n_keys, F = s.n_fresh, C = s.n_carry;
    const int c = h->carry_cur;
    return {
        {h->d_view.p, nullptr, N * 8}, {h->d_next_time.p, nullptr, N * 8}, {h->d_ill_end.p, nullptr, N * 8},
        {h->d_open_key.p, nullptr, N * 4}, {h->d_open_tin.p, nullptr, N * 8},
        {h->d_last_calc.p, nullptr, K * 8}, {h->bk_count.p, nullptr, (K + 1) * 4},
        {h->fresh_person.p, nullptr, F * 4}, {h->fresh_key.p, nullptr, F * 4}, {h->fresh_rank.p, nullptr, F * 4},
        {h->fresh_tin.p, nullptr, F * 8}, {h->fresh_tout.p, nullptr, F * 8},
        {h->carry_person[c].p, nullptr, C * 4}, {h->carry_key[c].p, nullptr, C * 4}, {h->carry_rank[c].p, nullptr, C * 4},
        {h->carry_tin[c].p, nullptr, C * 8}, {h->carry_tout[c].p, nullptr, C * 8},
        {h->inf_person.p, nullptr, (size_t) s.n_inf * 4}, {h->inf_gkey.p, nullptr, (size_t) s.n_inf * 8},
        {h->inf_t.p, nullptr, (size_t) s.n_inf * 8}, {h->inf_p.p, nullptr, (size_t) s.n_inf * 8},
        {h->tr_person.p, nullptr, (size_t) s.n_tr * 4}, {h->tr_phase.p, nullptr, (size_t) s.n_tr},
        {h->tr_time.p, nullptr, (size_t) s.n_tr * 8},
        {nullptr, h->policy.data(), (size_t) s.n_policy * sizeof(PolicyChange)},
        {nullptr, h->series_t.data(), (size_t) s.n_series * 8},
        {nullptr, h->series_counts.data(), (size_t) s.n_series * 8 * 8},
        {nullptr, h->h_closure.data(), (size_t) s.n_closure * 8},
    };
}

size_t padded(size_t b) { return (b + 15) & ~(size_t) 15; }

int32_t describe(heros_handle h, SnapshotHeader& s)
{
    using namespace heros_impl;
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (h->win_state != 0) return fail(h, HEROS_ERR_INVALID, "a window opened by heros_window_begin is still in progress");
    if (h->xc.connected) return fail(h, HEROS_ERR_UNSUPPORTED, "snapshots of an engine connected to a peer exchange are not supported");
    cudaSetDevice(h->cfg.device);
    if ((rc = sync_counters(h)) != HEROS_OK) return rc;
    memset(&s, 0, sizeof s);
    s.magic = SNAPSHOT_MAGIC; s.version = 3; s.model = (uint32_t) h->cfg.model; s.seed = h->cfg.seed;
    s.n_agents = h->n_agents; s.n_keys = h->n_keys; s.n_fresh = (uint32_t) h->h_ctr->n_fresh; s.n_carry = (uint32_t) h->h_ctr->n_carry;
    s.n_policy = (uint32_t) h->policy.size(); s.n_series = (uint32_t) h->series_t.size();
    s.n_closure = (uint32_t) h->h_closure.size(); s.closure_active = h->closure_active ? 1u : 0u;
    s.invalid_reported = h->invalid_reported;
    s.n_inf = std::min<unsigned long long>(h->h_ctr->n_inf_log, h->inf_person.n);
    s.n_tr = std::min<unsigned long long>(h->h_ctr->n_tr_log, h->tr_person.n);
    s.t_now = h->t_now;
    s.ctr = *h->h_ctr;
    return HEROS_OK;
}

}  // namespace

using namespace heros_impl;

extern "C" {

int32_t heros_snapshot_size(heros_handle h, int64_t* bytes_out)
{
    if (!h || !bytes_out) return HEROS_ERR_INVALID;
    SnapshotHeader s;
    const int32_t rc = describe(h, s);
    if (rc != HEROS_OK) return rc;
    size_t total = padded(sizeof s);
    for (const Piece& p : pieces(h, s)) total += padded(p.bytes);
    *bytes_out = (int64_t) total;
    return HEROS_OK;
}

int32_t heros_snapshot_save(heros_handle h, int64_t capacity, void* blob, int64_t* bytes_out)
{
    if (!h || !blob || !bytes_out || capacity < 0) return HEROS_ERR_INVALID;
    SnapshotHeader s;
    int32_t rc = describe(h, s);
    if (rc != HEROS_OK) return rc;
    size_t total = padded(sizeof s);
    const std::vector<Piece> ps = pieces(h, s);
    for (const Piece& p : ps) total += padded(p.bytes);
    *bytes_out = (int64_t) total;
    if ((size_t) capacity < total) return fail(h, HEROS_ERR_CAPACITY, "snapshot buffer too small (heros_snapshot_size)");
    unsigned char* out = static_cast<unsigned char*>(blob);
    memset(out, 0, padded(sizeof s));
    memcpy(out, &s, sizeof s);
    size_t at = padded(sizeof s);
    for (const Piece& p : ps)
    {
        if (p.bytes)
        {
            if (p.dev) CU(h, cudaMemcpyAsync(out + at, p.dev, p.bytes, cudaMemcpyDeviceToHost, h->stream));
            else memcpy(out + at, p.host, p.bytes);
        }
        at += padded(p.bytes);
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return HEROS_OK;
}

int32_t heros_snapshot_load(heros_handle h, int64_t bytes, const void* blob)
{
    if (!h || !blob) return HEROS_ERR_INVALID;
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (h->win_state != 0) return fail(h, HEROS_ERR_INVALID, "a window opened by heros_window_begin is still in progress");
    if (h->xc.connected) return fail(h, HEROS_ERR_UNSUPPORTED, "snapshots of an engine connected to a peer exchange are not supported");
    if ((size_t) bytes < sizeof(SnapshotHeader)) return fail(h, HEROS_ERR_INVALID, "not a snapshot");
    SnapshotHeader s;
    memcpy(&s, blob, sizeof s);
    if (s.magic != SNAPSHOT_MAGIC || s.version != 3) return fail(h, HEROS_ERR_INVALID, "not a snapshot of this engine version");
    // the static inputs (heros_set_*) are the host's to repeat; they must describe the same model
    if (s.n_agents != h->n_agents || s.n_keys != h->n_keys || s.model != (uint32_t) h->cfg.model || s.seed != h->cfg.seed)
        return fail(h, HEROS_ERR_INVALID, "the snapshot was taken with other persons, locations, model or seed");
    if (s.n_closure != 0 && s.n_closure != (uint32_t) h->n_types * 2) return fail(h, HEROS_ERR_INVALID, "the snapshot was taken with other location types");
    if (s.closure_active && !h->have_schedule) return fail(h, HEROS_ERR_INVALID, "the snapshot holds a closure policy: call heros_set_schedule first");
    if (s.n_inf > h->inf_person.n || s.n_tr > h->tr_person.n) return fail(h, HEROS_ERR_INVALID, "corrupt snapshot (log sizes)");
    cudaSetDevice(h->cfg.device);
    {
        // nothing of the engine is touched before the blob has been found complete
        size_t total = padded(sizeof s);
        for (const Piece& p : pieces(h, s)) total += padded(p.bytes);
        if ((size_t) bytes < total) return fail(h, HEROS_ERR_INVALID, "truncated snapshot");
    }
    if ((rc = sync_counters(h)) != HEROS_OK) return rc;  // nothing of the engine is in flight while its state is replaced
    if ((rc = ensure_fresh(h, std::max<size_t>(s.n_fresh, 1))) != HEROS_OK) return rc;
    {
        const int c = h->carry_cur;
        const size_t n = std::max<size_t>(s.n_carry, 1);
        CU(h, h->carry_person[c].ensure(n)); CU(h, h->carry_key[c].ensure(n)); CU(h, h->carry_rank[c].ensure(n));
        CU(h, h->carry_tin[c].ensure(n)); CU(h, h->carry_tout[c].ensure(n));
    }
    h->policy.assign(s.n_policy, PolicyChange{});
    h->series_t.assign(s.n_series, 0.0);
    h->series_counts.assign((size_t) s.n_series * 8, 0);
    h->h_closure.assign(s.n_closure, 1.0);
    const std::vector<Piece> ps = pieces(h, s);  // after the resizes: the host pieces point into the vectors
    const unsigned char* in = static_cast<const unsigned char*>(blob);
    size_t at = padded(sizeof s);
    for (const Piece& p : ps)
    {
        if (p.bytes)
        {
            if (p.dev) CU(h, cudaMemcpyAsync(p.dev, in + at, p.bytes, cudaMemcpyHostToDevice, h->stream));
            else memcpy(p.host, in + at, p.bytes);
        }
        at += padded(p.bytes);
    }
    s.ctr.timeline = h->d_timeline.p;  // a device pointer of the engine that took the snapshot means nothing here
    CU(h, cudaMemcpyAsync(h->d_ctr.p, &s.ctr, sizeof(Counters), cudaMemcpyHostToDevice, h->stream));
    if (s.n_policy)
    {
        CU(h, h->d_policy.ensure(s.n_policy));
        CU(h, cudaMemcpyAsync(h->d_policy.p, h->policy.data(), (size_t) s.n_policy * sizeof(PolicyChange), cudaMemcpyHostToDevice, h->stream));
    }
    if (s.n_closure)
    {
        if (!h->s_loc_type.p)
        {
            CU(h, h->s_loc_type.ensure(h->n_loc));
            CU(h, cudaMemcpyAsync(h->s_loc_type.p, h->h_loc_type.data(), h->n_loc, cudaMemcpyHostToDevice, h->stream));
        }
        CU(h, h->s_closure.ensure(h->n_types));
        CU(h, cudaMemcpyAsync(h->s_closure.p, h->h_closure.data(), (size_t) s.n_closure * 8, cudaMemcpyHostToDevice, h->stream));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    h->closure_active = s.closure_active != 0;
    h->fresh_ub = s.n_fresh; h->carry_ub = s.n_carry; h->t_now = s.t_now; h->invalid_reported = s.invalid_reported;
    cudaStreamSynchronize(h->meta_stream);
    for (bool& p : h->keep_pending) p = false;
    h->series_fetched = s.n_series;
    h->n_guest = 0;
    h->o_inf_n = h->o_tr_n = ~0ull;  // the sorted output caches are stale
    h->ctr_current = false;
    return sync_counters(h);
}

}  // extern "C"
