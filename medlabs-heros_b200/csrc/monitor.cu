// monitor.cu -- occupancy sampling for the hourly location dumps: heros_sample_occupancy.
#include "engine_state.cuh"

namespace heros {
namespace {

// HerosModel.locationDump (model/HerosModel.java:266-299) writes, every simulated hour, the number of persons in every
// location (Location.getAllPersonIds().size()) and in every sub-location
// (DiseaseTransmission.getNrPersonsInSublocation(location, index), :284).  Here the stays the engine holds for the
// coming window (pending closed stays + the open stay of every person) are sampled at the requested times: a stay
// counts at time t iff t_in <= t < t_out.  One thread per stay; the sample times are few (24 per day) and ascending.
__device__ __forceinline__ void sample_stay(uint32_t key, double tin, double tout, int n_times, const double* times,
                                            uint32_t n_keys, int32_t* per_sub)
{
    int lo = 0, hi = n_times;  // first sample time >= t_in
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (times[mid] < tin) lo = mid + 1; else hi = mid;
    }
    for (int j = lo; j < n_times && times[j] < tout; ++j) atomicAdd(&per_sub[(size_t) j * n_keys + key], 1);
}

__global__ void __launch_bounds__(TPB) k_sample_pool(const unsigned long long* n_dev, const uint32_t* key, const double* tin,
                                                     const double* tout, int n_times, const double* times, uint32_t n_keys,
                                                     int32_t* per_sub)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t) *n_dev) return;  // the fill count of the pool lives on the device
    const uint32_t k = key[i];
    if (k == KEY_NONE || k >= n_keys) return;  // retired slot / a stay in another shard's location
    sample_stay(k, tin[i], tout[i], n_times, times, n_keys, per_sub);
}

__global__ void __launch_bounds__(TPB) k_sample_open(uint32_t n_agents, const uint32_t* open_key, const double* open_tin,
                                                     int n_times, const double* times, uint32_t n_keys, int32_t* per_sub)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_agents) return;
    const uint32_t k = open_key[a];
    const double tin = open_tin[a];
    if (k == KEY_NONE || k >= n_keys || tin != tin) return;  // nobody in the slot
    sample_stay(k, tin, bits_f64(0x7ff0000000000000ull), n_times, times, n_keys, per_sub);
}

// persons per location = the sum over its sub-locations; one thread per (sample time, location)
__global__ void __launch_bounds__(TPB) k_sum_locations(uint32_t n_loc, const uint32_t* loc_base, int n_times, uint32_t n_keys,
                                                       const int32_t* per_sub, int32_t* per_loc)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) n_times * n_loc) return;
    const uint32_t loc = (uint32_t) (i % n_loc);
    const size_t j = i / n_loc;
    int32_t s = 0;
    for (uint32_t k = loc_base[loc]; k < loc_base[loc + 1]; ++k) s += per_sub[j * n_keys + k];
    per_loc[i] = s;
}

}  // namespace
}  // namespace heros

using namespace heros;
using namespace heros_impl;

extern "C" {

int32_t heros_sample_occupancy(heros_handle h, int32_t n_times, const double* times, int64_t n_locations,
                               int32_t* persons_per_location, int64_t n_sublocations, int32_t* persons_per_sublocation)
{
    int32_t rc = check_ready(h);
    if (rc != HEROS_OK) return rc;
    if (n_times <= 0 || !times) return fail(h, HEROS_ERR_INVALID, "no sample times");
    for (int i = 0; i < n_times; ++i)
        if (!(times[i] == times[i]) || (i > 0 && !(times[i] > times[i - 1])))
            return fail(h, HEROS_ERR_INVALID, "sample times must be ascending");
    if (h->win_state != 0) return fail(h, HEROS_ERR_INVALID, "a window opened by heros_window_begin is still in progress");
    if (h->person_base != 0 || h->location_base != 0)
        return fail(h, HEROS_ERR_UNSUPPORTED, "occupancy sampling runs on an unsharded engine");
    if (persons_per_location && n_locations != (int64_t) h->n_loc)
        return fail(h, HEROS_ERR_INVALID, "persons_per_location must hold n_times x (number of locations) counts");
    if (persons_per_sublocation && n_sublocations != (int64_t) h->n_keys)
        return fail(h, HEROS_ERR_INVALID, "persons_per_sublocation must hold n_times x sum(max(n_sub, 1)) counts");
    if (!persons_per_location && !persons_per_sublocation) return HEROS_OK;
    cudaSetDevice(h->cfg.device);
    cudaStream_t s = h->stream;
    const size_t n_sub_counts = (size_t) n_times * h->n_keys, n_loc_counts = (size_t) n_times * h->n_loc;
    DevBuf<double>& d_times = h->m_times;  // sampling buffers, kept for the next call
    DevBuf<int32_t>&d_sub = h->m_sub, &d_loc = h->m_loc;
    CU(h, d_times.ensure(n_times)); CU(h, d_sub.ensure(n_sub_counts));
    CU(h, cudaMemcpyAsync(d_times.p, times, (size_t) n_times * 8, cudaMemcpyHostToDevice, s));
    CU(h, cudaMemsetAsync(d_sub.p, 0, n_sub_counts * 4, s));
    { const int32_t rc_ = join_submissions(h); if (rc_ != HEROS_OK) return rc_; }  // what k_submit is still writing
    const PoolArrays F = fresh_arrays(h), Cy = carry_arrays(h, h->carry_cur);
    if (h->fresh_ub)
        k_sample_pool<<<blocks_for(h->fresh_ub), TPB, 0, s>>>(&h->d_ctr.p->n_fresh, F.key, F.tin, F.tout, n_times, d_times.p, h->n_keys, d_sub.p);
    if (h->carry_ub)
        k_sample_pool<<<blocks_for(h->carry_ub), TPB, 0, s>>>(&h->d_ctr.p->n_carry, Cy.key, Cy.tin, Cy.tout, n_times, d_times.p, h->n_keys, d_sub.p);
    k_sample_open<<<blocks_for(h->n_agents), TPB, 0, s>>>(h->n_agents, h->d_open_key.p, h->d_open_tin.p, n_times, d_times.p,
                                                         h->n_keys, d_sub.p);
    CU(h, cudaGetLastError());
    if (persons_per_location)
    {
        CU(h, d_loc.ensure(n_loc_counts));
        k_sum_locations<<<blocks_for(n_loc_counts), TPB, 0, s>>>(h->n_loc, h->d_loc_base.p, n_times, h->n_keys, d_sub.p, d_loc.p);
        CU(h, cudaGetLastError());
        CU(h, cudaMemcpyAsync(persons_per_location, d_loc.p, n_loc_counts * 4, cudaMemcpyDeviceToHost, s));
    }
    if (persons_per_sublocation)
        CU(h, cudaMemcpyAsync(persons_per_sublocation, d_sub.p, n_sub_counts * 4, cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    return HEROS_OK;
}

}  // extern "C"
