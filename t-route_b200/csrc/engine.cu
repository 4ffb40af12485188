/*
 * engine.cu -- host side of libtroute_b200.so: the C ABI of include/troute_b200.h.
 *
 * Owns what compute_network_structured sets up per call in Python/Cython objects
 * (/root/reference/src/troute-routing/troute/routing/fast_reach/mc_reach.pyx:287-378: binary_find per
 * reach, MC_Segment / MC_Reach / MC_Levelpool objects, the _Reach struct array :476-481) -- but builds it
 * ONCE per network as flat level-sorted arrays resident in HBM, and per routing call only moves the
 * forcing (qlat, q0) in and the (n_rows, 3*nsteps) result out.
 */
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/troute_b200.h"
#include "kernels.cuh"
#include "internal.h"

using namespace trt;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(TRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                      \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t count)
    {
        if (count <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) cap = count;
        return e;
    }
};

}  // namespace

struct trt_network {
    int device = 0;
    int64_t n = 0;                 // rows == segments
    int nlevels = 0;
    std::vector<int32_t> level_of_row, pos_of_row, row_of_pos, lvl_ptr;
    std::vector<uint8_t> kind_of_row;
    std::vector<int64_t> boundary_rows;                       // rows of kind TRT_KIND_BOUNDARY
    int64_t n_routed = 0;                                     // all other rows

    // device topology / parameters: tile records (kernels.cuh), CSR of upstream positions, level offsets
    int64_t n_tiles = 0;
    std::vector<uint32_t> h_rec;                              // host mirror of d_rec (flags / slots change after creation)
    DevBuf<unsigned> d_rec;
    DevBuf<int> d_lvl_ptr, d_up_idx, d_row_of_pos, d_lp_slot;

    // level pools
    int64_t n_lp = 0;
    DevBuf<int> d_lp_pos;
    std::vector<int32_t> host_lp_pos;
    DevBuf<float> d_lp_qd0, d_lp_h0;

    // per-call state
    int T = 0, qts = 1, nq = 0;
    bool uploaded = false, ran = false;
    DevBuf<float> d_qlat_in, d_q0_in, d_qlat_t, d_S, d_fvd, d_up_out, d_bnd_fvd, d_lp_in, d_carry;
    DevBuf<unsigned> d_fmask;
    DevBuf<int> d_bnd_pos, d_tmp_pos;
    DevBuf<unsigned long long> d_hash;
    bool sync_ready = false;                                  // modes 0 / 1: the flow state of the uploaded call is initialised
    bool carry_valid = false;                                 // d_carry holds the last column of the previous window
    int carry_T = 0;
    float carry_dt = 0.f;

    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t copy_stream = nullptr;                       // trt_route: results of chunk c go home while c + 1 runs
    std::vector<cudaEvent_t> chunk_events;
    int route_chunks = 0;                                     // time chunks of trt_run_download; 0 = chosen per call (auto_route_chunks)
    int host_shards = 1;                                      // GPUs of this host that route shards of the same call (they share the host's copy bandwidth)
    DevBuf<float> d_deep_fvd;                                 // [n_deep][3T] results of the marching rows (chunked trt_route)
    DevBuf<float> d_last;                                     // [n][3] last timestep of every row (trt_download_last_step)
    float* h_deep_fvd = nullptr;                              // pinned staging of d_deep_fvd
    size_t h_deep_cap = 0;
    std::vector<cudaEvent_t> copy_events;                     // chunk c has reached the host (TRT_TIMELINE read-out)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // dataflow schedule (mode 2)
    DevBuf<int> d_unit_ptr, d_gate_stage, d_done, d_ctrl;     // d_ctrl: [0] claim, [1] frontier, [2] abort
    DevBuf<unsigned char> d_unit_shift;
    DevBuf<unsigned long long> d_stage_time;                  // mode 2 + profile_stages: completion time of every stage
    int sched_T = -1, sched_short = -1, sched_gate = -1, sched_nstages = 0, sched_lw = -1;
    int gate = 0;                                             // 0 = adaptive run-ahead window, else fixed stages
    int gate_min = 12;
    int64_t gate_lanes = 16384;
    bool profile_stages = false;                              // mode 0: time every stage launch
    std::vector<cudaEvent_t> stage_events;
    std::vector<float> stage_ms;
    std::vector<int64_t> stage_width;
    // marching schedule (mode 3: every level; mode 4: levels >= deep_level_used after the dataflow kernel)
    DevBuf<int> d_march_start;
    DevBuf<unsigned char> d_march_cnt;
    int march_group = 0;                                      // positions per marching warp (1..32), 0 = auto
    int march_group_used = 0;
    int deep_level = -1;                                      // mode 4: first marching level (-1 = from deep_lanes)
    int64_t deep_lanes = 8192;                                // mode 4 auto: march as many of the deepest levels as fit
    int march_sched_first = -1, march_sched_group = -1, march_units = 0;
    bool march_profile = false;
    int poll_sleep = -1, march_prepare = 1;
    DevBuf<unsigned long long> d_march_prof;                  // [n][4] + 1
    cudaEvent_t ev_mid = nullptr;                             // between the dataflow and the marching kernel
    // "overlap_march": the marching kernel runs BESIDE the dataflow kernel (second stream, 128-thread CTAs in the register
    // space a fourth dataflow CTA per SM would take) instead of after it; its lanes poll for the flows of the last wide level
    int overlap_march = 0;
    // dataflow_park_kernel (routing_kernels.cu): "park_max" = unfinished solves of a tile that may be parked (0 = never),
    // in stages at least "park_min_tiles" wide (-1 = 6 tiles per resident warp); stages at most "early_max_tiles" wide
    // publish lane by lane (-1 = one tile per resident warp).  park_max = 0 and early_max_tiles = 0 (the defaults): the
    // plain dataflow_kernel -- the parking form is bit-identical and raises the busy lanes per instruction (20.5 -> 24.3)
    // but its larger hot code falls off the instruction cache (profiles/r02_v7_park), so it is an option, not the default.
    int park_max = 0;
    int64_t park_min_tiles = -1, early_max_tiles = 0;
    DevBuf<unsigned> d_park_pool;
    cudaStream_t march_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_wide_end = nullptr;
    bool overlapped_run = false;
    double wide_ms = 0.0, march_ms = 0.0;
    int deep_level_used = 0;
    bool prepared = false;                                    // sentinel reset done for the next run
    std::vector<int32_t> host_bnd_pos;                        // prescribed rows of the last upload (positions)
    int64_t n_bnd = 0;
    DevBuf<int> d_zero_pos;                                   // boundary rows nobody prescribes or imports
    int64_t n_zero = 0;

    // cut edges to / from other shards
    std::vector<uint8_t> imported;                            // [n] row is written by a peer
    DevBuf<int> d_exp_peer;
    DevBuf<long long> d_exp_pos;
    int64_t n_exp = 0;
    float* peer_q[TRT_MAX_PEERS] = {nullptr};
    long long peer_n[TRT_MAX_PEERS] = {0};

    // streamflow nudging
    int64_t n_gages = 0;
    int gage_max = 0;
    float gage_dt = 0.f, gage_decay = 0.f;
    std::vector<uint8_t> gage_flag;                           // [n] position carries an active gage
    std::vector<uint8_t> export_flag;                         // [n] position exports to a peer
    std::vector<int32_t> gage_slot, export_slot;              // [n] gage index / export slot of a flagged position
    DevBuf<int> d_gage_pos;
    DevBuf<unsigned char> d_gage_active;
    DevBuf<float> d_usgs, d_lastobs, d_lastobs_init, d_nudge;

    bool collect_trips = false;                               // sum the secant trips of every segment over a run
    int trip_buckets = 1;                                     // ... separately for this many equal time slices of the call
    int trip_buckets_ran = 0;                                 // layout of d_trip_sum after the last collecting run
    DevBuf<int> d_trip_sum;                                   // [trip_buckets + 1][n] by position (last row: over-bank steps)

    // options / stats
    int mode = 4;                  // 0 stage-per-launch, 1 persistent cooperative (grid.sync per stage), 2 dataflow,
                                   // 3 marching lanes, 4 dataflow for the wide shallow levels + marching for the deep ones
    int grid_blocks = 0;           // 0 = max co-resident
    double kernel_ms = 0.0;
    int64_t launches = 0, stages = 0, lane_steps = 0;

    NetDev netdev() const
    {
        NetDev d;
        d.n = (int)n; d.nlevels = nlevels; d.lvl_ptr = d_lvl_ptr.p; d.up_idx = d_up_idx.p; d.rec = d_rec.p;
        d.row_of_pos = d_row_of_pos.p; d.lp_slot = d_lp_slot.p;
        return d;
    }
    RunDev rundev(int short_ts) const
    {
        RunDev r;
        r.T = T; r.t_off = 0; r.Tc = T; r.qts = qts; r.nq = nq; r.short_ts = short_ts; r.qlat_t = d_qlat_t.p; r.S = d_S.p;
        r.trip_sum = collect_trips ? d_trip_sum.p : nullptr;
        r.trip_buckets = trip_buckets;
        r.gage.n_gages = (int)n_gages; r.gage.gmax = gage_max; r.gage.dt = gage_dt; r.gage.decay = gage_decay;
        r.gage.usgs = d_usgs.p; r.gage.lastobs = d_lastobs.p; r.gage.nudge = d_nudge.p;
        r.fmask = d_fmask.p; r.lp_in = d_lp_in.p;
        return r;
    }
};

int trt_internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }

extern "C" {

const char* trt_last_error(void) { return g_err.c_str(); }
int trt_version(void) { return 200; }   /* 2.0: tiled flow state + TMA-staged tile records, trt_continue, trt_result_hash */

int trt_device_count(void)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) return fail(TRT_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return c;
}

int trt_network_create(int device, int64_t n_rows, const int64_t* up_ptr, const int64_t* up_rows, const uint8_t* kind,
                       const float* data_values, int32_t ncols, const int32_t* scols, trt_network** out)
{
    return trt_network_create_ex(device, n_rows, up_ptr, up_rows, kind, data_values, ncols, scols, nullptr, out);
}

int trt_network_create_ex(int device, int64_t n_rows, const int64_t* up_ptr, const int64_t* up_rows, const uint8_t* kind,
                          const float* data_values, int32_t ncols, const int32_t* scols, const int32_t* levels_in,
                          trt_network** out)
{
    return trt_network_create_ordered(device, n_rows, up_ptr, up_rows, kind, data_values, ncols, scols, levels_in, nullptr,
                                      out);
}

int trt_network_create_ordered(int device, int64_t n_rows, const int64_t* up_ptr, const int64_t* up_rows,
                               const uint8_t* kind, const float* data_values, int32_t ncols, const int32_t* scols,
                               const int32_t* levels_in, const int32_t* order_key, trt_network** out)
{
    if (!out) return fail(TRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_rows < 0 || n_rows > 2000000000LL) return fail(TRT_ERR_INVALID, "n_rows out of range: %lld", (long long)n_rows);
    if (n_rows > 0 && (!up_ptr || !kind || !data_values || !scols))
        return fail(TRT_ERR_INVALID, "NULL array argument");
    if (ncols < 9) return fail(TRT_ERR_INVALID, "data_values needs at least 9 columns, got %d", ncols);
    for (int c = 0; c < 9; ++c)
        if (scols[c] < 0 || scols[c] >= ncols) return fail(TRT_ERR_INVALID, "scols[%d]=%d out of range", c, scols[c]);
    const int64_t n = n_rows;
    const int64_t E = n ? up_ptr[n] : 0;
    if (n && up_ptr[0] != 0) return fail(TRT_ERR_INVALID, "up_ptr[0] must be 0");
    if (E < 0 || E > 2000000000LL) return fail(TRT_ERR_INVALID, "edge count out of range");
    for (int64_t r = 0; r < n; ++r)
        if (up_ptr[r + 1] < up_ptr[r]) return fail(TRT_ERR_INVALID, "up_ptr not monotone at row %lld", (long long)r);
    for (int64_t e = 0; e < E; ++e)
        if (up_rows[e] < 0 || up_rows[e] >= n)
            return fail(TRT_ERR_INVALID, "upstream row %lld out of range (edge %lld)", (long long)up_rows[e], (long long)e);
    for (int64_t r = 0; r < n; ++r) {
        if (kind[r] > TRT_KIND_BOUNDARY) return fail(TRT_ERR_INVALID, "kind[%lld]=%d unknown", (long long)r, kind[r]);
        if (kind[r] == TRT_KIND_BOUNDARY && up_ptr[r + 1] != up_ptr[r])
            return fail(TRT_ERR_INVALID, "boundary row %lld must not have upstream rows", (long long)r);
    }

    trt_network* net = new (std::nothrow) trt_network();
    if (!net) return fail(TRT_ERR_NOMEM, "out of host memory");
    net->device = device;
    net->n = n;

    // ---- levels: longest path from a headwater (Kahn sweep over the downstream adjacency), or the caller's ----
    std::vector<int32_t>& level = net->level_of_row;
    level.assign((size_t)n, 0);
    if (levels_in) {
        // a shard of a larger network keeps the levels of the whole network, so that all shards walk the same stages
        for (int64_t r = 0; r < n; ++r) {
            if (levels_in[r] < 0) { delete net; return fail(TRT_ERR_INVALID, "level of row %lld is negative", (long long)r); }
            level[(size_t)r] = levels_in[r];
        }
        for (int64_t r = 0; r < n; ++r)
            for (int64_t e = up_ptr[r]; e < up_ptr[r + 1]; ++e)
                if (level[(size_t)up_rows[e]] >= level[(size_t)r]) {
                    delete net;
                    return fail(TRT_ERR_CYCLE, "level of row %lld does not exceed the level of its upstream row %lld",
                                (long long)r, (long long)up_rows[e]);
                }
    } else {
        std::vector<int64_t> down_ptr((size_t)n + 1, 0);
        for (int64_t e = 0; e < E; ++e) down_ptr[(size_t)up_rows[e] + 1]++;
        for (int64_t r = 0; r < n; ++r) down_ptr[(size_t)r + 1] += down_ptr[(size_t)r];
        std::vector<int64_t> down((size_t)E), fill(down_ptr.begin(), down_ptr.end() - 1);
        for (int64_t r = 0; r < n; ++r)
            for (int64_t e = up_ptr[r]; e < up_ptr[r + 1]; ++e) down[(size_t)fill[(size_t)up_rows[e]]++] = r;
        std::vector<int64_t> indeg((size_t)n), queue;
        queue.reserve((size_t)n);
        for (int64_t r = 0; r < n; ++r) {
            indeg[(size_t)r] = up_ptr[r + 1] - up_ptr[r];
            if (indeg[(size_t)r] == 0) queue.push_back(r);
        }
        size_t head = 0;
        while (head < queue.size()) {
            const int64_t r = queue[head++];
            for (int64_t e = down_ptr[(size_t)r]; e < down_ptr[(size_t)r + 1]; ++e) {
                const int64_t dn = down[(size_t)e];
                level[(size_t)dn] = std::max(level[(size_t)dn], level[(size_t)r] + 1);
                if (--indeg[(size_t)dn] == 0) queue.push_back(dn);
            }
        }
        if ((int64_t)queue.size() != n) {
            delete net;
            return fail(TRT_ERR_CYCLE, "upstream graph has a cycle (%lld of %lld rows reachable)",
                        (long long)queue.size(), (long long)n);
        }
    }
    int nlev = 0;
    for (int64_t r = 0; r < n; ++r) nlev = std::max(nlev, level[(size_t)r] + 1);
    net->nlevels = nlev;

    // ---- counting sort by level: engine position order ----
    net->lvl_ptr.assign((size_t)nlev + 1, 0);
    for (int64_t r = 0; r < n; ++r) net->lvl_ptr[(size_t)level[(size_t)r] + 1]++;
    for (int l = 0; l < nlev; ++l) net->lvl_ptr[(size_t)l + 1] += net->lvl_ptr[(size_t)l];
    net->pos_of_row.resize((size_t)n);
    net->row_of_pos.resize((size_t)n);
    {
        std::vector<int32_t> cursor(net->lvl_ptr.begin(), net->lvl_ptr.end() - (nlev ? 1 : 0));
        for (int64_t r = 0; r < n; ++r) {
            const int32_t p = cursor[(size_t)level[(size_t)r]]++;
            net->pos_of_row[(size_t)r] = p;
            net->row_of_pos[(size_t)p] = (int32_t)r;
        }
        if (order_key) {
            // inside a level any order is valid: sort by the caller's key (stable), e.g. the secant trip counts a previous
            // call collected, so that the lanes of a warp need the same number of trips
            for (int l = 0; l < nlev; ++l)
                std::stable_sort(net->row_of_pos.begin() + net->lvl_ptr[(size_t)l],
                                 net->row_of_pos.begin() + net->lvl_ptr[(size_t)l + 1],
                                 [&](int32_t a, int32_t b) { return order_key[a] < order_key[b]; });
            for (int64_t p = 0; p < n; ++p) net->pos_of_row[(size_t)net->row_of_pos[(size_t)p]] = (int32_t)p;
        }
    }
    net->kind_of_row.assign(kind, kind + n);
    net->imported.assign((size_t)n, 0);
    for (int64_t r = 0; r < n; ++r)
        if (kind[r] == TRT_KIND_BOUNDARY) net->boundary_rows.push_back(r); else net->n_routed++;

    // ---- position-space arrays: one 2 KB record per tile of 32 positions (kernels.cuh), CSR of upstream positions ----
    const int64_t n_tiles = (n + 31) / 32;
    net->n_tiles = n_tiles;
    std::vector<int32_t> h_up_idx((size_t)E);
    std::vector<uint32_t>& rec = net->h_rec;
    rec.assign((size_t)n_tiles * R_TILE_WORDS, 0u);
    auto f2u = [](float x) { uint32_t u; memcpy(&u, &x, 4); return u; };
    for (int64_t p = n; p < n_tiles * 32; ++p) {                      // padding lanes of the last tile: never routed
        rec[rec_idx(p, R_FLAGS)] = TRT_KIND_BOUNDARY;
        rec[rec_idx(p, R_UP0)] = rec[rec_idx(p, R_UP1)] = (uint32_t)-1;
    }
    {
        int64_t o = 0;
        for (int64_t p = 0; p < n; ++p) {
            const int64_t r = net->row_of_pos[(size_t)p];
            const int64_t cnt = up_ptr[r + 1] - up_ptr[r];
            if (cnt >= (1 << 24)) { delete net; return fail(TRT_ERR_INVALID, "row %lld has too many upstream rows", (long long)r); }
            const float* dv = data_values + (size_t)r * ncols;
            for (int c = 0; c < 9; ++c) rec[rec_idx(p, R_PAR + c)] = f2u(dv[scols[c]]);
            rec[rec_idx(p, R_LEVEL)] = (uint32_t)level[(size_t)r];
            rec[rec_idx(p, R_FLAGS)] = (uint32_t)kind[r] | ((uint32_t)cnt << 8);
            rec[rec_idx(p, R_ESTART)] = (uint32_t)o;
            rec[rec_idx(p, R_UP0)] = rec[rec_idx(p, R_UP1)] = (uint32_t)-1;
            for (int64_t e = up_ptr[r]; e < up_ptr[r + 1]; ++e) {
                const int32_t up = net->pos_of_row[(size_t)up_rows[e]];
                if (e - up_ptr[r] == 0) rec[rec_idx(p, R_UP0)] = (uint32_t)up;
                if (e - up_ptr[r] == 1) rec[rec_idx(p, R_UP1)] = (uint32_t)up;
                h_up_idx[(size_t)o++] = up;
            }
        }
    }

    // ---- upload ----
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = preload_routing_kernels();
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&net->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&net->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&net->ev1);
    if (e == cudaSuccess) e = net->d_lvl_ptr.reserve((size_t)nlev + 1);
    if (e == cudaSuccess) e = net->d_up_idx.reserve((size_t)E);
    if (e == cudaSuccess) e = net->d_row_of_pos.reserve((size_t)n);
    if (e == cudaSuccess) e = net->d_rec.reserve(rec.size());
#define UP(dst, src, count) \
    if (e == cudaSuccess && (count) > 0) e = cudaMemcpy(dst, src, (size_t)(count) * sizeof(*(src)), cudaMemcpyHostToDevice)
    UP(net->d_lvl_ptr.p, net->lvl_ptr.data(), nlev + 1);
    UP(net->d_up_idx.p, h_up_idx.data(), E);
    UP(net->d_row_of_pos.p, net->row_of_pos.data(), n);
    UP(net->d_rec.p, rec.data(), (int64_t)rec.size());
#undef UP
    if (e != cudaSuccess) {
        const int rc = fail(TRT_ERR_CUDA, "network upload failed: %s", cudaGetErrorString(e));
        trt_network_destroy(net);
        return rc;
    }
    *out = net;
    return TRT_OK;
}

int trt_network_destroy(trt_network* net)
{
    if (!net) return TRT_OK;
    cudaSetDevice(net->device);
    if (net->ev0) cudaEventDestroy(net->ev0);
    if (net->ev1) cudaEventDestroy(net->ev1);
    if (net->ev_mid) cudaEventDestroy(net->ev_mid);
    if (net->ev_fork) cudaEventDestroy(net->ev_fork);
    if (net->ev_join) cudaEventDestroy(net->ev_join);
    if (net->ev_wide_end) cudaEventDestroy(net->ev_wide_end);
    if (net->march_stream) cudaStreamDestroy(net->march_stream);
    for (cudaEvent_t ev : net->chunk_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : net->copy_events) cudaEventDestroy(ev);
    if (net->h_deep_fvd) cudaFreeHost(net->h_deep_fvd);
    for (cudaEvent_t ev : net->stage_events) cudaEventDestroy(ev);
    if (net->copy_stream) cudaStreamDestroy(net->copy_stream);
    if (net->stream && net->own_stream) cudaStreamDestroy(net->stream);
    delete net;
    return TRT_OK;
}

int trt_network_num_levels(const trt_network* net, int32_t* out)
{
    if (!net || !out) return fail(TRT_ERR_INVALID, "NULL argument");
    *out = net->nlevels;
    return TRT_OK;
}
int trt_network_get_levels(const trt_network* net, int32_t* level_of_row)
{
    if (!net || !level_of_row) return fail(TRT_ERR_INVALID, "NULL argument");
    std::copy(net->level_of_row.begin(), net->level_of_row.end(), level_of_row);
    return TRT_OK;
}
int trt_network_get_positions(const trt_network* net, int32_t* pos_of_row)
{
    if (!net || !pos_of_row) return fail(TRT_ERR_INVALID, "NULL argument");
    std::copy(net->pos_of_row.begin(), net->pos_of_row.end(), pos_of_row);
    return TRT_OK;
}

int trt_network_set_levelpools(trt_network* net, int64_t n_lp, const int64_t* lp_rows, const double* wbody_cols,
                               float routing_period)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (n_lp < 0 || (n_lp > 0 && (!lp_rows || !wbody_cols))) return fail(TRT_ERR_INVALID, "bad level-pool arguments");
    CU(cudaSetDevice(net->device));
    const int64_t n = net->n;
    std::vector<int32_t> pos((size_t)n_lp);
    std::vector<float> qd0((size_t)n_lp), h0((size_t)n_lp), par((size_t)n_lp * 9);
    for (int64_t i = 0; i < n_lp; ++i) {
        const int64_t r = lp_rows[i];
        if (r < 0 || r >= n) return fail(TRT_ERR_INVALID, "level-pool row %lld out of range", (long long)r);
        if (net->kind_of_row[(size_t)r] != TRT_KIND_LEVELPOOL)
            return fail(TRT_ERR_INVALID, "row %lld is not of kind TRT_KIND_LEVELPOOL", (long long)r);
        const double* a = wbody_cols + 11 * i;
        pos[(size_t)i] = net->pos_of_row[(size_t)r];
        // MC_Levelpool.__init__ argument order, levelpool.pyx:48-57 (double -> C float)
        const float area = (float)a[0], max_depth = (float)a[1], oa = (float)a[2], oc = (float)a[3], oe = (float)a[4],
                    wc = (float)a[5], we = (float)a[6], wl = (float)a[7], ifd = (float)a[8];
        const float we0 = (float)a[10];
        float H = we0;
        if (we0 < -900000000.0f) H = oe + ((max_depth - oe) * ifd);   // levelpool_structs.c:97-103
        qd0[(size_t)i] = (float)a[9];                                  // mc_reach.pyx:298
        h0[(size_t)i] = H;
        float* p = &par[(size_t)i * 9];
        // slot 0 is the routing period the reservoir runs with: the `dt` ARGUMENT of the call (routing_period,
        // mc_reach.pyx:272,553), not a table column -- lake rows carry NaN channel parameters (compute.py:1455-1457)
        p[0] = routing_period;
        p[1] = area; p[2] = max_depth; p[3] = oa; p[4] = oc; p[5] = oe; p[6] = wc; p[7] = we; p[8] = wl;
        for (int c = 0; c < 9; ++c) memcpy(&net->h_rec[rec_idx(pos[(size_t)i], R_PAR + c)], &p[c], 4);   // host mirror
    }
    CU(net->d_lp_pos.reserve((size_t)n_lp));
    CU(net->d_lp_qd0.reserve((size_t)n_lp));
    CU(net->d_lp_h0.reserve((size_t)n_lp));
    if (n_lp > 0) {
        CU(cudaMemcpy(net->d_lp_pos.p, pos.data(), (size_t)n_lp * sizeof(int32_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(net->d_lp_qd0.p, qd0.data(), (size_t)n_lp * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(net->d_lp_h0.p, h0.data(), (size_t)n_lp * sizeof(float), cudaMemcpyHostToDevice));
        DevBuf<float> d_par9;
        CU(d_par9.reserve((size_t)n_lp * 9));
        CU(cudaMemcpy(d_par9.p, par.data(), (size_t)n_lp * 9 * sizeof(float), cudaMemcpyHostToDevice));
        CU(launch_scatter_lp_params(net->d_lp_pos.p, d_par9.p, net->d_rec.p, (int)n_lp, net->stream));
        if (pos != net->host_lp_pos || !net->d_lp_slot.p) {
            // level-pool index of every level-pool position (the reservoir-inflow series is kept per reservoir)
            std::vector<int32_t> slot((size_t)std::max<int64_t>(n, 1), -1);
            for (int64_t i = 0; i < n_lp; ++i) slot[(size_t)pos[(size_t)i]] = (int32_t)i;
            CU(net->d_lp_slot.reserve((size_t)std::max<int64_t>(n, 1)));
            CU(cudaMemcpyAsync(net->d_lp_slot.p, slot.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, net->stream));
            CU(cudaStreamSynchronize(net->stream));
            net->host_lp_pos = pos;
        }
        CU(cudaStreamSynchronize(net->stream));
    }
    net->n_lp = n_lp;
    return TRT_OK;
}

// flag bits and the gage / export slot words of the tile records follow the host mirror
static int upload_kind(trt_network* net)
{
    const int64_t n = net->n;
    if (n == 0) return TRT_OK;
    for (int64_t p = 0; p < n; ++p) {
        uint32_t& w = net->h_rec[rec_idx(p, R_FLAGS)];
        w &= ~(uint32_t)(TRT_KIND_EXPORT_FLAG | TRT_KIND_GAGE_FLAG);
        if (!net->export_flag.empty() && net->export_flag[(size_t)p]) {
            w |= TRT_KIND_EXPORT_FLAG;
            net->h_rec[rec_idx(p, R_EXP)] = (uint32_t)net->export_slot[(size_t)p];
        }
        if (!net->gage_flag.empty() && net->gage_flag[(size_t)p]) {
            w |= TRT_KIND_GAGE_FLAG;
            net->h_rec[rec_idx(p, R_GAGE)] = (uint32_t)net->gage_slot[(size_t)p];
        }
    }
    CU(cudaMemcpy(net->d_rec.p, net->h_rec.data(), net->h_rec.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return TRT_OK;
}

int trt_network_set_gages(trt_network* net, int64_t n_gages, const int64_t* gage_rows, const uint8_t* active,
                          const float* usgs_values, int32_t gage_maxtimestep, const float* lastobs_values_init,
                          const float* time_since_lastobs_init, float da_decay_coefficient, float routing_period)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (n_gages < 0 || gage_maxtimestep < 0) return fail(TRT_ERR_INVALID, "bad gage arguments");
    if (n_gages > 0 && (!gage_rows || !active || !lastobs_values_init || !time_since_lastobs_init ||
                        (gage_maxtimestep > 0 && !usgs_values)))
        return fail(TRT_ERR_INVALID, "NULL gage array");
    CU(cudaSetDevice(net->device));
    const int64_t n = net->n;
    net->gage_flag.assign((size_t)n, 0);
    std::vector<int32_t> slot((size_t)std::max<int64_t>(n, 1), -1), pos((size_t)n_gages);
    std::vector<float> init((size_t)n_gages * 2);
    for (int64_t g = 0; g < n_gages; ++g) {
        const int64_t r = gage_rows[g];
        if (r < 0 || r >= n) return fail(TRT_ERR_INVALID, "gage row %lld out of range", (long long)r);
        const int32_t p = net->pos_of_row[(size_t)r];
        pos[(size_t)g] = p;
        if (active[g]) {
            if (net->kind_of_row[(size_t)r] == TRT_KIND_BOUNDARY)
                return fail(TRT_ERR_INVALID, "gage row %lld is a prescribed (boundary) row", (long long)r);
            if (slot[(size_t)p] >= 0) return fail(TRT_ERR_INVALID, "two active gages on row %lld", (long long)r);
            slot[(size_t)p] = (int32_t)g;
            net->gage_flag[(size_t)p] = 1;
        }
        init[(size_t)2 * g] = time_since_lastobs_init[g];         // lastobs_times[gage_i]  (mc_reach.pyx:397)
        init[(size_t)2 * g + 1] = lastobs_values_init[g];         // lastobs_values[gage_i] (:396)
    }
    net->n_gages = n_gages; net->gage_max = gage_maxtimestep;
    net->gage_decay = da_decay_coefficient; net->gage_dt = routing_period;
    net->gage_slot = slot;
    if (n_gages > 0) {
        CU(net->d_gage_pos.reserve((size_t)n_gages));
        CU(net->d_gage_active.reserve((size_t)n_gages));
        CU(net->d_usgs.reserve((size_t)n_gages * (size_t)std::max(1, gage_maxtimestep)));
        CU(net->d_lastobs.reserve((size_t)n_gages * 2));
        CU(net->d_lastobs_init.reserve((size_t)n_gages * 2));
        CU(cudaMemcpy(net->d_gage_pos.p, pos.data(), (size_t)n_gages * sizeof(int32_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(net->d_gage_active.p, active, (size_t)n_gages, cudaMemcpyHostToDevice));
        if (gage_maxtimestep > 0)
            CU(cudaMemcpy(net->d_usgs.p, usgs_values, (size_t)n_gages * gage_maxtimestep * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(net->d_lastobs_init.p, init.data(), init.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return upload_kind(net);
}

int trt_download_gages(trt_network* net, float* nudge, float* lastobs_times, float* lastobs_values)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (!net->ran) return fail(TRT_ERR_STATE, "trt_download_gages called before trt_run");
    const int64_t G = net->n_gages;
    if (G == 0) return TRT_OK;
    CU(cudaSetDevice(net->device));
    CU(cudaStreamSynchronize(net->stream));
    if (nudge)
        CU(cudaMemcpy(nudge, net->d_nudge.p, (size_t)G * (size_t)(net->T + 1) * sizeof(float), cudaMemcpyDeviceToHost));
    if (lastobs_times || lastobs_values) {
        std::vector<float> lo((size_t)G * 2);
        CU(cudaMemcpy(lo.data(), net->d_lastobs.p, lo.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (int64_t g = 0; g < G; ++g) {
            if (lastobs_times) lastobs_times[g] = lo[(size_t)2 * g];
            if (lastobs_values) lastobs_values[g] = lo[(size_t)2 * g + 1];
        }
    }
    return TRT_OK;
}

static cudaError_t prepare_state(trt_network* net, bool sentinel);

// initial state of the flow array: column 0 from the caller's q0 / the reservoir table, or -- after trt_continue -- from
// the last column of the previous window, kept on the device
static cudaError_t init_columns(trt_network* net)
{
    cudaStream_t st = net->stream;
    if (net->n == 0) return cudaSuccess;
    if (net->carry_valid)
        return launch_column_copy(net->d_S.p, net->T + 1, 0, net->d_carry.p, (int)net->n_tiles, 1, st);
    cudaError_t e = launch_init_state(net->d_q0_in.p, net->d_row_of_pos.p, net->d_S.p, (int)net->n, net->T, st);
    if (e == cudaSuccess)
        e = launch_init_levelpool(net->d_lp_pos.p, net->d_lp_qd0.p, net->d_lp_h0.p, net->d_S.p, net->T, (int)net->n_lp, st);
    return e;
}

static int upload_forcing(trt_network* net, int32_t nsteps, int32_t qts, const float* qlat, int32_t nqcols, const float* q0,
                          int64_t n_bnd, const int64_t* bnd_rows, const float* bnd_fvd, bool carry)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (nsteps < 0) return fail(TRT_ERR_INVALID, "nsteps < 0");
    if (qts < 1) return fail(TRT_ERR_INVALID, "qts_subdivisions must be >= 1");
    // mc_reach.pyx:246-247
    if ((double)nqcols < (double)nsteps / (double)qts)
        return fail(TRT_ERR_INVALID,
                    "Number of columns (timesteps) in Qlat is incorrect: expected at least (%g), got (%d)",
                    (double)nsteps / (double)qts, nqcols);
    const int64_t n = net->n;
    if (n > 0 && (!qlat || (!q0 && !carry))) return fail(TRT_ERR_INVALID, "NULL qlat / q0");
    if (n_bnd < 0 || (n_bnd > 0 && (!bnd_rows || !bnd_fvd))) return fail(TRT_ERR_INVALID, "bad boundary arguments");
    CU(cudaSetDevice(net->device));
    cudaStream_t st = net->stream;
    if (carry) {
        // the last column of the finished window and the gages' last observations become the initial state, on the device
        // (what AbstractNetwork.new_q0 / update_waterbody_water_elevation / DataAssimilation.new_lastobs do on the host
        // between two nwm_route calls, AbstractNetwork.py:177-198, DataAssimilation.py:1506-1551)
        if (!net->ran) return fail(TRT_ERR_STATE, "trt_continue needs a finished routing call on this handle");
        CU(net->d_carry.reserve((size_t)net->n_tiles * 64));
        CU(launch_column_copy(net->d_S.p, net->T + 1, net->T, net->d_carry.p, (int)net->n_tiles, 0, st));
        if (net->n_gages > 0)
            CU(launch_carry_gages(net->d_lastobs.p, net->d_lastobs_init.p, (int)net->n_gages,
                                  (float)net->T * net->gage_dt, st));
        net->carry_valid = true;
    } else {
        net->carry_valid = false;
    }
    net->T = nsteps; net->qts = qts; net->nq = nqcols;
    net->uploaded = false; net->ran = false;

    const size_t T1 = (size_t)nsteps + 1;
    CU(net->d_qlat_in.reserve((size_t)n * nqcols));
    if (!carry) CU(net->d_q0_in.reserve((size_t)n * 3));
    CU(net->d_qlat_t.reserve((size_t)n * nqcols));
    CU(net->d_S.reserve((size_t)net->n_tiles * T1 * 64));
    CU(net->d_fmask.reserve((size_t)net->n_tiles * T1));
    CU(net->d_lp_in.reserve((size_t)net->n_lp * T1));
    // allocate here, not in trt_run: an allocation may synchronise the device, and a peer handle on the same device may
    // already be running a kernel that waits for this handle's kernels
    if (net->n_gages > 0) CU(net->d_nudge.reserve((size_t)net->n_gages * T1));
    {
        // rows of marching segments are copied home from a compact buffer in the chunked route; the strided chunk copies
        // still sweep over their (then unwritten) rows of d_fvd, so give a fresh allocation defined contents once
        const float* before = net->d_fvd.p;
        CU(net->d_fvd.reserve((size_t)n * 3 * (size_t)nsteps));
        if (net->d_fvd.p != before && n > 0 && nsteps > 0)
            CU(cudaMemsetAsync(net->d_fvd.p, 0, net->d_fvd.cap * sizeof(float), st));
    }

    if (n > 0) {
        // cudaMemcpyDefault: the forcing may already live on the device (a caller that keeps several windows of lateral
        // inflow resident passes device pointers)
        CU(cudaMemcpyAsync(net->d_qlat_in.p, qlat, (size_t)n * nqcols * sizeof(float), cudaMemcpyDefault, st));
        if (!carry) CU(cudaMemcpyAsync(net->d_q0_in.p, q0, (size_t)n * 3 * sizeof(float), cudaMemcpyDefault, st));
        CU(launch_gather_qlat(net->d_qlat_in.p, net->d_row_of_pos.p, net->d_qlat_t.p, (int)n, nqcols, st));
    }
    if (n_bnd > 0) {
        std::vector<int32_t> pos((size_t)n_bnd);
        for (int64_t i = 0; i < n_bnd; ++i) {
            const int64_t r = bnd_rows[i];
            if (r < 0 || r >= n) return fail(TRT_ERR_INVALID, "boundary row %lld out of range", (long long)r);
            if (net->kind_of_row[(size_t)r] != TRT_KIND_BOUNDARY)
                return fail(TRT_ERR_INVALID, "row %lld is not of kind TRT_KIND_BOUNDARY", (long long)r);
            pos[(size_t)i] = net->pos_of_row[(size_t)r];
        }
        CU(net->d_bnd_pos.reserve((size_t)n_bnd));
        CU(net->d_bnd_fvd.reserve((size_t)n_bnd * 3 * (size_t)nsteps));
        CU(cudaMemcpyAsync(net->d_bnd_pos.p, pos.data(), (size_t)n_bnd * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(net->d_bnd_fvd.p, bnd_fvd, (size_t)n_bnd * 3 * (size_t)nsteps * sizeof(float),
                           cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));   // `pos` is a stack-owned staging vector
        net->host_bnd_pos = pos;
    } else {
        net->host_bnd_pos.clear();
    }
    net->n_bnd = n_bnd;
    {
        // boundary rows that are neither prescribed here nor written by a peer shard hold zero for every step
        std::vector<int64_t> covered(bnd_rows, bnd_rows + n_bnd);
        std::sort(covered.begin(), covered.end());
        std::vector<int32_t> zero_pos;
        for (int64_t r : net->boundary_rows)
            if (!std::binary_search(covered.begin(), covered.end(), r) && !net->imported[(size_t)r])
                zero_pos.push_back(net->pos_of_row[(size_t)r]);
        net->n_zero = (int64_t)zero_pos.size();
        if (net->n_zero > 0) {
            CU(net->d_zero_pos.reserve(zero_pos.size()));
            CU(cudaMemcpy(net->d_zero_pos.p, zero_pos.data(), zero_pos.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
    }
    net->uploaded = true;
    net->prepared = false;
    // The flow state itself is (re)built at the start of every run: prepare_state.  The bulk-synchronous schedules get it
    // here already, so that trt_import_boundary_flow can write prescribed series between the upload and the run.
    net->sync_ready = false;
    if (net->mode < 2) { CU(prepare_state(net, false)); net->sync_ready = true; }
    return TRT_OK;
}

int trt_upload_forcing(trt_network* net, int32_t nsteps, int32_t qts, const float* qlat, int32_t nqcols, const float* q0,
                       int64_t n_bnd, const int64_t* bnd_rows, const float* bnd_fvd)
{
    if (net && net->n > 0 && !q0) return fail(TRT_ERR_INVALID, "NULL qlat / q0");
    return upload_forcing(net, nsteps, qts, qlat, nqcols, q0, n_bnd, bnd_rows, bnd_fvd, false);
}

int trt_continue(trt_network* net, int32_t nsteps, int32_t qts, const float* qlat, int32_t nqcols, int64_t n_bnd,
                 const int64_t* bnd_rows, const float* bnd_fvd)
{
    return upload_forcing(net, nsteps, qts, qlat, nqcols, nullptr, n_bnd, bnd_rows, bnd_fvd, true);
}

int trt_network_update_gage_observations(trt_network* net, const float* usgs_values, int32_t gage_maxtimestep)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (gage_maxtimestep < 0 || (net->n_gages > 0 && gage_maxtimestep > 0 && !usgs_values))
        return fail(TRT_ERR_INVALID, "bad gage observation table");
    CU(cudaSetDevice(net->device));
    net->gage_max = gage_maxtimestep;
    if (net->n_gages > 0 && gage_maxtimestep > 0) {
        CU(net->d_usgs.reserve((size_t)net->n_gages * (size_t)gage_maxtimestep));
        CU(cudaMemcpy(net->d_usgs.p, usgs_values, (size_t)net->n_gages * gage_maxtimestep * sizeof(float), cudaMemcpyHostToDevice));
    }
    return TRT_OK;
}

static cudaError_t prepare_state(trt_network* net, bool sentinel)
{
    cudaStream_t st = net->stream;
    const size_t T = (size_t)net->T;
    if (net->n == 0 || T == 0) return cudaSuccess;
    // polling schedules: everything "not yet written" (0xFFFFFFFF == TRT_SENTINEL); then no flow bits, the initial state
    // and the prescribed series
    cudaError_t e = cudaMemsetAsync(net->d_S.p, sentinel ? 0xFF : 0, (size_t)net->n_tiles * (T + 1) * 64 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(net->d_fmask.p, 0, (size_t)net->n_tiles * (T + 1) * sizeof(unsigned), st);
    if (e == cudaSuccess) e = init_columns(net);
    if (e == cudaSuccess && net->n_bnd > 0)
        e = launch_fill_boundary(net->d_bnd_pos.p, net->d_bnd_fvd.p, net->d_S.p, (int)net->n_bnd, (int)T, st);
    if (e == cudaSuccess && net->n_zero > 0)
        e = launch_fill_zero_rows(net->d_zero_pos.p, net->d_S.p, (int)net->n_zero, (int)T, st);
    return e;
}

// Route the steps t_off + 1 .. t_off + Tc of the uploaded call (the whole call: 0, net->T).  Chunks must be issued in
// time order on the handle's stream; `first` resets the flow state, the nudging state and the statistics.
// phase: 0 = everything, 1 = only the dataflow (wide) levels, 2 = only the marching (deep) levels.
enum { PHASE_ALL = 0, PHASE_WIDE = 1, PHASE_DEEP = 2 };
static int run_chunk(trt_network* net, int32_t assume_short_ts, int t_off, int Tc, bool first, int phase = PHASE_ALL)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (!net->uploaded) return fail(TRT_ERR_STATE, "trt_run called before trt_upload_forcing");
    CU(cudaSetDevice(net->device));
    cudaStream_t st = net->stream;
    if (first) {
        // the flow state of this run: polling schedules start from "not yet written" everywhere (unless trt_prepare has done
        // it already: shards must all be reset before any of them runs), the bulk-synchronous ones from zeros
        if (net->mode >= 2) { if (!net->prepared) CU(prepare_state(net, true)); }
        else {
            if (!net->sync_ready) CU(prepare_state(net, false));
            else if (net->n > 0 && net->T > 0)
                CU(cudaMemsetAsync(net->d_fmask.p, 0, (size_t)net->n_tiles * (size_t)(net->T + 1) * sizeof(unsigned), st));
            net->sync_ready = true;
        }
        net->prepared = false;
        if (net->n_gages > 0) {
            // nudging state back to its initial values (after the flow-state reset: it rewrites t = 0)
            RunDev rg = net->rundev(0);
            CU(launch_reset_gages(rg.gage, net->d_gage_pos.p, net->d_gage_active.p, net->d_lastobs_init.p, net->d_S.p, net->T, st));
        }
    }
    if (net->collect_trips) CU(net->d_trip_sum.reserve((size_t)std::max<int64_t>(net->n, 1) * (size_t)(net->trip_buckets + 1)));
    CU(net->d_lp_in.reserve((size_t)net->n_lp * (size_t)(net->T + 1)));      // reservoirs declared after the upload
    const NetDev nd = net->netdev();
    RunDev rd = net->rundev(assume_short_ts ? 1 : 0);
    rd.t_off = t_off; rd.Tc = Tc;
    const int T = Tc;                          // steps this launch schedules
    const int L = assume_short_ts ? 1 : net->nlevels;
    if (first) { net->launches = 0; net->stages = 0; net->lane_steps = 0; net->kernel_ms = 0.0; }
    if (first && net->collect_trips && net->n > 0) {
        CU(cudaMemsetAsync(net->d_trip_sum.p, 0, (size_t)net->n * (size_t)(net->trip_buckets + 1) * sizeof(int), st));
        net->trip_buckets_ran = net->trip_buckets;
    }

    if (first) {
        CU(cudaEventRecord(net->ev0, st));
        net->launches += 3 + (net->n_lp > 0) + (net->n_bnd > 0) + (net->n_zero > 0) + (net->n_gages > 0);   // state reset
    }
    if (net->n > 0 && T > 0 && L > 0) {
        const int k_begin = 1, k_end = L + T;   // stages k = level + t, level in [0, L), t in [1, T]
        net->stages = k_end - k_begin;
        net->lane_steps += net->n_routed * T;
        if (net->mode >= 2) {
            // levels [0, Lw) go through the dataflow wavefront, levels [Lw, nlevels) march
            int Lw = net->nlevels;
            if (net->mode == 3) Lw = 0;
            else if (net->mode >= 4) {
                if (net->deep_level >= 0) Lw = std::min(net->deep_level, net->nlevels);
                else {
                    Lw = net->nlevels;
                    while (Lw > 0 && net->n - net->lvl_ptr[(size_t)Lw - 1] <= net->deep_lanes) --Lw;
                }
            }
            net->deep_level_used = Lw;
            const int pos_deep = net->lvl_ptr[(size_t)Lw];
            const int Lk = assume_short_ts ? (Lw > 0 ? 1 : 0) : Lw;   // levels the stage index runs over
            // (re)build the unit table of this (T, schedule) pair
            const int nstages = Lk > 0 ? Lk + T - 1 : 0;
            net->stages = nstages + (pos_deep < net->n ? T : 0);
            if (nstages > 0 && (net->sched_T != T || net->sched_short != (assume_short_ts ? 1 : 0) ||
                                net->sched_gate != net->gate || net->sched_nstages != nstages ||
                                net->sched_lw != Lw)) {
                std::vector<int32_t> unit_ptr((size_t)nstages + 1, 0), gate_stage((size_t)nstages, 0);
                std::vector<unsigned char> shift((size_t)nstages, 0);
                std::vector<int32_t> last_nonempty((size_t)nstages + 1, 0);   // last non-empty stage <= k
                std::vector<int64_t> cum((size_t)nstages + 1, 0);             // lanes in stages 1..k
                int64_t units = 0;
                int win = 1;                                                  // first stage of the run-ahead window
                for (int k = 1; k <= nstages; ++k) {
                    int64_t lo, hi;
                    if (assume_short_ts) { lo = 0; hi = pos_deep; }
                    else {
                        lo = net->lvl_ptr[(size_t)std::max(0, k - T)];
                        hi = net->lvl_ptr[(size_t)std::min(Lw, k)];
                    }
                    const int64_t width = hi - lo;
                    // A unit = 1, 2 or 4 whole tiles (32 aligned positions each).  4-tile units amortise the claim when a
                    // stage is many waves wide; narrower stages (a shard of a multi-GPU run, the medium-depth levels) use
                    // single tiles so that a stage is not four sequential tiles long.
                    const int sh = width >= (int64_t)1 << 20 ? 2 : (width >= 400000 ? 1 : 0);
                    const int64_t tiles = width > 0 ? ((hi - 1) >> 5) - (lo >> 5) + 1 : 0;
                    shift[(size_t)k - 1] = (unsigned char)sh;
                    units += (tiles + (1 << sh) - 1) >> sh;
                    if (units > 2000000000LL) return fail(TRT_ERR_INVALID, "too many work units");
                    unit_ptr[(size_t)k] = (int32_t)units;
                    last_nonempty[(size_t)k] = width > 0 ? k : last_nonempty[(size_t)k - 1];
                    cum[(size_t)k] = cum[(size_t)k - 1] + width;
                    // Run-ahead window of stage k: the stages (j, k] a unit of stage k may overtake.  Fixed gate:
                    // j = k - gate.  Adaptive (gate = 0): at least gate_min stages -- a wide stage must be able to
                    // overtake a lane stuck in the retry ladder for milliseconds -- and beyond that as many stages as
                    // hold <= gate_lanes lanes, which bounds the number of lanes polling in the narrow deep levels.
                    int j;
                    if (net->gate > 0) j = k - net->gate;
                    else {
                        while (win < k && cum[(size_t)k] - cum[(size_t)win] > net->gate_lanes) ++win;
                        j = std::min(win, k - net->gate_min);
                    }
                    gate_stage[(size_t)k - 1] = j >= 1 ? last_nonempty[(size_t)j] : 0;
                }
                CU(net->d_unit_ptr.reserve((size_t)nstages + 1));
                CU(net->d_gate_stage.reserve((size_t)nstages));
                CU(net->d_unit_shift.reserve((size_t)nstages));
                CU(net->d_done.reserve((size_t)nstages));
                CU(net->d_ctrl.reserve(8));
                CU(cudaMemcpyAsync(net->d_unit_ptr.p, unit_ptr.data(), ((size_t)nstages + 1) * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, st));
                CU(cudaMemcpyAsync(net->d_gate_stage.p, gate_stage.data(), (size_t)nstages * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, st));
                CU(cudaMemcpyAsync(net->d_unit_shift.p, shift.data(), (size_t)nstages, cudaMemcpyHostToDevice, st));
                CU(cudaStreamSynchronize(st));   // staging vectors go out of scope
                net->sched_T = T; net->sched_short = assume_short_ts ? 1 : 0; net->sched_gate = net->gate;
                net->sched_nstages = nstages; net->sched_lw = Lw;
            }
            // marching units over positions [pos_deep, n)
            int G = net->march_group;
            if (G == 0) {
                // Auto: as few segments per warp as possible (one lane per warp = the shortest link latency on the chain:
                // 27.9 -> 19.6 ms on the bench network, profiles/r02_v1_tiled_tma/lease1_ab_march_group1.json).  Residency of
                // EVERY unit is not needed -- units are claimed in position order and a segment at level l is busy from
                // link-time l to l + T -- but the segments that are LIVE at the same time must fit the resident warps,
                // or the chain stalls behind warps that are still walking an upstream segment through its T steps
                // (T = 2,016 with one lane per warp: 79 ms instead of ~35, profiles/r02_v6_overlap_onecall).
                int mg = 0;
                CU(march_max_grid(&mg));
                if (net->grid_blocks > 0) mg = std::min(mg, net->grid_blocks);
                int64_t warps = std::max<int64_t>(1, (int64_t)mg * 8);
                if (net->overlap_march && net->mode == 4 && phase == PHASE_ALL && net->grid_blocks <= 0) {
                    int sms_m = 0;                                     // beside the dataflow kernel: one 4-warp CTA per SM
                    CU(cudaDeviceGetAttribute(&sms_m, cudaDevAttrMultiProcessorCount, net->device));
                    warps = std::max<int64_t>(1, (int64_t)sms_m * 4);
                }
                // Units are claimed in position order and a warp keeps its unit for all T steps, so the warps work through
                // the marching segments at (segments / warps) x T links of occupancy while the wave needs `levels` links to
                // travel down the chain: as long as the first is the smaller number the chain never waits for a warp and one
                // lane per warp is right, whatever the widest window of levels holds.  Measured on the bench network
                // (8,192 marching segments, 2,703 levels, 2,368 resident warps): T = 288 -> 996 links of occupancy, one lane
                // per warp 18.6 ms against 22.4 ms with two (the rule of the largest T-level window asked for two);
                // T = 2,016 -> 6,974 links, one lane per warp 79 ms against ~35 ms with four.
                const int64_t n_march = net->n - pos_deep, lv_march = std::max(1, net->nlevels - Lw);
                G = 1;
                while (G < 32 && n_march * std::max(1, T) > (int64_t)G * warps * lv_march) G *= 2;
            }
            if (pos_deep < net->n && (net->march_sched_first != pos_deep || net->march_sched_group != G)) {
                std::vector<int32_t> start;
                std::vector<unsigned char> cnt;
                for (int64_t q = pos_deep; q < net->n; q += G) {
                    start.push_back((int32_t)q);
                    cnt.push_back((unsigned char)std::min<int64_t>(G, net->n - q));
                }
                CU(net->d_march_start.reserve(start.size()));
                CU(net->d_march_cnt.reserve(cnt.size()));
                CU(cudaMemcpy(net->d_march_start.p, start.data(), start.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
                CU(cudaMemcpy(net->d_march_cnt.p, cnt.data(), cnt.size(), cudaMemcpyHostToDevice));
                net->march_units = (int)start.size();
                net->march_sched_first = pos_deep; net->march_sched_group = G; net->march_group_used = G;
            }
            CU(net->d_ctrl.reserve(8));
            SchedDev sd;
            sd.wide_levels = Lk; sd.pos_end = pos_deep;
            sd.nstages = nstages; sd.T = T; sd.unit_ptr = net->d_unit_ptr.p; sd.unit_shift = net->d_unit_shift.p;
            sd.claim = (unsigned int*)net->d_ctrl.p; sd.frontier = net->d_ctrl.p + 1; sd.abort_flag = net->d_ctrl.p + 2;
            sd.done = net->d_done.p; sd.gate_stage = net->d_gate_stage.p;
            sd.stage_time = nullptr;
            if (net->profile_stages && nstages > 0) {
                CU(net->d_stage_time.reserve((size_t)nstages + 1));
                CU(cudaMemsetAsync(net->d_stage_time.p, 0, ((size_t)nstages + 1) * sizeof(unsigned long long), st));
                sd.stage_time = net->d_stage_time.p;
                net->stage_width.assign((size_t)nstages + 1, 0);
                for (int k = 1; k <= nstages; ++k)
                    net->stage_width[(size_t)k] = assume_short_ts ? pos_deep
                        : net->lvl_ptr[(size_t)std::min(Lw, k)] - net->lvl_ptr[(size_t)std::max(0, k - T)];
            }
            PeerDev pd;
            pd.exp_peer = net->d_exp_peer.p; pd.exp_pos = net->d_exp_pos.p;
            for (int i = 0; i < TRT_MAX_PEERS; ++i) pd.S[i] = net->peer_q[i];
            int grid = net->grid_blocks, max_grid = 0;
            CU(dataflow_max_grid(&max_grid));
            if (max_grid <= 0) return fail(TRT_ERR_CUDA, "dataflow kernel cannot be made resident");
            if (grid <= 0) grid = max_grid;
            {
                const int64_t warps = (int64_t)grid * 8;
                const int64_t pmin = net->park_min_tiles >= 0 ? net->park_min_tiles : 6 * warps;
                const int64_t emax = net->early_max_tiles >= 0 ? net->early_max_tiles : warps;
                sd.park_max = net->park_max;
                sd.park_min_tiles = (int)std::min<int64_t>(pmin, 0x7fffffff);
                sd.early_max_tiles = (int)std::min<int64_t>(emax, 0x7fffffff);
                sd.park_pool = nullptr;
                if (sd.park_max > 0 || sd.early_max_tiles > 0) {
                    CU(net->d_park_pool.reserve((size_t)warps * TRT_PARK_WORDS * TRT_PARK_SLOTS));
                    sd.park_pool = net->d_park_pool.p;
                }
            }
            CU(cudaMemsetAsync(net->d_ctrl.p, 0, 8 * sizeof(int), st));
            if (nstages > 0) CU(cudaMemsetAsync(net->d_done.p, 0, (size_t)nstages * sizeof(int), st));
            if (first) CU(cudaEventRecord(net->ev0, st));
            // Overlap: both kernels in flight at once.  Needs both phases in this call, and room: the dataflow kernel gives up
            // one CTA per SM (16 K registers), which holds one 128-thread marching CTA.
            int sms = 0;
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, net->device));
            const bool overlap = net->overlap_march && net->mode == 4 && phase == PHASE_ALL && nstages > 0 && pos_deep < net->n &&
                                 net->grid_blocks <= 0 && !net->march_profile && max_grid >= 2 * sms;
            net->overlapped_run = overlap;
            if (overlap) {
                if (!net->march_stream) CU(cudaStreamCreateWithFlags(&net->march_stream, cudaStreamNonBlocking));
                if (!net->ev_fork) CU(cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming));
                if (!net->ev_join) CU(cudaEventCreate(&net->ev_join));
                if (!net->ev_wide_end) CU(cudaEventCreate(&net->ev_wide_end));
                CU(cudaEventRecord(net->ev_fork, st));                 // state reset and control words are in place
                CU(cudaStreamWaitEvent(net->march_stream, net->ev_fork, 0));
                grid = max_grid - sms;
            }
            if (nstages > 0 && phase != PHASE_DEEP) {
                CU(launch_dataflow(nd, rd, sd, pd, grid, st));
                net->launches++;
                if (overlap) CU(cudaEventRecord(net->ev_wide_end, st));
            }
            if (pos_deep < net->n && phase != PHASE_WIDE) {
                MarchDev md;
                md.n_units = net->march_units;
                md.unit_start = net->d_march_start.p; md.unit_cnt = net->d_march_cnt.p;
                md.claim = (unsigned int*)net->d_ctrl.p + 4; md.abort_flag = net->d_ctrl.p + 2;
                md.prof = nullptr; md.t_start = nullptr; md.poll_sleep = net->poll_sleep; md.prepare = net->march_prepare;
                if (net->march_profile) {
                    CU(net->d_march_prof.reserve((size_t)net->n * 4 + 1));
                    CU(cudaMemsetAsync(net->d_march_prof.p, 0, (size_t)net->n * 4 * sizeof(unsigned long long), st));
                    CU(cudaMemsetAsync(net->d_march_prof.p + (size_t)net->n * 4, 0xFF, sizeof(unsigned long long), st));
                    md.prof = net->d_march_prof.p; md.t_start = net->d_march_prof.p + (size_t)net->n * 4;
                }
                if (!net->ev_mid) CU(cudaEventCreate(&net->ev_mid));
                if (!overlap) CU(cudaEventRecord(net->ev_mid, st));
                int mgrid = 0;
                CU(march_max_grid(&mgrid));
                if (mgrid <= 0) return fail(TRT_ERR_CUDA, "marching kernel cannot be made resident");
                if (net->grid_blocks > 0) mgrid = std::min(mgrid, net->grid_blocks);
                if (overlap) {
                    // one 4-warp CTA per SM beside three dataflow CTAs
                    mgrid = (int)std::min<int64_t>(sms, ((int64_t)md.n_units + 3) / 4);
                    CU(launch_march(nd, rd, md, pd, mgrid, net->march_stream, 128));
                    CU(cudaEventRecord(net->ev_join, net->march_stream));
                    CU(cudaStreamWaitEvent(st, net->ev_join, 0));       // the result pass needs both
                } else {
                    mgrid = (int)std::min<int64_t>(mgrid, ((int64_t)md.n_units + 7) / 8);
                    CU(launch_march(nd, rd, md, pd, mgrid, st));
                }
                net->launches++;
            }
        } else if (net->mode == 1) {
            int grid = net->grid_blocks;
            int max_grid = 0;
            CU(persistent_max_grid(&max_grid));
            if (max_grid <= 0) return fail(TRT_ERR_CUDA, "persistent kernel cannot be made resident");
            if (grid <= 0 || grid > max_grid) grid = max_grid;
            CU(launch_persistent(nd, rd, k_begin, k_end, grid, st));
            net->launches += 1;
        } else {
            if (net->profile_stages) {
                while ((int)net->stage_events.size() < k_end) {
                    cudaEvent_t ev; CU(cudaEventCreate(&ev)); net->stage_events.push_back(ev);
                }
                net->stage_width.assign((size_t)k_end, 0);
                CU(cudaEventRecord(net->stage_events[0], st));
            }
            for (int k = k_begin; k < k_end; ++k) {
                int lo, hi;
                if (assume_short_ts) { lo = 0; hi = (int)net->n; }
                else {
                    lo = net->lvl_ptr[(size_t)std::max(0, k - T)];
                    hi = net->lvl_ptr[(size_t)std::min(L, k)];
                }
                if (hi > lo) {
                    CU(launch_stage(nd, rd, k, lo, hi, st));
                    net->launches++;
                }
                if (net->profile_stages) {
                    net->stage_width[(size_t)k] = hi - lo;
                    CU(cudaEventRecord(net->stage_events[(size_t)k], st));
                }
            }
        }
    }
    CU(cudaEventRecord(net->ev1, st));
    const int* bp = net->d_bnd_pos.p;
    const float* bf = net->d_bnd_fvd.p;
    const int nb = (int)net->n_bnd;
    if (phase == PHASE_ALL) CU(launch_finalize(nd, rd, net->d_fvd.p, bp, bf, nb, st));
    else {
        // split result pass: the dataflow rows in place; the marching rows into a compact buffer that goes home on its own
        const int pos_deep = net->mode >= 2 ? net->lvl_ptr[(size_t)net->deep_level_used] : (int)net->n;
        if (phase == PHASE_WIDE) CU(launch_finalize(nd, rd, net->d_fvd.p, bp, bf, nb, st, 0, pos_deep, -1));
        else {
            CU(net->d_deep_fvd.reserve((size_t)(net->n - pos_deep) * 3 * (size_t)net->T));
            CU(launch_finalize(nd, rd, net->d_deep_fvd.p, bp, bf, nb, st, pos_deep, (int)net->n, pos_deep));
        }
    }
    net->launches += (net->n > 0 && T > 0) ? 1 + (nb > 0) : 0;
    net->ran = true;
    return TRT_OK;
}

static int run_async(trt_network* net, int32_t assume_short_ts)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    return run_chunk(net, assume_short_ts, 0, net->T, true);
}

int trt_run_async(trt_network* net, int32_t assume_short_ts) { return run_async(net, assume_short_ts); }

int trt_sync(trt_network* net)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    CU(cudaSetDevice(net->device));
    CU(cudaStreamSynchronize(net->stream));
    if (net->ran) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, net->ev0, net->ev1) == cudaSuccess) net->kernel_ms = ms;
        net->wide_ms = net->kernel_ms; net->march_ms = 0.0;
        if (net->mode >= 3 && net->ev_mid && net->march_units > 0 && net->deep_level_used < net->nlevels) {
            float a = 0.f, b = 0.f;
            if (net->overlapped_run) {
                // both kernels started together: wide_ms = until the dataflow kernel ended, march_ms = what the marching
                // kernel added after that
                if (cudaEventElapsedTime(&a, net->ev0, net->ev_wide_end) == cudaSuccess &&
                    cudaEventElapsedTime(&b, net->ev0, net->ev_join) == cudaSuccess) { net->wide_ms = a; net->march_ms = std::max(0.f, b - a); }
            } else if (cudaEventElapsedTime(&a, net->ev0, net->ev_mid) == cudaSuccess &&
                cudaEventElapsedTime(&b, net->ev_mid, net->ev1) == cudaSuccess) { net->wide_ms = a; net->march_ms = b; }
        }
        if (net->mode == 0 && net->profile_stages && net->stages > 0) {
            net->stage_ms.assign((size_t)net->stages + 1, 0.f);
            for (int64_t k = 1; k <= net->stages; ++k)
                cudaEventElapsedTime(&net->stage_ms[(size_t)k], net->stage_events[(size_t)k - 1], net->stage_events[(size_t)k]);
        }
        if (net->mode >= 2 && net->profile_stages && net->d_stage_time.p && net->sched_nstages > 0 &&
            net->deep_level_used > 0) {
            // dataflow stages overlap: report the time between consecutive stage completions
            const int64_t ns = net->sched_nstages;
            std::vector<unsigned long long> ts((size_t)ns + 1);
            CU(cudaMemcpy(ts.data(), net->d_stage_time.p, ts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            net->stage_ms.assign((size_t)ns + 1, 0.f);
            unsigned long long prev = ts[0];
            for (int64_t k = 1; k <= ns; ++k) {
                const unsigned long long tk = ts[(size_t)k] ? std::max(ts[(size_t)k], prev) : prev;   // empty stage
                net->stage_ms[(size_t)k] = (float)((double)(tk - prev) * 1e-6);
                prev = tk;
            }
        }
        if (net->mode >= 2 && net->d_ctrl.p && net->launches > 0) {
            int ctrl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            CU(cudaMemcpy(ctrl, net->d_ctrl.p, sizeof(ctrl), cudaMemcpyDeviceToHost));
            if (ctrl[2] != 0) {
                // which input never arrived: ctrl[2] = 1 dataflow lane, 2 run-ahead gate (stage in ctrl[5]), 3 marching lane;
                // ctrl[6..7] = address of the flow-state slot the lane was polling
                char where[256] = "";
                unsigned long long addr = 0;
                memcpy(&addr, ctrl + 6, sizeof(addr));
                const unsigned long long base = (unsigned long long)(uintptr_t)net->d_S.p;
                if (ctrl[2] == 2) snprintf(where, sizeof(where), "; the run-ahead gate waited for stage %d", ctrl[5]);
                else if (addr >= base && net->T >= 0) {
                    // S[tile][t][plane][lane]: 64 floats per (tile, t)
                    const unsigned long long off = (addr - base) / sizeof(float), T1 = (unsigned long long)(net->T + 1);
                    const long long pos = (long long)((off / (64ull * T1)) * 32ull + (off & 31ull));
                    const int t = (int)((off >> 6) % T1), c = (off & 32ull) ? 2 : 0;
                    if (pos >= 0 && pos < net->n) {
                        const int64_t r = net->row_of_pos[(size_t)pos];
                        snprintf(where, sizeof(where), "; a %s lane polled %s of row %lld (kind %d, level %d, %s) at step %d",
                                 ctrl[2] == 3 ? "marching" : "dataflow", c == 0 ? "q" : (c == 2 ? "depth" : "v"), (long long)r,
                                 (int)net->kind_of_row[(size_t)r], net->level_of_row[(size_t)r],
                                 net->imported[(size_t)r] ? "written by a peer shard" : "not imported", t);
                    }
                }
                return fail(TRT_ERR_STATE, "dataflow run aborted: a lane waited > 8 s for an input that never arrived "
                                           "(peer shard missing, or an unprescribed boundary row)%s", where);
            }
        }
    }
    return TRT_OK;
}

int trt_run(trt_network* net, int32_t assume_short_ts)
{
    const int rc = run_async(net, assume_short_ts);
    if (rc != TRT_OK) return rc;
    return trt_sync(net);
}

int trt_download_results(trt_network* net, float* fvd_out, float* upstream_out)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (!net->ran) return fail(TRT_ERR_STATE, "trt_download_results called before trt_run");
    CU(cudaSetDevice(net->device));
    cudaStream_t st = net->stream;
    const size_t n = (size_t)net->n, T = (size_t)net->T;
    if (fvd_out && n * T > 0)
        CU(cudaMemcpyAsync(fvd_out, net->d_fvd.p, n * 3 * T * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (upstream_out && n * T > 0) {
        CU(net->d_up_out.reserve(n * T));
        CU(cudaMemsetAsync(net->d_up_out.p, 0, n * T * sizeof(float), st));
        CU(launch_upstream_out(net->d_lp_pos.p, net->d_row_of_pos.p, net->d_lp_in.p, net->d_up_out.p, (int)net->n_lp, (int)T, st));
        CU(cudaMemcpyAsync(upstream_out, net->d_up_out.p, n * T * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return TRT_OK;
}

// Reservoir inflow series of the level pools alone, [n_lp, nsteps] in the order of trt_network_set_levelpools: what the
// upstream_array of compute_network_structured holds (mc_reach.pyx:710) without the n_rows x nsteps table of zeros around it
// (3.1 GB over PCIe for a CONUS day; the mirror fills its zero table on the host).
int trt_download_levelpool_inflow(trt_network* net, float* inflow_out)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (!net->ran) return fail(TRT_ERR_STATE, "trt_download_levelpool_inflow called before trt_run");
    const size_t n_lp = (size_t)net->n_lp, T = (size_t)net->T;
    if (n_lp * T == 0) return TRT_OK;
    if (!inflow_out) return fail(TRT_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(net->device));
    // lp_in is [n_lp][T + 1] with column 0 unused: a strided copy of columns 1 .. T
    CU(cudaMemcpy2DAsync(inflow_out, T * sizeof(float), net->d_lp_in.p + 1, (T + 1) * sizeof(float), T * sizeof(float), n_lp,
                         cudaMemcpyDeviceToHost, net->stream));
    CU(cudaStreamSynchronize(net->stream));
    return TRT_OK;
}

// (q, v, d) of the LAST timestep of every row: all a BMI-style caller reads back after a window (troute_model.py:318-330,
// _retrieve_last_output) -- 12 bytes per segment instead of 12 * nsteps
int trt_download_last_step(trt_network* net, float* qvd_out)
{
    if (!net || !qvd_out) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!net->ran) return fail(TRT_ERR_STATE, "trt_download_last_step called before trt_run");
    CU(cudaSetDevice(net->device));
    cudaStream_t st = net->stream;
    const size_t n = (size_t)net->n, T = (size_t)net->T;
    if (n * T == 0) return TRT_OK;
    CU(net->d_last.reserve(n * 3));
    // device-side gather of the last three columns into a contiguous [n, 3] buffer, then one contiguous copy
    CU(cudaMemcpy2DAsync(net->d_last.p, 3 * sizeof(float), net->d_fvd.p + 3 * (T - 1), 3 * T * sizeof(float), 3 * sizeof(float), n,
                         cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(qvd_out, net->d_last.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return TRT_OK;
}

// One routing call, host buffers in and out.  With a polling schedule (mode >= 2) and "route_chunks" > 1 the call is cut
// into time chunks: while the kernels of chunk c + 1 run, the finished columns of chunk c travel to the host on a second
// stream (a strided 2-D copy out of the [n_rows, 3T] result) -- the 9.4 GB result of a CONUS day takes longer to cross
// PCIe than to compute, so hiding one behind the other is worth more than any kernel tuning.  Every chunk pays the
// narrow tail of the wavefront once and narrower chunks make shorter DMA rows (below ~500 bytes the copy engine
// slows down), which bounds the useful number of chunks (auto_route_chunks below picks it per call).
int trt_route(trt_network* net, int32_t nsteps, int32_t qts, int32_t assume_short_ts, const float* qlat, int32_t nqcols,
              const float* q0, int64_t n_bnd, const int64_t* bnd_rows, const float* bnd_fvd, float* fvd_out,
              float* upstream_out)
{
    int rc = trt_upload_forcing(net, nsteps, qts, qlat, nqcols, q0, n_bnd, bnd_rows, bnd_fvd);
    if (rc != TRT_OK) return rc;
    return trt_run_download(net, assume_short_ts, fvd_out, upstream_out);
}

// How many time chunks trt_run_download cuts a call into when the caller does not say ("route_chunks" = 0).
// More chunks start the D2H copy earlier and leave a shorter last copy exposed; every chunk pays the narrow tail of the
// wavefront again (one stage latency per wide level), and rows shorter than ~500 bytes slow the copy engine down.
// Model, constants measured on B200 (profiles/r02_v3_smem_state: stage profile, e2e_timeline_by_chunks.txt, pcie_probe2.txt):
//   stage time = max(30 us, 0.11 ns x lanes of the stage);  copy rate 50 GB/s per GPU (43 GB/s below 576-byte rows), and
//   87 GB/s for all the GPUs of the host together (the e2e lines of 2 / 4 / 8 GPUs: 71 / 59 / 83 GB/s aggregate -- with
//   shards on several GPUs the call is bound by the host side of the copies, whatever the chunking);
//   call(C) = max(compute + (C - 1) x tail + copy / C,  compute / C + tail + copy).
// One GPU, 2.7 M segments x 288 steps: 6 chunks (measured 232 / 220 / 292 ms for 4 / 6 / 8).  An eighth of that network on
// each of 8 GPUs: the copy (108 ms at the host's rate) dwarfs the compute (~20 ms), so what counts is an early first chunk:
// again ~6 chunks (113 ms measured; one chunk would be compute + copy = ~130).
static int auto_route_chunks(const trt_network* net, int nsteps, int assume_short_ts)
{
    if (net->n <= 0 || nsteps <= 1) return 1;
    int Lw = net->nlevels;
    if (net->mode == 4) {
        if (net->deep_level >= 0) Lw = std::min(net->deep_level, net->nlevels);
        else { while (Lw > 0 && net->n - net->lvl_ptr[(size_t)Lw - 1] <= net->deep_lanes) --Lw; }
    } else if (net->mode == 3) return 1;
    const double levels = assume_short_ts ? 1.0 : (double)Lw;
    const double n_wide = (double)(Lw > 0 ? net->lvl_ptr[(size_t)Lw] : 0);
    const double t_lat = 30e-6, c_lane = 0.11e-9;
    const double compute = nsteps * std::max(t_lat, n_wide * c_lane) + levels * t_lat;
    const double tail = levels * t_lat;
    const double bytes = (double)net->n * 3.0 * nsteps * sizeof(float);
    const double host_rate = 87e9 / std::max(1, net->host_shards);  // "host_shards": GPUs of this host copying home at once
    int best = 1;
    double best_t = 1e30;
    for (int C : {1, 2, 3, 4, 6, 8, 12}) {
        if (C > nsteps) break;
        const double row_bytes = 3.0 * sizeof(float) * (double)(nsteps / C);
        const double copy = bytes / std::min(host_rate, row_bytes >= 576.0 ? 50e9 : 43e9 * std::min(1.0, row_bytes / 432.0));
        const double t = C == 1 ? compute + copy : std::max(compute + (C - 1) * tail + copy / C, compute / C + tail + copy);
        if (t < best_t * 0.98) { best_t = t; best = C; }            // prefer fewer chunks unless the gain is real
    }
    return best;
}

int trt_run_download(trt_network* net, int32_t assume_short_ts, float* fvd_out, float* upstream_out)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (!net->uploaded) return fail(TRT_ERR_STATE, "trt_run_download called before trt_upload_forcing");
    int rc;
    const int nsteps = net->T;
    const int want = net->route_chunks > 0 ? net->route_chunks : auto_route_chunks(net, nsteps, assume_short_ts);
    const int C = (net->mode >= 2 && fvd_out && net->n > 0) ? std::max(1, std::min(want, nsteps)) : 1;
    if (C <= 1) {
        rc = run_async(net, assume_short_ts);
        if (rc != TRT_OK) return rc;
        rc = trt_download_results(net, fvd_out, upstream_out);
        if (rc != TRT_OK) return rc;
        return trt_sync(net);
    }
    CU(cudaSetDevice(net->device));
    if (!net->copy_stream) CU(cudaStreamCreateWithFlags(&net->copy_stream, cudaStreamNonBlocking));
    while ((int)net->chunk_events.size() < C + 1) {
        cudaEvent_t ev; CU(cudaEventCreate(&ev)); net->chunk_events.push_back(ev);
        CU(cudaEventCreate(&ev)); net->copy_events.push_back(ev);
    }
    static const bool timeline = getenv("TRT_TIMELINE") != nullptr;   // debug: where the time of a chunked call goes
    const size_t n = (size_t)net->n, T = (size_t)nsteps;
    // Mode 4: only the dataflow (wide) levels are chunked -- their rows are 99 % of the result; the marching levels run
    // ONCE over all steps after the last wide chunk (a marching lane needs nothing from a later wide chunk), so the
    // latency-bound main stem is paid once, and its ~1 % of rows go home through a compact buffer + a host scatter.
    int Lw = 0;
    if (net->mode == 4) {
        if (net->deep_level >= 0) Lw = std::min(net->deep_level, net->nlevels);
        else { Lw = net->nlevels; while (Lw > 0 && net->n - net->lvl_ptr[(size_t)Lw - 1] <= net->deep_lanes) --Lw; }
    }
    const bool split = net->mode == 4 && Lw > 0 && Lw < net->nlevels;
    int t_off = 0;
    for (int c = 0; c < C; ++c) {
        const int Tc = (nsteps - t_off + (C - c) - 1) / (C - c);          // equal chunks, the longer ones first
        rc = run_chunk(net, assume_short_ts, t_off, Tc, c == 0, split ? PHASE_WIDE : PHASE_ALL);
        if (rc != TRT_OK) return rc;
        CU(cudaEventRecord(net->chunk_events[(size_t)c], net->stream));
        CU(cudaStreamWaitEvent(net->copy_stream, net->chunk_events[(size_t)c], 0));
        CU(cudaMemcpy2DAsync(fvd_out + (size_t)t_off * 3, 3 * T * sizeof(float), net->d_fvd.p + (size_t)t_off * 3,
                             3 * T * sizeof(float), (size_t)Tc * 3 * sizeof(float), n, cudaMemcpyDeviceToHost,
                             net->copy_stream));
        if (timeline) CU(cudaEventRecord(net->copy_events[(size_t)c], net->copy_stream));
        t_off += Tc;
    }
    if (split) {
        rc = run_chunk(net, assume_short_ts, 0, nsteps, false, PHASE_DEEP);
        if (rc != TRT_OK) return rc;
        const int pos_deep = net->lvl_ptr[(size_t)Lw];
        const size_t n_deep = n - (size_t)pos_deep;
        if (net->h_deep_cap < n_deep * 3 * T) {
            if (net->h_deep_fvd) cudaFreeHost(net->h_deep_fvd);
            net->h_deep_fvd = nullptr; net->h_deep_cap = 0;
            CU(cudaMallocHost(&net->h_deep_fvd, n_deep * 3 * T * sizeof(float)));
            net->h_deep_cap = n_deep * 3 * T;
        }
        if (timeline) CU(cudaEventRecord(net->chunk_events[(size_t)C], net->stream));
        CU(cudaMemcpyAsync(net->h_deep_fvd, net->d_deep_fvd.p, n_deep * 3 * T * sizeof(float), cudaMemcpyDeviceToHost,
                           net->stream));
        CU(cudaStreamSynchronize(net->stream));
        CU(cudaStreamSynchronize(net->copy_stream));          // the chunk copies wrote stale values into these rows
        for (size_t i = 0; i < n_deep; ++i)
            memcpy(fvd_out + (size_t)net->row_of_pos[(size_t)pos_deep + i] * 3 * T, net->h_deep_fvd + i * 3 * T,
                   3 * T * sizeof(float));
    }
    if (timeline) {
        CU(cudaStreamSynchronize(net->copy_stream));
        CU(cudaStreamSynchronize(net->stream));
        fprintf(stderr, "[trt timeline] ms after the state reset:");
        for (int c = 0; c < C; ++c) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, net->ev0, net->chunk_events[(size_t)c]);
            cudaEventElapsedTime(&b, net->ev0, net->copy_events[(size_t)c]);
            fprintf(stderr, " chunk %d computed %.1f home %.1f |", c, a, b);
        }
        if (split) { float a = 0.f; cudaEventElapsedTime(&a, net->ev0, net->chunk_events[(size_t)C]); fprintf(stderr, " marching rows computed %.1f", a); }
        fprintf(stderr, "\n");
    }
    if (upstream_out) {
        rc = trt_download_results(net, nullptr, upstream_out);
        if (rc != TRT_OK) return rc;
    }
    CU(cudaStreamSynchronize(net->copy_stream));
    return trt_sync(net);
}

static int rows_to_device_pos(trt_network* net, int64_t count, const int64_t* rows)
{
    std::vector<int32_t> pos((size_t)count);
    for (int64_t i = 0; i < count; ++i) {
        if (rows[i] < 0 || rows[i] >= net->n) return fail(TRT_ERR_INVALID, "row %lld out of range", (long long)rows[i]);
        pos[(size_t)i] = net->pos_of_row[(size_t)rows[i]];
    }
    CU(net->d_tmp_pos.reserve((size_t)count));
    if (count > 0)
        CU(cudaMemcpyAsync(net->d_tmp_pos.p, pos.data(), (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, net->stream));
    CU(cudaStreamSynchronize(net->stream));
    return TRT_OK;
}

int trt_export_flow_series(trt_network* net, int64_t count, const int64_t* rows, void* dst_device)
{
    if (!net || (count > 0 && (!rows || !dst_device))) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!net->uploaded) return fail(TRT_ERR_STATE, "no forcing uploaded");
    CU(cudaSetDevice(net->device));
    int rc = rows_to_device_pos(net, count, rows);
    if (rc != TRT_OK) return rc;
    CU(launch_export_series(net->d_tmp_pos.p, net->d_S.p, (float*)dst_device, (int)count, net->T, net->stream));
    CU(cudaStreamSynchronize(net->stream));
    return TRT_OK;
}

int trt_import_boundary_flow(trt_network* net, int64_t count, const int64_t* rows, const void* src_device)
{
    if (!net || (count > 0 && (!rows || !src_device))) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!net->uploaded) return fail(TRT_ERR_STATE, "no forcing uploaded");
    for (int64_t i = 0; i < count; ++i)
        if (rows[i] >= 0 && rows[i] < net->n && net->kind_of_row[(size_t)rows[i]] != TRT_KIND_BOUNDARY)
            return fail(TRT_ERR_INVALID, "row %lld is not of kind TRT_KIND_BOUNDARY", (long long)rows[i]);
    CU(cudaSetDevice(net->device));
    int rc = rows_to_device_pos(net, count, rows);
    if (rc != TRT_OK) return rc;
    CU(launch_import_series(net->d_tmp_pos.p, (const float*)src_device, net->d_S.p, (int)count, net->T, net->stream));
    CU(cudaStreamSynchronize(net->stream));
    return TRT_OK;
}

int trt_device_results(trt_network* net, void** fvd_device)
{
    if (!net || !fvd_device) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!net->ran) return fail(TRT_ERR_STATE, "trt_device_results called before trt_run");
    *fvd_device = net->d_fvd.p;
    return TRT_OK;
}

int trt_download_rows(trt_network* net, int64_t n_sel, const int64_t* rows, float* fvd_out)
{
    if (!net || (n_sel > 0 && (!rows || !fvd_out))) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!net->ran) return fail(TRT_ERR_STATE, "trt_download_rows called before trt_run");
    if (n_sel <= 0 || net->T == 0) return TRT_OK;
    CU(cudaSetDevice(net->device));
    CU(cudaStreamSynchronize(net->stream));
    const size_t w = 3 * (size_t)net->T * sizeof(float);
    // runs of consecutive rows go home as one strided copy each (a basin listed in row order is a handful of runs)
    int64_t i = 0;
    while (i < n_sel) {
        if (rows[i] < 0 || rows[i] >= net->n) return fail(TRT_ERR_INVALID, "row %lld out of range", (long long)rows[i]);
        int64_t j = i + 1;
        while (j < n_sel && rows[j] == rows[j - 1] + 1) ++j;
        CU(cudaMemcpyAsync((char*)fvd_out + (size_t)i * w, (const char*)net->d_fvd.p + (size_t)rows[i] * w, (size_t)(j - i) * w,
                           cudaMemcpyDeviceToHost, net->stream));
        i = j;
    }
    CU(cudaStreamSynchronize(net->stream));
    return TRT_OK;
}

int trt_result_hash(trt_network* net, int64_t n_sel, const int64_t* rows, const int64_t* ids, uint64_t* out)
{
    if (!net || !out) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!net->ran) return fail(TRT_ERR_STATE, "trt_result_hash called before trt_run");
    if (n_sel < 0) return fail(TRT_ERR_INVALID, "n_sel < 0");
    CU(cudaSetDevice(net->device));
    cudaStream_t st = net->stream;
    const int64_t count = rows ? n_sel : net->n;
    for (int64_t i = 0; rows && i < n_sel; ++i)
        if (rows[i] < 0 || rows[i] >= net->n) return fail(TRT_ERR_INVALID, "row %lld out of range", (long long)rows[i]);
    DevBuf<long long> d_rows, d_ids;
    CU(net->d_hash.reserve(1));
    CU(cudaMemsetAsync(net->d_hash.p, 0, sizeof(unsigned long long), st));
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t is long long");
    if (rows && count > 0) {
        CU(d_rows.reserve((size_t)count));
        CU(cudaMemcpyAsync(d_rows.p, rows, (size_t)count * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    }
    if (ids && count > 0) {
        CU(d_ids.reserve((size_t)count));
        CU(cudaMemcpyAsync(d_ids.p, ids, (size_t)count * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    }
    CU(launch_hash_rows(net->d_fvd.p, rows ? d_rows.p : nullptr, ids ? d_ids.p : nullptr, count, 3LL * net->T, net->d_hash.p, st));
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, net->d_hash.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *out = (uint64_t)h;
    return TRT_OK;
}

int trt_prepare(trt_network* net)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (!net->uploaded) return fail(TRT_ERR_STATE, "trt_prepare called before trt_upload_forcing");
    CU(cudaSetDevice(net->device));
    if (net->mode >= 2) {
        CU(prepare_state(net, true));
        CU(cudaStreamSynchronize(net->stream));
        net->prepared = true;
    }
    return TRT_OK;
}

int trt_network_set_imports(trt_network* net, int64_t count, const int64_t* rows)
{
    if (!net || (count > 0 && !rows)) return fail(TRT_ERR_INVALID, "NULL argument");
    std::fill(net->imported.begin(), net->imported.end(), 0);
    for (int64_t i = 0; i < count; ++i) {
        if (rows[i] < 0 || rows[i] >= net->n) return fail(TRT_ERR_INVALID, "import row %lld out of range", (long long)rows[i]);
        if (net->kind_of_row[(size_t)rows[i]] != TRT_KIND_BOUNDARY)
            return fail(TRT_ERR_INVALID, "import row %lld is not of kind TRT_KIND_BOUNDARY", (long long)rows[i]);
        net->imported[(size_t)rows[i]] = 1;
    }
    return TRT_OK;
}

int trt_network_set_exports(trt_network* net, int64_t count, const int64_t* rows, const int32_t* peer,
                            const int64_t* peer_pos)
{
    if (!net || (count > 0 && (!rows || !peer || !peer_pos))) return fail(TRT_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(net->device));
    const int64_t n = net->n;
    std::vector<int32_t> slot((size_t)std::max<int64_t>(n, 1), -1), h_peer((size_t)count);
    std::vector<long long> h_pos((size_t)count);
    net->export_flag.assign((size_t)n, 0);
    for (int64_t i = 0; i < count; ++i) {
        if (rows[i] < 0 || rows[i] >= n) return fail(TRT_ERR_INVALID, "export row %lld out of range", (long long)rows[i]);
        if (peer[i] < 0 || peer[i] >= TRT_MAX_PEERS) return fail(TRT_ERR_INVALID, "peer index %d out of range", peer[i]);
        if (net->kind_of_row[(size_t)rows[i]] == TRT_KIND_BOUNDARY)
            return fail(TRT_ERR_INVALID, "export row %lld is a boundary row", (long long)rows[i]);
        const int32_t pos = net->pos_of_row[(size_t)rows[i]];
        if (slot[(size_t)pos] >= 0) return fail(TRT_ERR_INVALID, "row %lld exported twice", (long long)rows[i]);
        slot[(size_t)pos] = (int32_t)i;
        h_peer[(size_t)i] = peer[i];
        h_pos[(size_t)i] = peer_pos[i];
        net->export_flag[(size_t)pos] = 1;
    }
    CU(net->d_exp_peer.reserve((size_t)count));
    CU(net->d_exp_pos.reserve((size_t)count));
    net->export_slot = slot;
    if (n > 0) {
        const int rc = upload_kind(net);
        if (rc != TRT_OK) return rc;
    }
    if (count > 0) {
        CU(cudaMemcpy(net->d_exp_peer.p, h_peer.data(), (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(net->d_exp_pos.p, h_pos.data(), (size_t)count * sizeof(long long), cudaMemcpyHostToDevice));
    }
    net->n_exp = count;
    return TRT_OK;
}

int trt_network_set_peer(trt_network* net, int32_t peer, void* peer_q_device, int64_t peer_n_rows)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (peer < 0 || peer >= TRT_MAX_PEERS) return fail(TRT_ERR_INVALID, "peer index %d out of range", peer);
    net->peer_q[peer] = (float*)peer_q_device;
    net->peer_n[peer] = peer_n_rows;
    return TRT_OK;
}

int trt_network_state_ptr(trt_network* net, void** q_device)
{
    if (!net || !q_device) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!net->uploaded) return fail(TRT_ERR_STATE, "no forcing uploaded: the flow array is allocated by trt_upload_forcing");
    *q_device = net->d_S.p;
    return TRT_OK;
}

int trt_ipc_get_handle(void* device_ptr, uint8_t handle[64])
{
    if (!device_ptr || !handle) return fail(TRT_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, device_ptr));
    memcpy(handle, &h, 64);
    return TRT_OK;
}

int trt_ipc_open_handle(int device, const uint8_t handle[64], void** device_ptr)
{
    if (!handle || !device_ptr) return fail(TRT_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TRT_OK;
}

int trt_ipc_close_handle(void* device_ptr)
{
    if (device_ptr) CU(cudaIpcCloseMemHandle(device_ptr));
    return TRT_OK;
}

int trt_set_option(trt_network* net, const char* key, int64_t value)
{
    if (!net || !key) return fail(TRT_ERR_INVALID, "NULL argument");
    if (!strcmp(key, "mode")) {
        if (value < 0 || value > 4)
            return fail(TRT_ERR_INVALID, "mode must be 0 (stage launches), 1 (persistent, grid.sync), 2 (dataflow), "
                                         "3 (marching) or 4 (dataflow + marching)");
        net->mode = (int)value;
    } else if (!strcmp(key, "grid_blocks")) {
        if (value < 0) return fail(TRT_ERR_INVALID, "grid_blocks must be >= 0");
        net->grid_blocks = (int)value;
    } else if (!strcmp(key, "profile_stages")) {
        net->profile_stages = value != 0;
    } else if (!strcmp(key, "gate")) {
        if (value < 0 || value > 1000000) return fail(TRT_ERR_INVALID, "gate must be >= 0");
        net->gate = (int)value;
        net->sched_T = -1;
    } else if (!strcmp(key, "gate_min")) {
        if (value < 1 || value > 1000000) return fail(TRT_ERR_INVALID, "gate_min must be >= 1");
        net->gate_min = (int)value;
        net->sched_T = -1;
    } else if (!strcmp(key, "gate_lanes")) {
        if (value < 1) return fail(TRT_ERR_INVALID, "gate_lanes must be >= 1");
        net->gate_lanes = value;
        net->sched_T = -1;
    } else if (!strcmp(key, "collect_trips")) {
        net->collect_trips = value != 0;
    } else if (!strcmp(key, "trip_buckets")) {
        if (value < 1 || value > 64) return fail(TRT_ERR_INVALID, "trip_buckets must be in 1..64");
        net->trip_buckets = (int)value;
    } else if (!strcmp(key, "route_chunks")) {
        if (value < 0 || value > 1024) return fail(TRT_ERR_INVALID, "route_chunks must be in 0..1024 (0 = chosen per call)");
        net->route_chunks = (int)value;
    } else if (!strcmp(key, "host_shards")) {
        if (value < 1 || value > 64) return fail(TRT_ERR_INVALID, "host_shards must be in 1..64");
        net->host_shards = (int)value;
    } else if (!strcmp(key, "march_group")) {
        if (value < 0 || value > 32) return fail(TRT_ERR_INVALID, "march_group must be in 0..32");
        net->march_group = (int)value;
    } else if (!strcmp(key, "park_max")) {
        if (value < 0 || value > 31) return fail(TRT_ERR_INVALID, "park_max must be in 0..31");
        net->park_max = (int)value;
    } else if (!strcmp(key, "park_min_tiles")) {
        if (value < -1) return fail(TRT_ERR_INVALID, "park_min_tiles must be >= -1");
        net->park_min_tiles = value;
    } else if (!strcmp(key, "early_max_tiles")) {
        if (value < -1) return fail(TRT_ERR_INVALID, "early_max_tiles must be >= -1");
        net->early_max_tiles = value;
    } else if (!strcmp(key, "overlap_march")) {
        net->overlap_march = value != 0;
    } else if (!strcmp(key, "march_prepare")) {
        net->march_prepare = value != 0;
    } else if (!strcmp(key, "poll_sleep")) {
        net->poll_sleep = (int)value;
    } else if (!strcmp(key, "march_profile")) {
        net->march_profile = value != 0;
    } else if (!strcmp(key, "deep_level")) {
        if (value < -1) return fail(TRT_ERR_INVALID, "deep_level must be >= -1");
        net->deep_level = (int)value;
    } else if (!strcmp(key, "deep_lanes")) {
        if (value < 0) return fail(TRT_ERR_INVALID, "deep_lanes must be >= 0");
        net->deep_lanes = value;
    } else if (!strcmp(key, "stream")) {
        // adopt a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream) so that the
        // caller's CUDA events bracket this handle's kernels; 0 restores the private stream
        CU(cudaSetDevice(net->device));
        CU(cudaStreamSynchronize(net->stream));
        if (value == 0) {
            if (!net->own_stream) {
                CU(cudaStreamCreateWithFlags(&net->stream, cudaStreamNonBlocking));
                net->own_stream = true;
            }
        } else {
            if (net->own_stream && net->stream) cudaStreamDestroy(net->stream);
            net->stream = (cudaStream_t)(uintptr_t)value;
            net->own_stream = false;
        }
    } else {
        return fail(TRT_ERR_INVALID, "unknown option '%s'", key);
    }
    return TRT_OK;
}

int trt_stage_profile(const trt_network* net, int64_t capacity, float* stage_ms, int64_t* stage_width, int64_t* count)
{
    if (!net || !count) return fail(TRT_ERR_INVALID, "NULL argument");
    const int64_t m = (int64_t)net->stage_ms.size();
    *count = m;
    for (int64_t k = 0; k < m && k < capacity; ++k) {
        if (stage_ms) stage_ms[k] = net->stage_ms[(size_t)k];
        if (stage_width) stage_width[k] = net->stage_width[(size_t)k];
    }
    return TRT_OK;
}

static int download_trip_sums(trt_network* net, std::vector<int32_t>& h)
{
    if (!net->collect_trips || !net->d_trip_sum.p || !net->ran || net->trip_buckets_ran < 1)
        return fail(TRT_ERR_STATE, "no trip counts: set option collect_trips = 1 before trt_run");
    CU(cudaSetDevice(net->device));
    CU(cudaStreamSynchronize(net->stream));
    h.resize((size_t)net->n * (size_t)(net->trip_buckets_ran + 1));
    if (!h.empty()) CU(cudaMemcpy(h.data(), net->d_trip_sum.p, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return TRT_OK;
}

int trt_trip_counts(trt_network* net, int32_t* trips_of_row)
{
    if (!net || !trips_of_row) return fail(TRT_ERR_INVALID, "NULL argument");
    std::vector<int32_t> h;
    const int rc = download_trip_sums(net, h);
    if (rc != TRT_OK) return rc;
    const size_t n = (size_t)net->n;
    for (size_t r = 0; r < n; ++r) {
        int32_t sum = 0;
        for (int b = 0; b < net->trip_buckets_ran; ++b) sum += h[(size_t)b * n + (size_t)net->pos_of_row[r]];
        trips_of_row[r] = sum;
    }
    return TRT_OK;
}

int trt_trip_counts_bucketed(trt_network* net, int32_t buckets, int32_t* trips /* [buckets][n_rows] */)
{
    if (!net || !trips) return fail(TRT_ERR_INVALID, "NULL argument");
    std::vector<int32_t> h;
    const int rc = download_trip_sums(net, h);
    if (rc != TRT_OK) return rc;
    if (buckets != net->trip_buckets_ran)
        return fail(TRT_ERR_INVALID, "the last collecting run used trip_buckets = %d, not %d", net->trip_buckets_ran, buckets);
    const size_t n = (size_t)net->n;
    for (int b = 0; b < buckets; ++b)
        for (size_t r = 0; r < n; ++r) trips[(size_t)b * n + r] = h[(size_t)b * n + (size_t)net->pos_of_row[r]];
    return TRT_OK;
}

int trt_overbank_counts(trt_network* net, int32_t* steps_of_row)
{
    if (!net || !steps_of_row) return fail(TRT_ERR_INVALID, "NULL argument");
    std::vector<int32_t> h;
    const int rc = download_trip_sums(net, h);
    if (rc != TRT_OK) return rc;
    const size_t n = (size_t)net->n, row = (size_t)net->trip_buckets_ran;
    for (size_t r = 0; r < n; ++r) steps_of_row[r] = h[row * n + (size_t)net->pos_of_row[r]];
    return TRT_OK;
}

int trt_last_run_phases(const trt_network* net, double* wide_ms, double* march_ms, int32_t* first_marching_level)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (wide_ms) *wide_ms = net->wide_ms;
    if (march_ms) *march_ms = net->march_ms;
    if (first_marching_level) *first_marching_level = net->deep_level_used;
    return TRT_OK;
}

int trt_march_profile(trt_network* net, int64_t capacity_rows, uint64_t* out4, int64_t* rows)
{
    if (!net || !rows) return fail(TRT_ERR_INVALID, "NULL argument");
    *rows = net->d_march_prof.p ? net->n : 0;
    if (!out4 || !net->d_march_prof.p) return TRT_OK;
    CU(cudaSetDevice(net->device));
    // device order is engine position; hand it back in the caller's row order
    std::vector<unsigned long long> h((size_t)net->n * 4);
    CU(cudaMemcpy(h.data(), net->d_march_prof.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < net->n && r < capacity_rows; ++r)
        for (int c = 0; c < 4; ++c) out4[(size_t)r * 4 + c] = h[(size_t)net->pos_of_row[(size_t)r] * 4 + c];
    return TRT_OK;
}

int trt_last_run_stats(const trt_network* net, double* kernel_ms, int64_t* launches, int64_t* stages, int64_t* lane_steps)
{
    if (!net) return fail(TRT_ERR_INVALID, "NULL network");
    if (kernel_ms) *kernel_ms = net->kernel_ms;
    if (launches) *launches = net->launches;
    if (stages) *stages = net->stages;
    if (lane_steps) *lane_steps = net->lane_steps;
    return TRT_OK;
}

// ---------------------------------------------------------------------------------------------------
// known-answer entry points
// ---------------------------------------------------------------------------------------------------
int trt_mc_segment_batch(int device, int64_t count, const float* in15, float* out6, int32_t* iters)
{
    if (count < 0 || (count > 0 && (!in15 || !out6))) return fail(TRT_ERR_INVALID, "bad arguments");
    if (count == 0) return TRT_OK;
    CU(cudaSetDevice(device));
    DevBuf<float> d_in, d_out;
    DevBuf<int> d_it;
    CU(d_in.reserve((size_t)count * 15));
    CU(d_out.reserve((size_t)count * 6));
    CU(d_it.reserve((size_t)count));
    CU(cudaMemcpy(d_in.p, in15, (size_t)count * 15 * sizeof(float), cudaMemcpyHostToDevice));
    CU(launch_mc_batch(d_in.p, d_out.p, d_it.p, count, 0));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out6, d_out.p, (size_t)count * 6 * sizeof(float), cudaMemcpyDeviceToHost));
    if (iters) CU(cudaMemcpy(iters, d_it.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost));
    return TRT_OK;
}

int trt_levelpool_series(int device, const double* a, int64_t nsteps, const float* inflow, float lateral_inflow,
                         float routing_period, float* outflow_series, float* elevation_series)
{
    if (!a || nsteps < 0 || (nsteps > 0 && (!inflow || !outflow_series || !elevation_series)))
        return fail(TRT_ERR_INVALID, "bad arguments");
    if (nsteps == 0) return TRT_OK;
    CU(cudaSetDevice(device));
    float lp9[9] = {(float)a[0], (float)a[1], (float)a[2], (float)a[3], (float)a[4], (float)a[5], (float)a[6], (float)a[7], 10.0f};
    const float ifd = (float)a[8], we0 = (float)a[10];
    float H = we0;
    if (we0 < -900000000.0f) H = lp9[4] + ((lp9[1] - lp9[4]) * ifd);
    DevBuf<float> d_lp, d_in, d_q, d_h;
    CU(d_lp.reserve(9)); CU(d_in.reserve((size_t)nsteps)); CU(d_q.reserve((size_t)nsteps)); CU(d_h.reserve((size_t)nsteps));
    CU(cudaMemcpy(d_lp.p, lp9, sizeof(lp9), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_in.p, inflow, (size_t)nsteps * sizeof(float), cudaMemcpyHostToDevice));
    CU(launch_levelpool_series(d_lp.p, H, d_in.p, lateral_inflow, routing_period, d_q.p, d_h.p, nsteps, 0));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(outflow_series, d_q.p, (size_t)nsteps * sizeof(float), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(elevation_series, d_h.p, (size_t)nsteps * sizeof(float), cudaMemcpyDeviceToHost));
    return TRT_OK;
}

int trt_powf_batch(int device, int64_t count, const float* x, const float* y, float* out)
{
    if (count < 0 || (count > 0 && (!x || !y || !out))) return fail(TRT_ERR_INVALID, "bad arguments");
    if (count == 0) return TRT_OK;
    CU(cudaSetDevice(device));
    DevBuf<float> d_x, d_y, d_o;
    CU(d_x.reserve((size_t)count)); CU(d_y.reserve((size_t)count)); CU(d_o.reserve((size_t)count));
    CU(cudaMemcpy(d_x.p, x, (size_t)count * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_y.p, y, (size_t)count * sizeof(float), cudaMemcpyHostToDevice));
    CU(launch_powf_batch(d_x.p, d_y.p, d_o.p, count, 0));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, d_o.p, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
    return TRT_OK;
}

int trt_fdiv_batch(int device, int64_t count, const float* a, const float* d, float* out, uint8_t* inside)
{
    if (count < 0 || (count > 0 && (!a || !d || !out || !inside))) return fail(TRT_ERR_INVALID, "bad arguments");
    if (count == 0) return TRT_OK;
    CU(cudaSetDevice(device));
    DevBuf<float> d_a, d_d, d_o;
    DevBuf<unsigned char> d_i;
    CU(d_a.reserve((size_t)count)); CU(d_d.reserve((size_t)count)); CU(d_o.reserve((size_t)count)); CU(d_i.reserve((size_t)count));
    CU(cudaMemcpy(d_a.p, a, (size_t)count * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_d.p, d, (size_t)count * sizeof(float), cudaMemcpyHostToDevice));
    CU(launch_fdiv_batch(d_a.p, d_d.p, d_o.p, d_i.p, count, 0));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, d_o.p, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(inside, d_i.p, (size_t)count, cudaMemcpyDeviceToHost));
    return TRT_OK;
}

int trt_host_alloc(void** ptr, uint64_t bytes)
{
    if (!ptr) return fail(TRT_ERR_INVALID, "NULL argument");
    CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return TRT_OK;
}
int trt_host_free(void* ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return TRT_OK;
}

}  // extern "C"
