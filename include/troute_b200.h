/*
 * troute_b200.h -- C ABI of the B200 channel-routing engine (libtroute_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path
 *     compute_nhd_routing_v02 (src/troute-routing/troute/routing/compute.py:507-1738)
 *       -> compute_network_structured (src/troute-routing/troute/routing/fast_reach/mc_reach.pyx:164-845)
 *         -> c_muskingcungenwm (src/kernel/muskingum/pyMCsingleSegStime_NoLoop.f90:8-21)
 *         -> run_lp           (src/kernel/reservoir/bind_lp.f90:52-90)
 * (paths relative to /root/reference).  The reference's own FFI for this path is the per-segment
 * `c_muskingcungenwm(float* x21)` / `run_lp(handle, float* x5)` pair declared in
 * fast_reach/fortran_wrappers.pxd:19-40 and reservoirs/levelpool/levelpool_structs.c:8-18; one call
 * per segment per timestep is meaningless for a GPU, so the boundary moves up one level: a network
 * handle holds the flattened river network on the device and one call routes every segment for
 * every timestep.  INTEGRATION.md shows the ctypes stub a maintainer registers in
 * compute.py:_compute_func_map (:21-26).
 *
 * Conventions: plain C, pointers + sizes, no torch types.  Every function returns 0 on success or a
 * negative trt_status; trt_last_error() returns the message of the last failure on the calling
 * thread.  The caller owns every host buffer.  "rows" are positions in the caller's sorted segment
 * index (data_idx of mc_reach.pyx:170); the engine keeps its own level-sorted device order and
 * converts at the boundary.  A handle owns one CUDA stream and is not re-entrant.
 */
#ifndef TROUTE_B200_H
#define TROUTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct trt_network trt_network;

typedef enum {
    TRT_OK = 0,
    TRT_ERR_INVALID = -1,     /* bad argument / shape mismatch (ValueError in mc_reach.pyx:243-250) */
    TRT_ERR_CYCLE = -2,       /* the upstream graph is not a DAG */
    TRT_ERR_CUDA = -3,        /* CUDA runtime failure; message holds cudaGetErrorString */
    TRT_ERR_NOMEM = -4,
    TRT_ERR_STATE = -5        /* call order violated (e.g. download before run) */
} trt_status;

/* segment kinds (reach types of mc_reach.pyx:291 / compute.py:41-47, plus prescribed rows) */
#define TRT_KIND_MC        0  /* Muskingum-Cunge segment */
#define TRT_KIND_LEVELPOOL 1  /* level-pool reservoir (RESERVOIR_LP, plain level pool) */
#define TRT_KIND_BOUNDARY  2  /* flow series prescribed by the caller (upstream_results, mc_reach.pyx:458-469) */

const char* trt_last_error(void);
int trt_version(void);
/* number of visible CUDA devices, or a negative trt_status */
int trt_device_count(void);

/*
 * Build a network on `device`.
 *   n_rows            rows of the caller's segment index
 *   up_ptr, up_rows   CSR: rows whose outflow enters row r are up_rows[up_ptr[r] .. up_ptr[r+1]), in the
 *                     order the reference sums them (mc_reach.pyx:499-502; inside a reach the single
 *                     upstream is the previous segment, mc_reach.pyx:133-138)
 *   kind              [n_rows] TRT_KIND_*
 *   data_values       [n_rows, ncols] float32 parameter table (param_df_sub.values, compute.py:1538)
 *   scols             [9] columns of dt, dx, bw, tw, twcc, n, ncc, cs, s0 in data_values
 *                     (column_mapper, mc_reach.pyx:150-162)
 * The graph is levelled (longest path from the headwaters), segments are renumbered level by level
 * and all arrays are uploaded as structure-of-arrays.
 */
int trt_network_create(int device, int64_t n_rows, const int64_t* up_ptr, const int64_t* up_rows,
                       const uint8_t* kind, const float* data_values, int32_t ncols, const int32_t* scols,
                       trt_network** out);
/* Same, with the wavefront level of every row given by the caller (NULL = compute).  A shard of a larger network
 * passes the levels of the WHOLE network so that every shard walks the same stages; level[row] must exceed the level
 * of each of its upstream rows. */
int trt_network_create_ex(int device, int64_t n_rows, const int64_t* up_ptr, const int64_t* up_rows,
                          const uint8_t* kind, const float* data_values, int32_t ncols, const int32_t* scols,
                          const int32_t* level_of_row, trt_network** out);
/* Same, with a sort key for the order of the segments INSIDE a wavefront level (NULL = caller's row order).  Any order is
 * valid and gives the same results; lanes of a warp run in lockstep, so rows with equal keys should behave alike -- e.g.
 * order_key = trt_trip_counts of a previous call on the same network (secant trip counts repeat from step to step). */
int trt_network_create_ordered(int device, int64_t n_rows, const int64_t* up_ptr, const int64_t* up_rows,
                               const uint8_t* kind, const float* data_values, int32_t ncols, const int32_t* scols,
                               const int32_t* level_of_row, const int32_t* order_key, trt_network** out);
/* option "collect_trips" = 1 before a run: secant trips of every row summed over the steps of that run (rows routed by the
 * marching kernel report 0) */
int trt_trip_counts(trt_network* net, int32_t* trips_of_row /* [n_rows] */);
/* Same, resolved in time: option "trip_buckets" = B (1..64, default 1) before the collecting run sums the trips of step t
 * into slice (t - 1) * B / nsteps.  Segments whose trip counts move together through the storm belong into the same warp;
 * troute_b200.network.order_key_from_trips turns this table into an order_key (tools/trip_order_study.py: 28.0 instead of
 * 25.9 busy lanes out of 32 on the bench network, against 22.4 in the caller's row order). */
int trt_trip_counts_bucketed(trt_network* net, int32_t buckets, int32_t* trips /* [buckets][n_rows] */);
/* Collected by the same run: the number of steps every row ended above its bankfull depth in a compound channel, i.e. took
 * the over-bank branch of the celerity (MCsingleSegStime_f2py_NOLOOP.f90:248-258, one more power) instead of the in-bank
 * one.  A warp whose lanes disagree executes both; flooding lasts for hours, so ordering by this count first leaves 25 %
 * of the warp-steps mixed instead of 55 % (tools/trip_order_study.py). */
int trt_overbank_counts(trt_network* net, int32_t* steps_of_row /* [n_rows] */);
int trt_network_destroy(trt_network* net);

/* topology queries: number of wavefront levels; level of every row; engine position of every row */
int trt_network_num_levels(const trt_network* net, int32_t* out);
int trt_network_get_levels(const trt_network* net, int32_t* level_of_row /* [n_rows] */);
int trt_network_get_positions(const trt_network* net, int32_t* pos_of_row /* [n_rows] */);

/*
 * Level-pool reservoirs.  wbody_cols rows are (LkArea, LkMxE, OrificeA, OrificeC, OrificeE, WeirC,
 * WeirE, WeirL, ifd, qd0, h0) as in compute.py:1416-1430 / levelpool.pyx:48-57; dam length is 10 m
 * (levelpool.pyx:66); h0 < -9e8 selects the cold-start elevation (levelpool_structs.c:97-103).
 * routing_period is the `dt` argument of compute_network_structured (run_lp_c(..., routing_period, ...),
 * mc_reach.pyx:272,553); lake rows of data_values hold NaN and are never read.
 * May be called before every trt_upload_forcing (the reference passes the table on every call).
 */
int trt_network_set_levelpools(trt_network* net, int64_t n_lp, const int64_t* lp_rows, const double* wbody_cols,
                               float routing_period);

/*
 * Streamflow nudging at gages (simple_da.pyx:21-128; set-up mc_reach.pyx:380-411, call site :761-796).
 *   gage_rows    [n_gages] row of every gage segment (usgs_positions)
 *   active       [n_gages] 1 = this gage is the one its reach assimilates (reach_has_gage, :398: the last gage listed
 *                for a reach wins); inactive gages only seed the initial flow with their first observation (:403-411)
 *   usgs_values  [n_gages, gage_maxtimestep] observations per routing step, NaN = missing
 *   lastobs_values_init / time_since_lastobs_init  [n_gages] last observation before the call (NaN = none)
 *   routing_period  the `dt` argument of compute_network_structured
 * An active gage must be the LAST segment of its reach -- T-Route's network builder breaks reaches at gages
 * (nhd_network.py:295-359); the caller checks.  n_gages = 0 switches nudging off.  State is reset at every trt_run.
 * trt_download_gages: nudge [n_gages, nsteps + 1] (column 0 unused, as in the reference), final last-observation
 * (time, value) per gage.
 */
int trt_network_set_gages(trt_network* net, int64_t n_gages, const int64_t* gage_rows, const uint8_t* active,
                          const float* usgs_values, int32_t gage_maxtimestep, const float* lastobs_values_init,
                          const float* time_since_lastobs_init, float da_decay_coefficient, float routing_period);
int trt_download_gages(trt_network* net, float* nudge, float* lastobs_times, float* lastobs_values);

/*
 * One routing call = upload, run, download.  Shapes follow compute_network_structured:
 *   qlat       [n_rows, nqcols] float32, nqcols >= nsteps/qts_subdivisions   (mc_reach.pyx:243-247)
 *   q0         [n_rows, 3] float32 (qu0, qd0, h0): column 0 seeds flow, column 2 seeds depth (:361)
 *   bnd_rows   [n_bnd] rows of kind TRT_KIND_BOUNDARY;  bnd_fvd [n_bnd, nsteps*3] their prescribed
 *              (q, v, d) series for steps 1..nsteps (tmp["results"], mc_reach.pyx:462-463)
 *   fvd_out    [n_rows, nsteps*3] float32 (q, v, d interleaved per step)      (:807-813)
 *   upstream_out [n_rows, nsteps] float32 or NULL: reservoir inflow on level-pool rows, 0 elsewhere
 *              (the reference leaves np.empty garbage on non-reservoir rows, mc_reach.pyx:487)
 * Host pointers may be pageable or pinned; pinned buffers make the copies asynchronous-capable.  qlat and q0 may also be
 * DEVICE pointers (forcing kept resident by the caller): the copies use cudaMemcpyDefault.
 */
int trt_upload_forcing(trt_network* net, int32_t nsteps, int32_t qts_subdivisions, const float* qlat,
                       int32_t nqcols, const float* q0, int64_t n_bnd, const int64_t* bnd_rows,
                       const float* bnd_fvd);
/*
 * The next routing window of the same network WITHOUT a host round trip of the model state: the last column of the
 * finished call (flow and depth of every segment, outflow and water elevation of every reservoir) becomes the initial state
 * on the device, and the gages' last observations are re-based to the new window start (time -= nsteps_prev * dt).
 * Replaces, for a device-resident model, what the reference does on the host between two nwm_route calls:
 * AbstractNetwork.new_q0 + update_waterbody_water_elevation (src/troute-network/troute/AbstractNetwork.py:177-198) and
 * DataAssimilation.new_lastobs (src/troute-network/troute/DataAssimilation.py:1506-1551); loops
 * src/troute-nwm/src/nwm_routing/__main__.py:258-266, src/troute_model.py:250-262 (BMI update_until).
 * Arguments as trt_upload_forcing minus q0; the new window may have a different number of steps.  Follow with trt_run /
 * trt_run_download.  trt_network_update_gage_observations replaces the observation table (usgs_values of the new window)
 * and keeps the last-observation state.  Sharded runs: every shard calls trt_continue, then trt_prepare, barrier, run.
 */
int trt_continue(trt_network* net, int32_t nsteps, int32_t qts_subdivisions, const float* qlat, int32_t nqcols,
                 int64_t n_bnd, const int64_t* bnd_rows, const float* bnd_fvd);
int trt_network_update_gage_observations(trt_network* net, const float* usgs_values, int32_t gage_maxtimestep);
/*
 * 64-bit checksum of the device-resident result of the last run: the SUM over the selected rows of
 * hash(ids[i], bits of result row rows[i]) -- independent of schedule, within-level order and sharding, so the checksums of
 * the shards of one network add up to the checksum of the unsharded run (bench.py `verify`).  rows = NULL: every row;
 * ids = NULL: the row numbers.  tests/helpers.py::result_hash is the same function in numpy.
 */
int trt_result_hash(trt_network* net, int64_t n_sel, const int64_t* rows, const int64_t* ids, uint64_t* out);
/* result rows rows[0 .. n_sel) of the last run -> fvd_out [n_sel, nsteps*3] (host), without moving the whole table */
int trt_download_rows(trt_network* net, int64_t n_sel, const int64_t* rows, float* fvd_out);
/* kernels only; everything is already resident in HBM.  Blocks until the device is done. */
int trt_run(trt_network* net, int32_t assume_short_ts);
/* trt_run without the final wait: enqueue on the handle's stream and return; trt_sync waits and
 * collects the run statistics.  Used with the "stream" option to time on the caller's stream. */
int trt_run_async(trt_network* net, int32_t assume_short_ts);
int trt_sync(trt_network* net);
int trt_download_results(trt_network* net, float* fvd_out, float* upstream_out);
/* the reservoir inflow series of the level pools alone: inflow_out [n_lp, nsteps], level pools in the order of
 * trt_network_set_levelpools -- the non-zero rows of upstream_array (mc_reach.pyx:710, :807-813) without the table of zeros */
int trt_download_levelpool_inflow(trt_network* net, float* inflow_out);
/* (q, v, d) of the LAST timestep of the last run, every row in caller order -> qvd_out [n_rows, 3] (host): what the BMI model
 * reads back after a window (src/troute_model.py:318-330 `_retrieve_last_output`); 12 bytes per segment cross PCIe. */
int trt_download_last_step(trt_network* net, float* qvd_out);
/* trt_run + trt_download_results with the two overlapped: the call is cut into "route_chunks" time chunks and the
 * finished columns of chunk c are copied to the host while chunk c + 1 is computed (see trt_route).  Shards of one
 * network must use the same "route_chunks". */
int trt_run_download(trt_network* net, int32_t assume_short_ts, float* fvd_out, float* upstream_out);
/* trt_upload_forcing + trt_run_download */
int trt_route(trt_network* net, int32_t nsteps, int32_t qts_subdivisions, int32_t assume_short_ts,
              const float* qlat, int32_t nqcols, const float* q0, int64_t n_bnd, const int64_t* bnd_rows,
              const float* bnd_fvd, float* fvd_out, float* upstream_out);

/*
 * Device-side access for multi-GPU hand-off and for checks without a host round trip.
 *   trt_export_flow_series   gathers q[rows, 0..nsteps] into dst (DEVICE pointer, [n, nsteps+1] float32)
 *   trt_import_boundary_flow writes prescribed q[rows, 1..nsteps] from src (DEVICE pointer,
 *                            [n, nsteps+1] float32, column 0 ignored); v and d of those rows stay 0
 *                            (bulk-synchronous schedules, "mode" 0 / 1, only: the polling schedules rebuild the flow
 *                            state at the start of every run)
 *   trt_device_results       device pointer of the [n_rows, nsteps*3] result after trt_run
 *                            (valid until the next upload)
 */
int trt_export_flow_series(trt_network* net, int64_t n, const int64_t* rows, void* dst_device);
int trt_import_boundary_flow(trt_network* net, int64_t n, const int64_t* rows, const void* src_device);
int trt_device_results(trt_network* net, void** fvd_device);

/*
 * Sub-basin sharding across GPUs (one handle per GPU / process).  A cut edge u -> s between shards is a row of kind
 * TRT_KIND_BOUNDARY in the downstream shard ("import") whose flow series is written, value by value as it is computed,
 * by the kernel of the upstream shard straight into the downstream GPU's flow array over NVLink peer memory ("export").
 * This replaces the pickled tail-water series the reference hands from one order of sub-networks to the next
 * (compute.py:882-900 -> mc_reach.pyx:458-469).  No collective is involved: a consumer lane polls the slot it reads.
 *   trt_network_state_ptr   device pointer of this handle's flow state ([tiles of 32 positions][nsteps+1][q|d][32] float32,
 *                           csrc/kernels.cuh), valid after trt_upload_forcing and until a later upload needs a larger array
 *   trt_ipc_get/open/close  CUDA IPC plumbing to map that array into the peer process
 *   trt_network_set_peer    flow array of peer shard `peer` (mapped pointer) and its row count
 *   trt_network_set_exports rows of this shard whose outflow enters peer shard peer[i] at engine position
 *                           peer_pos[i] (= trt_network_get_positions of the peer, for the import row)
 *   trt_network_set_imports rows of kind TRT_KIND_BOUNDARY that a peer writes (all other boundary rows that are not
 *                           prescribed by trt_upload_forcing hold zero)
 *   trt_prepare             reset the flow state for the next run (must complete on ALL shards before ANY shard calls
 *                           trt_run*, because peers write into it); single-GPU callers never need it
 */
int trt_network_state_ptr(trt_network* net, void** q_device);
int trt_ipc_get_handle(void* device_ptr, uint8_t handle[64]);
int trt_ipc_open_handle(int device, const uint8_t handle[64], void** device_ptr);
int trt_ipc_close_handle(void* device_ptr);
int trt_network_set_peer(trt_network* net, int32_t peer, void* peer_q_device, int64_t peer_n_rows);
int trt_network_set_exports(trt_network* net, int64_t count, const int64_t* rows, const int32_t* peer,
                            const int64_t* peer_pos);
int trt_network_set_imports(trt_network* net, int64_t count, const int64_t* rows);
int trt_prepare(trt_network* net);

/* run-time knobs:
 *   "mode"        0 = one launch per wavefront stage, 1 = persistent cooperative kernel with a grid barrier per stage,
 *                 2 = dataflow kernel: units claimed in stage order, lanes wait on the slots they read,
 *                 3 = marching kernel: a lane owns one segment and walks it through every timestep, waiting on the
 *                     flow slots of its upstream neighbours,
 *                 4 = (default) mode 2 for the wide shallow levels, then mode 3 for the deep levels
 *   "deep_level"  mode 4: first level that marches; -1 (default) = as many of the deepest levels as hold at most
 *                 "deep_lanes" segments (default 8192).  Shards of one network must use the SAME
 *                 deep_level (set it explicitly): a dataflow kernel must never wait for a value that another shard
 *                 produces only in its marching kernel
 *   "march_group" segments per marching warp, 1..32; 0 (default) = the smallest power of two for which the resident warps
 *                 get through the marching segments (segments x timesteps / warps links of occupancy) no slower than the
 *                 wave travels down the chain (levels links): 1 for a day at 300 s on the bench network (shortest link
 *                 latency; 16.1 ms against 20.0 ms with 2), 4 for a week routed as one call
 *   "gate"        mode 2 run-ahead bound: a unit of stage k starts once stage k - gate is complete; 0 (default) =
 *                 adaptive: max("gate_min" stages, as many stages as hold "gate_lanes" lanes)
 *   "collect_trips", "trip_buckets"  see trt_trip_counts / trt_trip_counts_bucketed
 *   "grid_blocks" CTAs of the persistent / dataflow kernel (0 = as many as are co-resident)
 *   "route_chunks" time chunks of trt_route / trt_run_download: the results of chunk c go home while chunk c + 1 is routed;
 *                 0 (default) = chosen per call from the width of the network, the number of steps and "host_shards";
 *                 1 = compute everything, then copy
 *   "host_shards" GPUs of this host that route shards of the same call (default 1): they share the host's copy bandwidth
 *   "park_max", "park_min_tiles", "early_max_tiles"  second form of the dataflow kernel (parked stragglers / early
 *                 publication, csrc/routing_kernels.cu: dataflow_park_kernel); defaults 0 / -1 / 0 = off.  Same bits either way
 *   "stream"      adopt a caller-owned cudaStream_t (passed as an integer; 0 = back to the private stream) */
int trt_set_option(trt_network* net, const char* key, int64_t value);
/* "profile_stages" = 1 with "mode" = 0: device time and width (lanes) of every wavefront stage of the last run;
 * entry k describes stage k (entry 0 unused); *count = entries available */
int trt_stage_profile(const trt_network* net, int64_t capacity, float* stage_ms, int64_t* stage_width, int64_t* count);
/* statistics of the last trt_run: device milliseconds of the wavefront kernels, number of kernel
 * launches, wavefront stages, lane-steps executed */
int trt_last_run_stats(const trt_network* net, double* kernel_ms, int64_t* launches, int64_t* stages,
                       int64_t* lane_steps);

/* the two phases of a mode-4 run: device milliseconds of the dataflow kernel (levels below first_marching_level) and of
 * the marching kernel (the levels from there on) */
int trt_last_run_phases(const trt_network* net, double* wide_ms, double* march_ms, int32_t* first_marching_level);
/* "march_profile" = 1: per row of the last run, {ns from kernel start until its first step was done, ns until its last
 * step was done, ns spent solving (inputs arrived -> flow published, summed over the steps), failed polls}; rows that did not march hold zeros.
 * *rows = rows available (0 when profiling was off); out4 may be NULL to query. */
int trt_march_profile(trt_network* net, int64_t capacity_rows, uint64_t* out4, int64_t* rows);

/*
 * Batch of independent single-segment solves on the device: the GPU twin of
 * reach.compute_reach_kernel (fast_reach/reach.pyx:66-103) for known-answer tests.
 *   in15  [count, 15] rows (dt,qup,quc,qdp,ql,dx,bw,tw,twcc,n,ncc,cs,s0,velp,depthp)
 *   out6  [count, 6]  rows (qdc, velc, depthc, ck, cn, X)       iters [count] or NULL
 */
int trt_mc_segment_batch(int device, int64_t count, const float* in15, float* out6, int32_t* iters);
/* level pool over an inflow series on the device (reservoir KATs); out = {outflow, elevation} per step */
int trt_levelpool_series(int device, const double* wbody_row, int64_t nsteps, const float* inflow,
                         float lateral_inflow, float routing_period, float* outflow_series,
                         float* elevation_series);
/* elementwise trt_powf_det on the device (numerics-contract test) */
int trt_powf_batch(int device, int64_t count, const float* x, const float* y, float* out);
/* elementwise fast-path division of the marching lanes (csrc/mc_device.cuh: McDivFast) on the device: out = the quotient the
 * inline sequence returns, inside = 1 where both operands lie in the window in which it must equal the IEEE quotient
 * (numerics-contract test; a / d itself is what every other division of the library compiles to) */
int trt_fdiv_batch(int device, int64_t count, const float* a, const float* d, float* out, uint8_t* inside);

/* pinned host memory helpers for callers that want asynchronous-capable buffers */
int trt_host_alloc(void** ptr, uint64_t bytes);
int trt_host_free(void* ptr);

/* ---------------------------------------------------------------------------------------------------------------------
 * Diffusive-wave mainstem solver (BASELINE.json configs[3]).
 *
 * trt_c_diffnw replaces, argument for argument, the reference's Fortran entry point
 *     void c_diffnw(double* timestep_ar_g, int* nts_ql_g, ... , double* q_ev_g, double* elv_ev_g, double* depth_ev_g)
 *     src/kernel/diffusive/pydiffusive.f90:8-52, declared for Cython in
 *     src/troute-routing/troute/routing/fast_reach/fortran_wrappers.pxd and called from fast_reach/diffusive.pyx:59-103
 * (42 arguments, every one by reference, arrays in Fortran column-major order: node arrays (mxncomp_g, nrch_g), time series
 * with time first, outputs (ntss_ev_g, mxncomp_g, nrch_g)).  The only difference is the int return value (the Fortran
 * subroutine returns nothing): 0 or a negative trt_status with trt_last_error().
 * Both cross-section kinds of the reference are supported: synthetic trapezoid + floodplain (mxnbathy_g = 0, readXsection)
 * and surveyed vertices (mxnbathy_g > 0, readXsection_natural_mann_vertices).  Unsupported inputs fail with TRT_ERR_INVALID
 * instead of computing something else: the refactored-hydrofabric crosswalk (cwnrow_g > 0; the reference's own input
 * builder no longer produces one, diffusive_utils_v02.py:1036-1041).  Array extents are trusted as in the Fortran.
 *
 * trt_diffnw_batch runs n_domains independent tailwater domains in ONE launch sequence, one CTA per domain (the reference
 * loops over them serially, compute.py:1764): argv holds n_domains x 42 pointers, the argument lists of c_diffnw one after
 * the other.  trt_diffusive_set_device selects the CUDA device (default 0); trt_diffusive_last_run returns the device time
 * of the table kernels and of the time-loop kernel of the last call (ms) and the number of kernel launches. */
int trt_c_diffnw(const double* timestep_ar_g, const int* nts_ql_g, const int* nts_ub_g, const int* nts_db_g,
                 const int* ntss_ev_g, const int* nts_qtrib_g, const int* nts_da_g, const int* mxncomp_g, const int* nrch_g,
                 const double* z_ar_g, const double* bo_ar_g, const double* traps_ar_g, const double* tw_ar_g,
                 const double* twcc_ar_g, const double* mann_ar_g, const double* manncc_ar_g, double* so_ar_g,
                 const double* dx_ar_g, const double* iniq, const int* frnw_col, const int* frnw_ar_g, const double* qlat_g,
                 const double* ubcd_g, const double* dbcd_g, const double* qtrib_g, const int* paradim,
                 const double* para_ar_g, const int* mxnbathy_g, const double* x_bathy_g, const double* z_bathy_g,
                 const double* mann_bathy_g, const int* size_bathy_g, const double* usgs_da_g, const int* usgs_da_reach_g,
                 const double* rdx_ar_g, const int* cwnrow_g, const int* cwncol_g, const double* crosswalk_g,
                 const double* z_thalweg_g, double* q_ev_g, double* elv_ev_g, double* depth_ev_g);
int trt_diffnw_batch(int n_domains, const void* const* argv);
int trt_diffusive_set_device(int device);
int trt_diffusive_last_run(double* table_ms, double* loop_ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* TROUTE_B200_H */
