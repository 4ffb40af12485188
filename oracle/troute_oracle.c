/*
 * troute_oracle.c -- TEST INFRASTRUCTURE.  CPU oracle for the B200 routing path.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load liboracle.  The product path
 * (t-route_b200/) never links or imports it.
 *
 * It restates, in plain C float32,
 *   - the single-segment Muskingum-Cunge solve   src/kernel/muskingum/MCsingleSegStime_f2py_NOLOOP.f90
 *   - the level-pool RK3 step                    src/kernel/reservoir/Level_Pool/module_levelpool.F:233-427
 *   - streamflow nudging                         src/troute-routing/troute/routing/fast_reach/simple_da.pyx:21-128
 *   - the network time loop                      src/troute-routing/troute/routing/fast_reach/mc_reach.pyx:164-845
 * (paths relative to /root/reference).  The reference itself cannot be built in this image (no
 * Fortran compiler: gfortran/flang/nvfortran/f2c are all absent), so the pin is the reference's own
 * known-answer vectors (tests/test_oracle_kat.py): mc_sseg_stime_NOLOOP_demo.py:173-244, the level
 * pool fixtures of reservoirs/test/test_compute_kernel.py, and routing/test_compute.py:33-42.
 *
 * Two arithmetic builds of every kernel live side by side:
 *   pow_mode 0 "libm": x**y is the platform's powf -- what a gfortran build of the reference does.
 *                      Used for the KAT pin and as the timed CPU baseline.
 *   pow_mode 1 "det" : x**y is trt_powf_det (include/trt_detmath.h), the bit-specified powf the CUDA
 *                      path uses.  GPU parity is bit-for-bit against this build; the distance between
 *                      the two builds is the reference's own sensitivity to its libm and is measured
 *                      in tests/test_oracle_network.py.
 *
 * Build: oracle/Makefile (-O2 -ffp-contract=off, as src/kernel/muskingum/makefile:6 has no FMA / fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/trt_detmath.h"

static const trt_u64 g_tl[2 * TRT_LOG2_TAB_N] = TRT_LOG2_TAB_INIT;
static const trt_u64 g_te[TRT_EXP2_TAB_N] = TRT_EXP2_TAB_INIT;

static inline float powf_det(float x, float y) { return trt_powf_det(x, y, g_tl, g_te); }

#define ORACLE_POWF(x, y) powf((x), (y))
#define ORACLE_SFX(name) name##_libm
#include "mc_kernel.inc"
#undef ORACLE_POWF
#undef ORACLE_SFX

#define ORACLE_POWF(x, y) powf_det((x), (y))
#define ORACLE_SFX(name) name##_det
#include "mc_kernel.inc"
#undef ORACLE_POWF
#undef ORACLE_SFX

/* ------------------------------------------------------------------------------------------ */
/* scalar entry points (KATs)                                                                  */
/* ------------------------------------------------------------------------------------------ */

float oracle_powf_det(float x, float y) { return powf_det(x, y); }

void oracle_powf_det_array(long n, const float* x, const float* y, float* out)
{
    for (long i = 0; i < n; ++i) out[i] = powf_det(x[i], y[i]);
}
void oracle_powf_libm_array(long n, const float* x, const float* y, float* out)
{
    for (long i = 0; i < n; ++i) out[i] = powf(x[i], y[i]);
}

/* c_muskingcungenwm (pyMCsingleSegStime_NoLoop.f90:8-21) through reach.muskingcunge (reach.pyx:7-64).
 * out6 = {qdc, velc, depthc, ck, cn, X}; returns the number of secant iterations. */
int oracle_mc_segment(int pow_mode, float dt, float qup, float quc, float qdp, float ql, float dx, float bw,
                      float tw, float twcc, float n, float ncc, float cs, float s0, float velp, float depthp,
                      float* out6)
{
    int iters = 0;
    if (pow_mode == 0)
        muskingcungenwm_libm(dt, qup, quc, qdp, ql, dx, bw, tw, twcc, n, ncc, cs, s0, velp, depthp,
                             &out6[0], &out6[1], &out6[2], &out6[3], &out6[4], &out6[5], &iters);
    else
        muskingcungenwm_det(dt, qup, quc, qdp, ql, dx, bw, tw, twcc, n, ncc, cs, s0, velp, depthp,
                            &out6[0], &out6[1], &out6[2], &out6[3], &out6[4], &out6[5], &iters);
    return iters;
}

/* batch of independent single-segment solves; in15 rows = (dt,qup,quc,qdp,ql,dx,bw,tw,twcc,n,ncc,cs,s0,velp,depthp) */
void oracle_mc_segment_batch(int pow_mode, long count, const float* in15, float* out6, int* iters)
{
    for (long i = 0; i < count; ++i) {
        const float* a = in15 + 15 * i;
        int it = oracle_mc_segment(pow_mode, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10],
                                   a[11], a[12], a[13], a[14], out6 + 6 * i);
        if (iters) iters[i] = it;
    }
}

/* Level pool state + parameters, as _MC_Levelpool (levelpool_structs.h:7-17) */
typedef struct {
    float dam_length, area, max_depth;
    float orifice_area, orifice_coefficient, orifice_elevation;
    float weir_coefficient, weir_elevation, weir_length;
    float initial_fractional_depth, water_elevation;
} lp_t;

/* MC_Levelpool.__init__ (levelpool.pyx:34-82: argument order :48-57, dam_length = 10 :66) and
 * init_levelpool_reach (levelpool_structs.c:72-120: cold-start elevation :97-103).
 * args = one row of wbody_cols: LkArea, LkMxE, OrificeA, OrificeC, OrificeE, WeirC, WeirE, WeirL, ifd, qd0, h0
 * (compute.py:1416-1430). */
static void lp_init(lp_t* lp, const double* args)
{
    lp->dam_length = 10.0f;
    lp->area = (float)args[0];
    lp->max_depth = (float)args[1];
    lp->orifice_area = (float)args[2];
    lp->orifice_coefficient = (float)args[3];
    lp->orifice_elevation = (float)args[4];
    lp->weir_coefficient = (float)args[5];
    lp->weir_elevation = (float)args[6];
    lp->weir_length = (float)args[7];
    lp->initial_fractional_depth = (float)args[8];
    float water_elevation = (float)args[10];
    if (water_elevation < -900000000.0f) {
        lp->water_elevation = lp->orifice_elevation
            + ((lp->max_depth - lp->orifice_elevation) * lp->initial_fractional_depth);
    } else {
        lp->water_elevation = water_elevation;
    }
}

/* run_lp_c -> route (levelpool_structs.c:148-153) -> run_lp (bind_lp.f90:52-90, passes `inflow` as both
 * previous_timestep_inflow and inflow :71-72) -> run_levelpool_reservoir (module_levelpool.F:162-227). */
static void lp_run(int pow_mode, lp_t* lp, float inflow, float lateral_inflow, float routing_period,
                   float* outflow, float* water_elevation)
{
    float H = lp->water_elevation;
    if (pow_mode == 0)
        levelpool_physics_libm(inflow, inflow, outflow, lateral_inflow, routing_period, &H, lp->area,
                               lp->weir_elevation, lp->max_depth, lp->weir_coefficient, lp->weir_length,
                               lp->dam_length, lp->orifice_elevation, lp->orifice_coefficient, lp->orifice_area);
    else
        levelpool_physics_det(inflow, inflow, outflow, lateral_inflow, routing_period, &H, lp->area,
                              lp->weir_elevation, lp->max_depth, lp->weir_coefficient, lp->weir_length,
                              lp->dam_length, lp->orifice_elevation, lp->orifice_coefficient, lp->orifice_area);
    lp->water_elevation = H;
    *water_elevation = H;
}

/* Run one level pool over an inflow series (the reservoir KATs: test_compute_kernel.py drives
 * MC_Levelpool.run(inflow, 0.0, dt) in a loop).  wbody_row has 11 doubles (see lp_init).
 * out2 = {last outflow, last water elevation}; series outputs optional. */
void oracle_levelpool_series(int pow_mode, const double* wbody_row, long nsteps, const float* inflow,
                             float lateral_inflow, float routing_period, float* out2,
                             float* outflow_series, float* elevation_series)
{
    lp_t lp;
    lp_init(&lp, wbody_row);
    float q = 0.0f, H = lp.water_elevation;
    for (long t = 0; t < nsteps; ++t) {
        lp_run(pow_mode, &lp, inflow[t], lateral_inflow, routing_period, &q, &H);
        if (outflow_series) outflow_series[t] = q;
        if (elevation_series) elevation_series[t] = H;
    }
    out2[0] = q;
    out2[1] = H;
}

/* simple_da.pyx:109-128.  `exp` is libc double exp on float operands promoted to double. */
static float obs_persist_shift_mode(int pow_mode, float last_valid_obs, float model_val, float minutes_since_last_valid,
                                    float decay_coeff)
{
    float da_weight, da_shift, da_weighted_shift;
    const double arg = fabs((double)minutes_since_last_valid) / -(double)decay_coeff;
    /* pow_mode 1: the bit-specified exponential the CUDA path uses (include/trt_detmath.h trt_expf_det) */
    da_weight = pow_mode == 0 ? (float)exp(arg) : trt_expf_det(arg, g_te);
    da_shift = last_valid_obs - model_val;
    da_weighted_shift = da_shift * da_weight;
    return da_weighted_shift;
}

static float obs_persist_shift(float last_valid_obs, float model_val, float minutes_since_last_valid, float decay_coeff)
{
    return obs_persist_shift_mode(0, last_valid_obs, model_val, minutes_since_last_valid, decay_coeff);
}

/* simple_da.pyx:92-107 (python-visible wrapper simple_da_with_decay_py :4-19; KAT routing/test_compute.py:33-42) */
float oracle_simple_da_with_decay(float last_valid_obs, float model_val, float minutes_since_last_valid, float decay_coeff)
{
    return model_val + obs_persist_shift(last_valid_obs, model_val, minutes_since_last_valid, decay_coeff);
}

/* simple_da.pyx:21-89.  out4 = {replacement_val, nudge_val, lastobs_time, lastobs_val} */
static void simple_da(int pow_mode, float timestep, float routing_period, float decay_coeff, float gage_maxtimestep,
                      float target_val, float model_val, float lastobs_time, float lastobs_val, float* out4)
{
    float replacement_val, nudge_val, da_weighted_shift, da_decay_minutes;
    if ((timestep <= gage_maxtimestep) && !isnan(target_val)) {
        replacement_val = target_val;
        nudge_val = target_val - model_val;
        lastobs_time = (timestep)*routing_period;
        lastobs_val = target_val;
    } else if (isnan(target_val) && isnan(lastobs_val)) {
        replacement_val = model_val;
        nudge_val = 0.0f;
        lastobs_val = NAN;
        lastobs_time = NAN;
    } else {
        da_decay_minutes = ((timestep)*routing_period - lastobs_time) / 60;
        da_weighted_shift = obs_persist_shift_mode(pow_mode, lastobs_val, model_val, da_decay_minutes, decay_coeff);
        nudge_val = da_weighted_shift;
        replacement_val = model_val + da_weighted_shift;
    }
    out4[0] = replacement_val; out4[1] = nudge_val; out4[2] = lastobs_time; out4[3] = lastobs_val;
}

/* ------------------------------------------------------------------------------------------ */
/* the network time loop: compute_network_structured, mc_reach.pyx:164-845, on flat arrays      */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int pow_mode;
    int nsteps;
    int qts_subdivisions;
    int assume_short_ts;
    float routing_period;
    int64_t n_rows;
    /* reaches in the caller's (upstream-first) order */
    int64_t n_reaches;
    const int64_t* reach_ptr;    /* [n_reaches+1] into reach_rows */
    const int64_t* reach_rows;   /* row (position in data_idx) of every segment, reach by reach */
    const int32_t* reach_type;   /* 0 = MC reach, 1 = level-pool reservoir (mc_reach.pyx:291, compute.py:41-47) */
    const int64_t* reach_up_ptr; /* [n_reaches+1] into reach_up_rows */
    const int64_t* reach_up_rows;/* rows of upstream_connections[reach[0]] (mc_reach.pyx:288-289), in list order */
    const int32_t* reach_wbody;  /* [n_reaches] row into wbody_cols, -1 for MC reaches */
    /* per-row data */
    const float* data_values;    /* [n_rows, ncols] */
    int ncols;
    const int32_t* scols;        /* [9] columns of dt,dx,bw,tw,twcc,n,ncc,cs,s0 (column_mapper, mc_reach.pyx:150-162) */
    const float* qlat;           /* [n_rows, nqcols] */
    int nqcols;
    /* DA (nudging) */
    int32_t n_gages;
    int32_t gage_maxtimestep;
    const float* usgs_values;            /* [n_gages, gage_maxtimestep] */
    const int32_t* usgs_positions;       /* [n_gages] row of the gage segment */
    const int32_t* reach_has_gage;       /* [n_reaches] gage index or INT32_MIN (mc_reach.pyx:388-401) */
    float da_decay_coefficient;
    float* lastobs_times;                /* [n_gages] in/out */
    float* lastobs_values;               /* [n_gages] in/out */
    float* nudge;                        /* [n_gages, nsteps+1] out */
    /* state / outputs */
    lp_t* lps;                   /* [n_reaches] (only type 1 entries used) */
    float* fvd;                  /* [n_rows, nsteps+1, 3] */
    float* upstream_array;       /* [n_rows, nsteps+1] */
    int64_t* iter_hist;          /* optional [16]: secant iteration histogram, buckets 0..7, 8-15, 16-31, 32-63, 64-127, 128-255, 256-511, 512+ (slot 15 unused) */
} net_t;

/* one timestep of one reach: mc_reach.pyx:493-796 */
static void route_reach_step(const net_t* N, int64_t i, int timestep, float* buf /* scratch [max_reach_len*?] unused */)
{
    (void)buf;
    const int64_t T1 = (int64_t)N->nsteps + 1;
    float* fvd = N->fvd;
#define FVD(row, t, c) fvd[((row) * T1 + (t)) * 3 + (c)]

    /* :496-505 */
    float upstream_flows = 0.0f;
    float previous_upstream_flows = 0.0f;
    for (int64_t u = N->reach_up_ptr[i]; u < N->reach_up_ptr[i + 1]; ++u) {
        int64_t id = N->reach_up_rows[u];
        upstream_flows += FVD(id, timestep, 0);
        previous_upstream_flows += FVD(id, timestep - 1, 0);
    }
    if (N->assume_short_ts) upstream_flows = previous_upstream_flows;

    const int64_t s0i = N->reach_ptr[i], s1i = N->reach_ptr[i + 1];

    if (N->reach_type[i] == 1) {
        /* RESERVOIR_LP, plain level pool: :550-553, :706-710 */
        const int64_t rid = N->reach_rows[s0i];
        float reservoir_outflow = 0.0f, reservoir_water_elevation = 0.0f;
        lp_run(N->pow_mode, &N->lps[i], upstream_flows, 0.0f, N->routing_period, &reservoir_outflow,
               &reservoir_water_elevation);
        FVD(rid, timestep, 0) = reservoir_outflow;
        FVD(rid, timestep, 1) = 0.0f;
        FVD(rid, timestep, 2) = reservoir_water_elevation;
        N->upstream_array[rid * T1 + timestep] = upstream_flows;
    } else {
        /* MC reach: buffer fill :721-735, compute_reach_kernel :70-138, copy out :743-750 */
        float qup = previous_upstream_flows, quc = upstream_flows;
        const int32_t* sc = N->scols;
        for (int64_t k = s0i; k < s1i; ++k) {
            const int64_t row = N->reach_rows[k];
            const float* dv = N->data_values + row * N->ncols;
            const float qlat = N->qlat[row * N->nqcols + (int)((timestep - 1) / N->qts_subdivisions)];
            const float qdp = FVD(row, timestep - 1, 0);
            const float velp = 0.0f;
            const float depthp = FVD(row, timestep - 1, 2);
            float o_q, o_v, o_d;
            int iters = 0;
            if (N->pow_mode == 0)
                muskingcungenwm_libm(dv[sc[0]], qup, quc, qdp, qlat, dv[sc[1]], dv[sc[2]], dv[sc[3]], dv[sc[4]],
                                     dv[sc[5]], dv[sc[6]], dv[sc[7]], dv[sc[8]], velp, depthp, &o_q, &o_v, &o_d,
                                     0, 0, 0, &iters);
            else
                muskingcungenwm_det(dv[sc[0]], qup, quc, qdp, qlat, dv[sc[1]], dv[sc[2]], dv[sc[3]], dv[sc[4]],
                                    dv[sc[5]], dv[sc[6]], dv[sc[7]], dv[sc[8]], velp, depthp, &o_q, &o_v, &o_d,
                                    0, 0, 0, &iters);
            if (N->iter_hist) {
                int b = iters;
                if (iters > 7) { b = 8; for (int v = iters >> 4; v && b < 14; v >>= 1) ++b; }
#ifdef _OPENMP
#pragma omp atomic
#endif
                N->iter_hist[b]++;
            }
            FVD(row, timestep, 0) = o_q;
            FVD(row, timestep, 1) = o_v;
            FVD(row, timestep, 2) = o_d;
            qup = qdp;                                   /* :133 */
            if (N->assume_short_ts) quc = qup;           /* :135-138 */
            else quc = o_q;
        }
    }

    /* streamflow nudging :761-796 */
    if (N->reach_has_gage && N->reach_has_gage[i] > -1) {
        const int32_t gage_i = N->reach_has_gage[i];
        const int64_t pos = N->usgs_positions[gage_i];
        float target = (timestep >= N->gage_maxtimestep)
                           ? NAN
                           : N->usgs_values[(int64_t)gage_i * N->gage_maxtimestep + timestep];
        float da_buf[4];
        simple_da(N->pow_mode, (float)timestep, N->routing_period, N->da_decay_coefficient, (float)N->gage_maxtimestep, target,
                  FVD(pos, timestep, 0), N->lastobs_times[gage_i], N->lastobs_values[gage_i], da_buf);
        FVD(pos, timestep, 0) = da_buf[0];
        N->nudge[(int64_t)gage_i * T1 + timestep] = da_buf[1];
        N->lastobs_times[gage_i] = da_buf[2];
        N->lastobs_values[gage_i] = da_buf[3];
    }
#undef FVD
}

/*
 * oracle_route_network: the reference loop order -- timestep outer, reach inner (mc_reach.pyx:492-493).
 *
 * Set-up restated from mc_reach.pyx:
 *   :253      flowveldepth zero-initialised, shape (n_rows, nsteps+1, 3)  -- the CALLER passes it zeroed,
 *             with any upstream_results rows already injected (:458-469), because injection is host logic;
 *   :359-361  MC reaches:   flowveldepth[rows, 0, :] = initial_conditions[rows, :]
 *   :298      reservoirs:   flowveldepth[row, 0, 0]  = wbody_cols[wb, 9]
 *   :405-411  gage rows:    flowveldepth[pos, 0, 0]  = usgs_values[g, 0] where not NaN
 *
 * Job decomposition (optional): when n_jobs > 0 the reaches are routed the way compute.py's
 * by-subnetwork-jit modes do it (compute.py:909-1209): orders run one after another, the jobs of one
 * order run in parallel (OpenMP threads stand in for joblib-loky workers), and every job runs ALL
 * timesteps of its own reaches; a job only reads rows of earlier orders, which are complete.  With
 * n_jobs == 0 the whole reach list is one serial job.  Returns 0, or a negative error code.
 */
int oracle_route_network(
    int pow_mode, int nsteps, float dt, int qts_subdivisions, int assume_short_ts,
    int64_t n_rows,
    int64_t n_reaches, const int64_t* reach_ptr, const int64_t* reach_rows, const int32_t* reach_type,
    const int64_t* reach_up_ptr, const int64_t* reach_up_rows,
    const int32_t* reach_wbody, const double* wbody_cols /* [n_wb, 11] */,
    const float* data_values, int ncols, const int32_t* scols,
    const float* initial_conditions /* [n_rows, 3] */,
    const float* qlat, int nqcols,
    int32_t n_gages, int32_t gage_maxtimestep, const float* usgs_values, const int32_t* usgs_positions,
    const int32_t* usgs_positions_reach, const int32_t* usgs_positions_gage,
    const float* lastobs_values_init, const float* time_since_lastobs_init, double da_decay_coefficient,
    float* lastobs_times_out, float* lastobs_values_out, float* nudge_out,
    int64_t n_orders, const int64_t* order_ptr /* [n_orders+1] into jobs */,
    const int64_t* job_ptr /* [n_jobs+1] into job_reaches */, const int64_t* job_reaches,
    int nthreads,
    float* flowveldepth, float* upstream_array, int64_t* iter_hist)
{
    if (nsteps < 0 || n_rows < 0 || n_reaches < 0) return -1;
    if ((float)nqcols < (float)nsteps / (float)qts_subdivisions) return -2;   /* mc_reach.pyx:246-247 */

    net_t N;
    memset(&N, 0, sizeof(N));
    N.pow_mode = pow_mode; N.nsteps = nsteps; N.qts_subdivisions = qts_subdivisions;
    N.assume_short_ts = assume_short_ts; N.routing_period = dt; N.n_rows = n_rows;
    N.n_reaches = n_reaches; N.reach_ptr = reach_ptr; N.reach_rows = reach_rows; N.reach_type = reach_type;
    N.reach_up_ptr = reach_up_ptr; N.reach_up_rows = reach_up_rows; N.reach_wbody = reach_wbody;
    N.data_values = data_values; N.ncols = ncols; N.scols = scols; N.qlat = qlat; N.nqcols = nqcols;
    N.fvd = flowveldepth; N.upstream_array = upstream_array; N.iter_hist = iter_hist;

    const int64_t T1 = (int64_t)nsteps + 1;
    N.lps = (lp_t*)calloc((size_t)(n_reaches > 0 ? n_reaches : 1), sizeof(lp_t));
    if (!N.lps) return -3;

    for (int64_t i = 0; i < n_reaches; ++i) {
        if (reach_type[i] == 1) {
            const int64_t row = reach_rows[reach_ptr[i]];
            const double* wb = wbody_cols + 11 * (int64_t)reach_wbody[i];
            flowveldepth[(row * T1 + 0) * 3 + 0] = (float)wb[9];              /* :298 */
            lp_init(&N.lps[i], wb);
        } else {
            for (int64_t k = reach_ptr[i]; k < reach_ptr[i + 1]; ++k) {       /* :359-361 */
                const int64_t row = reach_rows[k];
                flowveldepth[(row * T1 + 0) * 3 + 0] = initial_conditions[row * 3 + 0];
                flowveldepth[(row * T1 + 0) * 3 + 1] = initial_conditions[row * 3 + 1];
                flowveldepth[(row * T1 + 0) * 3 + 2] = initial_conditions[row * 3 + 2];
            }
        }
    }

    int32_t* reach_has_gage = 0;
    if (n_gages > 0) {                                                        /* :380-411 */
        reach_has_gage = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_reaches);
        if (!reach_has_gage) { free(N.lps); return -3; }
        for (int64_t i = 0; i < n_reaches; ++i) reach_has_gage[i] = INT32_MIN;
        for (int32_t g = 0; g < n_gages; ++g) {
            lastobs_values_out[g] = lastobs_values_init[g];
            lastobs_times_out[g] = time_since_lastobs_init[g];
            reach_has_gage[usgs_positions_reach[g]] = usgs_positions_gage[g];
        }
        if (gage_maxtimestep > 0) {
            for (int32_t g = 0; g < n_gages; ++g) {
                float v0 = usgs_values[(int64_t)g * gage_maxtimestep + 0];
                if (!isnan(v0)) flowveldepth[((int64_t)usgs_positions[g] * T1 + 0) * 3 + 0] = v0;
            }
        }
        N.n_gages = n_gages; N.gage_maxtimestep = gage_maxtimestep; N.usgs_values = usgs_values;
        N.usgs_positions = usgs_positions; N.reach_has_gage = reach_has_gage;
        N.da_decay_coefficient = (float)da_decay_coefficient;                 /* simple_da(const float decay_coeff) */
        N.lastobs_times = lastobs_times_out; N.lastobs_values = lastobs_values_out; N.nudge = nudge_out;
    }

    if (n_orders <= 0) {
        /* reference order: while timestep < nsteps+1: for i in range(num_reaches)   :492-493 */
        for (int timestep = 1; timestep < nsteps + 1; ++timestep)
            for (int64_t i = 0; i < n_reaches; ++i) route_reach_step(&N, i, timestep, 0);
    } else {
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#else
        (void)nthreads;
#endif
        for (int64_t o = 0; o < n_orders; ++o) {
            const int64_t j0 = order_ptr[o], j1 = order_ptr[o + 1];
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
            for (int64_t j = j0; j < j1; ++j) {
                for (int timestep = 1; timestep < nsteps + 1; ++timestep)
                    for (int64_t r = job_ptr[j]; r < job_ptr[j + 1]; ++r)
                        route_reach_step(&N, job_reaches[r], timestep, 0);
            }
        }
    }

    free(reach_has_gage);
    free(N.lps);
    return 0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
