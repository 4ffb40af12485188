/*
 * routing_kernels.cu -- sm_100a kernels of the routing path.
 *
 * Replaces the time-outer / reach-inner double loop of compute_network_structured
 * (/root/reference/src/troute-routing/troute/routing/fast_reach/mc_reach.pyx:492-800) and the
 * per-reach segment walk of compute_reach_kernel (:70-138).
 *
 * Schedule.  Segment s at step t needs q[u, t], q[u, t-1] of its upstream neighbours u and its own
 * q[s, t-1], depth[s, t-1] (mc_reach.pyx:499-502, :721-735), nothing else.  With level(s) = longest
 * path from a headwater, all pairs (s, t) with level(s) + t == k are mutually independent, so the
 * network is routed as a wavefront over k = 1 .. L + T - 1 instead of T * L dependent level-steps.
 * Positions are sorted by level, hence the active set of stage k is ONE contiguous position range
 * [lvl_ptr[max(0, k-T)], lvl_ptr[min(L, k)]) and every load below is a unit-stride sweep of an SoA
 * array (the upstream gather is the only indexed access).  With assume_short_ts every segment of a
 * step is independent (quc := qup) and the same kernel runs with L = 1.
 *
 * One lane = one segment-step; 32 consecutive positions per warp.  No tensor cores: the work is
 * ~2k dependent scalar FP32/FP64 instructions per lane, not a contraction.
 *
 * Compile with -fmad=false (see mc_device.cuh).
 */
#include <algorithm>
#include <cooperative_groups.h>
#include "kernels.cuh"
#include "mc_device.cuh"
#include "../../include/troute_b200.h"

namespace cg = cooperative_groups;

namespace trt {

__device__ const trt_u64 g_log2_tab[2 * TRT_LOG2_TAB_N] = TRT_LOG2_TAB_INIT;
__device__ const trt_u64 g_exp2_tab[TRT_EXP2_TAB_N] = TRT_EXP2_TAB_INIT;

constexpr int kBlock = 256;

struct SmemTabs {
    __align__(16) trt_u64 tl[2 * TRT_LOG2_TAB_N];
    trt_u64 te[TRT_EXP2_TAB_N];
};

__device__ __forceinline__ PowTabs stage_tables(SmemTabs& s)
{
    for (int i = threadIdx.x; i < 2 * TRT_LOG2_TAB_N; i += blockDim.x) s.tl[i] = g_log2_tab[i];
    for (int i = threadIdx.x; i < TRT_EXP2_TAB_N; i += blockDim.x) s.te[i] = g_exp2_tab[i];
    __syncthreads();
    PowTabs t; t.tl = s.tl; t.te = s.te;
    return t;
}

// ---- loads / stores of the flow state --------------------------------------------------------------------
// Bulk-synchronous schedules read finished rows with ld.cg.  The dataflow schedule reads slots that another warp
// (or another GPU, over NVLink) may not have written yet: they hold TRT_SENTINEL until the one 4-byte store that
// publishes the value lands in L2, so a volatile (L1-bypassing) poll is the whole synchronisation.
__device__ __forceinline__ unsigned ld_volatile_u32(const float* p)
{
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <bool WAIT>
__device__ __forceinline__ float ld_state(const float* p, int* abort_flag)
{
    if (!WAIT) return __ldcg(p);
    unsigned v = ld_volatile_u32(p);
    if (v == TRT_SENTINEL) {
        unsigned spins = 0;
        do {
            __nanosleep(spins < 16 ? 40 : 400);
            v = ld_volatile_u32(p);
            if ((++spins & 0x3FFF) == 0) {
                // ~6 ms of waiting per check; bail out if somebody flagged an error, or after ~8 s on our own
                if (*reinterpret_cast<volatile int*>(abort_flag) != 0) return __uint_as_float(0x7FC00000u);
                if (spins > (1u << 24)) {
                    // the first lane to give up records WHICH slot never arrived (abort_flag = ctrl + 2, address in ctrl[6..7])
                    if (atomicCAS(abort_flag, 0, 1) == 0)
                        *reinterpret_cast<volatile unsigned long long*>(abort_flag + 4) = (unsigned long long)p;
                    return __uint_as_float(0x7FC00000u);
                }
            }
        } while (v == TRT_SENTINEL);
    }
    return __uint_as_float(v);
}

template <bool WAIT>
__device__ __forceinline__ void st_state(float* p, float x)
{
    if (WAIT) {
        unsigned b = __float_as_uint(x);
        if (b == TRT_SENTINEL) b = 0x7FC00000u;      // a NaN payload that happens to equal the sentinel
        asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(b) : "memory");
    } else {
        *p = x;
    }
}

__device__ __forceinline__ void prefetch_l2(const float* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// q[s, t] becomes visible to every consumer (same GPU: the poll of a downstream lane; other GPU: the import row of the
// downstream shard, written over NVLink peer memory) with ONE 4-byte store each.
__device__ __forceinline__ void publish_flow(float* own, float q, unsigned kflags, int s, int t, size_t T1,
                                             const PeerDev& peers)
{
    unsigned b = __float_as_uint(q);
    if (b == TRT_SENTINEL) b = 0x7FC00000u;      // a NaN payload that happens to equal the sentinel
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(own), "r"(b) : "memory");
    if (kflags & TRT_KIND_EXPORT_FLAG) {
        const int x = __ldg(peers.exp_slot + s);
        const int pr = __ldg(peers.exp_peer + x);
        float* dst = peers.S[pr] + ((size_t)__ldg(peers.exp_pos + x) * T1 + (size_t)t) * 3;
        asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(dst), "r"(b) : "memory");
    }
}

// simple_da (simple_da.pyx:21-89) for gage `g` at step `t`: returns the value that replaces the modelled flow, records the
// nudge and updates the last-observation state.  Float expressions keep the operand order of the Cython source; the decay
// weight is trt_expf_det (include/trt_detmath.h).  Ordering: the lane of step t reads the state the lane of step t - 1
// wrote; that lane fences before it publishes its flow / depth and this one fences after it has seen them.
__device__ __forceinline__ float apply_nudging(const RunDev& run, int s, int t, float model_val, const PowTabs& tabs)
{
    const GageDev& G = run.gage;
    const int g = __ldg(G.slot + s);
    __threadfence();
    float lastobs_time = __uint_as_float(ld_volatile_u32(G.lastobs + 2 * g));
    float lastobs_val = __uint_as_float(ld_volatile_u32(G.lastobs + 2 * g + 1));
    const float timestep = (float)t, gage_maxtimestep = (float)G.gmax;
    const float target_val = (t >= G.gmax) ? __uint_as_float(0x7FC00000u) : __ldg(G.usgs + (size_t)g * G.gmax + t);
    float replacement_val, nudge_val;
    if ((timestep <= gage_maxtimestep) && !(target_val != target_val)) {           // :47-55
        replacement_val = target_val;
        nudge_val = target_val - model_val;
        lastobs_time = (timestep) * G.dt;
        lastobs_val = target_val;
    } else if ((target_val != target_val) && (lastobs_val != lastobs_val)) {       // :58-62
        replacement_val = model_val;
        nudge_val = 0.0f;
        lastobs_val = __uint_as_float(0x7FC00000u);
        lastobs_time = __uint_as_float(0x7FC00000u);
    } else {                                                                       // :66-75, obs_persist_shift :109-128
        const float da_decay_minutes = ((timestep) * G.dt - lastobs_time) / 60;
        const double arg = fabs((double)da_decay_minutes) / -(double)G.decay;
        const float da_weight = trt_expf_det(arg, tabs.te);
        const float da_shift = lastobs_val - model_val;
        const float da_weighted_shift = da_shift * da_weight;
        nudge_val = da_weighted_shift;
        replacement_val = model_val + da_weighted_shift;
    }
    G.nudge[(size_t)g * (run.T + 1) + t] = nudge_val;
    asm volatile("st.volatile.global.f32 [%0], %1;" ::"l"(G.lastobs + 2 * g), "f"(lastobs_time) : "memory");
    asm volatile("st.volatile.global.f32 [%0], %1;" ::"l"(G.lastobs + 2 * g + 1), "f"(lastobs_val) : "memory");
    __threadfence();
    return replacement_val;
}

// route segment `s` (engine position) at step `t`.  sync_mask != 0 (polling schedules only): the lanes named in it -- exactly
// the lanes of the warp that call this function for a routed (non-boundary) segment -- meet at a __syncwarp after their
// inputs have arrived and before the solve.
template <bool WAIT>
__device__ __forceinline__ void route_lane(const NetDev& net, const RunDev& run, int s, int t, const PowTabs& tabs,
                                           const PeerDev* peers, int* abort_flag, unsigned sync_mask = 0)
{
    const unsigned kflags = net.kind[s];
    const unsigned kind = kflags & 0x0F;
    if (kind == TRT_KIND_BOUNDARY) return;            // prescribed rows are never computed

    const size_t n = (size_t)net.n;
    const size_t T1 = (size_t)run.T + 1;
    float* own = run.S + ((size_t)s * T1 + (size_t)t) * 3;   // (q, v, d) of (s, t); own - 3 is (s, t-1)

    // parameters first: they do not depend on anybody's results, so their latency overlaps the waits below
    const float* par = net.par + s;
    const float p0 = __ldg(par + 0 * n), p1 = __ldg(par + 1 * n), p2 = __ldg(par + 2 * n), p3 = __ldg(par + 3 * n),
                p4 = __ldg(par + 4 * n), p5 = __ldg(par + 5 * n), p6 = __ldg(par + 6 * n), p7 = __ldg(par + 7 * n),
                p8 = __ldg(par + 8 * n);
    const int e0 = __ldg(net.up_ptr + s), e1 = __ldg(net.up_ptr + s + 1);

    // upstream gather in reference order: upstream_flows += ..., previous_upstream_flows += ...  (mc_reach.pyx:496-505)
    float quc = 0.0f, qup = 0.0f;
    if (run.short_ts) {
        for (int e = e0; e < e1; ++e) {
            const float* up = run.S + ((size_t)__ldg(net.up_idx + e) * T1 + (size_t)t) * 3;
            qup += ld_state<WAIT>(up - 3, abort_flag);
        }
        quc = qup;
    } else {
        for (int e = e0; e < e1; ++e) {
            const float* up = run.S + ((size_t)__ldg(net.up_idx + e) * T1 + (size_t)t) * 3;
            quc += ld_state<WAIT>(up, abort_flag);
            qup += ld_state<WAIT>(up - 3, abort_flag);
        }
    }
    // depth (MC) / water elevation (level pool) at t-1
    const float statep = ld_state<WAIT>(own - 1, abort_flag);
    float ql = 0.0f, qdp = 0.0f;
    if (kind != TRT_KIND_LEVELPOOL) {
        ql = __ldg(run.qlat_t + (size_t)((t - 1) / run.qts) * n + s);              // :723
        qdp = ld_state<WAIT>(own - 3, abort_flag);                                  // :733
    }
    // Every input is here.  The polls above are spin loops with a sleep in them; the warp does not reliably come back
    // together behind them on its own (ncu, profiles/r01_v5_trip_order: 2.46e7 unit iterations but 3.12e7 executions of the
    // code from here on, at 25 of 32 lanes -- a quarter of the warps run the whole solve in two pieces).
    if (WAIT && sync_mask) __syncwarp(sync_mask);

    float o_q, o_v, o_d;
    bool write_v = true;
    if (kind == TRT_KIND_LEVELPOOL) {
        // run_lp_c(r, upstream_flows, 0.0, routing_period, ...)  mc_reach.pyx:553; results :706-710
        LpParams lp;
        lp.area = p1; lp.max_depth = p2; lp.orifice_area = p3; lp.orifice_coefficient = p4; lp.orifice_elevation = p5;
        lp.weir_coefficient = p6; lp.weir_elevation = p7; lp.weir_length = p8; lp.dam_length = 10.0f;
        float H = statep, outflow;
        trt_levelpool_step(lp, quc, 0.0f, p0, H, outflow, tabs);
        o_q = outflow;
        o_v = quc;      // velocity slot carries the reservoir inflow (upstream_array, :710); finalize writes 0 for v
        o_d = H;
    } else {
        // polling schedules (WAIT): the velocity slot keeps TRT_SENTINEL and the result pass fills it in from the depth
        const McResult r = trt_mc_segment<false, !WAIT>(p0, qup, quc, qdp, ql, p1, p2, p3, p4, p5, p6, p7, p8, statep, tabs);
        o_q = r.qdc; o_v = r.velc; o_d = r.depthc;
        if (run.trip_sum) {
            atomicAdd(run.trip_sum + (size_t)(((t - 1) * run.trip_buckets) / run.T) * n + s, r.iters);
            if (r.over) atomicAdd(run.trip_sum + (size_t)run.trip_buckets * n + s, 1);       // row trip_buckets: over-bank steps
        }
        write_v = !WAIT || !(ql > 0.0f || qup > 0.0f || quc > 0.0f || qdp > 0.0f);   // no-flow branch: v = 0 (:171-178)
    }
    if (kflags & TRT_KIND_GAGE_FLAG) o_q = apply_nudging(run, s, t, o_q, tabs);    // mc_reach.pyx:761-796
    if (write_v) own[1] = o_v;
    st_state<WAIT>(own + 2, o_d);
    st_state<WAIT>(own, o_q);
    if (WAIT && (kflags & TRT_KIND_EXPORT_FLAG)) {
        // this segment drains into another shard: scatter its outflow into that GPU's inflow slot (peer memory)
        const int x = __ldg(peers->exp_slot + s);
        const int pr = __ldg(peers->exp_peer + x);
        float* dst = peers->S[pr] + ((size_t)__ldg(peers->exp_pos + x) * T1 + (size_t)t) * 3;
        unsigned b = __float_as_uint(o_q);
        if (b == TRT_SENTINEL) b = 0x7FC00000u;
        asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(dst), "r"(b) : "memory");
    }
}

__device__ __forceinline__ int lane_step(const NetDev& net, const RunDev& run, int k, int s)
{
    return run.short_ts ? k : k - __ldg(net.level + s);
}

__global__ void __launch_bounds__(kBlock) stage_kernel(NetDev net, RunDev run, int k, int lo, int hi)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    const int s = lo + blockIdx.x * kBlock + threadIdx.x;
    if (s >= hi) return;
    const int t = lane_step(net, run, k, s);
    if (t < 1 || t > run.Tc) return;
    route_lane<false>(net, run, s, t + run.t_off, tabs, nullptr, nullptr);
}

__global__ void __launch_bounds__(kBlock) persistent_kernel(NetDev net, RunDev run, int k_begin, int k_end)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    cg::grid_group grid = cg::this_grid();
    const int L = run.short_ts ? 1 : net.nlevels;
    const int gstride = gridDim.x * kBlock;
    const int gtid = blockIdx.x * kBlock + threadIdx.x;
    for (int k = k_begin; k < k_end; ++k) {
        int lo, hi;
        if (run.short_ts) { lo = 0; hi = net.n; }
        else {
            lo = __ldg(net.lvl_ptr + max(0, k - run.Tc));
            hi = __ldg(net.lvl_ptr + min(L, k));
        }
        for (int s = lo + gtid; s < hi; s += gstride) {
            const int t = lane_step(net, run, k, s);
            if (t >= 1 && t <= run.Tc) route_lane<false>(net, run, s, t + run.t_off, tabs, nullptr, nullptr);
        }
        grid.sync();
    }
}

// ---------------------------------------------------------------------------------------------------------
// dataflow schedule (see kernels.cuh): persistent warps claim units in stage order, lanes wait on their own inputs
// ---------------------------------------------------------------------------------------------------------
// 4 CTAs per SM = 64 registers per thread (60 bytes of spill stores, 32 warps per SM); 3 would give 80 registers, no spills
// and 24 warps -- an A/B for a GPU session: make EXTRA=-DTRT_DATAFLOW_MIN_BLOCKS=3
#ifndef TRT_DATAFLOW_MIN_BLOCKS
#define TRT_DATAFLOW_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(kBlock, TRT_DATAFLOW_MIN_BLOCKS) dataflow_kernel(NetDev net, RunDev run, SchedDev sc, PeerDev peers)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    const int lane = threadIdx.x & 31;
    const unsigned total = (unsigned)__ldg(sc.unit_ptr + sc.nstages);
    const int L = sc.wide_levels;
    const int pos_end = sc.pos_end;
    int cursor = 0;                                   // stage index (k - 1) of this warp's previous unit
    if (sc.stage_time && blockIdx.x == 0 && threadIdx.x == 0) sc.stage_time[0] = globaltimer_ns();
    for (;;) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(sc.claim, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= total) break;
        // stage of unit u: last index i >= cursor with unit_ptr[i] <= u
        int lo_i = cursor, hi_i = sc.nstages;         // invariant: unit_ptr[lo_i] <= u < unit_ptr[hi_i]
        while (hi_i - lo_i > 1) {
            const int mid = (lo_i + hi_i) >> 1;
            if ((unsigned)__ldg(sc.unit_ptr + mid) <= u) lo_i = mid; else hi_i = mid;
        }
        cursor = lo_i;
        const int k = lo_i + 1;
        int lo, hi;
        if (run.short_ts) { lo = 0; hi = pos_end; }
        else {
            lo = __ldg(net.lvl_ptr + max(0, k - run.Tc));
            hi = __ldg(net.lvl_ptr + min(L, k));
        }
        const int shift = __ldg(sc.unit_shift + lo_i);
        const int p0 = lo + (int)((u - (unsigned)__ldg(sc.unit_ptr + lo_i)) << shift);
        const int p1 = min(hi, p0 + (1 << shift));

        // run-ahead gate: do not start polling individual slots before stage k - gate is complete
        const int need = __ldg(sc.gate_stage + lo_i);   // last non-empty stage <= k - gate (0 = none)
        if (need >= 1) {
            if (lane == 0) {
                unsigned spins = 0;
                while (*reinterpret_cast<volatile int*>(sc.frontier) < need) {
                    __nanosleep(200);
                    if ((++spins & 0x3FFF) == 0) {
                        if (*reinterpret_cast<volatile int*>(sc.abort_flag) != 0) break;
                        if (spins > (1u << 25)) { if (atomicCAS(sc.abort_flag, 0, 2) == 0) sc.abort_flag[3] = need; break; }
                    }
                }
            }
            __syncwarp();
        }

        for (int base = p0; base < p1; base += 32) {
            const int s = base + lane;
            bool live = s < p1;
            unsigned mask = 0;
            if (sc.resync) {
                // all 32 lanes are here together: name the ones that will route a segment for the __syncwarp in route_lane
                live = live && (net.kind[s] & 0x0F) != TRT_KIND_BOUNDARY;
                mask = __ballot_sync(0xffffffffu, live);
            }
            if (live) {
                const int t = (run.short_ts ? k : k - __ldg(net.level + s)) + run.t_off;
                route_lane<true>(net, run, s, t, tabs, &peers, sc.abort_flag, mask);
            }
            if (sc.resync) __syncwarp();
        }
        __syncwarp();
        if (lane == 0) {
            // stage bookkeeping for the gate: the warp that finishes the last unit of stage k advances the frontier
            const int units_k = __ldg(sc.unit_ptr + lo_i + 1) - __ldg(sc.unit_ptr + lo_i);
            __threadfence();
            if (atomicAdd(sc.done + lo_i, 1) + 1 == units_k) {
                atomicMax(sc.frontier, k);
                if (sc.stage_time) sc.stage_time[k] = globaltimer_ns();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// marching schedule (see kernels.cuh): a lane owns ONE segment and walks it through all T timesteps
// ---------------------------------------------------------------------------------------------------------
// Where the wavefront is narrow (the deep main stems: a few hundred segments per stage for thousands of stages) the
// run time is the length of the dependency chain times the latency of one link.  A link here costs: one L2 round trip
// (the upstream lane's st.volatile of q[u][t], this lane's ld.volatile poll) plus one secant solve whose channel
// geometry, previous flow and previous depth are already in registers.  Lanes of a warp are independent state
// machines (WAIT for inputs of step t / ITERATE one secant trip / DONE), so a lane stalled on its upstream neighbour
// or in the retry ladder never holds up the other lanes, and lanes that are iterating execute the same instructions
// whatever timestep each of them is at.  Units (<= 32 consecutive positions) are claimed in position order: every
// upstream position is lower, hence claimed earlier and resident or finished -- no deadlock.
enum { MARCH_WAIT = 0, MARCH_ITER = 1, MARCH_DONE = 2 };

#ifndef TRT_MARCH_MIN_BLOCKS
#define TRT_MARCH_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(kBlock, TRT_MARCH_MIN_BLOCKS) march_kernel(NetDev net, RunDev run, MarchDev mk, PeerDev peers)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    const int lane = threadIdx.x & 31;
    const size_t n = (size_t)net.n;
    const int T = run.T;
    const size_t T1 = (size_t)T + 1;
    if (mk.prof && blockIdx.x == 0 && threadIdx.x == 0) atomicMin(mk.t_start, globaltimer_ns());
    int cursor = 0;                               // stage of this warp's previous wide unit
    for (;;) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(mk.claim, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= (unsigned)(mk.n_wide_units + mk.n_units)) break;
        int p, t_first = run.t_off + 1, t_last = run.t_off + run.Tc;
        bool mine;
        if (u < (unsigned)mk.n_wide_units) {
            // stage of wide unit u: last K >= cursor with wide_unit_ptr[K] <= u
            int lo_i = cursor, hi_i = mk.nstages;
            while (hi_i - lo_i > 1) {
                const int mid = (lo_i + hi_i) >> 1;
                if ((unsigned)__ldg(mk.wide_unit_ptr + mid) <= u) lo_i = mid; else hi_i = mid;
            }
            cursor = lo_i;
            const int K = lo_i;
            const int lo = __ldg(net.lvl_ptr + max(0, K - mk.nblocks + 1));
            const int hi = __ldg(net.lvl_ptr + min(mk.wide_levels, K + 1));
            p = lo + (int)((u - (unsigned)__ldg(mk.wide_unit_ptr + K)) << 5) + lane;
            mine = p < hi;
            if (mine) {
                const int b = K - __ldg(net.level + p);           // 0 <= b < nblocks by construction of [lo, hi)
                t_first = run.t_off + b * mk.Tb + 1;
                t_last = min(run.t_off + run.Tc, t_first + mk.Tb - 1);
            }
        } else {
            const unsigned v = u - (unsigned)mk.n_wide_units;
            p = __ldg(mk.unit_start + v) + lane;
            mine = lane < (int)__ldg(mk.unit_cnt + v);
        }
        unsigned long long prof_first = 0, prof_wait = 0, prof_fail = 0;
        long long wait_since = 0;

        int state = MARCH_DONE;
        unsigned kflags = 0, kind = TRT_KIND_BOUNDARY;
        if (mine) { kflags = net.kind[p]; kind = kflags & 0x0F; }
        float* row = run.S;                       // (q, v, d) series of this lane's segment
        float p0 = 0.f, p1 = 1.f, p2 = 1.f, p3 = 1.f, p4 = 0.f, p5 = 1.f, p6 = 0.f, p7 = 1.f, p8 = 1.f;
        int e0 = 0, e1 = 0;
        if (kind != TRT_KIND_BOUNDARY) {          // prescribed rows are never computed
            state = MARCH_WAIT;
            row = run.S + (size_t)p * T1 * 3;
            const float* par = net.par + p;
            p0 = __ldg(par + 0 * n); p1 = __ldg(par + 1 * n); p2 = __ldg(par + 2 * n); p3 = __ldg(par + 3 * n);
            p4 = __ldg(par + 4 * n); p5 = __ldg(par + 5 * n); p6 = __ldg(par + 6 * n); p7 = __ldg(par + 7 * n);
            p8 = __ldg(par + 8 * n);
            e0 = __ldg(net.up_ptr + p); e1 = __ldg(net.up_ptr + p + 1);
        }
        const bool is_lp = kind == TRT_KIND_LEVELPOOL;
        const McChannel c = mc_channel(p0, p1, p2, p3, p4, p5, p6, p7, p8);
        McSolve s;
        s.have0 = false; s.have1 = false;
        int t = t_first;
        float qdp = 0.f, statep = 0.f, upsum_prev = 0.f, ql = 0.f;
        int ql_left = 0;
        unsigned waited = 0;
        int e_cur = e0;                           // next upstream slot to read for the current step
        float psum = 0.0f;                        // flows of the slots before e_cur, summed in order
        if (state == MARCH_WAIT) {
            // State at t_first - 1: the initial condition (t_first == 1; init_state_kernel / init_levelpool_kernel) or
            // the last step of this segment's previous block, written by a unit of the previous stage.  That unit and
            // the units of the upstream segments were claimed before this one, so waiting here cannot deadlock.
            const float* prev = row + (size_t)(t_first - 1) * 3;
            qdp = ld_state<true>(prev, mk.abort_flag);
            statep = ld_state<true>(prev + 2, mk.abort_flag);
            for (int e = e0; e < e1; ++e)         // previous_upstream_flows of step t_first  (mc_reach.pyx:499-502)
                upsum_prev += ld_state<true>(run.S + ((size_t)__ldg(net.up_idx + e) * T1 + (size_t)(t_first - 1)) * 3,
                                             mk.abort_flag);
            if (t_last < t_first) state = MARCH_DONE;
        }

        for (;;) {
            if (state == MARCH_WAIT) {
                // inputs of step t: every upstream flow at step t (t - 1 with assume_short_ts), summed in reference order.
                // Main-stem segments can have dozens of tributaries: the slots are read eight at a time (independent
                // loads, one L2 round trip) and the walk resumes where it stopped, so a poll re-reads only the slot it
                // is waiting for and the flows behind it are fetched after the awaited one has arrived.
                const int ti = run.short_ts ? t - 1 : t;
                bool ok = true;
                while (e_cur < e1) {
                    unsigned v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = min(e_cur + j, e1 - 1);
                        const float* up = run.S + ((size_t)__ldg(net.up_idx + e) * T1 + (size_t)ti) * 3;
                        if (mk.poll_mode == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[j]) : "l"(up) : "memory");
                        else v[j] = ld_volatile_u32(up);
                    }
                    const int m = min(8, e1 - e_cur);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (ok && j < m) {
                            if (v[j] == TRT_SENTINEL) ok = false;
                            else { psum += __uint_as_float(v[j]); ++e_cur; }
                        }
                    }
                    if (!ok) break;
                }
                const float sum = psum;
                if (mk.prof) {
                    if (!ok) ++prof_fail;
                    else wait_since = (long long)globaltimer_ns();      // start of the busy part of this step
                }
                if (ok) {
                    waited = 0;
                    e_cur = e0; psum = 0.0f;
                    if (ql_left == 0) {                                              // :723
                        ql = is_lp ? 0.0f : __ldg(run.qlat_t + (size_t)((t - 1) / run.qts) * n + p);
                        ql_left = run.qts - ((t - 1) % run.qts);
                    }
                    --ql_left;
                    const float quc = sum;
                    const float qup = run.short_ts ? sum : upsum_prev;
                    upsum_prev = sum;
                    mc_begin<true>(s, qup, quc, qdp, ql, statep);
                    state = MARCH_ITER;
                    if (!is_lp && !s.flow) {                                         // :171-178
                        float* own = row + (size_t)t * 3;
                        own[1] = 0.0f;
                        float q = 0.0f;
                        if (kflags & TRT_KIND_GAGE_FLAG) q = apply_nudging(run, p, t, q, tabs);
                        st_state<true>(own + 2, 0.0f);
                        publish_flow(own, q, kflags, p, t, T1, peers);
                        qdp = q; statep = 0.0f;
                        s.have0 = false; s.have1 = false;
                        ++t;
                        state = t > t_last ? MARCH_DONE : MARCH_WAIT;
                    }
                } else if ((++waited & 0xFFF) == 0) {
                    // every 4096 failed polls: somebody flagged an error, or this lane has been starving for seconds
                    if (*reinterpret_cast<volatile int*>(mk.abort_flag) != 0) state = MARCH_DONE;
                    else if (waited >= (1u << 26)) {
                        if (atomicCAS(mk.abort_flag, 0, 3) == 0) {
                            const int ti2 = run.short_ts ? t - 1 : t;
                            *reinterpret_cast<volatile unsigned long long*>(mk.abort_flag + 4) =
                                (unsigned long long)(run.S + ((size_t)__ldg(net.up_idx + e_cur) * T1 + (size_t)ti2) * 3);
                        }
                        state = MARCH_DONE;
                    }
                }
            }
            if (state == MARCH_ITER) {
                float* own = row + (size_t)t * 3;
                if (is_lp) {
                    // run_lp_c(r, upstream_flows, 0.0, routing_period, ...)  mc_reach.pyx:553; results :706-710
                    LpParams lp;
                    lp.area = p1; lp.max_depth = p2; lp.orifice_area = p3; lp.orifice_coefficient = p4;
                    lp.orifice_elevation = p5; lp.weir_coefficient = p6; lp.weir_elevation = p7; lp.weir_length = p8;
                    lp.dam_length = 10.0f;
                    float H = statep, outflow;
                    trt_levelpool_step(lp, s.quc, 0.0f, p0, H, outflow, tabs);
                    if (kflags & TRT_KIND_GAGE_FLAG) outflow = apply_nudging(run, p, t, outflow, tabs);
                    publish_flow(own, outflow, kflags, p, t, T1, peers);
                    own[1] = s.quc;             // reservoir inflow rides in the velocity slot (upstream_array, :710)
                    st_state<true>(own + 2, H);
                    qdp = outflow; statep = H;
                    ++t;
                    state = t > t_last ? MARCH_DONE : MARCH_WAIT;
                } else if (mc_iterate(c, s, tabs)) {
                    float q = mc_outflow(s);
                    if (kflags & TRT_KIND_GAGE_FLAG) q = apply_nudging(run, p, t, q, tabs);
                    publish_flow(own, q, kflags, p, t, T1, peers);   // downstream lanes are waiting for this
                    if (mk.prof) prof_wait += globaltimer_ns() - (unsigned long long)wait_since;
                    st_state<true>(own + 2, s.h);                    // own[1] (velocity): result pass, from this depth
                    if ((t & 1) == 0) {
                        // cold tributary rows (finished long ago, evicted from L2): pull the sectors of the coming steps
                        // in, off the critical path.  One 32-byte sector holds 2.67 steps of (q, v, d).
                        const int tp = min(t + 8, T);      // T: last column of the flow state
                        for (int e = e0; e < e1; ++e)
                            prefetch_l2(run.S + ((size_t)__ldg(net.up_idx + e) * T1 + (size_t)tp) * 3);
                    }
                    qdp = q; statep = s.h;
                    if (mk.prof && t == 1) prof_first = globaltimer_ns();
                    ++t;
                    state = t > t_last ? MARCH_DONE : MARCH_WAIT;
                    if (state == MARCH_WAIT && mk.prepare) mc_prepare(c, s, statep, tabs);   // while the lane would only poll
                }
            }
            const unsigned iterating = __ballot_sync(0xffffffffu, state == MARCH_ITER);
            if (iterating == 0) {
                if (__all_sync(0xffffffffu, state == MARCH_DONE)) {
                    if (mk.prof && mine && u >= (unsigned)mk.n_wide_units) {
                        const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long*>(mk.t_start);
                        unsigned long long* o = mk.prof + (size_t)p * 4;
                        o[0] = prof_first ? prof_first - t0 : 0; o[1] = globaltimer_ns() - t0; o[2] = prof_wait; o[3] = prof_fail;
                    }
                    break;
                }
                // nobody has work: every live lane polls.  Back off a little so that thousands of waiting warps do not
                // crowd the L2 slices the producers are writing to.
                if (mk.poll_sleep < 0) {
                    if (__all_sync(0xffffffffu, state != MARCH_WAIT || waited > 8)) __nanosleep(waited > 64 ? 256 : 32);
                } else if (mk.poll_sleep > 0) __nanosleep(mk.poll_sleep);
            }
        }
    }
}

cudaError_t march_max_grid(int* blocks)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, march_kernel, kBlock, 0);
    if (e != cudaSuccess) return e;
    *blocks = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_march(const NetDev& net, const RunDev& run, const MarchDev& march, const PeerDev& peers,
                         int grid_blocks, cudaStream_t st)
{
    march_kernel<<<grid_blocks, kBlock, 0, st>>>(net, run, march, peers);
    return cudaGetLastError();
}

cudaError_t dataflow_max_grid(int* blocks)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dataflow_kernel, kBlock, 0);
    if (e != cudaSuccess) return e;
    *blocks = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_dataflow(const NetDev& net, const RunDev& run, const SchedDev& sched, const PeerDev& peers,
                            int grid_blocks, cudaStream_t st)
{
    dataflow_kernel<<<grid_blocks, kBlock, 0, st>>>(net, run, sched, peers);
    return cudaGetLastError();
}

cudaError_t launch_stage(const NetDev& net, const RunDev& run, int k, int lo, int hi, cudaStream_t st)
{
    if (hi <= lo) return cudaSuccess;
    const int blocks = (hi - lo + kBlock - 1) / kBlock;
    stage_kernel<<<blocks, kBlock, 0, st>>>(net, run, k, lo, hi);
    return cudaGetLastError();
}

cudaError_t persistent_max_grid(int* blocks)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_kernel, kBlock, 0);
    if (e != cudaSuccess) return e;
    *blocks = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_persistent(const NetDev& net, const RunDev& run, int k_begin, int k_end, int grid_blocks,
                              cudaStream_t st)
{
    NetDev n = net; RunDev r = run;
    void* args[] = {&n, &r, &k_begin, &k_end};
    return cudaLaunchCooperativeKernel((void*)persistent_kernel, dim3(grid_blocks), dim3(kBlock), args, 0, st);
}

// ------------------------------------------------------------------------------------------------
// boundary conversions between the caller's row-major tables and the time-major engine arrays
// ------------------------------------------------------------------------------------------------

// qlat_t[c][pos] = qlat_rows[row_of_pos[pos]][c]
__global__ void gather_qlat_kernel(const float* __restrict__ in, const int* __restrict__ row_of_pos,
                                   float* __restrict__ out, int n, int nq)
{
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const float* src = in + (size_t)row_of_pos[pos] * nq;
    for (int c = 0; c < nq; ++c) out[(size_t)c * n + pos] = __ldg(src + c);
}

// state[pos][0][:] = initial_conditions[row][:]  (flowveldepth_nd[ids, 0] = init_array[ids], mc_reach.pyx:361)
__global__ void init_state_kernel(const float* __restrict__ q0, const int* __restrict__ row_of_pos, float* __restrict__ S,
                                  int n, int T1)
{
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const float* src = q0 + (size_t)row_of_pos[pos] * 3;
    float* dst = S + (size_t)pos * T1 * 3;
    // a NaN whose bits happen to equal TRT_SENTINEL would read as "not yet written": canonicalise it
    for (int c = 0; c < 3; ++c) {
        unsigned b = __float_as_uint(src[c]);
        if (b == TRT_SENTINEL) b = 0x7FC00000u;
        dst[c] = __uint_as_float(b);
    }
}

// reservoirs: flowveldepth[row, 0, 0] = qd0 (mc_reach.pyx:298); the elevation state lives in the depth slot
__global__ void init_levelpool_kernel(const int* __restrict__ lp_pos, const float* __restrict__ qd0,
                                      const float* __restrict__ h0, float* S, int T1, int n_lp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lp) return;
    float* dst = S + (size_t)lp_pos[i] * T1 * 3;
    dst[0] = qd0[i]; dst[1] = 0.0f; dst[2] = h0[i];
}

// overlay the routing period and the 8 reservoir parameters on the 9 parameter slots of the level-pool positions
__global__ void scatter_lp_params_kernel(const int* __restrict__ lp_pos, const float* __restrict__ par9, float* par, int n,
                                         int n_lp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lp * 9) return;
    const int l = i / 9, c = i % 9;
    par[(size_t)c * n + lp_pos[l]] = par9[i];
}

// prescribed rows: flowveldepth[row, t, :] = results[(t-1)*3 + :]  (mc_reach.pyx:462-463) -- one contiguous series
__global__ void fill_boundary_kernel(const int* __restrict__ bnd_pos, const float* __restrict__ bnd_fvd, float* S,
                                     int n_bnd, int T)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long w = 3LL * T;
    if (i >= (long long)n_bnd * w) return;
    const int b = (int)(i / w);
    const long long c = i % w;
    S[((size_t)bnd_pos[b] * (T + 1) + 1) * 3 + c] = bnd_fvd[(size_t)b * w + c];
}

// boundary rows nobody prescribes stay zero for every step (flowveldepth is zero-initialised, mc_reach.pyx:253)
__global__ void fill_zero_rows_kernel(const int* __restrict__ pos, float* S, int count, int T)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long w = 3LL * T;
    if (i >= (long long)count * w) return;
    S[((size_t)pos[i / w] * (T + 1) + 1) * 3 + (i % w)] = 0.0f;
}

// Result in the reference's layout (mc_reach.pyx:807-813): fvd[row][3*(t-1) + c] = state[pos][t][c], t = 1..T.
// The engine already keeps every segment's series contiguous, so this is one 12*T-byte copy per segment -- a
// permutation from level-sorted positions to the caller's rows, one warp per segment.  The polling schedules leave the
// velocity slot of a Muskingum-Cunge step at TRT_SENTINEL: velocity is a function of the final depth and the channel
// alone (:163-169), nobody downstream reads it, and here one warp = one segment evaluates it with uniform parameters and
// no divergence instead of inside the branchy solve.
// Positions [p_begin, p_end); compact_from >= 0: row (p - compact_from) of a compact buffer instead of the caller's row
// (the marching rows of a time-chunked trt_route go home separately, see engine.cu).
__global__ void __launch_bounds__(256) permute_rows_kernel(NetDev net, RunDev run, float* __restrict__ fvd, int p_begin,
                                                           int p_end, int compact_from)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    const int warps_per_block = 256 / 32;
    const int lane = threadIdx.x & 31;
    const size_t n = (size_t)net.n;
    for (long long p = p_begin + (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < p_end;
         p += (long long)gridDim.x * warps_per_block) {
        const float* src = run.S + ((size_t)p * (run.T + 1) + 1) * 3;
        float* dst = fvd + (size_t)(compact_from >= 0 ? (int)p - compact_from : net.row_of_pos[p]) * 3 * run.T;
        const unsigned kind = net.kind[p] & 0x0F;
        const float* par = net.par + p;
        const McChannel c = mc_channel(__ldg(par + 0 * n), __ldg(par + 1 * n), __ldg(par + 2 * n), __ldg(par + 3 * n),
                                       __ldg(par + 4 * n), __ldg(par + 5 * n), __ldg(par + 6 * n), __ldg(par + 7 * n),
                                       __ldg(par + 8 * n));
        for (int t = run.t_off + lane; t < run.t_off + run.Tc; t += 32) {
            const float q = __ldcs(src + 3 * t), d = __ldcs(src + 3 * t + 2);
            float v = __ldcs(src + 3 * t + 1);
            if (kind == TRT_KIND_LEVELPOOL) v = 0.0f;      // flowveldepth[r.id, t, 1] = 0.0  (:708)
            else if (kind == TRT_KIND_MC && __float_as_uint(v) == TRT_SENTINEL) v = mc_velocity(c, d, tabs);
            __stcs(dst + 3 * t, q); __stcs(dst + 3 * t + 1, v); __stcs(dst + 3 * t + 2, d);
        }
    }
}

// upstream_array[row, t] = reservoir inflow (mc_reach.pyx:710), carried in the velocity slot of level-pool rows
__global__ void upstream_out_kernel(const int* __restrict__ lp_pos, const int* __restrict__ row_of_pos,
                                    const float* __restrict__ S, float* __restrict__ up, int n_lp, int T)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_lp * T) return;
    const int l = (int)(i / T), t = (int)(i % T) + 1;
    const int pos = lp_pos[l];
    up[(size_t)row_of_pos[pos] * T + (t - 1)] = S[((size_t)pos * (T + 1) + t) * 3 + 1];
}

// start of a run: last-observation state back to its initial values, nudge cleared, and the initial flow of every gage
// segment with an observation at step 0 replaced by it (mc_reach.pyx:403-411; inactive gages included)
__global__ void reset_gages_kernel(GageDev g, const int* __restrict__ gage_pos, const unsigned char* __restrict__ active,
                                   const float* __restrict__ lastobs_init, float* S, int T)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long w = T + 1;
    if (i < (long long)g.n_gages * w) g.nudge[i] = 0.0f;
    if (i < g.n_gages) {
        g.lastobs[2 * i] = lastobs_init[2 * i];
        g.lastobs[2 * i + 1] = lastobs_init[2 * i + 1];
        if (g.gmax > 0) {
            const float v = g.usgs[(size_t)i * g.gmax];
            if (!(v != v)) S[(size_t)gage_pos[i] * w * 3] = v;
        }
    }
}

__global__ void export_series_kernel(const int* __restrict__ pos, const float* __restrict__ S, float* __restrict__ dst,
                                     int count, int T)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)count * (T + 1)) return;
    const int c = (int)(i / (T + 1)), t = (int)(i % (T + 1));
    dst[i] = S[((size_t)pos[c] * (T + 1) + t) * 3];
}

__global__ void import_series_kernel(const int* __restrict__ pos, const float* __restrict__ src, float* __restrict__ S,
                                     int count, int T)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)count * (T + 1)) return;
    const int c = (int)(i / (T + 1)), t = (int)(i % (T + 1));
    if (t == 0) return;
    S[((size_t)pos[c] * (T + 1) + t) * 3] = src[i];
}

#define TRT_GRID1D(total, block) (unsigned)(((total) + (block)-1) / (block))

cudaError_t launch_gather_qlat(const float* in, const int* row_of_pos, float* out, int n, int nq, cudaStream_t st)
{
    if (n == 0 || nq == 0) return cudaSuccess;
    gather_qlat_kernel<<<TRT_GRID1D(n, 256), 256, 0, st>>>(in, row_of_pos, out, n, nq);
    return cudaGetLastError();
}
cudaError_t launch_init_state(const float* q0, const int* row_of_pos, float* S, int n, int T, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    init_state_kernel<<<TRT_GRID1D(n, 256), 256, 0, st>>>(q0, row_of_pos, S, n, T + 1);
    return cudaGetLastError();
}
cudaError_t launch_init_levelpool(const int* lp_pos, const float* qd0, const float* h0, float* S, int T, int n_lp,
                                  cudaStream_t st)
{
    if (n_lp == 0) return cudaSuccess;
    init_levelpool_kernel<<<TRT_GRID1D(n_lp, 128), 128, 0, st>>>(lp_pos, qd0, h0, S, T + 1, n_lp);
    return cudaGetLastError();
}
cudaError_t launch_scatter_lp_params(const int* lp_pos, const float* par9, float* par, int n, int n_lp, cudaStream_t st)
{
    if (n_lp == 0) return cudaSuccess;
    scatter_lp_params_kernel<<<TRT_GRID1D(n_lp * 9, 128), 128, 0, st>>>(lp_pos, par9, par, n, n_lp);
    return cudaGetLastError();
}
cudaError_t launch_fill_boundary(const int* bnd_pos, const float* bnd_fvd, float* S, int n_bnd, int T, cudaStream_t st)
{
    const long long total = 3LL * n_bnd * T;
    if (total == 0) return cudaSuccess;
    fill_boundary_kernel<<<TRT_GRID1D(total, 256), 256, 0, st>>>(bnd_pos, bnd_fvd, S, n_bnd, T);
    return cudaGetLastError();
}
cudaError_t launch_fill_zero_rows(const int* pos, float* S, int count, int T, cudaStream_t st)
{
    const long long total = 3LL * count * T;
    if (total == 0) return cudaSuccess;
    fill_zero_rows_kernel<<<TRT_GRID1D(total, 256), 256, 0, st>>>(pos, S, count, T);
    return cudaGetLastError();
}
cudaError_t launch_finalize(const NetDev& net, const RunDev& run, float* fvd, cudaStream_t st, int p_begin, int p_end,
                            int compact_from)
{
    if (p_end < 0) p_end = net.n;
    if (p_end <= p_begin || run.Tc == 0) return cudaSuccess;
    const long long rows_per_block = 8;
    const unsigned blocks = (unsigned)std::min<long long>((p_end - p_begin + rows_per_block - 1) / rows_per_block, 148LL * 32);
    permute_rows_kernel<<<blocks, 256, 0, st>>>(net, run, fvd, p_begin, p_end, compact_from);
    return cudaGetLastError();
}
cudaError_t launch_upstream_out(const int* lp_pos, const int* row_of_pos, const float* S, float* up, int n_lp, int T,
                                cudaStream_t st)
{
    const long long total = (long long)n_lp * T;
    if (total == 0) return cudaSuccess;
    upstream_out_kernel<<<TRT_GRID1D(total, 256), 256, 0, st>>>(lp_pos, row_of_pos, S, up, n_lp, T);
    return cudaGetLastError();
}
cudaError_t launch_reset_gages(const GageDev& g, const int* gage_pos, const unsigned char* gage_active,
                               const float* lastobs_init, float* S, int T, cudaStream_t st)
{
    if (g.n_gages == 0) return cudaSuccess;
    const long long total = (long long)g.n_gages * (T + 1);
    reset_gages_kernel<<<TRT_GRID1D(total, 256), 256, 0, st>>>(g, gage_pos, gage_active, lastobs_init, S, T);
    return cudaGetLastError();
}
cudaError_t launch_export_series(const int* pos, const float* S, float* dst, int count, int T, cudaStream_t st)
{
    const long long total = (long long)count * (T + 1);
    if (total == 0) return cudaSuccess;
    export_series_kernel<<<TRT_GRID1D(total, 256), 256, 0, st>>>(pos, S, dst, count, T);
    return cudaGetLastError();
}
cudaError_t launch_import_series(const int* pos, const float* src, float* S, int count, int T, cudaStream_t st)
{
    const long long total = (long long)count * (T + 1);
    if (total == 0) return cudaSuccess;
    import_series_kernel<<<TRT_GRID1D(total, 256), 256, 0, st>>>(pos, src, S, count, T);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// known-answer / numerics-contract kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) mc_batch_kernel(const float* __restrict__ in15, float* __restrict__ out6,
                                                          int* __restrict__ iters, long long count)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= count) return;
    const float* a = in15 + 15 * i;
    // (dt,qup,quc,qdp,ql,dx,bw,tw,twcc,n,ncc,cs,s0,velp,depthp) -- reach.pyx:66-81
    const McResult r = trt_mc_segment<true>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11],
                                            a[12], a[14], tabs);
    float* o = out6 + 6 * i;
    o[0] = r.qdc; o[1] = r.velc; o[2] = r.depthc; o[3] = r.ck; o[4] = r.cn; o[5] = r.X;
    if (iters) iters[i] = r.iters;
}

__global__ void levelpool_series_kernel(const float* __restrict__ lp9, float h0, const float* __restrict__ inflow,
                                        float ql, float dt, float* __restrict__ outflow, float* __restrict__ elev,
                                        long long nsteps)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    LpParams lp;
    lp.area = lp9[0]; lp.max_depth = lp9[1]; lp.orifice_area = lp9[2]; lp.orifice_coefficient = lp9[3];
    lp.orifice_elevation = lp9[4]; lp.weir_coefficient = lp9[5]; lp.weir_elevation = lp9[6]; lp.weir_length = lp9[7];
    lp.dam_length = lp9[8];
    float H = h0;
    for (long long t = 0; t < nsteps; ++t) {
        float q;
        trt_levelpool_step(lp, inflow[t], ql, dt, H, q, tabs);
        outflow[t] = q; elev[t] = H;
    }
}

__global__ void __launch_bounds__(kBlock) powf_batch_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            float* __restrict__ out, long long count)
{
    __shared__ SmemTabs smem;
    const PowTabs tabs = stage_tables(smem);
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i < count) out[i] = dpow(x[i], y[i], tabs);
}

cudaError_t launch_mc_batch(const float* in15, float* out6, int* iters, long long count, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    mc_batch_kernel<<<TRT_GRID1D(count, kBlock), kBlock, 0, st>>>(in15, out6, iters, count);
    return cudaGetLastError();
}
cudaError_t launch_levelpool_series(const float* lp9, float h0, const float* inflow, float ql, float dt, float* outflow,
                                    float* elev, long long nsteps, cudaStream_t st)
{
    levelpool_series_kernel<<<1, 32, 0, st>>>(lp9, h0, inflow, ql, dt, outflow, elev, nsteps);
    return cudaGetLastError();
}
cudaError_t launch_powf_batch(const float* x, const float* y, float* out, long long count, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    powf_batch_kernel<<<TRT_GRID1D(count, kBlock), kBlock, 0, st>>>(x, y, out, count);
    return cudaGetLastError();
}

}  // namespace trt
