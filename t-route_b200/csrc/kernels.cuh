/*
 * kernels.cuh -- device-side view of a routing network and the launchers of routing_kernels.cu.
 *
 * HBM layout (all float32 unless noted; `pos` = level-sorted engine position, n = segments):
 *   par      [9][n]      structure-of-arrays channel geometry: dt, dx, bw, tw, twcc, n, ncc, cs, s0
 *                        (for level-pool rows the same 9 slots hold dt, LkArea, LkMxE, OrificeA, OrificeC,
 *                         OrificeE, WeirC, WeirE, WeirL -- a reservoir has no channel geometry)
 *   kind     [n] u8      TRT_KIND_*
 *   level    [n] i32     wavefront level; positions are sorted by it, lvl_ptr[L+1] delimits the levels
 *   up_ptr   [n+1] i32, up_idx [E] i32   CSR of upstream positions, reference summation order
 *   qlat_t   [nq][n]     lateral inflow, time-major
 *   S        [n][T+1][3] flow / velocity / depth of every position for t = 0..T (t = 0 = initial state), each
 *                        segment's series contiguous.  This is the only state of the model: (s, t) reads (s, t-1) and
 *                        (u, t), (u, t-1) of its upstream neighbours.  A time-major [T+1][n] layout coalesces better in
 *                        the wide headwater levels but makes the ~600 lanes of a deep-mainstem stage (one segment per
 *                        level, each at a different t) touch ~1200 distinct 2 MB pages per stage: measured 95 us per
 *                        stage of TLB misses instead of 12 us (profiles/r01_*).  Position-major keeps such a stage inside
 *                        one or two pages, and it IS the reference's result layout, so the final transpose
 *                        degenerates into a row permutation.
 *   fvd      [n_rows][3T] the reference's result layout (mc_reach.pyx:807-813), caller row order
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace trt {

struct NetDev {
    int n;                     // segments
    int nlevels;               // wavefront levels of the dependent (assume_short_ts = false) schedule
    const int* lvl_ptr;        // [nlevels + 1]
    const int* level;          // [n]
    const int* up_ptr;         // [n + 1]
    const int* up_idx;         // [E]
    const unsigned char* kind; // [n]
    const float* par;          // [9][n]
    const int* row_of_pos;     // [n]
};

// Streamflow nudging (simple_da.pyx:21-128, call site mc_reach.pyx:761-796): the flow of a gage segment is replaced, right
// after it is computed and before anybody downstream reads it, by the observation of that step or -- outside the
// observation window -- by the model value pulled towards the last observation with an exponentially decaying weight.
// State per gage: (time, value) of the last observation; written and read only by the lanes that route the gage segment,
// in timestep order.
#define TRT_KIND_GAGE_FLAG 0x20
struct GageDev {
    int n_gages;                 // 0: no nudging
    int gmax;                    // observation columns per gage (gage_maxtimestep)
    float dt;                    // routing period of the call
    float decay;                 // da_decay_coefficient
    const int* slot;             // [n] gage index of a position whose kind carries TRT_KIND_GAGE_FLAG
    const float* usgs;           // [n_gages][gmax] observations (NaN = missing)
    float* lastobs;              // [n_gages][2] (time, value)
    float* nudge;                // [n_gages][T + 1]
};

struct RunDev {
    GageDev gage;
    int* trip_sum;  // NULL, or [trip_buckets + 1][n]: secant trips of every position, summed over the steps of each of
                    // trip_buckets equal slices of the call ("collect_trips"; step t falls into slice (t-1)*buckets/T);
                    // last row: number of steps the segment ended above bankfull depth
    int trip_buckets;
    int T;          // timesteps of the call (the flow state holds T + 1 columns)
    int t_off;      // this launch routes steps t_off + 1 .. t_off + Tc (a time chunk of the call; whole call: 0, T)
    int Tc;
    int qts;        // qts_subdivisions
    int nq;         // qlat columns
    int short_ts;   // assume_short_ts
    const float* qlat_t;
    float* S;       // flow state [n][T + 1][3] = (q, v, d) of every position for t = 0 .. T (t = 0: initial state)
};

// Dataflow schedule (mode 2).  Work = the stages of the wavefront, cut into units of 32 or 128 consecutive positions
// and claimed IN ORDER from one counter; a unit never waits on a stage barrier, its lanes wait on exactly the values
// they read (a not-yet-written q / d slot holds TRT_SENTINEL).  Because units are claimed in stage order, everything a
// claimed unit waits for has been claimed earlier by a warp that is running, so the earliest unfinished unit always
// progresses: no deadlock, and a lane stuck in the 750-iteration retry ladder delays only its own dependents.
#define TRT_SENTINEL 0xFFFFFFFFu

struct SchedDev {
    int nstages;                      // stages k = 1 .. nstages
    int T;
    int wide_levels;                  // levels [0, wide_levels) are routed by this schedule (the rest march, mode 4);
                                      // 1 with assume_short_ts: every segment of a step is independent
    int pos_end;                      // positions [0, pos_end) belong to this schedule
    const int* unit_ptr;              // [nstages + 1] first unit of stage k is unit_ptr[k - 1]
    const unsigned char* unit_shift;  // [nstages] log2 of the unit width of that stage (5 or 7)
    unsigned int* claim;              // [1] next unit to hand out
    int* done;                        // [nstages] finished units per stage
    int* frontier;                    // [1] highest stage known to be complete (run-ahead gate)
    int* abort_flag;                  // [1] set when a wait timed out
    const int* gate_stage;            // [nstages] stage that must be complete before a unit of stage k starts:
                                      // the last non-empty stage <= k - gate (0 = no wait)
    unsigned long long* stage_time;   // [nstages + 1] or NULL: %globaltimer (ns) when stage k completed, in slot k;
                                      // slot 0 = kernel start ("profile_stages" option)
    int resync;                       // 1: the lanes of a unit meet at a __syncwarp between their input polls and the solve
                                      // ("warp_resync" option, see dataflow_kernel)
};

// cut edges to other shards: lane s with (kind & TRT_KIND_EXPORT_FLAG) stores q also to peer memory
#define TRT_KIND_EXPORT_FLAG 0x10
#define TRT_MAX_PEERS 16
struct PeerDev {
    const int* exp_slot;              // [n] index into exp_peer / exp_pos, valid where the flag is set
    const int* exp_peer;              // [n_exp]
    const long long* exp_pos;         // [n_exp] position in the peer's arrays
    float* S[TRT_MAX_PEERS];          // peer flow-state arrays (mapped peer memory)
};

// Marching schedule (mode 3, and the deep levels of mode 4): units of <= 32 consecutive positions, claimed in position
// order; every lane walks its segment through all T timesteps, waiting on the q slots of its upstream neighbours.
//
// Mode 5 puts the wide shallow levels through the same lanes, in pieces: a wide unit is 32 consecutive positions x one
// BLOCK of Tb consecutive timesteps, and units are handed out in order of stage K = level + block index, the wavefront of
// mode 2 with blocks in place of steps.  Everything a unit reads was produced by a unit of a lower stage (upstream
// segments: lower level, same block; its own previous block), i.e. by a unit claimed earlier.  Compared with one step
// per unit the channel geometry is loaded and pre-processed once per Tb steps, flow and depth of the previous step stay
// in registers, the previous upstream sum is reused as qup, and the lanes of a warp drift apart in time so that a lane
// needing 5 secant trips does not hold up 31 lanes that needed 2.  The wide units come first in the queue, the deep
// marching units (all T steps) after them: the deep lanes start while the last wide stages drain.
struct MarchDev {
    int n_wide_units;                 // units [0, n_wide_units) are wide units
    int wide_levels;                  // levels [0, wide_levels) are routed by wide units
    int nblocks;                      // time blocks per segment = ceil(T / Tb)
    int Tb;                           // timesteps per block
    int nstages;                      // wide stages K = 0 .. nstages - 1 = wide_levels + nblocks - 1
    const int* wide_unit_ptr;         // [nstages + 1] first unit of stage K
    int n_units;                      // deep units follow: unit n_wide_units + i is deep unit i
    const int* unit_start;            // [n_units] first position of the unit
    const unsigned char* unit_cnt;    // [n_units] lanes in use (1..32)
    unsigned int* claim;              // [1] next unit to hand out
    int* abort_flag;                  // [1]
    unsigned long long* prof;         // NULL, or [n][4] per position: ns from kernel start to its first / last step
                                      // done, ns between the inputs of a step arriving and its flow being published (summed over steps), failed polls ("march_profile" option)
    unsigned long long* t_start;      // [1] %globaltimer at kernel start (prof only)
    int prepare;                      // 1: evaluate the first-trip phase A of the next step right after a step is done
    int poll_mode;                    // experiment: 0 ld.volatile, 1 ld.relaxed.gpu, 2 atomicOr(p, 0)
    int poll_sleep;                   // experiment: ns of back-off between polls of an idle warp (-1 = adaptive)
};
cudaError_t march_max_grid(int* blocks);
cudaError_t launch_march(const NetDev& net, const RunDev& run, const MarchDev& march, const PeerDev& peers,
                         int grid_blocks, cudaStream_t st);

// wavefront: stage k routes every (segment s, step t) with level(s) + t == k
cudaError_t launch_stage(const NetDev& net, const RunDev& run, int k, int lo, int hi, cudaStream_t st);
// persistent cooperative kernel over stages [k_begin, k_end)
cudaError_t launch_persistent(const NetDev& net, const RunDev& run, int k_begin, int k_end, int grid_blocks,
                              cudaStream_t st);
// largest co-resident grid (blocks) of the persistent kernel on the current device
cudaError_t persistent_max_grid(int* blocks);
cudaError_t dataflow_max_grid(int* blocks);
cudaError_t launch_dataflow(const NetDev& net, const RunDev& run, const SchedDev& sched, const PeerDev& peers,
                            int grid_blocks, cudaStream_t st);
cudaError_t launch_fill_zero_rows(const int* pos, float* S, int count, int T, cudaStream_t st);

cudaError_t launch_gather_qlat(const float* qlat_rows, const int* row_of_pos, float* qlat_t, int n, int nq,
                               cudaStream_t st);
cudaError_t launch_init_state(const float* q0_rows, const int* row_of_pos, float* S, int n, int T, cudaStream_t st);
cudaError_t launch_init_levelpool(const int* lp_pos, const float* lp_qd0, const float* lp_h0, float* S, int T, int n_lp,
                                  cudaStream_t st);
cudaError_t launch_scatter_lp_params(const int* lp_pos, const float* par9, float* par, int n, int n_lp, cudaStream_t st);
cudaError_t launch_fill_boundary(const int* bnd_pos, const float* bnd_fvd, float* S, int n_bnd, int T, cudaStream_t st);
// result pass over the steps of run's time chunk
// positions [p_begin, p_end) (p_end < 0: all); compact_from >= 0: destination row = position - compact_from
cudaError_t launch_finalize(const NetDev& net, const RunDev& run, float* fvd_rows, cudaStream_t st, int p_begin = 0,
                            int p_end = -1, int compact_from = -1);
cudaError_t launch_upstream_out(const int* lp_pos, const int* row_of_pos, const float* S, float* up_rows, int n_lp, int T,
                                cudaStream_t st);
cudaError_t launch_reset_gages(const GageDev& g, const int* gage_pos, const unsigned char* gage_active,
                               const float* lastobs_init, float* S, int T, cudaStream_t st);
cudaError_t launch_export_series(const int* pos, const float* S, float* dst, int count, int T, cudaStream_t st);
cudaError_t launch_import_series(const int* pos, const float* src, float* S, int count, int T, cudaStream_t st);

cudaError_t launch_mc_batch(const float* in15, float* out6, int* iters, long long count, cudaStream_t st);
cudaError_t launch_levelpool_series(const float* lp9, float h0, const float* inflow, float ql, float dt,
                                    float* outflow, float* elev, long long nsteps, cudaStream_t st);
cudaError_t launch_powf_batch(const float* x, const float* y, float* out, long long count, cudaStream_t st);

}  // namespace trt
