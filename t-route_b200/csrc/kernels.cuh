/*
 * kernels.cuh -- device-side view of a routing network and the launchers of routing_kernels.cu.
 *
 * HBM layout (`pos` = level-sorted engine position, n = segments, a TILE = 32 consecutive positions, n_tiles = ceil(n/32)):
 *
 *   rec      [n_tiles][16][32] 32-bit words   the STATIC record of a tile, 2 KB, lane-interleaved (word w of position p sits
 *                        at rec[(p >> 5) * 512 + w * 32 + (p & 31)]).  One `cp.async.bulk` (TMA bulk copy, completion on an
 *                        mbarrier) brings everything a warp needs to know about its 32 segments into shared memory, one
 *                        work item ahead of the solve:
 *                          0..8   dt, dx, bw, tw, twcc, n, ncc, cs, s0   (level-pool rows: dt, LkArea, LkMxE, OrificeA,
 *                                 OrificeC, OrificeE, WeirC, WeirE, WeirL -- a reservoir has no channel geometry)
 *                          9      wavefront level
 *                          10     kind | TRT_KIND_EXPORT_FLAG | TRT_KIND_GAGE_FLAG | (number of upstream positions << 8)
 *                          11,12  the first two upstream positions (reference summation order; 97 % of NHD segments have
 *                                 at most two), -1 = none
 *                          13     offset of this position's upstream list in up_idx (used when there are more than two)
 *                          14     gage index (TRT_KIND_GAGE_FLAG)      15  export slot (TRT_KIND_EXPORT_FLAG)
 *   up_idx   [E] i32     CSR of upstream positions, reference summation order (mc_reach.pyx:499-502)
 *   lvl_ptr  [L+1] i32   positions are sorted by level
 *   qlat_t   [nq][n]     lateral inflow, time-major
 *   S        [n_tiles][T+1][2][32] f32   flow q (plane 0) and depth d (plane 1) of every position for t = 0..T (t = 0 =
 *                        initial state): THE model state -- (s, t) reads (s, t-1) and (u, t), (u, t-1) of its upstream
 *                        neighbours.  A warp routes the 32 positions of a tile at one timestep, so its own-state read and
 *                        its result write are two full 128-byte lines each (no partial sectors, nothing to merge), and a
 *                        segment's series stays inside one 2 MB page (the T+1 lines of a tile are contiguous: 74 KB for a
 *                        day at 300 s), which is what the marching lanes of the deep main stem need (a time-major [T+1][n]
 *                        array costs them a TLB miss per step: measured 95 us instead of 12 us per stage in round 1).
 *                        A not-yet-written slot holds TRT_SENTINEL; publishing a value is ONE 4-byte (lane) / 128-byte
 *                        (warp) store.  Velocity is NOT state: nobody downstream reads it, the result pass evaluates it.
 *   fmask    [n_tiles][T+1] u32   bit (p & 31) set when the Muskingum-Cunge solve of (p, t) took the flow branch, i.e.
 *                        velocity = f(depth) (MCsingleSegStime_f2py_NOLOOP.f90:163-169); clear = the no-flow branch (:171-178), v = 0
 *   lp_in    [n_lp][T+1] reservoir inflow of every level pool (upstream_array, mc_reach.pyx:710)
 *   fvd      [n_rows][3T] the reference's result layout (mc_reach.pyx:807-813), caller row order
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace trt {

enum { R_PAR = 0, R_LEVEL = 9, R_FLAGS = 10, R_UP0 = 11, R_UP1 = 12, R_ESTART = 13, R_GAGE = 14, R_EXP = 15,
       R_WORDS = 16, R_TILE_WORDS = 16 * 32 };

// index of q[pos, t] in S; depth is 32 floats further on, the previous timestep 64 floats back
__host__ __device__ __forceinline__ size_t s_idx(long long pos, int t, int T1)
{
    return (((size_t)(pos >> 5) * (size_t)T1 + (size_t)t) << 6) + (size_t)(pos & 31);
}
__host__ __device__ __forceinline__ size_t rec_idx(long long pos, int word)
{
    return (size_t)(pos >> 5) * R_TILE_WORDS + (size_t)word * 32 + (size_t)(pos & 31);
}

struct NetDev {
    int n;                     // segments
    int nlevels;               // wavefront levels of the dependent (assume_short_ts = false) schedule
    const int* lvl_ptr;        // [nlevels + 1]
    const int* up_idx;         // [E]
    const unsigned* rec;       // [n_tiles][16][32]
    const int* row_of_pos;     // [n]
    const int* lp_slot;        // [n] level-pool index of a level-pool position (NULL without reservoirs)
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
    float* S;       // flow state, see the file header
    unsigned* fmask;
    float* lp_in;
};

// Dataflow schedule (mode 2).  Work = the stages of the wavefront, cut into units of 1, 2 or 4 tiles and claimed IN ORDER
// from one counter; a unit never waits on a stage barrier, its lanes wait on exactly the values they read (a
// not-yet-written q / d slot holds TRT_SENTINEL).  Because units are claimed in stage order, everything a claimed unit
// waits for has been claimed earlier by a warp that is running (or that is finishing the one unit it claimed before),
// so the earliest unfinished unit always progresses: no deadlock, and a lane stuck in the 750-iteration retry ladder
// delays only its own dependents.
#define TRT_SENTINEL 0xFFFFFFFFu

struct SchedDev {
    int nstages;                      // stages k = 1 .. nstages
    int T;
    int wide_levels;                  // levels [0, wide_levels) are routed by this schedule (the rest march, mode 4);
                                      // 1 with assume_short_ts: every segment of a step is independent
    int pos_end;                      // positions [0, pos_end) belong to this schedule
    const int* unit_ptr;              // [nstages + 1] first unit of stage k is unit_ptr[k - 1]
    const unsigned char* unit_shift;  // [nstages] log2 of the tiles per unit of that stage (0, 1 or 2)
    unsigned int* claim;              // [1] next unit to hand out
    int* done;                        // [nstages] finished units per stage
    int* frontier;                    // [1] highest stage known to be complete (run-ahead gate)
    int* abort_flag;                  // [1] set when a wait timed out
    const int* gate_stage;            // [nstages] stage that must be complete before a unit of stage k starts:
                                      // the last non-empty stage <= k - gate (0 = no wait)
    unsigned long long* stage_time;   // [nstages + 1] or NULL: %globaltimer (ns) when stage k completed, in slot k;
                                      // slot 0 = kernel start ("profile_stages" option)
    // dataflow_park_kernel (routing_kernels.cu): parked stragglers and early publication
    unsigned* park_pool;              // [warps of the grid][TRT_PARK_WORDS][TRT_PARK_SLOTS] contexts of parked solves (L2-resident)
    int park_max;                     // park the unfinished solves of a tile once at most this many are left (0 = never)
    int park_min_tiles;               // ... in stages at least this many tiles wide (throughput regime)
    int early_max_tiles;              // stages at most this many tiles wide: every lane publishes when its own solve ends
};
#define TRT_PARK_WORDS 40
#define TRT_PARK_SLOTS 64

// cut edges to other shards: lane s with (kind & TRT_KIND_EXPORT_FLAG) stores q also to peer memory
#define TRT_KIND_EXPORT_FLAG 0x10
#define TRT_MAX_PEERS 16
struct PeerDev {
    const int* exp_peer;              // [n_exp]
    const long long* exp_pos;         // [n_exp] position in the peer's arrays
    float* S[TRT_MAX_PEERS];          // peer flow-state arrays (mapped peer memory, same layout and T as ours)
};

// Marching schedule (mode 3, and the deep levels of mode 4): units of <= 32 consecutive positions, claimed in position
// order; every lane walks its segment through all T timesteps, waiting on the q slots of its upstream neighbours.
struct MarchDev {
    int n_units;
    const int* unit_start;            // [n_units] first position of the unit
    const unsigned char* unit_cnt;    // [n_units] lanes in use (1..32)
    unsigned int* claim;              // [1] next unit to hand out
    int* abort_flag;                  // [1]
    unsigned long long* prof;         // NULL, or [n][4] per position: ns from kernel start to its first / last step
                                      // done, ns between the inputs of a step arriving and its flow being published (summed over steps), failed polls ("march_profile" option)
    unsigned long long* t_start;      // [1] %globaltimer at kernel start (prof only)
    int prepare;                      // 1: evaluate the first-trip phase A of the next step right after a step is done
    int poll_sleep;                   // ns of back-off between polls of an idle warp (-1 = adaptive)
};

// force the module of every kernel of the library onto the current device (see routing_kernels.cu)
cudaError_t preload_routing_kernels();

cudaError_t march_max_grid(int* blocks);
// block_threads: 0 = the kernel's 256; 128 = the small CTA that fits beside three dataflow CTAs per SM ("overlap_march")
cudaError_t launch_march(const NetDev& net, const RunDev& run, const MarchDev& march, const PeerDev& peers,
                         int grid_blocks, cudaStream_t st, int block_threads = 0);

// wavefront: stage k routes every (segment s, step t) with level(s) + t == k
cudaError_t launch_stage(const NetDev& net, const RunDev& run, int k, int lo, int hi, cudaStream_t st);
// persistent cooperative kernel over stages [k_begin, k_end)
cudaError_t launch_persistent(const NetDev& net, const RunDev& run, int k_begin, int k_end, int grid_blocks,
                              cudaStream_t st);
// largest co-resident grid (blocks) of the persistent kernel on the current device
cudaError_t persistent_max_grid(int* blocks);
cudaError_t dataflow_max_grid(int* blocks);
// sched.park_pool != NULL selects dataflow_park_kernel (parked stragglers / early publication)
cudaError_t launch_dataflow(const NetDev& net, const RunDev& run, const SchedDev& sched, const PeerDev& peers,
                            int grid_blocks, cudaStream_t st);
cudaError_t launch_fill_zero_rows(const int* pos, float* S, int count, int T, cudaStream_t st);

cudaError_t launch_gather_qlat(const float* qlat_rows, const int* row_of_pos, float* qlat_t, int n, int nq,
                               cudaStream_t st);
cudaError_t launch_init_state(const float* q0_rows, const int* row_of_pos, float* S, int n, int T, cudaStream_t st);
cudaError_t launch_init_levelpool(const int* lp_pos, const float* lp_qd0, const float* lp_h0, float* S, int T, int n_lp,
                                  cudaStream_t st);
cudaError_t launch_scatter_lp_params(const int* lp_pos, const float* par9, unsigned* rec, int n_lp, cudaStream_t st);
cudaError_t launch_fill_boundary(const int* bnd_pos, const float* bnd_fvd, float* S, int n_bnd, int T, cudaStream_t st);
// Device-resident hand-off between consecutive routing windows (trt_continue): column t of the flow state <-> a compact
// [n_tiles][2][32] buffer.  The last column of one window becomes column 0 of the next (whose T may differ).
cudaError_t launch_column_copy(float* S, int T1, int t, float* col, int n_tiles, int to_state, cudaStream_t st);
// last-observation state of the gages after a window -> initial state of the next: times shifted by -(nsteps * dt), as
// compute_network_structured returns them (mc_reach.pyx:822-836)
cudaError_t launch_carry_gages(const float* lastobs, float* lastobs_init, int n_gages, float shift, cudaStream_t st);
// result pass over the steps of run's time chunk
// positions [p_begin, p_end) (p_end < 0: all); compact_from >= 0: destination row = position - compact_from
cudaError_t launch_finalize(const NetDev& net, const RunDev& run, float* fvd_rows, const int* bnd_pos, const float* bnd_fvd,
                            int n_bnd, cudaStream_t st, int p_begin = 0, int p_end = -1, int compact_from = -1);
cudaError_t launch_upstream_out(const int* lp_pos, const int* row_of_pos, const float* lp_in, float* up_rows, int n_lp, int T,
                                cudaStream_t st);
cudaError_t launch_reset_gages(const GageDev& g, const int* gage_pos, const unsigned char* gage_active,
                               const float* lastobs_init, float* S, int T, cudaStream_t st);
cudaError_t launch_export_series(const int* pos, const float* S, float* dst, int count, int T, cudaStream_t st);
cudaError_t launch_import_series(const int* pos, const float* src, float* S, int count, int T, cudaStream_t st);
// 64-bit checksum of a result table in row order: sum over rows of hash(row, bits of the row) -- independent of how the
// rows were sharded or scheduled (bench.py `verify`)
// rows: result rows to hash (NULL = 0 .. n_rows - 1); row_ids: the id each of them is hashed under (NULL = the row number)
cudaError_t launch_hash_rows(const float* fvd, const long long* rows, const long long* row_ids, long long n_rows,
                             long long row_len, unsigned long long* out, cudaStream_t st);

cudaError_t launch_mc_batch(const float* in15, float* out6, int* iters, long long count, cudaStream_t st);
cudaError_t launch_levelpool_series(const float* lp9, float h0, const float* inflow, float ql, float dt,
                                    float* outflow, float* elev, long long nsteps, cudaStream_t st);
cudaError_t launch_powf_batch(const float* x, const float* y, float* out, long long count, cudaStream_t st);
cudaError_t launch_fdiv_batch(const float* a, const float* d, float* out, unsigned char* inside, long long count, cudaStream_t st);

}  // namespace trt
