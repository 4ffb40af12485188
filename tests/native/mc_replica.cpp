/*
 * mc_replica.cpp -- TEST INFRASTRUCTURE.  The product's per-lane Muskingum-Cunge / level-pool physics
 * (t-route_b200/csrc/mc_device.cuh) compiled for the HOST with g++ (tests/native/shim/cuda_runtime.h stands in for the CUDA
 * header), so that a change to the device source can be checked against oracle/mc_kernel.inc bit for bit before any GPU time
 * is spent (tests/test_mc_replica.py).  Never linked into libtroute_b200.so.
 */
#include "../../t-route_b200/csrc/mc_device.cuh"

static const trt_u64 k_log2_tab[2 * TRT_LOG2_TAB_N] = TRT_LOG2_TAB_INIT;
static const trt_u64 k_exp2_tab[TRT_EXP2_TAB_N] = TRT_EXP2_TAB_INIT;

extern "C" void trt_replica_mc_segment_batch(long n, const float* in15, float* out6, int* iters, int resumable)
{
    trt::PowTabs T; T.tl = k_log2_tab; T.te = k_exp2_tab;
    for (long i = 0; i < n; ++i) {
        const float* a = in15 + 15 * i;   /* dt,qup,quc,qdp,ql,dx,bw,tw,twcc,n,ncc,cs,s0,velp,depthp */
        trt::McResult r;
        if (resumable == 0) {
            r = trt::trt_mc_segment<true, true>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[14], T);
        } else if (resumable == 2) {
            /* the dataflow kernel's form: channel and inflows read through the shared-memory views (here: plain arrays with
             * the words 32 floats apart, lane 5 of a tile) */
            static float rec[16 * 32], dv[trt::MC_DV_WORDS * 32], inw[trt::MC_IN_WORDS * 32];
            const int lane = 5;
            const int map[9] = {0, 5, 6, 7, 8, 9, 10, 11, 12};   /* record words dt dx bw tw twcc n ncc cs s0 <- row columns */
            for (int w = 0; w < 9; ++w) rec[w * 32 + lane] = a[map[w]];
            inw[trt::MC_IN_QUP * 32 + lane] = a[1]; inw[trt::MC_IN_QUC * 32 + lane] = a[2];
            inw[trt::MC_IN_QDP * 32 + lane] = a[3]; inw[trt::MC_IN_QL * 32 + lane] = a[4];
            const trt::McChannelSm c = trt::mc_channel_to_shared(rec + lane, dv + lane);
            trt::McInSm in; in.p = inw + lane;
            r = trt::trt_mc_solve<true, true>(c, in, a[14], T);
        } else if (resumable == 3) {
            /* the marching kernel's decomposition with the fast-path division (McDivFast): prepare, begin, one trip per call;
             * a trip (or the prepare) that met a division outside the window restarts the step with IEEE divisions, exactly
             * as march_kernel does.  out6[3] reports whether the step stayed on the fast path. */
            const trt::McChannel c = trt::mc_channel(a[0], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]);
            trt::McSolve s;
            s.have0 = false; s.have1 = false;
            trt::McDivFast fd;
            trt::mc_prepare(c, s, a[14], T, fd);
            trt::McIn in; in.qup_ = a[1]; in.quc_ = a[2]; in.qdp_ = a[3]; in.ql_ = a[4];
            bool slow = !fd.good();
            if (slow) trt::mc_begin<false>(s, in, a[14]); else trt::mc_begin<true>(s, in, a[14]);
            r.ck = r.cn = 0.0f;
            if (s.flow) {
                for (;;) {
                    bool done;
                    if (slow) done = trt::mc_iterate(c, in, s, T);
                    else {
                        trt::McDivFast f2;
                        done = trt::mc_iterate(c, in, s, T, f2);
                        if (!f2.good()) { slow = true; trt::mc_begin<false>(s, in, a[14]); continue; }
                    }
                    if (done) break;
                }
                r.qdc = trt::mc_outflow(s, in); r.depthc = s.h; r.velc = trt::mc_velocity(c, s.h, T); r.X = s.k.X;
            } else { r.qdc = r.velc = r.depthc = r.X = 0.0f; }
            r.ck = slow ? 0.0f : 1.0f;
            r.iters = trt::mc_total_trips(s);
        } else {
            /* the marching kernel's decomposition: prepare (phase A of the first trip from the previous depth), begin,
             * one trip per call, outflow, velocity from the final depth */
            const trt::McChannel c = trt::mc_channel(a[0], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]);
            trt::McSolve s;
            s.have0 = false; s.have1 = false;
            trt::mc_prepare(c, s, a[14], T);
            trt::McIn in; in.qup_ = a[1]; in.quc_ = a[2]; in.qdp_ = a[3]; in.ql_ = a[4];
            trt::mc_begin<true>(s, in, a[14]);
            r.ck = r.cn = 0.0f;
            if (s.flow) {
                while (!trt::mc_iterate(c, in, s, T)) {}
                r.qdc = trt::mc_outflow(s, in); r.depthc = s.h; r.velc = trt::mc_velocity(c, s.h, T); r.X = s.k.X;
            } else { r.qdc = r.velc = r.depthc = r.X = 0.0f; }
            r.iters = trt::mc_total_trips(s);
        }
        float* o = out6 + 6 * i;
        o[0] = r.qdc; o[1] = r.velc; o[2] = r.depthc; o[3] = r.ck; o[4] = r.cn; o[5] = r.X;
        if (iters) iters[i] = r.iters;
    }
}

/* trt_div_fastpath against the IEEE quotient on n (a, d) pairs; returns the number of pairs inside the window whose bits differ
 * and stores how many pairs were inside the window */
extern "C" long trt_replica_fdiv_check(long n, const float* a, const float* d, long* inside)
{
    long bad = 0, in = 0;
    for (long i = 0; i < n; ++i) {
        trt::McDivFast f;
        const float q = f(a[i], d[i]);
        if (!f.good()) continue;
        ++in;
        const float w = a[i] / d[i];
        unsigned x, y; memcpy(&x, &q, 4); memcpy(&y, &w, 4);
        bad += x != y;
    }
    *inside = in;
    return bad;
}

/* wbody_row: the 11 doubles of compute.py:1416-1430 (LkArea, LkMxE, OrificeA, OrificeC, OrificeE, WeirC, WeirE, WeirL, ifd,
 * qd0, h0), converted and initialised exactly like trt_levelpool_series (engine.cu) */
extern "C" void trt_replica_levelpool_series(const double* a, long nsteps, const float* inflow, float ql, float dt, float* outflow,
                                             float* elev)
{
    trt::PowTabs T; T.tl = k_log2_tab; T.te = k_exp2_tab;
    trt::LpParams p;
    p.area = (float)a[0]; p.max_depth = (float)a[1]; p.orifice_area = (float)a[2]; p.orifice_coefficient = (float)a[3];
    p.orifice_elevation = (float)a[4]; p.weir_coefficient = (float)a[5]; p.weir_elevation = (float)a[6]; p.weir_length = (float)a[7];
    p.dam_length = 10.0f;
    const float ifd = (float)a[8], we0 = (float)a[10];
    float H = we0;
    if (we0 < -900000000.0f) H = p.orifice_elevation + ((p.max_depth - p.orifice_elevation) * ifd);
    for (long t = 0; t < nsteps; ++t) {
        float q;
        trt::trt_levelpool_step(p, inflow[t], ql, dt, H, q, T);
        outflow[t] = q; elev[t] = H;
    }
}
