/*
 * mc_wrfhydro.c -- TEST INFRASTRUCTURE.  A SECOND, independent C restatement of the Muskingum-Cunge single-segment solve:
 * the WRF-Hydro original the reference carries next to its own kernel,
 *     /root/reference/src/kernel/muskingum/MUSKINGCUNGE.f90:9-334  (subroutine submuskingcunge)
 * written down statement by statement, with no re-ordering, hoisting or sharing of sub-expressions (oracle/mc_kernel.inc,
 * the restatement of MCsingleSegStime_f2py_NOLOOP.f90 the GPU path is checked against, was written separately and
 * factors the secant evaluations differently).
 *
 * Why it exists ("widen the oracle pin"): the reference holds exactly one known-answer vector for the
 * NOLOOP kernel, in-bank and without retries.  The two Fortran files descend from the same routine and perform the SAME
 * float32 operations in the same order wherever their logic coincides: in-bank and compound-channel evaluations with a
 * positive celerity, first attempt of the retry ladder.  tests/test_oracle_wrfhydro.py requires bit equality between the
 * two restatements on every such row of the reference's 5,000-row randomized suite and lists the rows where the algorithms
 * themselves differ, with the reason.  Known differences of the two Fortran sources (flags returned below):
 *   1  retry ladder entered: the original resets Qj_0 = 0 at label 110 (:100), NOLOOP never initialises / resets it
 *      
 *   2  depth above bankfull in a channel without a floodplain (twcc <= 0 or ncc <= 0): NOLOOP extends the trapezoid
 *      (hydraulic_geometry :400-403, "NWM 3.0 exception") and uses the in-bank celerity; the original always switches to
 *      the compound formulas
 *   4  compound evaluation with zero celerity: the original divides by it in X (:144, :222), NOLOOP sets X = 0.5
 *   8  zero wetted perimeter: the original keeps the previous Qj / Qj_0 (:168, :249), NOLOOP sets it to 0
 *   16 only quc positive: the original does not route (:96 tests ql, qup, qdp), NOLOOP does (:73-74 also tests quc)
 * Build: part of liboracle.so (oracle/Makefile), same flags (-O2 -ffp-contract=off, no fast-math).
 */
#include <math.h>
#include <stdint.h>

#include "../include/trt_detmath.h"

static const trt_u64 w_tl[2 * TRT_LOG2_TAB_N] = TRT_LOG2_TAB_INIT;
static const trt_u64 w_te[TRT_EXP2_TAB_N] = TRT_EXP2_TAB_INIT;

static float w_pow(int pow_mode, float x, float y) { return pow_mode ? trt_powf_det(x, y, w_tl, w_te) : powf(x, y); }

#define POWF(x, y) w_pow(pow_mode, (x), (y))

/* in15: dt,qup,quc,qdp,ql,dx,bw,tw,twcc,n,ncc,cs,s0,velp,depthp (the argument order of reach.compute_reach_kernel);
 * out3: qdc, velc, depthc; returns the flag mask described above */
static int wrf_submuskingcunge(int pow_mode, const float* in15, float* out3, int* iters_out)
{
    const float dt = in15[0], qup = in15[1], quc = in15[2], qdp = in15[3], ql = in15[4], dx = in15[5], Bw = in15[6],
                Tw = in15[7], TwCC = in15[8], n = in15[9], nCC = in15[10], Cs = in15[11], So = in15[12], depthp = in15[14];
    float C1 = 0.0f, C2 = 0.0f, C3 = 0.0f, C4 = 0.0f, Km, X, Ck, Twl, AREA = 0.0f, AREAC = 0.0f, z, R = 0.0f, WP = 0.0f, WPC = 0.0f;
    float h, h_0, h_1, bfd, Qj_0 = 0.0f, Qj = 0.0f, D, aerror, rerror, qdc, velc, depthc;
    int iter = 0, maxiter = 100, tries = 0, flags = 0, total = 0;
    const float mindepth = 0.01f;
    const int floodplain = (TwCC > 0.0f) && (nCC > 0.0f);

    aerror = 0.01f;
    rerror = 1.0f;
    if (Cs == 0.0f) z = 1.0f; else z = 1.0f / Cs;                                   /* :69-73 */
    if (Bw > Tw) bfd = Bw / 0.00001f;                                               /* :75-81 */
    else if (Bw == Tw) bfd = Bw / (2.0f * z);
    else bfd = (Tw - Bw) / (2.0f * z);

    depthc = fmaxf(depthp, 0.0f);                                                   /* :92-94 */
    h = (depthc * 1.33f) + mindepth;
    h_0 = (depthc * 0.67f);

    if (!(ql > 0.0f || qup > 0.0f || qdp > 0.0f)) {                                 /* :96, :326-329 */
        if (quc > 0.0f) flags |= 16;
        out3[0] = 0.0f; out3[1] = 0.0f; out3[2] = 0.0f;   /* velc is left undefined by the Fortran; 0 here */
        if (iters_out) *iters_out = 0;
        return flags;
    }
    for (;;) {                                                                      /* label 110 */
        Qj_0 = 0.0f;                                                                /* :100 */
        iter = 0;
        while (rerror > 0.01f && aerror >= mindepth && iter <= maxiter) {           /* :105 */
            AREAC = 0.0f; WPC = 0.0f;
            /* ---- lower interval :110-170 ---- */
            Twl = Bw + 2.0f * z * h_0;
            if (h_0 > bfd) {
                if (!floodplain) flags |= 2;
                AREA = (Bw + bfd * z) * bfd;
                AREAC = (TwCC * (h_0 - bfd));
                WP = (Bw + 2.0f * bfd * sqrtf(1.0f + z * z));
                WPC = TwCC + (2.0f * (h_0 - bfd));
                R = (AREA + AREAC) / (WP + WPC);
            } else {
                AREA = (Bw + h_0 * z) * h_0;
                WP = (Bw + 2.0f * h_0 * sqrtf(1.0f + z * z));
                if (WP > 0.0f) R = AREA / WP; else R = 0.0f;
            }
            if (h_0 > bfd) {
                Ck = fmaxf(0.0f, ((sqrtf(So) / n) * ((5.0f / 3.0f) * POWF(R, 2.0f / 3.0f) -
                     ((2.0f / 3.0f) * POWF(R, 5.0f / 3.0f) * (2.0f * sqrtf(1.0f + z * z) / (Bw + 2.0f * bfd * z)))) * AREA
                     + ((sqrtf(So) / (nCC)) * (5.0f / 3.0f) * POWF(h_0 - bfd, 2.0f / 3.0f)) * AREAC) / (AREA + AREAC));
            } else if (h_0 > 0.0f) {
                Ck = fmaxf(0.0f, (sqrtf(So) / n) * ((5.0f / 3.0f) * POWF(R, 2.0f / 3.0f) -
                     ((2.0f / 3.0f) * POWF(R, 5.0f / 3.0f) * (2.0f * sqrtf(1.0f + z * z) / (Bw + 2.0f * h_0 * z)))));
            } else {
                Ck = 0.0f;
            }
            if (Ck > 0.0f) Km = fmaxf(dt, dx / Ck); else Km = dt;
            if (h_0 > bfd) {
                if (!(Ck > 0.0f)) flags |= 4;
                X = fminf(0.5f, fmaxf(0.0f, 0.5f * (1 - (Qj_0 / (2.0f * TwCC * So * Ck * dx)))));
            } else if (Ck > 0.0f) {
                X = fminf(0.5f, fmaxf(0.0f, 0.5f * (1 - (Qj_0 / (2.0f * Twl * So * Ck * dx)))));
            } else {
                X = 0.5f;
            }
            D = (Km * (1.0f - X) + dt / 2.0f);
            C1 = (Km * X + dt / 2.0f) / D;
            C2 = (dt / 2.0f - Km * X) / D;
            C3 = (Km * (1.0f - X) - dt / 2.0f) / D;
            C4 = (ql * dt) / D;
            if ((WP + WPC) > 0.0f) {
                Qj_0 = ((C1 * qup) + (C2 * quc) + (C3 * qdp) + C4) - ((1 / (((WP * n) + (WPC * nCC)) / (WP + WPC))) *
                       (AREA + AREAC) * (POWF(R, 2.0f / 3.0f)) * sqrtf(So));
            } else flags |= 8;

            AREAC = 0.0f; WPC = 0.0f;
            /* ---- upper interval :175-252 ---- */
            Twl = Bw + 2.0f * z * h;
            if (h > bfd) {
                if (!floodplain) flags |= 2;
                AREA = (Bw + bfd * z) * bfd;
                AREAC = (TwCC * (h - bfd));
                WP = (Bw + 2.0f * bfd * sqrtf(1.0f + z * z));
                WPC = TwCC + (2.0f * (h - bfd));
                R = (AREA + AREAC) / (WP + WPC);
            } else {
                AREA = (Bw + h * z) * h;
                WP = (Bw + 2.0f * h * sqrtf(1.0f + z * z));
                if (WP > 0.0f) R = AREA / WP; else R = 0.0f;
            }
            if (h > bfd) {
                Ck = fmaxf(0.0f, ((sqrtf(So) / n) * ((5.0f / 3.0f) * POWF(R, 2.0f / 3.0f) -
                     ((2.0f / 3.0f) * POWF(R, 5.0f / 3.0f) * (2.0f * sqrtf(1.0f + z * z) / (Bw + 2.0f * bfd * z)))) * AREA
                     + ((sqrtf(So) / (nCC)) * (5.0f / 3.0f) * POWF(h - bfd, 2.0f / 3.0f)) * AREAC) / (AREA + AREAC));
            } else if (h > 0.0f) {
                Ck = fmaxf(0.0f, (sqrtf(So) / n) * ((5.0f / 3.0f) * POWF(R, 2.0f / 3.0f) -
                     ((2.0f / 3.0f) * POWF(R, 5.0f / 3.0f) * (2.0f * sqrtf(1.0f + z * z) / (Bw + 2.0f * h * z)))));
            } else {
                Ck = 0.0f;
            }
            if (Ck > 0.0f) Km = fmaxf(dt, dx / Ck); else Km = dt;
            if (h > bfd) {
                if (!(Ck > 0.0f)) flags |= 4;
                X = fminf(0.5f, fmaxf(0.25f, 0.5f * (1 - (((C1 * qup) + (C2 * quc) + (C3 * qdp) + C4) / (2.0f * TwCC * So * Ck * dx)))));
            } else if (Ck > 0.0f) {
                X = fminf(0.5f, fmaxf(0.25f, 0.5f * (1 - (((C1 * qup) + (C2 * quc) + (C3 * qdp) + C4) / (2.0f * Twl * So * Ck * dx)))));
            } else {
                X = 0.5f;
            }
            D = (Km * (1 - X) + dt / 2);
            C1 = (Km * X + dt / 2.0f) / D;
            C2 = (dt / 2.0f - Km * X) / D;
            C3 = (Km * (1.0f - X) - dt / 2.0f) / D;
            C4 = (ql * dt) / D;
            if ((C4 < 0.0f) && (fabsf(C4) > (C1 * qup) + (C2 * quc) + (C3 * qdp))) C4 = -((C1 * qup) + (C2 * quc) + (C3 * qdp));
            if ((WP + WPC) > 0.0f) {
                Qj = ((C1 * qup) + (C2 * quc) + (C3 * qdp) + C4) - ((1.0f / (((WP * n) + (WPC * nCC)) / (WP + WPC))) *
                     (AREA + AREAC) * (POWF(R, 2.0f / 3.0f)) * sqrtf(So));
            } else flags |= 8;

            if (Qj_0 - Qj != 0.0f) {                                                /* :254-261 */
                h_1 = h - ((Qj * (h_0 - h)) / (Qj_0 - Qj));
                if (h_1 < 0.0f) h_1 = h;
            } else {
                h_1 = h;
            }
            if (h > 0.0f) {                                                         /* :263-269 */
                rerror = fabsf((h_1 - h) / h);
                aerror = fabsf(h_1 - h);
            } else {
                rerror = 0.0f;
                aerror = 0.9f;
            }
            h_0 = fmaxf(0.0f, h);
            h = fmaxf(0.0f, h_1);
            iter = iter + 1;
            total++;
            if (h < mindepth) break;                                                /* :275-277, goto 111 */
        }
        if (iter >= maxiter) {                                                      /* :283-292 */
            tries = tries + 1;
            if (tries <= 4) {
                flags |= 1;
                h = h * 1.33f;
                h_0 = h_0 * 0.67f;
                maxiter = maxiter + 25;
                continue;
            }
        }
        break;
    }
    if (((C1 * qup) + (C2 * quc) + (C3 * qdp) + C4) < 0.0f) {                       /* :307-319 */
        if ((C4 < 0.0f) && (fabsf(C4) > (C1 * qup) + (C2 * quc) + (C3 * qdp))) qdc = 0.0f;
        else qdc = fmaxf(((C1 * qup) + (C2 * quc) + C4), ((C1 * qup) + (C3 * qdp) + C4));
    } else {
        qdc = ((C1 * qup) + (C2 * quc) + (C3 * qdp) + C4);
    }
    Twl = Bw + (2.0f * z * h);                                                      /* :321-324 */
    R = (h * (Bw + Twl) / 2.0f) / (Bw + 2.0f * POWF(POWF((Twl - Bw) / 2.0f, 2.0f) + h * h, 0.5f));
    velc = (1.0f / n) * (POWF(R, 2.0f / 3.0f)) * sqrtf(So);
    depthc = h;
    out3[0] = qdc; out3[1] = velc; out3[2] = depthc;
    if (iters_out) *iters_out = total;
    return flags;
}

void oracle_wrfhydro_mc_batch(int pow_mode, long long count, const float* in15, float* out3, int* flags, int* iters)
{
    for (long long i = 0; i < count; ++i) flags[i] = wrf_submuskingcunge(pow_mode, in15 + 15 * i, out3 + 3 * i, iters ? iters + i : 0);
}
