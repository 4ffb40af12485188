/*
 * diffusive_oracle.c -- CPU restatement of T-Route's diffusive-wave solver.  TEST INFRASTRUCTURE ONLY: nothing in the
 * product package (t-route_b200/) may import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may.
 *
 * PARITY UNPINNED: the reference ships no known-answer vectors for this solver (its CI only checks exit codes,
 * SURVEY.md section 4) and its Fortran cannot be compiled in this image (no gfortran).  What pins this file is the
 * line-by-line correspondence below and the physical property tests of tests/test_diffusive_oracle.py (steady uniform
 * flow is a fixed point, mass conservation of the routed hydrograph, table monotonicity).
 *
 * Follows /root/reference/src/kernel/diffusive/diffusive.f90 (module `diffusive`, double precision throughout):
 *   diffnw                      :75-940     set-up, initial backward sweep, time loop, crosswalk refactoring
 *   calculateDT                 :942-991    adaptive time step from the CFL bound, clipped to the save interval
 *   mesh_diffusive_forward      :1108-1355  Crank-Nicolson + Hermite interpolation; Thomas elimination of (ppi,qqi,rri|ssi)
 *                                           and (...|sxi) along a reach, back substitution from a ghost node
 *   mesh_diffusive_backward     :1357-1553  water-surface elevation node by node, downstream to upstream (rtsafe), celerity
 *                                           and diffusivity of the reach
 *   rtsafe / funcd_diffdepth    :1555-1711  safeguarded Newton on the diffusive momentum equation
 *   intp_xsec_tab, locate       :1713-1748, :2701-2742
 *   readXsection                :2093-2443  synthetic (trapezoid + floodplain) cross-section look-up tables, nel = 501 rows
 *   cal_* helpers, r_interpol, LInterpol, intp_y   :2445-2699
 * entered through c_diffnw (/root/reference/src/kernel/diffusive/pydiffusive.f90:8-52): every argument by reference,
 * arrays in Fortran (column-major) order.
 *
 *   readXsection_natural_mann_vertices :1756-2091  tables of a surveyed cross section (taken when mxnbathy_g > 0)
 * Not restated: the data-assimilation branch, which is commented out in the Fortran itself (:1283-1306).
 *
 * Conventions frozen here (the Fortran leaves them open):
 *   * single-precision literals: `0.3`, `0.1`, `1e-4` ... in a double-precision expression are REAL(4) constants
 *     promoted to double; they are written (double)0.3f etc. below.
 *   * x**2.0 is x*x (GCC folds pow(x, 2.0) at -O2 without fast-math); x**3.0 and fractional powers are libm pow() --
 *     or trt_pow64_det when the GPU parity build is selected (trt_oracle_diffnw_pow_mode).
 *   * arrays the Fortran allocates without initialising (oldY, lateralFlow(ncomp, j), bo, ...) start at 0 here.
 *   * writes past ntss_ev_g rows of the output (a bounds-check abort in the reference build, -fbounds-check) are dropped.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (gfortran -O2 on baseline x86-64 emits no FMA).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/trt_detmath64.h"

/* x**y: 0 = platform libm pow (what a gfortran build of the reference computes; the default), 1 = trt_pow64_det
 * (include/trt_detmath64.h), the bit-specified pow of the CUDA path -- the GPU parity target */
static int g_pow_mode = 0;
void trt_oracle_diffnw_pow_mode(int mode) { g_pow_mode = mode; }
static double P(double x, double y) { return g_pow_mode ? trt_pow64_det(x, y) : pow(x, y); }
void trt_oracle_pow64_det_array(long n, const double* x, const double* y, double* out)
{
    for (long i = 0; i < n; ++i) out[i] = trt_pow64_det(x[i], y[i]);
}

#define NEL 501

typedef struct {
    int mxncomp, nlinks, nel;
    const int* frnw;              /* (nlinks, frnw_col) */
    int frnw_col;
    double dtini, dtini_min, cfl, theta, C_llm, D_llm, D_ulm, q_llm, so_llm;
    double *z, *dx, *bo, *pere, *qp, *qpx, *sk;
    double *oldQ, *newQ, *oldArea, *newArea, *oldY, *newY, *lateralFlow, *celerity, *diffusivity;
    double *celerity2, *diffusivity2, *eei, *ffi, *exi, *fxi, *co;
    double* xsec_tab;             /* (11, nel, mxncomp, nlinks) */
} Dw;

#define A2(p, i, j) ((p)[((i) - 1) + (size_t)((j) - 1) * (size_t)S->mxncomp])
#define FRNW(j, c) (S->frnw[((j) - 1) + (size_t)((c) - 1) * (size_t)S->nlinks])
#define A2P(p, n, j, ld) ((p)[((n) - 1) + (size_t)((j) - 1) * (size_t)(ld)])
#define TAB(c, e, i, j) (S->xsec_tab[((c) - 1) + 11 * ((size_t)((e) - 1) + (size_t)NEL * ((size_t)((i) - 1) + (size_t)S->mxncomp * (size_t)((j) - 1)))])

/* locate :2701-2742 (1-based result; xx has stride `st` doubles) */
static int locate(const double* xx, int n, size_t st, double x)
{
#define XX(k) xx[(size_t)((k) - 1) * st]
    const int ascnd = XX(n) >= XX(1);
    int jl = 0, ju = n + 1;
    for (;;) {
        if (ju - jl <= 1) break;
        const int jm = (ju + jl) / 2;
        if (ascnd == (x >= XX(jm))) jl = jm; else ju = jm;
    }
    if (x == XX(1)) return 1;
    if (x == XX(n)) return n - 1;
    return jl;
#undef XX
}

/* LInterpol :2650-2669 */
static double linterpol(double x1, double y1, double x2, double y2, double x)
{
    if (fabs(x2 - x1) < (double)0.0001f) return 0.5 * (y1 + y2);
    return (y2 - y1) / (x2 - x1) * (x - x1) + y1;
}

/* intp_y :2671-2699 */
static double intp_y(int nrow, const double* xarr, const double* yarr, double x)
{
    int irow = locate(xarr, nrow, 1, x);
    if (irow == 0) irow = 1;
    if (irow == nrow) irow = nrow - 1;
    return linterpol(xarr[irow - 1], yarr[irow - 1], xarr[irow], yarr[irow], x);
}

/* intp_xsec_tab :1713-1748 */
static double intp_xsec_tab(const Dw* S, int i, int j, int xcol, int ycol, double x)
{
    const double* xa = &TAB(xcol, 1, i, j);
    const double* ya = &TAB(ycol, 1, i, j);
    int irow = locate(xa, NEL, 11, x);
    if (irow == 0) irow = 1;
    if (irow == NEL) irow = NEL - 1;
    return linterpol(xa[(size_t)(irow - 1) * 11], ya[(size_t)(irow - 1) * 11], xa[(size_t)irow * 11], ya[(size_t)irow * 11], x);
}

/* r_interpol :2553-2594; x, y with stride 11 (table columns) unless sq_z is given: then x(k) = (elev(k) - z)**2.
 * `*yt` keeps its previous value when no interval matches, as the Fortran's intent(out) argument does in practice. */
static void r_interpol_tab(const Dw* S, int i, int j, int xcol, int ycol, int square, double zsub, double xrt, double* yt)
{
    const double* xa = &TAB(xcol, 1, i, j);
    const double* ya = &TAB(ycol, 1, i, j);
#define XV(k) (square ? ((xa[(size_t)((k) - 1) * 11] - zsub) * (xa[(size_t)((k) - 1) * 11] - zsub)) : xa[(size_t)((k) - 1) * 11])
#define YV(k) ya[(size_t)((k) - 1) * 11]
    double xmax = XV(1), xmin = XV(1), ymin = YV(1);
    for (int k = 2; k <= NEL; ++k) {
        const double v = XV(k);
        if (v > xmax) xmax = v;
        if (v < xmin) xmin = v;
        if (YV(k) < ymin) ymin = YV(k);
    }
    if (xrt <= xmax && xrt >= xmin) {
        for (int k = 1; k <= NEL - 1; ++k) {
            if ((XV(k) - xrt) * (XV(k + 1) - xrt) <= 0.0) {
                *yt = (xrt - XV(k)) / (XV(k + 1) - XV(k)) * (YV(k + 1) - YV(k)) + YV(k);
                break;
            }
        }
    } else if (xrt >= xmax) {
        *yt = (xrt - XV(NEL - 1)) / (XV(NEL) - XV(NEL - 1)) * (YV(NEL) - YV(NEL - 1)) + YV(NEL - 1);
    } else {
        *yt = ymin;
    }
#undef XV
#undef YV
}

/* ---- readXsection :2093-2443 and its helpers :2445-2551 ------------------------------------------------------------ */
static double cal_tri_area(double el, double x0, double x1, double y1) { return fabs(0.5 * (x1 - x0) * (el - y1)); }
static double cal_trap_area(double el, double x1, double y1, double x2, double y2)
{
    return fabs(0.5 * (x2 - x1) * (el - y1 + el - y2));
}
static double cal_multi_area(double el, const double* xx, const double* yy, int i1, int i2)
{
    double area = 0.0;
    for (int i = i1; i <= i2 - 1; ++i) area = area + cal_trap_area(el, xx[i - 1], yy[i - 1], xx[i], yy[i]);
    return area;
}
static double cal_dist(double x1, double y1, double x2, double y2)
{
    return sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (double)1.e-32f);
}
static double cal_perimeter(const double* xx, const double* yy, int i1, int i2)
{
    double p = 0.0;
    for (int i = i1; i <= i2 - 1; ++i) p = p + cal_dist(xx[i - 1], yy[i - 1], xx[i], yy[i]);
    return p;
}

static void read_xsection(Dw* S, int k, double lftBnkMann, double rmanning_main, double rgtBnkMann, double leftBnkX_given,
                          double rghtBnkX_given, double timesDepth, int num_reach, const double* z_ar, const double* bo_ar,
                          const double* traps_ar, const double* tw_ar, const double* twcc_ar)
{
    enum { MT = 8 };
    static const double TOL = (double)1e-8f;                       /* TOLERANCE :19 */
    double xcs[MT + 1], ycs[MT + 1];                               /* 1-based */
    double allX[MT + 1][4], allY[MT + 1][4];
    static double el1[NEL + 1][4], a1[NEL + 1][4], peri1[NEL + 1][4], redi1[NEL + 1][4], conv1[NEL + 1][4], tpW1[NEL + 1][4],
        diffArea[NEL + 1][4], newI1[NEL + 1][4];
    double elev[NEL + 1];
    int i_start[NEL + 1], i_end[NEL + 1];
    const int totalNodes[4] = {0, 5, 7, 5};
    const double f2m = 1.0;
    double leftBnkX = leftBnkX_given, rghtBnkX = rghtBnkX_given;
    const double z_g = A2(z_ar, k, num_reach), bo_g = A2(bo_ar, k, num_reach), traps_g = A2(traps_ar, k, num_reach),
                 tw_g = A2(tw_ar, k, num_reach), twcc_g = A2(twcc_ar, k, num_reach);
    const double hbf = (tw_g - bo_g) / (2.0 * traps_g);            /* bankfull depth :2150 */
    memset(allX, 0, sizeof allX); memset(allY, 0, sizeof allY);
    for (int i = 1; i <= MT; ++i) {                                /* :2156-2193 */
        double x1 = 0.0, y1 = 0.0;
        if (i == 1) { x1 = 0.0; y1 = z_g + timesDepth * hbf; }
        else if (i == 2) { x1 = 0.0; y1 = z_g + hbf; }
        else if (i == 3) { x1 = (twcc_g - tw_g) / 2.0; y1 = z_g + hbf; }
        else if (i == 4) { x1 = xcs[3] + traps_g * hbf; y1 = z_g; }
        else if (i == 5) { x1 = xcs[4] + bo_g; y1 = z_g; }
        else if (i == 6) { x1 = xcs[5] + traps_g * hbf; y1 = z_g + hbf; }
        else if (i == 7) { x1 = twcc_g; y1 = z_g + hbf; }
        else if (i == 8) { x1 = xcs[7]; y1 = z_g + timesDepth * hbf; }
        xcs[i] = x1 * f2m;
        ycs[i] = y1 * f2m;
    }
    const int mainChanStrt = 3, mainChanEnd = 6;                   /* :2195-2196 */
    (void)mainChanEnd;
    int num = MT + 1;                                              /* the do-variable after the loop :2197 */
    {
        double mn = xcs[2], mx = xcs[2];
        for (int i = 3; i <= num - 1; ++i) { if (xcs[i] < mn) mn = xcs[i]; if (xcs[i] > mx) mx = xcs[i]; }
        if (leftBnkX < mn) leftBnkX = mn;                          /* bank stations are not used below (:2201-2206) */
        if (rghtBnkX > mx) rghtBnkX = mx;
        (void)leftBnkX; (void)rghtBnkX;
    }
    double el_min = 99999., el_max = -99999.;
    for (int i = 2; i <= num - 1; ++i) {
        if (ycs[i] < el_min) el_min = ycs[i];
        if (ycs[i] > el_max) el_max = ycs[i];
    }
    const double el_range = (el_max - el_min) * 2.0;
    const double wall = el_min + el_range + 1.;
    /* left overbank :2217-2224 */
    for (int i = 1; i <= 3; ++i) { allX[i + 1][1] = xcs[i]; allY[i + 1][1] = ycs[i]; }
    allX[1][1] = xcs[1]; allY[1][1] = wall;
    allX[mainChanStrt + 2][1] = xcs[3]; allY[mainChanStrt + 2][1] = wall;
    /* main channel :2226-2237 */
    for (int i = 3; i <= 4; ++i) { allX[i - 1][2] = xcs[i]; allY[i - 1][2] = ycs[i]; }
    for (int i = 5; i <= 6; ++i) { allX[i][2] = xcs[i]; allY[i][2] = ycs[i]; }
    allX[1][2] = xcs[3]; allY[1][2] = wall;
    allX[7][2] = xcs[6]; allY[7][2] = wall;
    /* right overbank :2239-2247 */
    for (int i = 6; i <= 8; ++i) { allX[i - 4][3] = xcs[i]; allY[i - 4][3] = ycs[i]; }
    allX[1][3] = allX[2][3]; allY[1][3] = wall;
    allX[5][3] = allX[4][3]; allY[5][3] = wall;
    /* a 1 cm notch at the channel centre :2253-2254 */
    allX[4][2] = (allX[3][2] + allX[5][2]) / 2.0;
    allY[4][2] = allY[3][2] - (double)0.01f;
    el_min = allY[4][2];
    elev[1] = el_min;                                              /* :2258-2266 */
    elev[2] = el_min + (double)(0.01f / 4.f);
    elev[3] = el_min + (double)(0.01f / 4.f * 2.f);
    elev[4] = el_min + (double)(0.01f / 4.f * 3.f);
    elev[5] = el_min + (double)0.01f;
    const double el_incr = el_range / (double)(float)(NEL - 6.0f);
    for (int kkk = 6; kkk <= NEL; ++kkk) elev[kkk] = elev[5] + el_incr * (double)(float)(kkk - 5);
    memset(newI1, 0, sizeof newI1);
    for (int kkk = 1; kkk <= 3; ++kkk) {                           /* :2272-2391 */
        num = totalNodes[kkk];
        for (int i = 1; i <= MT; ++i) { xcs[i] = 0.; ycs[i] = 0.; }
        for (int i = 1; i <= num; ++i) { xcs[i] = allX[i][kkk]; ycs[i] = allY[i][kkk]; }
        const double rmanning = kkk == 1 ? lftBnkMann : (kkk == 2 ? rmanning_main : rgtBnkMann);
        double ymin_nodes = ycs[1];
        for (int i = 2; i <= num; ++i) if (ycs[i] < ymin_nodes) ymin_nodes = ycs[i];
        for (int j = 1; j <= NEL; ++j) {
            double el_now = elev[j];
            if (fabs(el_now - el_min) < TOL) el_now = el_now + (double)0.00001f;
            i_start[1] = -999; i_end[1] = -999;
            int i_area = 0, i_find = 0;
            for (int i = 1; i <= num - 1; ++i) {
                const double y1 = ycs[i], y2 = ycs[i + 1];
                if ((el_now <= y1) && (el_now > y2) && (i_find == 0)) { i_find = 1; i_area = i_area + 1; i_start[i_area] = i; }
                if ((el_now > y1) && (el_now <= y2) && (i_find == 1)) { i_find = 0; i_end[i_area] = i; }
            }
            double cal_area = 0., cal_peri = 0., cal_topW = 0.;
            for (int i = 1; i <= i_area; ++i) {
                double x1 = xcs[i_start[i]], x2 = xcs[i_start[i] + 1], y1 = ycs[i_start[i]], y2 = ycs[i_start[i] + 1];
                double x_start, x_end;
                if (y1 == y2) x_start = x1; else x_start = x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
                x1 = xcs[i_end[i]]; x2 = xcs[i_end[i] + 1]; y1 = ycs[i_end[i]]; y2 = ycs[i_end[i] + 1];
                if (y1 == y2) x_end = x1; else x_end = x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
                cal_topW = x_end - x_start + cal_topW;
                const int i1 = i_start[i], i2 = i_end[i];
                cal_area = cal_area + cal_tri_area(el_now, x_start, xcs[i1 + 1], ycs[i1 + 1])
                         + cal_multi_area(el_now, xcs + 1, ycs + 1, i1 + 1, i2)
                         + cal_tri_area(el_now, x_end, xcs[i2], ycs[i2]);
                cal_peri = cal_peri + cal_dist(x_start, el_now, xcs[i1 + 1], ycs[i1 + 1])
                         + cal_perimeter(xcs + 1, ycs + 1, i1 + 1, i2)
                         + cal_dist(x_end, el_now, xcs[i2], ycs[i2]);
                if (i1 == 1) cal_peri = cal_peri - cal_dist(x_start, el_now, xcs[i1 + 1], ycs[i1 + 1]);
                if (i2 == (num - 1)) cal_peri = cal_peri - cal_dist(x_end, el_now, xcs[i2], ycs[i2]);
            }
            el1[j][kkk] = el_now;
            a1[j][kkk] = cal_area;
            peri1[j][kkk] = cal_peri;
            redi1[j][kkk] = a1[j][kkk] / peri1[j][kkk];
            conv1[j][kkk] = 1. / rmanning * a1[j][kkk] * P(redi1[j][kkk], (double)(2.f / 3.f));
            if (peri1[j][kkk] <= TOL) { redi1[j][kkk] = 0.0; conv1[j][kkk] = 0.0; }
            tpW1[j][kkk] = cal_topW;
            /* (diffPere :2367-2377 is computed by the Fortran and never read) */
            if (j == 1) diffArea[j][kkk] = a1[j][kkk];
            else if (el_now <= ymin_nodes) diffArea[j][kkk] = a1[j][kkk];
            else diffArea[j][kkk] = a1[j][kkk] - a1[j - 1][kkk];
            const double waterElev = el1[j][kkk];
            for (int jj = 2; jj <= j; ++jj) {
                const double diffAreaCenter = el1[jj][kkk] - (el1[jj][kkk] - el1[jj - 1][kkk]) * 0.5;
                newI1[j][kkk] = newI1[j][kkk] + diffArea[jj][kkk] * (waterElev - diffAreaCenter);
            }
        }
    }
    for (int j = 1; j <= NEL; ++j) {                               /* :2393-2426 */
        const double sa = a1[j][1] + a1[j][2] + a1[j][3], sp = peri1[j][1] + peri1[j][2] + peri1[j][3],
                     sc = conv1[j][1] + conv1[j][2] + conv1[j][3];
        double newdPdA, newdKdA;
        if (j == 1) { newdPdA = sp / sa; newdKdA = sc / sa; }
        else {
            const double sa0 = a1[j - 1][1] + a1[j - 1][2] + a1[j - 1][3], sp0 = peri1[j - 1][1] + peri1[j - 1][2] + peri1[j - 1][3],
                         sc0 = conv1[j - 1][1] + conv1[j - 1][2] + conv1[j - 1][3];
            newdPdA = (sp - sp0) / (sa - sa0);
            newdKdA = (sc - sc0) / (sa - sa0);
        }
        const double compoundMann = sqrt((fabs(peri1[j][1]) * (lftBnkMann * lftBnkMann) + fabs(peri1[j][2]) * (rmanning_main * rmanning_main) +
                                          fabs(peri1[j][3]) * (rgtBnkMann * rgtBnkMann)) /
                                         (fabs(peri1[j][1]) + fabs(peri1[j][2]) + fabs(peri1[j][3])));
        TAB(1, j, k, num_reach) = el1[j][1];
        TAB(2, j, k, num_reach) = sa;
        TAB(3, j, k, num_reach) = sp;
        TAB(4, j, k, num_reach) = sa / sp;
        TAB(5, j, k, num_reach) = sc;
        TAB(6, j, k, num_reach) = fabs(tpW1[j][1]) + fabs(tpW1[j][2]) + fabs(tpW1[j][3]);
        TAB(7, j, k, num_reach) = newI1[j][1] + newI1[j][2] + newI1[j][3];
        TAB(8, j, k, num_reach) = newdPdA;
        TAB(9, j, k, num_reach) = newdKdA;
        TAB(11, j, k, num_reach) = 1. / compoundMann;
    }
    A2(S->z, k, num_reach) = el_min;                               /* :2428 */
}

/* ---- readXsection_natural_mann_vertices :1756-2091 ------------------------------------------------------------------ */
/* surveyed cross section: size_bathy(node, reach) vertices (x, z, Manning n), closed by a vertical wall on either side */
static int read_xsection_natural(Dw* S, int idx_node, int idx_reach, double timesDepth, const double* x_bathy,
                                 const double* z_bathy, const double* mann_bathy, const int* size_bathy, int mxnbathy)
{
    static const double TOL = (double)1e-8f;
    const int nb = size_bathy[(idx_node - 1) + (size_t)(idx_reach - 1) * (size_t)S->mxncomp];
    if (nb < 2 || nb > mxnbathy) return -6;
#define BATHY(p, ic) (p)[((ic) - 1) + (size_t)mxnbathy * ((size_t)(idx_node - 1) + (size_t)S->mxncomp * (size_t)(idx_reach - 1))]
    const int mt = nb + 2, num = mt;
    double* buf = (double*)malloc(sizeof(double) * (size_t)(3 * (mt + 1) + 8 * (NEL + 1)));
    if (!buf) return -3;
    double *xcs = buf, *ycs = xcs + (mt + 1), *manncs = ycs + (mt + 1);                      /* 1-based */
    double *el1 = manncs + (mt + 1), *a1 = el1 + (NEL + 1), *peri1 = a1 + (NEL + 1), *conv1 = peri1 + (NEL + 1),
           *tpW1 = conv1 + (NEL + 1), *newdKdA = tpW1 + (NEL + 1), *skk = newdKdA + (NEL + 1), *redi1 = skk + (NEL + 1);
    const double f2m = 1.0;
    for (int ic = 2; ic <= nb + 1; ++ic) {                                                   /* :1799-1815 */
        const double x1 = -BATHY(x_bathy, 1) + BATHY(x_bathy, ic - 1);
        xcs[ic] = x1 * f2m;
        ycs[ic] = BATHY(z_bathy, ic - 1) * f2m;
        manncs[ic] = BATHY(mann_bathy, ic - 1);
        if (manncs[ic] > (double)0.15f) manncs[ic] = (double)0.15f;
    }
    double el_min = 99999., el_max = -99999.;
    for (int ic = 2; ic <= num - 1; ++ic) {
        if (ycs[ic] < el_min) el_min = ycs[ic];
        if (ycs[ic] > el_max) el_max = ycs[ic];
    }
    const double el_range = (el_max - el_min) * timesDepth;
    const double el_incr = el_range / (double)(float)(NEL - 1.0f);
    xcs[1] = xcs[2]; ycs[1] = el_min + el_range + 1.0;
    xcs[num] = xcs[num - 1]; ycs[num] = el_min + el_range + 1.0;
    manncs[1] = 0.0; manncs[num - 1] = 0.0; manncs[num] = 0.0;
    for (int iel = 1; iel <= NEL; ++iel) {                                                   /* :1840-1925 */
        double el_now = el_min + (double)(float)(iel - 1) * el_incr;
        if (fabs(el_now - el_min) < TOL) el_now = el_now + (double)0.00001f;
        double cal_area = 0.0, cal_peri = 0.0, cal_topW = 0.0, cal_equiv_mann = 0.0;
        int i_find = 0, i_start = -999;
        for (int ic = 1; ic <= num - 1; ++ic) {
            const double ya = ycs[ic], yb = ycs[ic + 1];
            if ((el_now <= ya) && (el_now > yb) && (i_find == 0)) { i_find = 1; i_start = ic; }
            if ((el_now > ya) && (el_now <= yb) && (i_find == 1)) {
                /* the Fortran first lists the wetted pockets (i_start, i_end) and then sums them in the same order */
                i_find = 0;
                const int i1 = i_start, i2 = ic;
                double x1 = xcs[i1], x2 = xcs[i1 + 1], y1 = ycs[i1], y2 = ycs[i1 + 1];
                const double x_start = (y1 == y2) ? x1 : x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
                x1 = xcs[i2]; x2 = xcs[i2 + 1]; y1 = ycs[i2]; y2 = ycs[i2 + 1];
                const double x_end = (y1 == y2) ? x1 : x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
                cal_topW = x_end - x_start + cal_topW;
                cal_area = cal_area + cal_tri_area(el_now, x_start, xcs[i1 + 1], ycs[i1 + 1])
                         + cal_multi_area(el_now, xcs + 1, ycs + 1, i1 + 1, i2)
                         + cal_tri_area(el_now, x_end, xcs[i2], ycs[i2]);
                cal_peri = cal_peri + cal_dist(x_start, el_now, xcs[i1 + 1], ycs[i1 + 1])
                         + cal_perimeter(xcs + 1, ycs + 1, i1 + 1, i2)
                         + cal_dist(x_end, el_now, xcs[i2], ycs[i2]);
                double pxm = 0.0;                                                            /* cal_peri_x_mann :2060-2089 */
                for (int i = i1 + 1; i <= i2 - 1; ++i) pxm = pxm + cal_dist(xcs[i], ycs[i], xcs[i + 1], ycs[i + 1]) * P(manncs[i], 1.50);
                cal_equiv_mann = cal_equiv_mann + cal_dist(x_start, el_now, xcs[i1 + 1], ycs[i1 + 1]) * P(manncs[i1], 1.50)
                               + pxm + cal_dist(x_end, el_now, xcs[i2], ycs[i2]) * P(manncs[i2], 1.50);
                if (i1 == 1) cal_peri = cal_peri - cal_dist(x_start, el_now, xcs[i1 + 1], ycs[i1 + 1]);
                if (i2 == (num - 1)) cal_peri = cal_peri - cal_dist(x_end, el_now, xcs[i2], ycs[i2]);
            }
        }
        el1[iel] = el_now; a1[iel] = cal_area; peri1[iel] = cal_peri; tpW1[iel] = cal_topW;
        redi1[iel] = a1[iel] / peri1[iel];
        const double equiv_mann = P(cal_equiv_mann / cal_peri, (double)(2.0f / 3.0f));
        conv1[iel] = (1.0 / equiv_mann) * a1[iel] * P(redi1[iel], (double)(2.0f / 3.0f));
        if (peri1[iel] <= TOL) { redi1[iel] = 0.0; conv1[iel] = 0.0; }
        if (iel == 1) newdKdA[iel] = conv1[iel] / a1[iel];
        else newdKdA[iel] = (conv1[iel] - conv1[iel - 1]) / (a1[iel] - a1[iel - 1]);
        skk[iel] = 1.0 / equiv_mann;
    }
    /* conveyance made monotone in elevation :1951-1986 */
    const double incr_rate = (double)0.01f;
    int iel_start = 2;
    for (int iel = iel_start; iel <= NEL; ++iel) {
        /* NB: the Fortran's `iel_start = iel_incr_start` inside the loop does not change the do-loop's bounds */
        if (conv1[iel] <= conv1[iel - 1]) {
            int ii = iel;
            while ((conv1[ii] < conv1[iel - 1]) && (ii < NEL)) ii = ii + 1;
            const int inc0 = ii;
            if ((inc0 >= NEL) && (conv1[inc0] < conv1[iel - 1])) conv1[inc0] = (1.0 + incr_rate) * conv1[iel - 1];
            const double pos_slope = (conv1[inc0] - conv1[iel - 1]) / (el1[inc0] - el1[iel - 1]);
            for (ii = iel; ii <= inc0 - 1; ++ii) conv1[ii] = conv1[iel - 1] + pos_slope * (el1[ii] - el1[iel - 1]);
            for (ii = iel; ii <= inc0 - 1; ++ii) {
                if (ii == 1) newdKdA[ii] = conv1[ii] / a1[ii];
                else newdKdA[ii] = (conv1[ii] - conv1[ii - 1]) / (a1[ii] - a1[ii - 1]);
            }
        }
    }
    /* dK/dA made monotone :1988-2008 */
    for (int iel = 2; iel <= NEL; ++iel) {
        if (newdKdA[iel] <= newdKdA[iel - 1]) {
            int ii = iel;
            while ((newdKdA[ii] < newdKdA[iel - 1]) && (ii < NEL)) ii = ii + 1;
            const int inc0 = ii;
            if ((inc0 >= NEL) && (newdKdA[inc0] < newdKdA[iel - 1])) newdKdA[inc0] = (1.0 + incr_rate) * newdKdA[iel - 1];
            const double pos_slope = (newdKdA[inc0] - newdKdA[iel - 1]) / (el1[inc0] - el1[iel - 1]);
            for (ii = iel; ii <= inc0 - 1; ++ii) newdKdA[ii] = newdKdA[iel - 1] + pos_slope * (el1[ii] - el1[iel - 1]);
        }
    }
    for (int iel = 1; iel <= NEL; ++iel) {                                                   /* :2010-2019 */
        TAB(1, iel, idx_node, idx_reach) = el1[iel];
        TAB(2, iel, idx_node, idx_reach) = a1[iel];
        TAB(3, iel, idx_node, idx_reach) = peri1[iel];
        TAB(4, iel, idx_node, idx_reach) = redi1[iel];
        TAB(5, iel, idx_node, idx_reach) = conv1[iel];
        TAB(6, iel, idx_node, idx_reach) = tpW1[iel];
        TAB(9, iel, idx_node, idx_reach) = newdKdA[iel];
        TAB(11, iel, idx_node, idx_reach) = skk[iel];
    }
    A2(S->z, idx_node, idx_reach) = el_min;                                                  /* :2021 */
    free(buf);
    return 0;
#undef BATHY
}

/* ---- funcd_diffdepth :1664-1711, rtsafe :1555-1662 ------------------------------------------------------------------ */
static void funcd_diffdepth(const Dw* S, int i, int j, double Q_cur, double Q_ds, double z_cur, double z_ds, double y_cur,
                            double y_ds, double* f, double* df)
{
    const double elv_ds = y_ds + z_ds;
    const double conv_ds = intp_xsec_tab(S, i + 1, j, 1, 5, elv_ds);
    const double sf_ds = fabs(Q_ds) * Q_ds / (conv_ds * conv_ds);
    const double elv_cur = y_cur + z_cur;
    const double conv_cur = intp_xsec_tab(S, i, j, 1, 5, elv_cur);
    const double sf_cur = fabs(Q_cur) * Q_cur / (conv_cur * conv_cur);
    double slope = (A2(S->z, i, j) - A2(S->z, i + 1, j)) / A2(S->dx, i, j);
    slope = fmax(slope, S->so_llm);
    *f = y_cur - y_ds + slope * A2(S->dx, i, j) - 0.50 * (sf_cur + sf_ds) * A2(S->dx, i, j);
    const double dKdA_cur = intp_xsec_tab(S, i, j, 1, 9, elv_cur);
    const double topw_cur = intp_xsec_tab(S, i, j, 1, 6, elv_cur);
    *df = 1.0 + (fabs(Q_cur) * Q_cur / P(conv_cur, 3.0)) * A2(S->dx, i, j) * topw_cur * dKdA_cur;
}

static double rtsafe(const Dw* S, int i, int j, double Q_cur, double Q_ds, double z_cur, double z_ds, double y_ds)
{
    const int maxit = 40;
    const double xacc = (double)1e-4f;
    const double y_ulm_multi = 2.0, y_llm_multi = (double)0.1f;
    double df, dxx, dxold, f, fh, fl, temp, xh, xl, r;
    const double elv_norm = intp_xsec_tab(S, i, j, 10, 1, fabs(Q_cur));   /* normal elevation */
    const double y_norm = elv_norm - A2(S->z, i, j);
    const double y_old = A2(S->oldY, i, j) - A2(S->z, i, j);
    const double x1 = 0.5 * (y_norm + y_old) * y_llm_multi;
    const double x2 = 0.5 * (y_norm + y_old) * y_ulm_multi;
    funcd_diffdepth(S, i, j, Q_cur, Q_ds, z_cur, z_ds, x1, y_ds, &fl, &df);
    funcd_diffdepth(S, i, j, Q_cur, Q_ds, z_cur, z_ds, x2, y_ds, &fh, &df);
    if ((fl > 0.0 && fh > 0.0) || (fl < 0.0 && fh < 0.0)) return y_norm;
    if (fl == 0.0) return x1;
    else if (fh == 0.0) return x2;
    else if (fl < 0.0) { xl = x1; xh = x2; }
    else { xh = x1; xl = x2; }
    r = 0.50 * (x1 + x2);
    dxold = fabs(x2 - x1);
    dxx = dxold;
    funcd_diffdepth(S, i, j, Q_cur, Q_ds, z_cur, z_ds, r, y_ds, &f, &df);
    for (int iter = 1; iter <= maxit; ++iter) {
        if (((r - xh) * df - f) * ((r - xl) * df - f) > 0.0 || fabs(2.0 * f) > fabs(dxold * df)) {
            dxold = dxx;
            dxx = 0.50 * (xh - xl);
            r = xl + dxx;
            if (xl == r) return r;
        } else {
            dxold = dxx;
            dxx = f / df;
            temp = r;
            r = r - dxx;
            if (temp == r) return r;
        }
        if (fabs(dxx) < xacc) return r;
        funcd_diffdepth(S, i, j, Q_cur, Q_ds, z_cur, z_ds, r, y_ds, &f, &df);
        if (f < 0.0) xl = r; else xh = r;
    }
    return y_norm;
}

/* ---- mesh_diffusive_forward :1108-1355 ----------------------------------------------------------------------------- */
static void hermite_cn(const Dw* S, double dxm, double cel, double* a, double* b, double* dd, double* h)
{
    /* the a1..a4, b1..b4, dd1..dd4, h1..h4 block that appears at :1173-1196 and again at :1235-1258 */
    const double cour = S->dtini / dxm;
    const double cour2 = fabs(cel) * cour;
    const double c2 = cour2 * cour2, c3 = P(cour2, 3.0);
    a[1] = 3.0 * c2 - 2.0 * c3;
    a[2] = 1 - a[1];
    a[3] = (c2 - c3) * dxm;
    a[4] = (-1.0 * cour2 + 2.0 * c2 - c3) * dxm;
    b[1] = (6.0 * cour2 - 6.0 * c2) / (-1.0 * dxm);
    b[2] = -b[1];
    b[3] = (2.0 * cour2 - 3.0 * c2) * (-1.0);
    b[4] = (-1.0 + 4.0 * cour2 - 3.0 * c2) * (-1.0);
    dd[1] = (6.0 - 12.0 * cour2) / (dxm * dxm);
    dd[2] = -dd[1];
    dd[3] = (2.0 - 6.0 * cour2) / dxm;
    dd[4] = (4.0 - 6.0 * cour2) / dxm;
    h[1] = 12.0 / P(dxm, 3.0);
    h[2] = -h[1];
    h[3] = 6.0 / (dxm * dxm);
    h[4] = h[3];
}

static void mesh_diffusive_forward(Dw* S, int j)
{
    const int ncomp = FRNW(j, 1);
    double *eei = S->eei, *ffi = S->ffi, *exi = S->exi, *fxi = S->fxi;       /* 1-based below */
    double a[5], b[5], dd[5], h[5];
    for (int i = 1; i <= S->mxncomp; ++i) { eei[i] = -999.; ffi[i] = -999.; exi[i] = -999.; fxi[i] = -999.; }
    eei[1] = 1.; ffi[1] = 0.; exi[1] = 0.; fxi[1] = 0.;
    double allqlat = 0.0;                                          /* :1162-1165 */
    for (int i = 2; i <= ncomp - 1; ++i) allqlat = allqlat + A2(S->lateralFlow, i, j) * A2(S->dx, i, j);
    for (int i = 2; i <= ncomp; ++i) {                             /* :1171-1229 */
        const double dxm = A2(S->dx, i - 1, j);
        hermite_cn(S, dxm, A2(S->celerity, i, j), a, b, dd, h);
        const double alpha = (i == ncomp) ? 1.0 : A2(S->dx, i, j) / dxm;
        const double oq0 = A2(S->oldQ, i - 1, j), oq1 = A2(S->oldQ, i, j), px0 = A2(S->qpx, i - 1, j), px1 = A2(S->qpx, i, j);
        const double qy = a[1] * oq0 + a[2] * oq1 + a[3] * px0 + a[4] * px1;
        const double qxy = b[1] * oq0 + b[2] * oq1 + b[3] * px0 + b[4] * px1;
        const double qxxy = dd[1] * oq0 + dd[2] * oq1 + dd[3] * px0 + dd[4] * px1;
        const double qxxxy = h[1] * oq0 + h[2] * oq1 + h[3] * px0 + h[4] * px1;
        const double dif = A2(S->diffusivity, i, j);
        const double ppi = -S->theta * dif * S->dtini / (dxm * dxm) * 2.0 / (alpha * (alpha + 1.0)) * alpha;
        const double qqi = 1.0 - ppi * (alpha + 1.0) / alpha;
        const double rri = ppi / alpha;
        const double ssi = qy + S->dtini * dif * (1.0 - S->theta) * qxxy;
        const double sxi = qxy + S->dtini * dif * (1.0 - S->theta) * qxxxy;
        eei[i] = -1.0 * rri / (ppi * eei[i - 1] + qqi);
        ffi[i] = (ssi - ppi * ffi[i - 1]) / (ppi * eei[i - 1] + qqi);
        exi[i] = -1.0 * rri / (ppi * exi[i - 1] + qqi);
        fxi[i] = (sxi - ppi * fxi[i - 1]) / (ppi * exi[i - 1] + qqi);
    }
    /* the ghost point (:1231-1279) is evaluated by the Fortran but only qp_ghost and qpx_ghost reach a result */
    const double qp_ghost = A2(S->oldQ, ncomp - 1, j);
    const double qpx_ghost = 0.;
    A2(S->qp, ncomp, j) = eei[ncomp] * qp_ghost + ffi[ncomp];      /* :1297, :1303 */
    A2(S->qpx, ncomp, j) = exi[ncomp] * qpx_ghost + fxi[ncomp];
    for (int i = ncomp - 1; i >= 1; --i) {                         /* :1309-1312 */
        A2(S->qp, i, j) = eei[i] * A2(S->qp, i + 1, j) + ffi[i];
        A2(S->qpx, i, j) = exi[i] * A2(S->qpx, i + 1, j) + fxi[i];
    }
    A2(S->qp, 1, j) = A2(S->newQ, 1, j);                           /* :1316-1317 */
    A2(S->qp, 1, j) = A2(S->qp, 1, j) + allqlat;
    for (int i = 1; i <= ncomp; ++i)
        if (fabs(A2(S->qp, i, j)) < S->q_llm) A2(S->qp, i, j) = S->q_llm;
    for (int i = 1; i <= ncomp; ++i) A2(S->newQ, i, j) = A2(S->qp, i, j);
}

/* ---- mesh_diffusive_backward :1357-1553 ---------------------------------------------------------------------------- */
static void mesh_diffusive_backward(Dw* S, int j)
{
    const int ncomp = FRNW(j, 1);
    const double q_sk_multi = 1.0;
    r_interpol_tab(S, ncomp, j, 1, 2, 0, 0.0, A2(S->newY, ncomp, j), &A2(S->newArea, ncomp, j));   /* :1416-1417 */
    r_interpol_tab(S, ncomp, j, 1, 6, 0, 0.0, A2(S->newY, ncomp, j), &A2(S->bo, ncomp, j));        /* :1424-1425 */
    for (int i = ncomp; i >= 1; --i) {
        const double xt = A2(S->newY, i, j);
        const double zz = A2(S->z, i, j);
        r_interpol_tab(S, i, j, 1, 5, 1, zz, (xt - zz) * (xt - zz), &S->co[i]);                   /* :1446-1448 */
        S->co[i] = q_sk_multi * S->co[i];
        r_interpol_tab(S, i, j, 1, 2, 0, 0.0, xt, &A2(S->newArea, i, j));
        r_interpol_tab(S, i, j, 1, 3, 0, 0.0, xt, &A2(S->pere, i, j));
        r_interpol_tab(S, i, j, 1, 6, 0, 0.0, xt, &A2(S->bo, i, j));
        r_interpol_tab(S, i, j, 1, 11, 0, 0.0, xt, &A2(S->sk, i, j));
        const double q = A2(S->qp, i, j);
        const double sfi = q * fabs(q) / (S->co[i] * S->co[i]);                                   /* :1475 */
        S->celerity2[i] = (double)(5.0f / 3.0f) * P(fabs(sfi), (double)0.3f) * P(fabs(q), (double)0.4f)
                          / P(A2(S->bo, i, j), (double)0.4f)
                          / P(1. / (A2(S->sk, i, j) * q_sk_multi), (double)0.6f);               /* :1479-1481 */
        const double C_ulm = (i > 1) ? S->cfl * A2(S->dx, i - 1, j) / S->dtini_min : S->cfl * A2(S->dx, i, j) / S->dtini_min;
        if (S->celerity2[i] > C_ulm) S->celerity2[i] = C_ulm;
        S->diffusivity2[i] = fabs(q) / 2.0 / A2(S->bo, i, j) / fabs(sfi);                         /* :1495 */
        if (i > 1) {                                                                              /* :1498-1530 */
            const double Q_cur = A2(S->qp, i - 1, j), Q_ds = q, z_cur = A2(S->z, i - 1, j), z_ds = zz;
            double y_ds = xt - zz;
            y_ds = fmax(y_ds, (double)0.005f);
            const double y_cur = rtsafe(S, i - 1, j, Q_cur, Q_ds, z_cur, z_ds, y_ds);
            A2(S->newY, i - 1, j) = y_cur + z_cur;
            if (A2(S->newY, i - 1, j) > 100000.0) A2(S->newY, i - 1, j) = 100000.0;
        }
    }
    double sc = 0.0, sd = 0.0;                                     /* :1535-1545 */
    for (int i = 1; i <= ncomp; ++i) sc = sc + S->celerity2[i];
    for (int i = 1; i <= ncomp; ++i) sd = sd + S->diffusivity2[i];
    for (int i = 1; i <= ncomp; ++i) A2(S->celerity, i, j) = sc / ncomp;
    if (A2(S->celerity, 1, j) < S->C_llm)
        for (int i = 1; i <= ncomp; ++i) A2(S->celerity, i, j) = S->C_llm;
    for (int i = 1; i <= ncomp; ++i) {
        double d = sd / ncomp;
        if (d > S->D_ulm) d = S->D_ulm;
        if (d < S->D_llm) d = S->D_llm;
        A2(S->diffusivity, i, j) = d;
    }
}

static int is_mainstem(const int* mstem, int nm, int j)
{
    for (int i = 0; i < nm; ++i) if (mstem[i] == j) return 1;
    return 0;
}

/* ---- diffnw :75-940 behind c_diffnw (pydiffusive.f90:8-52) --------------------------------------------------------- */
int trt_oracle_diffnw(const double* timestep_ar_g, const int* nts_ql_g, const int* nts_ub_g, const int* nts_db_g,
                      const int* ntss_ev_g, const int* nts_qtrib_g, const int* nts_da_g, const int* mxncomp_g,
                      const int* nrch_g, const double* z_ar_g, const double* bo_ar_g, const double* traps_ar_g,
                      const double* tw_ar_g, const double* twcc_ar_g, const double* mann_ar_g, const double* manncc_ar_g,
                      double* so_ar_g, const double* dx_ar_g, const double* iniq, const int* frnw_col, const int* frnw_ar_g,
                      const double* qlat_g, const double* ubcd_g, const double* dbcd_g, const double* qtrib_g,
                      const int* paradim, const double* para_ar_g, const int* mxnbathy_g, const double* x_bathy_g,
                      const double* z_bathy_g, const double* mann_bathy_g, const int* size_bathy_g, const double* usgs_da_g,
                      const int* usgs_da_reach_g, const double* rdx_ar_g, const int* cwnrow_g, const int* cwncol_g,
                      const double* crosswalk_g, const double* z_thalweg_g, double* q_ev_g, double* elv_ev_g,
                      double* depth_ev_g)
{
    (void)nts_ub_g; (void)so_ar_g; (void)ubcd_g; (void)paradim; (void)usgs_da_g; (void)usgs_da_reach_g; (void)nts_da_g;
    static const double TOL = (double)1e-8f;
    Dw Sv, *S = &Sv;
    memset(S, 0, sizeof *S);
    const int mx = *mxncomp_g, nl = *nrch_g, nql = *nts_ql_g, ndb = *nts_db_g, nev = *ntss_ev_g, nqt = *nts_qtrib_g;
    const int mxnbathy = *mxnbathy_g;                              /* > 0: natural cross sections (:433-454) */
    if (mx < 2 || nl < 1) return -1;
    S->mxncomp = mx; S->nlinks = nl; S->nel = NEL; S->frnw = frnw_ar_g; S->frnw_col = *frnw_col;
    double dtini = timestep_ar_g[0];
    const double t0 = timestep_ar_g[1], tfin = timestep_ar_g[2], saveInterval = timestep_ar_g[3], dt_ql = timestep_ar_g[4],
                 dt_db = timestep_ar_g[6], dt_qtrib = timestep_ar_g[7];
    const double dtini_divisor = timestep_ar_g[9];
    S->dtini = dtini; S->dtini_min = dtini / dtini_divisor;
    const double timesDepth = 4.0;
    S->cfl = para_ar_g[0]; S->C_llm = para_ar_g[1]; S->D_llm = para_ar_g[2]; S->D_ulm = para_ar_g[3];
    S->q_llm = para_ar_g[7]; S->so_llm = para_ar_g[8]; S->theta = para_ar_g[9];
    const int dsbc_option = (int)para_ar_g[10];
    const double mindepth_nstab = (double)0.1f;

    const size_t n2 = (size_t)mx * nl;
    double* pool = (double*)calloc(n2 * 16 + (size_t)(mx + 1) * 7 + (size_t)11 * NEL * n2, sizeof(double));
    if (!pool) return -3;
    double* p = pool;
#define TAKE(field, cnt) S->field = p; p += (cnt)
    TAKE(z, n2); TAKE(dx, n2); TAKE(bo, n2); TAKE(pere, n2); TAKE(qp, n2); TAKE(qpx, n2); TAKE(sk, n2); TAKE(oldQ, n2);
    TAKE(newQ, n2); TAKE(oldArea, n2); TAKE(newArea, n2); TAKE(oldY, n2); TAKE(newY, n2); TAKE(lateralFlow, n2);
    TAKE(celerity, n2); TAKE(diffusivity, n2);
    TAKE(celerity2, mx + 1); TAKE(diffusivity2, mx + 1); TAKE(eei, mx + 1); TAKE(ffi, mx + 1); TAKE(exi, mx + 1);
    TAKE(fxi, mx + 1); TAKE(co, mx + 1);
    TAKE(xsec_tab, (size_t)11 * NEL * n2);
#undef TAKE
    int* mstem = (int*)malloc(sizeof(int) * (size_t)nl);
    double* tarr_ql = (double*)malloc(sizeof(double) * (size_t)(nql + 1) * 2);
    double* varr_ql = tarr_ql + (nql + 1);
    double* tarr_qtrib = (double*)malloc(sizeof(double) * (size_t)(nqt > 0 ? nqt : 1) * 2);
    double* varr_qtrib = tarr_qtrib + (nqt > 0 ? nqt : 1);
    double* tarr_db = (double*)malloc(sizeof(double) * (size_t)(ndb > 0 ? ndb : 1) * 2);
    double* varr_db = tarr_db + (ndb > 0 ? ndb : 1);
    if (!mstem || !tarr_ql || !tarr_qtrib || !tarr_db) { free(pool); free(mstem); free(tarr_ql); free(tarr_qtrib); free(tarr_db); return -3; }
    for (int i = 0; i < (ndb > 0 ? ndb : 1); ++i) varr_db[i] = 0.0;

    memcpy(S->z, z_ar_g, n2 * sizeof(double));                     /* :372 */
    for (size_t i = 0; i < n2; ++i) { S->newQ[i] = -999; S->newY[i] = -999; }
    double t = t0 * 60.0;
    memcpy(S->oldQ, iniq, n2 * sizeof(double));                    /* :383-385 */
    memcpy(S->newQ, S->oldQ, n2 * sizeof(double));
    memcpy(S->qp, S->oldQ, n2 * sizeof(double));
    const size_t nout = (size_t)nev * n2;
    for (size_t i = 0; i < nout; ++i) { q_ev_g[i] = 0.0; elv_ev_g[i] = 0.0; depth_ev_g[i] = 0.0; }
#define EV(arr, ts, i, j) (arr)[((ts) - 1) + (size_t)nev * ((size_t)((i) - 1) + (size_t)mx * (size_t)((j) - 1))]

    int nm = 0;                                                    /* mainstem reaches :403-417 */
    for (int j = 1; j <= nl; ++j) {
        const int nusrch = FRNW(j, 3);
        if (FRNW(j, 3 + nusrch + 1) == 555) mstem[nm++] = j;
    }
    if (nm == 0) { free(pool); free(mstem); free(tarr_ql); free(tarr_qtrib); free(tarr_db); return -4; }
    double minDx = 1e10;
    for (int jm = 0; jm < nm; ++jm) {                              /* :421-430 */
        const int j = mstem[jm], ncomp = FRNW(j, 1);
        for (int i = 1; i <= ncomp - 1; ++i) {
            A2(S->dx, i, j) = A2(dx_ar_g, i, j);
            if (A2(S->dx, i, j) < minDx) minDx = A2(S->dx, i, j);
        }
    }
    for (int jm = 0; jm < nm && mxnbathy != 0; ++jm) {             /* natural cross sections :441-454 */
        const int j = mstem[jm];
        for (int i = 1; i <= FRNW(j, 1); ++i) {
            const int rc = read_xsection_natural(S, i, j, timesDepth, x_bathy_g, z_bathy_g, mann_bathy_g, size_bathy_g, mxnbathy);
            if (rc != 0) { free(pool); free(mstem); free(tarr_ql); free(tarr_qtrib); free(tarr_db); return rc; }
        }
    }
    for (int jm = 0; jm < nm && mxnbathy == 0; ++jm) {             /* synthetic cross sections :456-483 */
        const int j = mstem[jm], ncomp = FRNW(j, 1);
        for (int i = 1; i <= ncomp; ++i) {
            const double leftBank = (A2(twcc_ar_g, i, j) - A2(tw_ar_g, i, j)) / 2.0;
            const double rightBank = (A2(twcc_ar_g, i, j) - A2(tw_ar_g, i, j)) / 2.0 + A2(tw_ar_g, i, j);
            const double skLeft = 1.0 / A2(manncc_ar_g, i, j), skRight = 1.0 / A2(manncc_ar_g, i, j),
                         skMain = 1.0 / A2(mann_ar_g, i, j);
            read_xsection(S, i, 1.0 / skLeft, 1.0 / skMain, 1.0 / skRight, leftBank, rightBank, timesDepth, j, z_ar_g,
                          bo_ar_g, traps_ar_g, tw_ar_g, twcc_ar_g);
        }
    }
    for (int jm = 0; jm < nm; ++jm) {                              /* uniform-flow column :487-506 */
        const int j = mstem[jm], ncomp = FRNW(j, 1);
        for (int i = 1; i <= ncomp; ++i) {
            for (int iel = 1; iel <= NEL; ++iel) {
                const double convey = TAB(5, iel, i, j);
                double slope;
                if (i < ncomp) slope = (A2(S->z, i, j) - A2(S->z, i + 1, j)) / A2(S->dx, i, j);
                else slope = (A2(S->z, i - 1, j) - A2(S->z, i, j)) / A2(S->dx, i - 1, j);
                if (slope <= S->so_llm) slope = S->so_llm;
                TAB(10, iel, i, j) = convey * P(slope, 0.50);
            }
        }
    }
    for (int n = 1; n <= nql; ++n) tarr_ql[n] = t0 * 60.0 + dt_ql * (double)n / 60.0;      /* :512-516 */
    tarr_ql[0] = t0 * 60;
    for (int n = 1; n <= nqt; ++n) tarr_qtrib[n - 1] = t0 * 60.0 + dt_qtrib * (double)(n - 1) / 60.0;
    for (int n = 1; n <= ndb; ++n) tarr_db[n - 1] = t0 * 60.0 + dt_db * (double)(n - 1) / 60.0;

    /* initial water surface, downstream to upstream :550-606 */
    for (int jm = nm; jm >= 1; --jm) {
        const int j = mstem[jm - 1], ncomp = FRNW(j, 1);
        if (FRNW(j, 2) < 0) {
            if (dsbc_option == 1) {
                for (int n = 1; n <= ndb; ++n) varr_db[n - 1] = dbcd_g[n - 1] + A2(S->z, ncomp, j);
                t = t0 * 60.0;
                A2(S->oldY, ncomp, j) = intp_y(ndb, tarr_db, varr_db, t);
                A2(S->newY, ncomp, j) = A2(S->oldY, ncomp, j);
                if ((A2(S->newY, ncomp, j) - A2(S->z, ncomp, j)) < mindepth_nstab)
                    A2(S->newY, ncomp, j) = mindepth_nstab + A2(S->z, ncomp, j);
            } else if (dsbc_option == 2) {
                A2(S->oldY, ncomp, j) = intp_xsec_tab(S, ncomp, j, 10, 1, A2(S->oldQ, ncomp, j));
                A2(S->newY, ncomp, j) = A2(S->oldY, ncomp, j);
            }
        } else {
            const int linknb = FRNW(j, 2);
            A2(S->newY, ncomp, j) = A2(S->newY, 1, linknb);
        }
        const double wdepth = A2(S->newY, ncomp, j) - A2(S->z, ncomp, j);
        for (int i = 1; i <= ncomp - 1; ++i) A2(S->oldY, i, j) = wdepth + A2(S->z, i, j);
        mesh_diffusive_backward(S, j);
        for (int i = 1; i <= ncomp; ++i) {
            A2(S->oldY, i, j) = A2(S->newY, i, j);
            /* :602 reads oldY(ncomp, nlinks) with ncomp of the CURRENT reach j */
            if (A2(S->oldY, i, j) < A2(S->oldY, ncomp, nl)) A2(S->oldY, i, j) = A2(S->oldY, ncomp, nl);
        }
    }

    /* tributary flows at the output times :611-633 */
    {
        int ts_ev = 1;
        t = t0 * 60.0;
        while (t <= tfin * 60.0) {
            if ((fmod((t - t0 * 60.) * 60., saveInterval) <= TOL) || (t == tfin * 60.)) {
                for (int j = 1; j <= nl; ++j) {
                    if (!is_mainstem(mstem, nm, j)) {
                        for (int n = 1; n <= nqt; ++n) varr_qtrib[n - 1] = A2P(qtrib_g, n, j, nqt);
                        if (ts_ev <= nev) {
                            EV(q_ev_g, ts_ev, FRNW(j, 1), j) = intp_y(nqt, tarr_qtrib, varr_qtrib, t);
                            EV(q_ev_g, ts_ev, 1, j) = EV(q_ev_g, ts_ev, FRNW(j, 1), j);
                        }
                    }
                }
                ts_ev = ts_ev + 1;
            }
            t = t + dtini / 60.;
        }
    }
    /* qpx = 0 (:636); pool is calloc'ed */
    double maxCelDx = 1.0 / minDx;                                 /* maxCelerity / minDx :638-639 */
    int ts_ev = 1;
    t = t0 * 60.0;
    long guard = 0;
    while (t < tfin * 60.) {                                       /* time loop :655-857 */
        if (++guard > 50000000L || !(dtini > 0.0)) {               /* the reference would spin forever on dtini <= 0 / NaN */
            free(pool); free(mstem); free(tarr_ql); free(tarr_qtrib); free(tarr_db);
            return -5;
        }
        for (int jm = 1; jm <= nm; ++jm) {                         /* predictor, upstream to downstream */
            const int j = mstem[jm - 1], ncomp = FRNW(j, 1);
            if (j == mstem[0]) {                                   /* calculateDT :942-991 */
                dtini = S->cfl / maxCelDx;
                const int a = (int)floor((t - t0 * 60.) / (saveInterval / 60.));
                const int b = (int)floor(((t - t0 * 60.) + dtini / 60.) / (saveInterval / 60.));
                if (b > a) dtini = (a + 1) * (saveInterval) - (t - t0 * 60.) * 60.;
                if (t + dtini / 60. > tfin * 60.) dtini = (tfin * 60. - t) * 60.;
                S->dtini = dtini;
            }
            for (int i = 1; i <= ncomp - 1; ++i) {                 /* lateral inflow at t :660-666 */
                for (int n = 1; n <= nql; ++n) varr_ql[n] = qlat_g[(n - 1) + (size_t)nql * ((size_t)(i - 1) + (size_t)mx * (size_t)(j - 1))];
                varr_ql[0] = qlat_g[(size_t)nql * ((size_t)(i - 1) + (size_t)mx * (size_t)(j - 1))];
                A2(S->lateralFlow, i, j) = intp_y(nql + 1, tarr_ql, varr_ql, t);
            }
            if (FRNW(j, 3) > 0) {                                  /* junction inflow :669-690 */
                A2(S->newQ, 1, j) = 0.0;
                for (int k = 1; k <= FRNW(j, 3); ++k) {
                    const int usrchj = FRNW(j, 3 + k);
                    double q_usrch;
                    if (is_mainstem(mstem, nm, usrchj)) q_usrch = A2(S->newQ, FRNW(usrchj, 1), usrchj);
                    else {
                        for (int n = 1; n <= nqt; ++n) varr_qtrib[n - 1] = A2P(qtrib_g, n, usrchj, nqt);
                        const double tf0 = t + dtini / 60.;
                        q_usrch = intp_y(nqt, tarr_qtrib, varr_qtrib, tf0);
                    }
                    A2(S->newQ, 1, j) = A2(S->newQ, 1, j) + q_usrch;
                }
            } else {
                A2(S->newQ, 1, j) = 0.0;
            }
            A2(S->newQ, 1, j) = A2(S->newQ, 1, j) + A2(S->lateralFlow, 1, j) * A2(S->dx, 1, j);
            mesh_diffusive_forward(S, j);
        }
        for (int jm = nm; jm >= 1; --jm) {                         /* corrector, downstream to upstream :698-752 */
            const int j = mstem[jm - 1], ncomp = FRNW(j, 1);
            if (FRNW(j, 2) >= 0) {
                const int linknb = FRNW(j, 2);
                A2(S->newY, ncomp, j) = A2(S->newY, 1, linknb);
            } else if (dsbc_option == 1) {
                A2(S->newY, ncomp, j) = intp_y(ndb, tarr_db, varr_db, t + dtini / 60.);
                if ((A2(S->newY, ncomp, j) - A2(S->z, ncomp, j)) < mindepth_nstab)
                    A2(S->newY, ncomp, j) = mindepth_nstab + A2(S->z, ncomp, j);
                A2(S->newArea, ncomp, j) = intp_xsec_tab(S, ncomp, j, 1, 2, A2(S->newY, ncomp, j));
            } else if (dsbc_option == 2) {
                A2(S->newY, ncomp, j) = intp_xsec_tab(S, ncomp, j, 10, 1, fabs(A2(S->newQ, ncomp, j)));
                A2(S->newArea, ncomp, j) = intp_xsec_tab(S, ncomp, j, 1, 2, A2(S->newY, ncomp, j));
            }
            mesh_diffusive_backward(S, j);
            if (jm == 1) {
                maxCelDx = 0.;
                for (int i = 1; i <= nm; ++i)
                    for (int kkk = 1; kkk <= FRNW(mstem[i - 1], 1) - 1; ++kkk)
                        maxCelDx = fmax(maxCelDx, A2(S->celerity, kkk, mstem[i - 1]) / A2(S->dx, kkk, mstem[i - 1]));
            }
        }
        t = t + dtini / 60.;
        if ((fmod((t - t0 * 60.) * 60., saveInterval) <= TOL) || (t == tfin * 60.)) {     /* :773-798 */
            if (ts_ev + 1 <= nev) {
                for (int jm = 1; jm <= nm; ++jm) {
                    const int j = mstem[jm - 1], ncomp = FRNW(j, 1);
                    for (int i = 1; i <= ncomp; ++i) {
                        EV(q_ev_g, ts_ev + 1, i, j) = A2(S->newQ, i, j);
                        EV(elv_ev_g, ts_ev + 1, i, j) = A2(S->newY, i, j);
                        EV(depth_ev_g, ts_ev + 1, i, j) = EV(elv_ev_g, ts_ev + 1, i, j) - A2(S->z, i, j);
                    }
                    for (int k = 1; k <= FRNW(j, 3); ++k) {
                        const int usrchj = FRNW(j, 3 + k);
                        if (!is_mainstem(mstem, nm, usrchj)) {
                            const double wdepth = A2(S->newY, 1, j) - A2(S->z, 1, j);
                            EV(elv_ev_g, ts_ev + 1, FRNW(usrchj, 1), usrchj) = A2(S->newY, 1, j);
                            EV(depth_ev_g, ts_ev + 1, FRNW(usrchj, 1), usrchj) = wdepth;
                        }
                    }
                }
            }
            ts_ev = ts_ev + 1;
        }
        if (t == t0 + dtini / 60.) {                               /* initial state :801-821 (true only when t0 == 0) */
            for (int jm = 1; jm <= nm; ++jm) {
                const int j = mstem[jm - 1], ncomp = FRNW(j, 1);
                for (int i = 1; i <= ncomp; ++i) {
                    EV(q_ev_g, 1, i, j) = A2(S->oldQ, i, j);
                    EV(elv_ev_g, 1, i, j) = A2(S->oldY, i, j);
                    EV(depth_ev_g, 1, i, j) = EV(elv_ev_g, 1, i, j) - A2(S->z, i, j);
                }
                for (int k = 1; k <= FRNW(j, 3); ++k) {
                    const int usrchj = FRNW(j, 3 + k);
                    if (!is_mainstem(mstem, nm, usrchj)) {
                        const double wdepth = A2(S->oldY, 1, j) - A2(S->z, 1, j);
                        EV(elv_ev_g, 1, FRNW(usrchj, 1), usrchj) = A2(S->oldY, 1, j);
                        EV(depth_ev_g, 1, FRNW(usrchj, 1), usrchj) = wdepth;
                    }
                }
            }
        }
        memcpy(S->oldY, S->newY, n2 * sizeof(double));             /* :824-830 */
        memcpy(S->oldQ, S->newQ, n2 * sizeof(double));
        memcpy(S->oldArea, S->newArea, n2 * sizeof(double));
        for (size_t i = 0; i < n2; ++i) { S->newY[i] = -999; S->newQ[i] = -999; S->newArea[i] = -999; S->pere[i] = -999; }
    }

    if (*cwnrow_g > 0) {                                           /* crosswalk to the unrefactored hydrofabric :837-903 */
        const int cwn = *cwnrow_g;
        (void)cwncol_g;
        const double equiv_one = (double)0.99f;
        double* tq = (double*)malloc(sizeof(double) * nout * 2);
        double* used = (double*)malloc(sizeof(double) * n2);
        int* flag = (int*)malloc(sizeof(int) * n2);
        if (!tq || !used || !flag) { free(tq); free(used); free(flag); free(pool); free(mstem); free(tarr_ql); free(tarr_qtrib); free(tarr_db); return -3; }
        double* te = tq + nout;
        memcpy(tq, q_ev_g, nout * sizeof(double));
        memcpy(te, elv_ev_g, nout * sizeof(double));
        for (size_t i = 0; i < nout; ++i) { q_ev_g[i] = 0.0; elv_ev_g[i] = 0.0; }
#define CW(r, c) crosswalk_g[((r) - 1) + (size_t)cwn * (size_t)((c) - 1)]
        for (int ts = 1; ts <= nev; ++ts) {
            for (size_t i = 0; i < n2; ++i) { used[i] = 0.0; flag[i] = 0; }
            for (int cwrow = 1; cwrow <= cwn; ++cwrow) {
                const int ri = (int)CW(cwrow, 1), rj = (int)CW(cwrow, 2), nlnk = (int)CW(cwrow, 3);
                const double slopeQ = (EV(tq, ts, ri + 1, rj) - EV(tq, ts, ri, rj)) / A2(rdx_ar_g, ri, rj);
                const double intcQ = EV(tq, ts, ri, rj);
                const double slopeD = ((EV(te, ts, ri + 1, rj) - A2(S->z, ri + 1, rj)) - (EV(te, ts, ri, rj) - A2(S->z, ri, rj))) / A2(rdx_ar_g, ri, rj);
                const double intcD = EV(te, ts, ri, rj) - A2(S->z, ri, rj);
                double dst_lnk = 0.0;
                for (int lnk = 1; lnk <= nlnk; ++lnk) {
                    const int oi = (int)CW(cwrow, 4 + 3 * (lnk - 1)), oj = (int)CW(cwrow, 5 + 3 * (lnk - 1));
                    const double lfrac = CW(cwrow, 6 + 3 * (lnk - 1));
                    const double dst_top = dst_lnk;
                    dst_lnk = dst_lnk + A2(dx_ar_g, oi, oj) * lfrac;
                    const double dst_btm = dst_lnk;
                    A2(used, oi, oj) = A2(used, oi, oj) + lfrac;
                    if (A2(used, oi, oj) < equiv_one) A2(flag, oi, oj) = A2(flag, oi, oj) + 1;
                    if ((A2(used, oi, oj) >= equiv_one) && (A2(flag, oi, oj) == 0)) {
                        EV(q_ev_g, ts, oi, oj) = intcQ + slopeQ * dst_top;
                        EV(q_ev_g, ts, oi + 1, oj) = intcQ + slopeQ * dst_btm;
                        EV(elv_ev_g, ts, oi, oj) = intcD + slopeD * dst_top + A2(z_thalweg_g, oi, oj);
                        EV(elv_ev_g, ts, oi + 1, oj) = intcD + slopeD * dst_btm + A2(z_thalweg_g, oi + 1, oj);
                    } else if ((A2(used, oi, oj) < equiv_one) && (A2(flag, oi, oj) == 1)) {
                        EV(q_ev_g, ts, oi, oj) = intcQ + slopeQ * dst_top;
                        EV(elv_ev_g, ts, oi, oj) = intcD + slopeD * dst_top + A2(z_thalweg_g, oi, oj);
                    } else if ((A2(used, oi, oj) >= equiv_one) && (A2(flag, oi, oj) >= 1)) {
                        EV(q_ev_g, ts, oi + 1, oj) = intcQ + slopeQ * dst_btm;
                        EV(elv_ev_g, ts, oi + 1, oj) = intcD + slopeD * dst_btm + A2(z_thalweg_g, oi + 1, oj);
                        A2(flag, oi, oj) = 0;
                    }
                }
            }
        }
#undef CW
        free(tq); free(used); free(flag);
    }
    free(pool); free(mstem); free(tarr_ql); free(tarr_qtrib); free(tarr_db);
    return 0;
}
