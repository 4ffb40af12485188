/*
 * diffusive_device.cuh -- the diffusive-wave mainstem solver of the B200 routing path (sm_100a device code).
 *
 * Replaces the Fortran module `diffusive` (/root/reference/src/kernel/diffusive/diffusive.f90) behind its C entry
 * point c_diffnw (pydiffusive.f90:8-52): Crank-Nicolson + Hermite-interpolation discharge sweep per reach
 * (mesh_diffusive_forward :1108-1355), water-surface sweep by safeguarded Newton (mesh_diffusive_backward :1357-1553,
 * rtsafe :1555-1662, funcd_diffdepth :1664-1711), CFL-adaptive time step (calculateDT :942-991) and the synthetic
 * cross-section tables (readXsection :2093-2443).  All arithmetic is IEEE binary64 as in the Fortran, operand order kept
 * expression by expression; x**y is trt_pow64_det (include/trt_detmath64.h), bit-identical on CPU and GPU; compile with
 * -fmad=false (gfortran -O2 emits no FMA on baseline x86-64).
 *
 * NOT a translation.  What is different, and why:
 *   * ONE kernel launch runs the whole time loop of a domain in one CTA (`dw_time_loop`); the reference's loop over
 *     tailwater domains (compute.py:1764, "TODO by-network parallel loop") becomes the grid: one CTA per domain.
 *   * The discharge sweep has no dependency between reaches inside a step: the elimination coefficients are functions of
 *     the previous step only, and the new junction inflow is written into node 1 AFTER the back substitution
 *     (diffusive.f90:1309-1318).  All reaches are therefore swept concurrently (thread per reach for the two-term
 *     recurrences, thread per node for the Hermite/CN coefficients that hold the powers and divisions).
 *   * The water-surface sweep IS a dependency chain from the tailwater to the heads (y of node i-1 needs y of node i);
 *     the arms above a confluence are independent chains and run in different warps.  Only the Newton iterations stay on
 *     the chain: the normal-depth search, the bracket and the friction slope at its ends (dw_bracket) run thread-per-node
 *     before it, conveyance / area / width / roughness interpolation and the celerity and diffusivity of every node (the
 *     pow-heavy part, :1440-1496) thread-per-node after it.
 *   * Table look-ups: the reference scans 501 rows linearly three times per r_interpol (maxval, minval, search) and
 *     bisects in intp_xsec_tab.  Elevation tables are piecewise uniform by construction (:2258-2266), so the row is
 *     computed arithmetically and verified against the table (`dw_locate_hint`, same result as the bisection for a
 *     monotone column); the uniform-flow column is searched from the row of the previous time step.
 *   * Tables are stored column-contiguous per node, [node][8 columns][512] (4 KB per column, 32 KB per node), so a node's
 *     working set is one contiguous block (L2-resident: 126 MB of L2 hold ~4,000 nodes; TMA-stageable).
 *
 * Every function below is `TRT_HD`: the same source compiles for the device (nvcc) and, in tests/ only, for the host
 * (tests/native/diffusive_replica.cpp) so that the solver logic can be checked against oracle/diffusive_oracle.c without a
 * GPU.  The product library never runs the host instantiation.
 */
#pragma once
#include <math.h>
#include <stddef.h>

#include "../../include/trt_detmath64.h"

#if defined(__CUDACC__)
#define TRT_HD __host__ __device__ __forceinline__
#else
#define TRT_HD static inline
#endif

/* thread geometry of the cooperating group that runs one domain: a CTA on the device, one thread on the host */
#if defined(__CUDA_ARCH__)
#define DW_TID ((int)threadIdx.x)
#define DW_NT ((int)blockDim.x)
#define DW_SYNC() __syncthreads()
#define DW_WARP ((int)(threadIdx.x >> 5))
#define DW_NWARP ((int)(blockDim.x >> 5))
#define DW_LANE ((int)(threadIdx.x & 31))
#elif defined(DW_EMULATE_CTA)
/* tests only (tests/native/diffusive_cta_main.cpp): a CTA emulated by host threads that meet at a barrier, to run the SPMD
 * phase structure of dw_time_loop under ThreadSanitizer -- a missing DW_SYNC() shows up as a data race */
extern thread_local int dw_emu_tid;
extern int dw_emu_nt;
void dw_emu_sync();
#define DW_TID dw_emu_tid
#define DW_NT dw_emu_nt
#define DW_SYNC() dw_emu_sync()
#define DW_WARP (dw_emu_tid >> 5)
#define DW_NWARP (dw_emu_nt >> 5)
#define DW_LANE (dw_emu_tid & 31)
#else
#define DW_TID 0
#define DW_NT 1
#define DW_SYNC() ((void)0)
#define DW_WARP 0
#define DW_NWARP 1
#define DW_LANE 0
#endif

namespace trtdw {

enum { NEL = 501, LD = 512, NCOL = 8 };
/* columns kept of xsec_tab(11, nel, ., .): the ones the time loop reads */
enum { C_ELEV = 0,   /* xsec_tab(1)  water-surface elevation of the row */
       C_AREA = 1,   /* xsec_tab(2)  flow area */
       C_PERI = 2,   /* xsec_tab(3)  wetted perimeter */
       C_CONV = 3,   /* xsec_tab(5)  conveyance */
       C_TOPW = 4,   /* xsec_tab(6)  top width */
       C_DKDA = 5,   /* xsec_tab(9)  dK/dA */
       C_QNRM = 6,   /* xsec_tab(10) uniform-flow discharge = conveyance * sqrt(slope) */
       C_SKK = 7 };  /* xsec_tab(11) 1 / compound Manning n */

/* REAL(4) literals of the Fortran, promoted the way gfortran promotes them */
#define DW_F(x) ((double)(x##f))

struct Dom {
    int mx, nl, nm, nev, nql, nqt, ndb, dsbc_option, frnw_col;
    double cfl, C_llm, D_llm, D_ulm, q_llm, so_llm, theta;
    double dtini_given, dtini_min, t0, tfin, saveInterval, dt_ql, dt_db, dt_qtrib;
    const int* frnw;            /* (nl, frnw_col) column-major, as given */
    const int* mstem;           /* [nm] 1-based indices of the mainstem reaches, upstream to downstream */
    const unsigned char* is_main; /* [nl + 1] */
    /* node geometry as given (mx, nl) */
    const double *z_in, *bo_in, *traps_in, *tw_in, *twcc_in, *mann_in, *manncc_in, *dx_in;
    /* surveyed ("natural") cross sections, mxnbathy > 0: vertices (mxnbathy, mx, nl), counts (mx, nl) */
    int mxnbathy;
    const double *x_bathy, *z_bathy, *mann_bathy;
    const int* size_bathy;
    double* mann15;             /* (mxnbathy, mx, nl) min(n, 0.15)**1.5 of every vertex (dw_nat_prep) */
    /* forcing */
    const double* qlat;         /* (nql, mx, nl) */
    const double* qtrib;        /* (nqt, nl) */
    const double* dbcd;         /* (ndb) */
    const double* iniq;         /* (mx, nl) */
    const double *tarr_ql, *tarr_qtrib, *tarr_db;   /* [nql + 1], [nqt], [ndb] minutes (diffusive.f90:512-540) */
    double* rmax;               /* [nl] celerity / dx maximum of a reach (for the CFL bound) */
    /* state, all (mx, nl) column-major */
    double *z, *dx, *bo, *pere, *qp, *qpx, *sk, *co, *oldQ, *newQ, *oldArea, *newArea, *oldY, *newY, *lateralFlow,
        *celerity, *diffusivity, *celerity2, *diffusivity2, *eei, *ffi, *exi, *fxi, *c_ppi, *c_qqi, *c_rri, *c_ssi, *c_sxi;
    double *b_ynorm, *b_x1, *b_x2, *b_sf1, *b_sf2;   /* (mx, nl) Newton bracket of every node, see dw_bracket */
    int nlev;                   /* water-surface sweep: mainstem reaches grouped by their distance (in reaches) from the tailwater */
    const int* lvl_ptr;         /* [nlev + 1] */
    const int* lvl_reach;       /* [nm] 1-based reach indices, level by level */
    double* varr_db;            /* [ndb] tailwater elevation series (dsbc_option 1) */
    int* hint_q;                /* (mx, nl) row of the last uniform-flow look-up */
    double* tab;                /* [(j-1)*mx + (i-1)][NCOL][LD] */
    double* tabmin;             /* [(j-1)*mx + (i-1)][NCOL] minimum of every column (r_interpol's `minval(y)`) */
    double* scal;               /* [8] per-domain scalars shared by the cooperating threads: [0] maxCelDx */
    int* status;                /* [1] 0 ok, -5 time step collapsed */
    double *q_ev, *elv_ev, *depth_ev;   /* (nev, mx, nl) */
};

#define DW_A2(p, i, j) ((p)[((i) - 1) + (size_t)((j) - 1) * (size_t)D.mx])
#define DW_FRNW(j, c) (D.frnw[((j) - 1) + (size_t)((c) - 1) * (size_t)D.nl])
#define DW_EV(p, ts, i, j) ((p)[((ts) - 1) + (size_t)D.nev * ((size_t)((i) - 1) + (size_t)D.mx * (size_t)((j) - 1))])

TRT_HD const double* dw_col(const Dom& D, int i, int j, int col)
{
    return D.tab + (((size_t)(j - 1) * D.mx + (size_t)(i - 1)) * NCOL + (size_t)col) * LD;
}
TRT_HD double* dw_col_w(const Dom& D, int i, int j, int col)
{
    return D.tab + (((size_t)(j - 1) * D.mx + (size_t)(i - 1)) * NCOL + (size_t)col) * LD;
}

/* ---- look-ups ------------------------------------------------------------------------------------------------------ */
/* `locate` (diffusive.f90:2701-2742) on an ascending column xx[0..n-1]: number of rows <= x, i.e. the 1-based row jl with
 * xx(jl) <= x < xx(jl+1) (0 below the table, n above), with the two end-point special cases.  Searched from `hint` (a
 * 1-based guess) by galloping, then bisection: for a monotone column the result is the bisection's. */
TRT_HD int dw_locate_hint(const double* xx, int n, double x, int hint)
{
    int lo, hi;                      /* invariant: xx(lo) <= x < xx(hi), with xx(0) = -inf, xx(n+1) = +inf (1-based) */
    if (!(x >= xx[0])) lo = 0, hi = 1;            /* below the table, or NaN (every comparison false: the bisection ends at 0) */
    else if (x >= xx[n - 1]) lo = n, hi = n + 1;
    else {
        int g = hint < 1 ? 1 : (hint > n - 1 ? n - 1 : hint);
        int step = 1;
        if (xx[g - 1] <= x) {
            lo = g; hi = g + 1;
            while (hi <= n && xx[hi - 1] <= x) { lo = hi; step <<= 1; hi = lo + step; }
            if (hi > n) hi = n;                    /* xx(n) > x holds here */
        } else {
            hi = g; lo = g - 1;
            while (lo >= 1 && xx[lo - 1] > x) { hi = lo; step <<= 1; lo = hi - step; }
            if (lo < 1) lo = 1;                    /* xx(1) <= x holds here */
        }
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (xx[mid - 1] <= x) lo = mid; else hi = mid;
        }
    }
    if (x == xx[0]) return 1;
    if (x == xx[n - 1]) return n - 1;
    return lo;
}

/* LInterpol :2650-2669 */
TRT_HD double dw_linterpol(double x1, double y1, double x2, double y2, double x)
{
    if (fabs(x2 - x1) < DW_F(0.0001)) return 0.5 * (y1 + y2);
    return (y2 - y1) / (x2 - x1) * (x - x1) + y1;
}

/* intp_y :2671-2699 on a short time series */
TRT_HD double dw_intp_y(int nrow, const double* xarr, const double* yarr, size_t ystride, double x, int hint)
{
    int irow = dw_locate_hint(xarr, nrow, x, hint);
    if (irow == 0) irow = 1;
    if (irow == nrow) irow = nrow - 1;
    return dw_linterpol(xarr[irow - 1], yarr[(size_t)(irow - 1) * ystride], xarr[irow], yarr[(size_t)irow * ystride], x);
}

/* arithmetic guess of the row of elevation `el` in a node's elevation column: rows 6..NEL are uniform (:2263-2266; surveyed
 * sections: every row, :1826).  The two numbers that define the grid are read once per node and kept in registers. */
struct ElevGrid { double e6, inc; };
TRT_HD ElevGrid dw_grid(const double* elev)
{
    ElevGrid g;
    g.e6 = elev[5];
    g.inc = elev[NEL - 1] - elev[NEL - 2];
    return g;
}
TRT_HD int dw_guess(const ElevGrid& g, double el)
{
    if (!(g.inc > 0.0)) return NEL / 2;
    const double r = (el - g.e6) / g.inc;
    if (!(r > -6.0)) return 1;
    if (r > (double)NEL) return NEL - 1;
    return 6 + (int)r;
}
TRT_HD int dw_elev_guess(const double* elev, double el) { return dw_guess(dw_grid(elev), el); }

/* intp_xsec_tab(i, j, nel, 1, ycol, x) :1713-1748 for an elevation argument */
TRT_HD int dw_row_of_elev(const double* elev, double el)
{
    int irow = dw_locate_hint(elev, NEL, el, dw_elev_guess(elev, el));
    if (irow == 0) irow = 1;
    if (irow == NEL) irow = NEL - 1;
    return irow;
}
TRT_HD double dw_interp_row(const double* xa, const double* ya, int irow, double x)
{
    return dw_linterpol(xa[irow - 1], ya[irow - 1], xa[irow], ya[irow], x);
}

/* r_interpol :2553-2594 on an ascending x column.  The Fortran takes the FIRST interval k with
 * (x(k) - xrt) * (x(k+1) - xrt) <= 0 when min(x) <= xrt <= max(x): for ascending x that is k = max(1, #rows < xrt)
 * (an xrt that sits exactly on row k+1 belongs to interval k, unlike `locate`).  Above the table: linear extrapolation of
 * the last interval; below (or NaN): the minimum of y.  `sq` selects x(k) = (elev(k) - zsub)**2 (the conveyance look-up by
 * squared depth, :1444-1448); elev(k) >= zsub for every row, so the squared column is ascending too. */
TRT_HD double dw_r_interpol(const double* elev, const double* ya, double ymin, bool sq, double zsub, double xrt, double prev)
{
#define DW_XV(k) (sq ? ((elev[(k) - 1] - zsub) * (elev[(k) - 1] - zsub)) : elev[(k) - 1])
    const double xmin = DW_XV(1), xmax = DW_XV(NEL);
    if (xrt <= xmax && xrt >= xmin) {
        /* rows strictly below xrt: search in the elevation domain (x -> (x - zsub)**2 is monotone for x >= zsub), then
         * settle the row with the comparisons the Fortran makes on the column it scans */
        const double el = sq ? (zsub + sqrt(xrt)) : xrt;
        int k = dw_locate_hint(elev, NEL, el, dw_elev_guess(elev, el));
        if (k < 1) k = 1;
        if (k > NEL - 1) k = NEL - 1;
        while (k > 1 && !(DW_XV(k) < xrt)) --k;              /* first interval whose left end is not above xrt ... */
        while (k < NEL - 1 && DW_XV(k + 1) < xrt) ++k;        /* ... and whose right end is not below it */
        /* now (k == 1 or x(k) < xrt) and x(k+1) >= xrt: the first interval the Fortran's scan accepts */
        if ((DW_XV(k) - xrt) * (DW_XV(k + 1) - xrt) <= 0.0)
            return (xrt - DW_XV(k)) / (DW_XV(k + 1) - DW_XV(k)) * (ya[k] - ya[k - 1]) + ya[k - 1];
        return prev;                                         /* no interval matched: the output argument is left alone */
    } else if (xrt >= xmax) {
        return (xrt - DW_XV(NEL - 1)) / (DW_XV(NEL) - DW_XV(NEL - 1)) * (ya[NEL - 1] - ya[NEL - 2]) + ya[NEL - 2];
    }
    return ymin;
#undef DW_XV
}

/* ---- synthetic cross-section tables: readXsection :2093-2443 -------------------------------------------------------- */
struct Xs {                 /* the three sub-sections (left overbank, main channel, right overbank) of one node */
    double X[3][8], Y[3][8];    /* 1-based node k of sub-section s is X[s][k] (k = 1..7) */
    int num[3];
    double man[3];
    double el_min, el_incr, e5;
};

TRT_HD void dw_xs_setup(Xs& S, double z_g, double bo_g, double traps_g, double tw_g, double twcc_g, double mann, double manncc)
{
    const double timesDepth = 4.0;
    const double hbf = (tw_g - bo_g) / (2.0 * traps_g);                       /* :2150 */
    double xcs[9], ycs[9];
    xcs[1] = 0.0;                        ycs[1] = z_g + timesDepth * hbf;     /* :2156-2193 */
    xcs[2] = 0.0;                        ycs[2] = z_g + hbf;
    xcs[3] = (twcc_g - tw_g) / 2.0;      ycs[3] = z_g + hbf;
    xcs[4] = xcs[3] + traps_g * hbf;     ycs[4] = z_g;
    xcs[5] = xcs[4] + bo_g;              ycs[5] = z_g;
    xcs[6] = xcs[5] + traps_g * hbf;     ycs[6] = z_g + hbf;
    xcs[7] = twcc_g;                     ycs[7] = z_g + hbf;
    xcs[8] = xcs[7];                     ycs[8] = z_g + timesDepth * hbf;
    double el_min = 99999., el_max = -99999.;                                 /* :2208-2213 over nodes 2..8 */
    for (int i = 2; i <= 8; ++i) {
        if (ycs[i] < el_min) el_min = ycs[i];
        if (ycs[i] > el_max) el_max = ycs[i];
    }
    const double el_range = (el_max - el_min) * 2.0;
    const double wall = el_min + el_range + 1.;
    for (int s = 0; s < 3; ++s) for (int k = 0; k < 8; ++k) { S.X[s][k] = 0.0; S.Y[s][k] = 0.0; }
    /* left overbank :2217-2224 */
    S.X[0][1] = xcs[1]; S.Y[0][1] = wall;
    S.X[0][2] = xcs[1]; S.Y[0][2] = ycs[1];
    S.X[0][3] = xcs[2]; S.Y[0][3] = ycs[2];
    S.X[0][4] = xcs[3]; S.Y[0][4] = ycs[3];
    S.X[0][5] = xcs[3]; S.Y[0][5] = wall;
    /* main channel :2226-2237, with the 1 cm notch at its centre :2253-2254 */
    S.X[1][1] = xcs[3]; S.Y[1][1] = wall;
    S.X[1][2] = xcs[3]; S.Y[1][2] = ycs[3];
    S.X[1][3] = xcs[4]; S.Y[1][3] = ycs[4];
    S.X[1][5] = xcs[5]; S.Y[1][5] = ycs[5];
    S.X[1][6] = xcs[6]; S.Y[1][6] = ycs[6];
    S.X[1][7] = xcs[6]; S.Y[1][7] = wall;
    S.X[1][4] = (S.X[1][3] + S.X[1][5]) / 2.0;
    S.Y[1][4] = S.Y[1][3] - DW_F(0.01);
    /* right overbank :2239-2247 */
    S.X[2][2] = xcs[6]; S.Y[2][2] = ycs[6];
    S.X[2][3] = xcs[7]; S.Y[2][3] = ycs[7];
    S.X[2][4] = xcs[8]; S.Y[2][4] = ycs[8];
    S.X[2][1] = S.X[2][2]; S.Y[2][1] = wall;
    S.X[2][5] = S.X[2][4]; S.Y[2][5] = wall;
    S.num[0] = 5; S.num[1] = 7; S.num[2] = 5;
    S.man[0] = manncc; S.man[1] = mann; S.man[2] = manncc;
    S.el_min = S.Y[1][4];                                                     /* :2256 */
    S.e5 = S.el_min + DW_F(0.01);                                             /* :2262 */
    S.el_incr = el_range / (double)(float)(NEL - 6.0f);                      /* :2263 */
}

/* elevation of table row j (1-based) :2258-2266, with the offset of the first row :2283-2285 */
TRT_HD double dw_xs_elev(const Xs& S, int j)
{
    double e;
    if (j == 1) e = S.el_min;
    else if (j == 2) e = S.el_min + (double)(0.01f / 4.f);
    else if (j == 3) e = S.el_min + (double)(0.01f / 4.f * 2.f);
    else if (j == 4) e = S.el_min + (double)(0.01f / 4.f * 3.f);
    else if (j == 5) e = S.e5;
    else e = S.e5 + S.el_incr * (double)(float)(j - 5);
    if (fabs(e - S.el_min) < DW_F(1e-8)) e = e + DW_F(0.00001);
    return e;
}

TRT_HD double dw_dist(double x1, double y1, double x2, double y2)
{
    return sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + DW_F(1.e-32));
}

/* wetted area, perimeter, top width and conveyance of sub-section s at elevation el_now :2286-2356 */
TRT_HD void dw_xs_eval(const Xs& S, int s, double el_now, double& area, double& peri, double& topw, double& conv)
{
    const double* xcs = S.X[s];
    const double* ycs = S.Y[s];
    const int num = S.num[s];
    double cal_area = 0., cal_peri = 0., cal_topW = 0.;
    int i_find = 0, i_start = -999;
    for (int i = 1; i <= num - 1; ++i) {
        const double ya = ycs[i], yb = ycs[i + 1];
        if ((el_now <= ya) && (el_now > yb) && (i_find == 0)) { i_find = 1; i_start = i; }
        if ((el_now > ya) && (el_now <= yb) && (i_find == 1)) {
            i_find = 0;
            /* one wetted pocket from segment i_start to segment i (the Fortran collects them first and sums them in the
             * same order afterwards) */
            const int i1 = i_start, i2 = i;
            double x1 = xcs[i1], x2 = xcs[i1 + 1], y1 = ycs[i1], y2 = ycs[i1 + 1];
            const double x_start = (y1 == y2) ? x1 : x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
            x1 = xcs[i2]; x2 = xcs[i2 + 1]; y1 = ycs[i2]; y2 = ycs[i2 + 1];
            const double x_end = (y1 == y2) ? x1 : x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
            cal_topW = x_end - x_start + cal_topW;
            double multi_area = 0.0, perim = 0.0;
            for (int k = i1 + 1; k <= i2 - 1; ++k) {
                multi_area = multi_area + fabs(0.5 * (xcs[k + 1] - xcs[k]) * (el_now - ycs[k] + el_now - ycs[k + 1]));
                perim = perim + dw_dist(xcs[k], ycs[k], xcs[k + 1], ycs[k + 1]);
            }
            cal_area = cal_area + fabs(0.5 * (xcs[i1 + 1] - x_start) * (el_now - ycs[i1 + 1])) + multi_area
                     + fabs(0.5 * (xcs[i2] - x_end) * (el_now - ycs[i2]));
            cal_peri = cal_peri + dw_dist(x_start, el_now, xcs[i1 + 1], ycs[i1 + 1]) + perim
                     + dw_dist(x_end, el_now, xcs[i2], ycs[i2]);
            if (i1 == 1) cal_peri = cal_peri - dw_dist(x_start, el_now, xcs[i1 + 1], ycs[i1 + 1]);
            if (i2 == (num - 1)) cal_peri = cal_peri - dw_dist(x_end, el_now, xcs[i2], ycs[i2]);
        }
    }
    area = cal_area; peri = cal_peri; topw = cal_topW;
    const double redi = area / peri;
    conv = 1. / S.man[s] * area * trt_pow64_det(redi, (double)(2.f / 3.f));
    if (peri <= DW_F(1e-8)) conv = 0.0;
}

struct XsRow { double el, area, peri, conv, topw, skk; };

TRT_HD XsRow dw_xs_row(const Xs& S, int j)
{
    XsRow r;
    double a[3], p[3], t[3], c[3];
    const double el_now = dw_xs_elev(S, j);
    for (int s = 0; s < 3; ++s) dw_xs_eval(S, s, el_now, a[s], p[s], t[s], c[s]);
    r.el = el_now;
    r.area = a[0] + a[1] + a[2];                                              /* :2411-2420 */
    r.peri = p[0] + p[1] + p[2];
    r.conv = c[0] + c[1] + c[2];
    r.topw = fabs(t[0]) + fabs(t[1]) + fabs(t[2]);
    const double compoundMann = sqrt((fabs(p[0]) * (S.man[0] * S.man[0]) + fabs(p[1]) * (S.man[1] * S.man[1]) +
                                      fabs(p[2]) * (S.man[2] * S.man[2])) / (fabs(p[0]) + fabs(p[1]) + fabs(p[2])));
    r.skk = 1. / compoundMann;
    return r;
}

/* table pass 1: one (node, row) pair -- elevation, area, perimeter, conveyance, top width, 1/n; row 1 also lowers the
 * node's bed to the notch (z(k, num_reach) = el_min, :2428) */
TRT_HD void dw_table_pass1(Dom& D, int i, int j, int row)
{
    Xs S;
    /* readXsection receives 1/skLeft etc. with sk = 1/n (:470-476): a double rounding that is kept */
    const double mann = 1.0 / (1.0 / DW_A2(D.mann_in, i, j)), manncc = 1.0 / (1.0 / DW_A2(D.manncc_in, i, j));
    dw_xs_setup(S, DW_A2(D.z_in, i, j), DW_A2(D.bo_in, i, j), DW_A2(D.traps_in, i, j), DW_A2(D.tw_in, i, j),
                DW_A2(D.twcc_in, i, j), mann, manncc);
    const XsRow r = dw_xs_row(S, row);
    dw_col_w(D, i, j, C_ELEV)[row - 1] = r.el;
    dw_col_w(D, i, j, C_AREA)[row - 1] = r.area;
    dw_col_w(D, i, j, C_PERI)[row - 1] = r.peri;
    dw_col_w(D, i, j, C_CONV)[row - 1] = r.conv;
    dw_col_w(D, i, j, C_TOPW)[row - 1] = r.topw;
    dw_col_w(D, i, j, C_SKK)[row - 1] = r.skk;
    if (row == 1) DW_A2(D.z, i, j) = S.el_min;
}

/* table pass 2 (after pass 1 of the whole domain): dK/dA (:2395-2402) and the uniform-flow column (:487-506) */
TRT_HD void dw_table_pass2(Dom& D, int i, int j, int row)
{
    const double* A = dw_col(D, i, j, C_AREA);
    const double* K = dw_col(D, i, j, C_CONV);
    if (D.mxnbathy == 0) {                 /* surveyed sections: dw_nat_smooth has written the column */
        double dkda;
        if (row == 1) dkda = K[0] / A[0];
        else dkda = (K[row - 1] - K[row - 2]) / (A[row - 1] - A[row - 2]);
        dw_col_w(D, i, j, C_DKDA)[row - 1] = dkda;
    }
    const int ncomp = DW_FRNW(j, 1);
    double slope;
    if (i < ncomp) slope = (DW_A2(D.z, i, j) - DW_A2(D.z, i + 1, j)) / DW_A2(D.dx, i, j);
    else slope = (DW_A2(D.z, i - 1, j) - DW_A2(D.z, i, j)) / DW_A2(D.dx, i - 1, j);
    if (slope <= D.so_llm) slope = D.so_llm;
    dw_col_w(D, i, j, C_QNRM)[row - 1] = K[row - 1] * trt_pow64_det(slope, 0.50);
}

/* ---- surveyed cross sections: readXsection_natural_mann_vertices :1756-2091 ------------------------------------------ */
#define DW_BATHY(p, ic, i, j) ((p)[((ic) - 1) + (size_t)D.mxnbathy * ((size_t)((i) - 1) + (size_t)D.mx * (size_t)((j) - 1))])

/* vertex ic (1-based) of the closed polygon of node (i, j): the surveyed points 2..nb+1 between two vertical walls */
struct NatXs { int nb; double x0, el_min, el_range, el_incr, wall; };

TRT_HD void dw_nat_setup(const Dom& D, int i, int j, NatXs& S)
{
    const double timesDepth = 4.0;
    S.nb = D.size_bathy[(i - 1) + (size_t)(j - 1) * D.mx];
    S.x0 = DW_BATHY(D.x_bathy, 1, i, j);
    double el_min = 99999., el_max = -99999.;                                 /* :1818-1823 */
    for (int ic = 1; ic <= S.nb; ++ic) {
        const double y = DW_BATHY(D.z_bathy, ic, i, j) * 1.0;
        if (y < el_min) el_min = y;
        if (y > el_max) el_max = y;
    }
    S.el_min = el_min;
    S.el_range = (el_max - el_min) * timesDepth;
    S.el_incr = S.el_range / (double)(float)(NEL - 1.0f);
    S.wall = el_min + S.el_range + 1.0;
}
/* coordinates and n**1.5 of polygon vertex k = 1 .. nb + 2 (:1799-1835) */
TRT_HD double dw_nat_x(const Dom& D, const NatXs& S, int i, int j, int k)
{
    const int ic = k <= 1 ? 1 : (k >= S.nb + 2 ? S.nb : k - 1);
    return (-S.x0 + DW_BATHY(D.x_bathy, ic, i, j)) * 1.0;
}
TRT_HD double dw_nat_y(const Dom& D, const NatXs& S, int i, int j, int k)
{
    if (k <= 1 || k >= S.nb + 2) return S.wall;
    return DW_BATHY(D.z_bathy, k - 1, i, j) * 1.0;
}
TRT_HD double dw_nat_m15(const Dom& D, const NatXs& S, int i, int j, int k)
{
    if (k <= 1 || k >= S.nb + 1) return 0.0;      /* manncs(1) = manncs(num-1) = manncs(num) = 0 and 0**1.5 = 0 */
    return DW_BATHY(D.mann15, k - 1, i, j);
}

/* thread per (node, vertex): the roughness term of every polygon side, min(n, 0.15)**1.5 (:1811-1814, :2047, :2084) */
TRT_HD void dw_nat_prep(Dom& D, int i, int j, int ic)
{
    double m = DW_BATHY(D.mann_bathy, ic, i, j);
    if (m > DW_F(0.15)) m = DW_F(0.15);
    DW_BATHY(D.mann15, ic, i, j) = trt_pow64_det(m, 1.50);
}

/* thread per (node, table row): one elevation of the surveyed section (:1840-1925), conveyance not yet made monotone */
TRT_HD void dw_nat_pass1(Dom& D, int i, int j, int row)
{
    NatXs S;
    dw_nat_setup(D, i, j, S);
    const int num = S.nb + 2;
    double el_now = S.el_min + (double)(float)(row - 1) * S.el_incr;
    if (fabs(el_now - S.el_min) < DW_F(1e-8)) el_now = el_now + DW_F(0.00001);
    double cal_area = 0.0, cal_peri = 0.0, cal_topW = 0.0, cal_equiv_mann = 0.0;
    int i_find = 0, i_start = -999;
    double ya = dw_nat_y(D, S, i, j, 1);
    for (int ic = 1; ic <= num - 1; ++ic) {
        const double yb = dw_nat_y(D, S, i, j, ic + 1);
        if ((el_now <= ya) && (el_now > yb) && (i_find == 0)) { i_find = 1; i_start = ic; }
        if ((el_now > ya) && (el_now <= yb) && (i_find == 1)) {
            i_find = 0;
            const int i1 = i_start, i2 = ic;
            double x1 = dw_nat_x(D, S, i, j, i1), x2 = dw_nat_x(D, S, i, j, i1 + 1), y1 = dw_nat_y(D, S, i, j, i1),
                   y2 = dw_nat_y(D, S, i, j, i1 + 1);
            const double x_start = (y1 == y2) ? x1 : x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
            const double xs1 = x2, ys1 = y2;                                  /* vertex i1 + 1 */
            x1 = dw_nat_x(D, S, i, j, i2); x2 = dw_nat_x(D, S, i, j, i2 + 1); y1 = dw_nat_y(D, S, i, j, i2);
            y2 = dw_nat_y(D, S, i, j, i2 + 1);
            const double x_end = (y1 == y2) ? x1 : x1 + (el_now - y1) / (y2 - y1) * (x2 - x1);
            const double xe2 = x1, ye2 = y1;                                  /* vertex i2 */
            cal_topW = x_end - x_start + cal_topW;
            double multi_area = 0.0, perim = 0.0, pxm = 0.0;
            double xk = xs1, yk = ys1;
            for (int k = i1 + 1; k <= i2 - 1; ++k) {
                const double xn = dw_nat_x(D, S, i, j, k + 1), yn = dw_nat_y(D, S, i, j, k + 1);
                multi_area = multi_area + fabs(0.5 * (xn - xk) * (el_now - yk + el_now - yn));
                const double dk = dw_dist(xk, yk, xn, yn);
                perim = perim + dk;
                pxm = pxm + dk * dw_nat_m15(D, S, i, j, k);
                xk = xn; yk = yn;
            }
            const double d_s = dw_dist(x_start, el_now, xs1, ys1), d_e = dw_dist(x_end, el_now, xe2, ye2);
            cal_area = cal_area + fabs(0.5 * (xs1 - x_start) * (el_now - ys1)) + multi_area + fabs(0.5 * (xe2 - x_end) * (el_now - ye2));
            cal_peri = cal_peri + d_s + perim + d_e;
            cal_equiv_mann = cal_equiv_mann + d_s * dw_nat_m15(D, S, i, j, i1) + pxm + d_e * dw_nat_m15(D, S, i, j, i2);
            if (i1 == 1) cal_peri = cal_peri - d_s;
            if (i2 == (num - 1)) cal_peri = cal_peri - d_e;
        }
        ya = yb;
    }
    const double redi = cal_area / cal_peri;
    const double equiv_mann = trt_pow64_det(cal_equiv_mann / cal_peri, (double)(2.0f / 3.0f));
    double conv = (1.0 / equiv_mann) * cal_area * trt_pow64_det(redi, (double)(2.0f / 3.0f));
    if (cal_peri <= DW_F(1e-8)) conv = 0.0;
    dw_col_w(D, i, j, C_ELEV)[row - 1] = el_now;
    dw_col_w(D, i, j, C_AREA)[row - 1] = cal_area;
    dw_col_w(D, i, j, C_PERI)[row - 1] = cal_peri;
    dw_col_w(D, i, j, C_CONV)[row - 1] = conv;
    dw_col_w(D, i, j, C_TOPW)[row - 1] = cal_topW;
    dw_col_w(D, i, j, C_SKK)[row - 1] = 1.0 / equiv_mann;
    if (row == 1) DW_A2(D.z, i, j) = S.el_min;                                /* :2021 */
}

/* thread per node: dK/dA of the raw table, then conveyance and dK/dA made monotone in elevation (:1919-1924, :1951-2008);
 * sequential over the 501 rows by construction */
TRT_HD void dw_nat_smooth(Dom& D, int i, int j)
{
    const double* el1 = dw_col(D, i, j, C_ELEV) - 1;          /* 1-based views */
    const double* a1 = dw_col(D, i, j, C_AREA) - 1;
    double* conv1 = dw_col_w(D, i, j, C_CONV) - 1;
    double* dkda = dw_col_w(D, i, j, C_DKDA) - 1;
    const double incr_rate = DW_F(0.01);
    dkda[1] = conv1[1] / a1[1];
    for (int iel = 2; iel <= NEL; ++iel) dkda[iel] = (conv1[iel] - conv1[iel - 1]) / (a1[iel] - a1[iel - 1]);
    for (int iel = 2; iel <= NEL; ++iel) {
        if (conv1[iel] <= conv1[iel - 1]) {
            int ii = iel;
            while ((conv1[ii] < conv1[iel - 1]) && (ii < NEL)) ii = ii + 1;
            const int inc0 = ii;
            if ((inc0 >= NEL) && (conv1[inc0] < conv1[iel - 1])) conv1[inc0] = (1.0 + incr_rate) * conv1[iel - 1];
            const double pos_slope = (conv1[inc0] - conv1[iel - 1]) / (el1[inc0] - el1[iel - 1]);
            for (ii = iel; ii <= inc0 - 1; ++ii) conv1[ii] = conv1[iel - 1] + pos_slope * (el1[ii] - el1[iel - 1]);
            for (ii = iel; ii <= inc0 - 1; ++ii) dkda[ii] = (conv1[ii] - conv1[ii - 1]) / (a1[ii] - a1[ii - 1]);
        }
    }
    for (int iel = 2; iel <= NEL; ++iel) {
        if (dkda[iel] <= dkda[iel - 1]) {
            int ii = iel;
            while ((dkda[ii] < dkda[iel - 1]) && (ii < NEL)) ii = ii + 1;
            const int inc0 = ii;
            if ((inc0 >= NEL) && (dkda[inc0] < dkda[iel - 1])) dkda[inc0] = (1.0 + incr_rate) * dkda[iel - 1];
            const double pos_slope = (dkda[inc0] - dkda[iel - 1]) / (el1[inc0] - el1[iel - 1]);
            for (ii = iel; ii <= inc0 - 1; ++ii) dkda[ii] = dkda[iel - 1] + pos_slope * (el1[ii] - el1[iel - 1]);
        }
    }
}

/* minimum of one column of one node (what r_interpol returns below the table) */
TRT_HD void dw_table_min(Dom& D, int i, int j, int col)
{
    const double* y = dw_col(D, i, j, col);
    double m = y[0];
    for (int k = 1; k < NEL; ++k) if (y[k] < m) m = y[k];
    D.tabmin[((size_t)(j - 1) * D.mx + (size_t)(i - 1)) * NCOL + (size_t)col] = m;
}

/* intp_xsec_tab(i, j, nel, 10, 1, q): elevation at which node (i, j) carries q as uniform flow */
TRT_HD double dw_normal_elev(const Dom& D, int i, int j, double q)
{
    const double* qn = dw_col(D, i, j, C_QNRM);
    int* hint = &D.hint_q[(i - 1) + (size_t)(j - 1) * D.mx];
    int irow = dw_locate_hint(qn, NEL, q, *hint);
    if (irow == 0) irow = 1;
    if (irow == NEL) irow = NEL - 1;
    *hint = irow;
    return dw_interp_row(qn, dw_col(D, i, j, C_ELEV), irow, q);
}

/* ---- water-surface solve ------------------------------------------------------------------------------------------- */
/* funcd_diffdepth :1664-1711 with the downstream friction slope (constant during a solve) passed in.
 * Latency matters here (this runs on the water-surface dependency chain): the row is guessed arithmetically and the two
 * bracketing rows of all four columns are loaded in ONE round of independent loads; the guess is then verified against the
 * elevations just loaded, and only a wrong guess (first five rows, table ends) pays for the search.  Same row, same values,
 * same arithmetic as intp_xsec_tab either way. */
TRT_HD void dw_funcd(const Dom& D, int i, int j, const ElevGrid& grid, double Q_cur, double sf_ds, double z_cur, double y_cur,
                     double y_ds, double slope_dx, double dxi, double& f, double& df)
{
    const double* elev = dw_col(D, i, j, C_ELEV);
    const double* convc = dw_col(D, i, j, C_CONV);
    const double* dkdac = dw_col(D, i, j, C_DKDA);
    const double* topwc = dw_col(D, i, j, C_TOPW);
    const double elv_cur = y_cur + z_cur;
    int irow = dw_guess(grid, elv_cur);
    irow = irow < 1 ? 1 : (irow > NEL - 1 ? NEL - 1 : irow);
    double x1 = elev[irow - 1], x2 = elev[irow];
    double c1 = convc[irow - 1], c2 = convc[irow], k1 = dkdac[irow - 1], k2 = dkdac[irow], t1 = topwc[irow - 1], t2 = topwc[irow];
    if (!(x1 <= elv_cur && elv_cur < x2)) {
        irow = dw_row_of_elev(elev, elv_cur);
        x1 = elev[irow - 1]; x2 = elev[irow];
        c1 = convc[irow - 1]; c2 = convc[irow]; k1 = dkdac[irow - 1]; k2 = dkdac[irow]; t1 = topwc[irow - 1]; t2 = topwc[irow];
    }
    const double conv_cur = dw_linterpol(x1, c1, x2, c2, elv_cur);
    const double sf_cur = fabs(Q_cur) * Q_cur / (conv_cur * conv_cur);
    f = y_cur - y_ds + slope_dx - 0.50 * (sf_cur + sf_ds) * dxi;
    const double dKdA_cur = dw_linterpol(x1, k1, x2, k2, elv_cur);
    const double topw_cur = dw_linterpol(x1, t1, x2, t2, elv_cur);
    df = 1.0 + (fabs(Q_cur) * Q_cur / trt_pow64_det(conv_cur, 3.0)) * dxi * topw_cur * dKdA_cur;
}

/* The part of rtsafe (:1583-1594) that does not depend on the downstream depth: normal depth of the node for its new
 * discharge, the bracket [x1, x2] around the mean of normal and previous depth, and the friction slope of the node at both
 * ends.  Evaluated for every node at once (thread per node) before the water-surface chain starts. */
TRT_HD void dw_bracket(Dom& D, int i, int j)
{
    const double y_ulm_multi = 2.0, y_llm_multi = DW_F(0.1);
    const double Q_cur = DW_A2(D.qp, i, j), z_cur = DW_A2(D.z, i, j);
    const double y_norm = dw_normal_elev(D, i, j, fabs(Q_cur)) - z_cur;
    const double y_old = DW_A2(D.oldY, i, j) - z_cur;
    const double x1 = 0.5 * (y_norm + y_old) * y_llm_multi;
    const double x2 = 0.5 * (y_norm + y_old) * y_ulm_multi;
    const double* elev = dw_col(D, i, j, C_ELEV);
    const double* conv = dw_col(D, i, j, C_CONV);
    const double e1 = x1 + z_cur, e2 = x2 + z_cur;
    const double c1 = dw_interp_row(elev, conv, dw_row_of_elev(elev, e1), e1);
    const double c2 = dw_interp_row(elev, conv, dw_row_of_elev(elev, e2), e2);
    DW_A2(D.b_ynorm, i, j) = y_norm;
    DW_A2(D.b_x1, i, j) = x1;
    DW_A2(D.b_x2, i, j) = x2;
    DW_A2(D.b_sf1, i, j) = fabs(Q_cur) * Q_cur / (c1 * c1);
    DW_A2(D.b_sf2, i, j) = fabs(Q_cur) * Q_cur / (c2 * c2);
}

/* rtsafe :1555-1662 for node i of reach j (the node upstream of the one whose depth y_ds is known), from the bracket of
 * dw_bracket.  The function values at the bracket ends (:1591-1592) are completed here with the downstream terms; their
 * derivatives are never read by the Fortran. */
TRT_HD double dw_rtsafe(const Dom& D, int i, int j, double Q_cur, double Q_ds, double z_cur, double z_ds, double y_ds)
{
    const int maxit = 40;
    const double xacc = DW_F(1e-4);
    double df, dxx, dxold, f, temp, xh, xl, r;
    const double y_norm = DW_A2(D.b_ynorm, i, j), x1 = DW_A2(D.b_x1, i, j), x2 = DW_A2(D.b_x2, i, j);
    const ElevGrid grid = dw_grid(dw_col(D, i, j, C_ELEV));
    /* loop-invariant parts of funcd_diffdepth: downstream friction slope (:1689-1692) and bed-slope term (:1699-1700) */
    const double elv_ds = y_ds + z_ds;
    const double* elev_ds = dw_col(D, i + 1, j, C_ELEV);
    const double conv_ds = dw_interp_row(elev_ds, dw_col(D, i + 1, j, C_CONV), dw_row_of_elev(elev_ds, elv_ds), elv_ds);
    const double sf_ds = fabs(Q_ds) * Q_ds / (conv_ds * conv_ds);
    const double dxi = DW_A2(D.dx, i, j);
    double slope = (DW_A2(D.z, i, j) - DW_A2(D.z, i + 1, j)) / dxi;
    slope = fmax(slope, D.so_llm);
    const double slope_dx = slope * dxi;
    const double fl = x1 - y_ds + slope_dx - 0.50 * (DW_A2(D.b_sf1, i, j) + sf_ds) * dxi;
    const double fh = x2 - y_ds + slope_dx - 0.50 * (DW_A2(D.b_sf2, i, j) + sf_ds) * dxi;
    if ((fl > 0.0 && fh > 0.0) || (fl < 0.0 && fh < 0.0)) return y_norm;
    if (fl == 0.0) return x1;
    else if (fh == 0.0) return x2;
    else if (fl < 0.0) { xl = x1; xh = x2; }
    else { xh = x1; xl = x2; }
    r = 0.50 * (x1 + x2);
    dxold = fabs(x2 - x1);
    dxx = dxold;
    dw_funcd(D, i, j, grid, Q_cur, sf_ds, z_cur, r, y_ds, slope_dx, dxi, f, df);
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
        dw_funcd(D, i, j, grid, Q_cur, sf_ds, z_cur, r, y_ds, slope_dx, dxi, f, df);
        if (f < 0.0) xl = r; else xh = r;
    }
    return y_norm;
}

/* the dependency chain of mesh_diffusive_backward (:1498-1530): depths of reach j from its last node upstream */
TRT_HD void dw_backward_chain(Dom& D, int j)
{
    const int ncomp = DW_FRNW(j, 1);
    for (int i = ncomp; i >= 2; --i) {
        const double Q_cur = DW_A2(D.qp, i - 1, j), Q_ds = DW_A2(D.qp, i, j);
        const double z_cur = DW_A2(D.z, i - 1, j), z_ds = DW_A2(D.z, i, j);
        double y_ds = DW_A2(D.newY, i, j) - z_ds;
        y_ds = fmax(y_ds, DW_F(0.005));
        const double y_cur = dw_rtsafe(D, i - 1, j, Q_cur, Q_ds, z_cur, z_ds, y_ds);
        double ny = y_cur + z_cur;
        if (ny > 100000.0) ny = 100000.0;
        DW_A2(D.newY, i - 1, j) = ny;
    }
}

/* everything mesh_diffusive_backward computes per node that nothing on the chain reads (:1440-1496) */
TRT_HD void dw_node_props(Dom& D, int i, int j)
{
    const size_t nd = (size_t)(j - 1) * D.mx + (size_t)(i - 1);
    const double* elev = dw_col(D, i, j, C_ELEV);
    const double* mn = D.tabmin + nd * NCOL;
    const double xt = DW_A2(D.newY, i, j), zz = DW_A2(D.z, i, j);
    const double q_sk_multi = 1.0;
    double co = dw_r_interpol(elev, dw_col(D, i, j, C_CONV), mn[C_CONV], true, zz, (xt - zz) * (xt - zz), DW_A2(D.co, i, j));
    co = q_sk_multi * co;
    DW_A2(D.co, i, j) = co;
    DW_A2(D.newArea, i, j) = dw_r_interpol(elev, dw_col(D, i, j, C_AREA), mn[C_AREA], false, 0.0, xt, DW_A2(D.newArea, i, j));
    DW_A2(D.pere, i, j) = dw_r_interpol(elev, dw_col(D, i, j, C_PERI), mn[C_PERI], false, 0.0, xt, DW_A2(D.pere, i, j));
    const double bo = dw_r_interpol(elev, dw_col(D, i, j, C_TOPW), mn[C_TOPW], false, 0.0, xt, DW_A2(D.bo, i, j));
    DW_A2(D.bo, i, j) = bo;
    const double sk = dw_r_interpol(elev, dw_col(D, i, j, C_SKK), mn[C_SKK], false, 0.0, xt, DW_A2(D.sk, i, j));
    DW_A2(D.sk, i, j) = sk;
    const double q = DW_A2(D.qp, i, j);
    const double sfi = q * fabs(q) / (co * co);                                                    /* :1475 */
    double cel = (double)(5.0f / 3.0f) * trt_pow64_det(fabs(sfi), DW_F(0.3)) * trt_pow64_det(fabs(q), DW_F(0.4)) / trt_pow64_det(bo, DW_F(0.4))
                 / trt_pow64_det(1. / (sk * q_sk_multi), DW_F(0.6));                                         /* :1479-1481 */
    const double C_ulm = (i > 1) ? D.cfl * DW_A2(D.dx, i - 1, j) / D.dtini_min : D.cfl * DW_A2(D.dx, i, j) / D.dtini_min;
    if (cel > C_ulm) cel = C_ulm;
    DW_A2(D.celerity2, i, j) = cel;
    DW_A2(D.diffusivity2, i, j) = fabs(q) / 2.0 / bo / fabs(sfi);                                  /* :1495 */
}

/* reach averages :1535-1545 (sequential sums, as the Fortran's sum() intrinsic) */
TRT_HD void dw_reach_means(Dom& D, int j)
{
    const int ncomp = DW_FRNW(j, 1);
    double sc = 0.0, sd = 0.0;
    for (int i = 1; i <= ncomp; ++i) sc = sc + DW_A2(D.celerity2, i, j);
    for (int i = 1; i <= ncomp; ++i) sd = sd + DW_A2(D.diffusivity2, i, j);
    double c = sc / ncomp;
    if (c < D.C_llm) c = D.C_llm;
    double d = sd / ncomp;
    if (d > D.D_ulm) d = D.D_ulm;
    if (d < D.D_llm) d = D.D_llm;
    for (int i = 1; i <= ncomp; ++i) { DW_A2(D.celerity, i, j) = c; DW_A2(D.diffusivity, i, j) = d; }
}

/* ---- discharge sweep ----------------------------------------------------------------------------------------------- */
/* Hermite / Crank-Nicolson coefficients of node i >= 2 of reach j (:1171-1224): everything with a power or a division */
TRT_HD void dw_forward_coef(Dom& D, int i, int j, double dtini)
{
    const int ncomp = DW_FRNW(j, 1);
    const double dxm = DW_A2(D.dx, i - 1, j);
    const double cour = dtini / dxm;
    const double cour2 = fabs(DW_A2(D.celerity, i, j)) * cour;
    const double c2 = cour2 * cour2, c3 = trt_pow64_det(cour2, 3.0);
    const double a1 = 3.0 * c2 - 2.0 * c3;
    const double a2 = 1 - a1;
    const double a3 = (c2 - c3) * dxm;
    const double a4 = (-1.0 * cour2 + 2.0 * c2 - c3) * dxm;
    const double b1 = (6.0 * cour2 - 6.0 * c2) / (-1.0 * dxm);
    const double b2 = -b1;
    const double b3 = (2.0 * cour2 - 3.0 * c2) * (-1.0);
    const double b4 = (-1.0 + 4.0 * cour2 - 3.0 * c2) * (-1.0);
    const double dd1 = (6.0 - 12.0 * cour2) / (dxm * dxm);
    const double dd2 = -dd1;
    const double dd3 = (2.0 - 6.0 * cour2) / dxm;
    const double dd4 = (4.0 - 6.0 * cour2) / dxm;
    const double h1 = 12.0 / trt_pow64_det(dxm, 3.0);
    const double h2 = -h1;
    const double h3 = 6.0 / (dxm * dxm);
    const double h4 = h3;
    const double alpha = (i == ncomp) ? 1.0 : DW_A2(D.dx, i, j) / dxm;
    const double oq0 = DW_A2(D.oldQ, i - 1, j), oq1 = DW_A2(D.oldQ, i, j), px0 = DW_A2(D.qpx, i - 1, j), px1 = DW_A2(D.qpx, i, j);
    const double qy = a1 * oq0 + a2 * oq1 + a3 * px0 + a4 * px1;
    const double qxy = b1 * oq0 + b2 * oq1 + b3 * px0 + b4 * px1;
    const double qxxy = dd1 * oq0 + dd2 * oq1 + dd3 * px0 + dd4 * px1;
    const double qxxxy = h1 * oq0 + h2 * oq1 + h3 * px0 + h4 * px1;
    const double dif = DW_A2(D.diffusivity, i, j);
    const double ppi = -D.theta * dif * dtini / (dxm * dxm) * 2.0 / (alpha * (alpha + 1.0)) * alpha;
    DW_A2(D.c_ppi, i, j) = ppi;
    DW_A2(D.c_qqi, i, j) = 1.0 - ppi * (alpha + 1.0) / alpha;
    DW_A2(D.c_rri, i, j) = ppi / alpha;
    DW_A2(D.c_ssi, i, j) = qy + dtini * dif * (1.0 - D.theta) * qxxy;
    DW_A2(D.c_sxi, i, j) = qxy + dtini * dif * (1.0 - D.theta) * qxxxy;
}

/* elimination, back substitution and the lower flow limit of reach j (:1226-1229, :1281-1330).  Reads oldQ and the
 * coefficients only; node 1 is NOT set here -- it takes the junction inflow of this step afterwards (dw_forward_head). */
TRT_HD void dw_forward_solve(Dom& D, int j)
{
    const int ncomp = DW_FRNW(j, 1);
    double e = 1., f = 0., ex = 0., fx = 0.;
    DW_A2(D.eei, 1, j) = e; DW_A2(D.ffi, 1, j) = f; DW_A2(D.exi, 1, j) = ex; DW_A2(D.fxi, 1, j) = fx;
    for (int i = 2; i <= ncomp; ++i) {
        const double ppi = DW_A2(D.c_ppi, i, j), qqi = DW_A2(D.c_qqi, i, j), rri = DW_A2(D.c_rri, i, j);
        const double e1 = -1.0 * rri / (ppi * e + qqi);
        const double f1 = (DW_A2(D.c_ssi, i, j) - ppi * f) / (ppi * e + qqi);
        const double ex1 = -1.0 * rri / (ppi * ex + qqi);
        const double fx1 = (DW_A2(D.c_sxi, i, j) - ppi * fx) / (ppi * ex + qqi);
        e = e1; f = f1; ex = ex1; fx = fx1;
        DW_A2(D.eei, i, j) = e; DW_A2(D.ffi, i, j) = f; DW_A2(D.exi, i, j) = ex; DW_A2(D.fxi, i, j) = fx;
    }
    const double qp_ghost = DW_A2(D.oldQ, ncomp - 1, j);
    const double qpx_ghost = 0.;
    double q = e * qp_ghost + f;
    double qx = ex * qpx_ghost + fx;
    DW_A2(D.qp, ncomp, j) = q; DW_A2(D.qpx, ncomp, j) = qx;
    for (int i = ncomp - 1; i >= 1; --i) {
        q = DW_A2(D.eei, i, j) * q + DW_A2(D.ffi, i, j);
        qx = DW_A2(D.exi, i, j) * qx + DW_A2(D.fxi, i, j);
        DW_A2(D.qp, i, j) = q; DW_A2(D.qpx, i, j) = qx;
    }
    for (int i = 2; i <= ncomp; ++i) {
        if (fabs(DW_A2(D.qp, i, j)) < D.q_llm) DW_A2(D.qp, i, j) = D.q_llm;
        DW_A2(D.newQ, i, j) = DW_A2(D.qp, i, j);
    }
}

/* node 1 of reach j: junction inflow of this step + lateral inflow of the first segment (:669-690), then the lateral
 * inflow of the interior segments (:1162-1165, :1316-1317) and the lower limit */
TRT_HD void dw_forward_head(Dom& D, int j, double t, double dtini, const double* tarr_qtrib)
{
    const int ncomp = DW_FRNW(j, 1);
    double q1 = 0.0;
    for (int k = 1; k <= DW_FRNW(j, 3); ++k) {
        const int us = DW_FRNW(j, 3 + k);
        double q_us;
        if (D.is_main[us]) q_us = DW_A2(D.newQ, DW_FRNW(us, 1), us);
        else q_us = dw_intp_y(D.nqt, tarr_qtrib, D.qtrib + (size_t)(us - 1) * D.nqt, 1, t + dtini / 60., 1);
        q1 = q1 + q_us;
    }
    q1 = q1 + DW_A2(D.lateralFlow, 1, j) * DW_A2(D.dx, 1, j);
    double allqlat = 0.0;
    for (int i = 2; i <= ncomp - 1; ++i) allqlat = allqlat + DW_A2(D.lateralFlow, i, j) * DW_A2(D.dx, i, j);
    double q = q1 + allqlat;
    if (fabs(q) < D.q_llm) q = D.q_llm;
    DW_A2(D.qp, 1, j) = q;
    DW_A2(D.newQ, 1, j) = q;
}

/* lateral inflow of segment i of reach j at time t (:660-666): the series is shifted by one row (varr_ql(1) = varr_ql(2)) */
TRT_HD double dw_lateral(const Dom& D, int i, int j, double t)
{
    const int n = D.nql + 1;
    int irow = dw_locate_hint(D.tarr_ql, n, t, 1);
    if (irow == 0) irow = 1;
    if (irow == n) irow = n - 1;
    const double* q = D.qlat + (size_t)D.nql * ((size_t)(i - 1) + (size_t)D.mx * (size_t)(j - 1));
    const double y1 = (irow == 1) ? q[0] : q[irow - 2];
    const double y2 = q[irow - 1];
    return dw_linterpol(D.tarr_ql[irow - 1], y1, D.tarr_ql[irow], y2, t);
}

/* tailwater / junction water-surface elevation of the last node of reach j (:700-728) */
TRT_HD void dw_downstream_stage(Dom& D, int j, double t_next)
{
    const int ncomp = DW_FRNW(j, 1);
    const double mindepth_nstab = DW_F(0.1);
    if (DW_FRNW(j, 2) >= 0) {
        DW_A2(D.newY, ncomp, j) = DW_A2(D.newY, 1, DW_FRNW(j, 2));
    } else if (D.dsbc_option == 1) {
        double y = dw_intp_y(D.ndb, D.tarr_db, D.varr_db, 1, t_next, 1);
        if ((y - DW_A2(D.z, ncomp, j)) < mindepth_nstab) y = mindepth_nstab + DW_A2(D.z, ncomp, j);
        DW_A2(D.newY, ncomp, j) = y;
    } else if (D.dsbc_option == 2) {
        DW_A2(D.newY, ncomp, j) = dw_normal_elev(D, ncomp, j, fabs(DW_A2(D.newQ, ncomp, j)));
    }
}

/* record one output row (:773-798 and :801-821) for reach j: q / elevation of row `ts` from (Q, Y) */
TRT_HD void dw_record(Dom& D, int j, int ts, const double* Q, const double* Y)
{
    const int ncomp = DW_FRNW(j, 1);
    for (int i = 1; i <= ncomp; ++i) {
        DW_EV(D.q_ev, ts, i, j) = DW_A2(Q, i, j);
        DW_EV(D.elv_ev, ts, i, j) = DW_A2(Y, i, j);
        DW_EV(D.depth_ev, ts, i, j) = DW_EV(D.elv_ev, ts, i, j) - DW_A2(D.z, i, j);
    }
    for (int k = 1; k <= DW_FRNW(j, 3); ++k) {
        const int us = DW_FRNW(j, 3 + k);
        if (!D.is_main[us]) {
            const double wdepth = DW_A2(Y, 1, j) - DW_A2(D.z, 1, j);
            DW_EV(D.elv_ev, ts, DW_FRNW(us, 1), us) = DW_A2(Y, 1, j);
            DW_EV(D.depth_ev, ts, DW_FRNW(us, 1), us) = wdepth;
        }
    }
}

/* The whole simulation of one domain after the tables are built: diffnw :550-832.  Runs in ONE cooperating thread group
 * (a CTA); every thread carries the scalar clock (t, dtini, ts_ev) redundantly, phases are separated by DW_SYNC(). */
TRT_HD void dw_time_loop(Dom& D)
{
    const int nm = D.nm, mx = D.mx, nl = D.nl;
    const double TOL = DW_F(1e-8);
    /* Upper bound on the time steps of a run: the celerity cap (:1484-1492) keeps the CFL step >= dtini_min, the save
     * interval adds at most one clipped step per row.  Exceeding it means the clock has stopped advancing (NaN forcing,
     * dtini far below the resolution of t): the reference would spin forever, here the run ends with status -5. */
    const double span = (D.tfin - D.t0) * 3600.0;
    const long max_steps = (long)fmin(4.0e7, 64.0 * (span / D.dtini_min + span / D.saveInterval) + 1000.0);
    /* ---- initial water surface, tailwater to heads (:550-606); one-off, one thread */
    if (DW_TID == 0) {
        for (int jm = nm; jm >= 1; --jm) {
            const int j = D.mstem[jm - 1], ncomp = DW_FRNW(j, 1);
            if (DW_FRNW(j, 2) < 0) {
                if (D.dsbc_option == 1) {
                    for (int n = 1; n <= D.ndb; ++n) D.varr_db[n - 1] = D.dbcd[n - 1] + DW_A2(D.z, ncomp, j);
                    double y = dw_intp_y(D.ndb, D.tarr_db, D.varr_db, 1, D.t0 * 60.0, 1);
                    DW_A2(D.oldY, ncomp, j) = y;
                    if ((y - DW_A2(D.z, ncomp, j)) < DW_F(0.1)) y = DW_F(0.1) + DW_A2(D.z, ncomp, j);
                    DW_A2(D.newY, ncomp, j) = y;
                } else if (D.dsbc_option == 2) {
                    DW_A2(D.oldY, ncomp, j) = dw_normal_elev(D, ncomp, j, DW_A2(D.oldQ, ncomp, j));
                    DW_A2(D.newY, ncomp, j) = DW_A2(D.oldY, ncomp, j);
                }
            } else {
                DW_A2(D.newY, ncomp, j) = DW_A2(D.newY, 1, DW_FRNW(j, 2));
            }
            const double wdepth = DW_A2(D.newY, ncomp, j) - DW_A2(D.z, ncomp, j);
            for (int i = 1; i <= ncomp - 1; ++i) DW_A2(D.oldY, i, j) = wdepth + DW_A2(D.z, i, j);
            for (int i = 1; i <= ncomp - 1; ++i) dw_bracket(D, i, j);
            dw_backward_chain(D, j);
            for (int i = 1; i <= ncomp; ++i) dw_node_props(D, i, j);
            dw_reach_means(D, j);
            for (int i = 1; i <= ncomp; ++i) {
                DW_A2(D.oldY, i, j) = DW_A2(D.newY, i, j);
                if (DW_A2(D.oldY, i, j) < DW_A2(D.oldY, ncomp, nl)) DW_A2(D.oldY, i, j) = DW_A2(D.oldY, ncomp, nl);
            }
        }
    }
    /* ---- tributary flows at the output times (:611-633): thread per tributary reach, each walks the clock */
    for (int j = 1 + DW_TID; j <= nl; j += DW_NT) {
        if (D.is_main[j]) continue;
        int ts_ev = 1;
        double t = D.t0 * 60.0;
        const int ncomp = DW_FRNW(j, 1);
        long walk = 0;
        while (t <= D.tfin * 60.0 && ++walk <= max_steps) {
            if ((fmod((t - D.t0 * 60.) * 60., D.saveInterval) <= TOL) || (t == D.tfin * 60.)) {
                if (ts_ev <= D.nev) {
                    const double q = dw_intp_y(D.nqt, D.tarr_qtrib, D.qtrib + (size_t)(j - 1) * D.nqt, 1, t, ts_ev);
                    DW_EV(D.q_ev, ts_ev, ncomp, j) = q;
                    DW_EV(D.q_ev, ts_ev, 1, j) = q;
                }
                ts_ev = ts_ev + 1;
            }
            t = t + D.dtini_given / 60.;
        }
    }
    DW_SYNC();

    /* ---- time loop (:655-832) */
    double t = D.t0 * 60.0, dtini = D.dtini_given;
    int ts_ev = 1;
    long guard = 0;
    while (t < D.tfin * 60.) {
        /* calculateDT :942-991 */
        const double maxCelDx = D.scal[0];
        dtini = D.cfl / maxCelDx;
        {
            const int a = (int)floor((t - D.t0 * 60.) / (D.saveInterval / 60.));
            const int b = (int)floor(((t - D.t0 * 60.) + dtini / 60.) / (D.saveInterval / 60.));
            if (b > a) dtini = (a + 1) * (D.saveInterval) - (t - D.t0 * 60.) * 60.;
            if (t + dtini / 60. > D.tfin * 60.) dtini = (D.tfin * 60. - t) * 60.;
        }
        if (!(dtini > 0.0) || ++guard > max_steps) {          /* the reference would never leave the loop */
            if (DW_TID == 0) *D.status = -5;
            break;
        }
        /* discharge sweep, phase 1: lateral inflow at t and the Hermite / CN coefficients, thread per node */
        for (int idx = DW_TID; idx < nm * mx; idx += DW_NT) {
            const int j = D.mstem[idx / mx], i = idx % mx + 1, ncomp = DW_FRNW(j, 1);
            if (i <= ncomp - 1) DW_A2(D.lateralFlow, i, j) = dw_lateral(D, i, j, t);
            if (i >= 2 && i <= ncomp) dw_forward_coef(D, i, j, dtini);
        }
        DW_SYNC();
        /* phase 2: elimination + back substitution, thread per reach */
        for (int jm = DW_TID; jm < nm; jm += DW_NT) dw_forward_solve(D, D.mstem[jm]);
        DW_SYNC();
        /* phase 3: node 1 of every reach takes the junction inflow of this step */
        for (int jm = DW_TID; jm < nm; jm += DW_NT) dw_forward_head(D, D.mstem[jm], t, dtini, D.tarr_qtrib);
        DW_SYNC();
        /* water-surface sweep.  First what does not depend on the downstream depth, thread per node ... */
        for (int idx = DW_TID; idx < nm * mx; idx += DW_NT) {
            const int j = D.mstem[idx / mx], i = idx % mx + 1;
            if (i <= DW_FRNW(j, 1) - 1) dw_bracket(D, i, j);
        }
        DW_SYNC();
        /* ... then the dependency chain from the tailwater to the heads: reaches of equal distance from the tailwater
         * (the arms above a confluence) are independent and go to different warps, one lane each */
        for (int lev = 0; lev < D.nlev; ++lev) {
            for (int k = D.lvl_ptr[lev] + DW_WARP; k < D.lvl_ptr[lev + 1]; k += DW_NWARP) {
                if (DW_LANE == 0) {
                    const int j = D.lvl_reach[k];
                    dw_downstream_stage(D, j, t + dtini / 60.);
                    dw_backward_chain(D, j);
                }
            }
            DW_SYNC();
        }
        /* ... and everything off the chain: hydraulic properties, celerity, diffusivity of every node */
        for (int idx = DW_TID; idx < nm * mx; idx += DW_NT) {
            const int j = D.mstem[idx / mx], i = idx % mx + 1;
            if (i <= DW_FRNW(j, 1)) dw_node_props(D, i, j);
        }
        DW_SYNC();
        for (int jm = DW_TID; jm < nm; jm += DW_NT) {
            const int j = D.mstem[jm], ncomp = DW_FRNW(j, 1);
            dw_reach_means(D, j);
            double m = 0.;
            for (int k = 1; k <= ncomp - 1; ++k) m = fmax(m, DW_A2(D.celerity, k, j) / DW_A2(D.dx, k, j));
            D.rmax[j - 1] = m;
        }
        DW_SYNC();
        if (DW_TID == 0) {                                     /* :742-750 */
            double m = 0.;
            for (int jm = 0; jm < nm; ++jm) m = fmax(m, D.rmax[D.mstem[jm] - 1]);
            D.scal[0] = m;
        }
        t = t + dtini / 60.;
        if ((fmod((t - D.t0 * 60.) * 60., D.saveInterval) <= TOL) || (t == D.tfin * 60.)) {
            if (ts_ev + 1 <= D.nev)
                for (int jm = DW_TID; jm < nm; jm += DW_NT) dw_record(D, D.mstem[jm], ts_ev + 1, D.newQ, D.newY);
            ts_ev = ts_ev + 1;
        }
        if (t == D.t0 + dtini / 60.)                           /* :801 (true only for t0 == 0 and a first step of dtini) */
            for (int jm = DW_TID; jm < nm; jm += DW_NT) dw_record(D, D.mstem[jm], 1, D.oldQ, D.oldY);
        DW_SYNC();
        for (size_t k = (size_t)DW_TID; k < (size_t)mx * nl; k += (size_t)DW_NT) {     /* :824-830 */
            D.oldY[k] = D.newY[k]; D.newY[k] = -999;
            D.oldQ[k] = D.newQ[k]; D.newQ[k] = -999;
            D.oldArea[k] = D.newArea[k]; D.newArea[k] = -999;
            D.pere[k] = -999;
        }
        DW_SYNC();
    }
}

}  // namespace trtdw
