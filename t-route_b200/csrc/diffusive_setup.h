/*
 * diffusive_setup.h -- host-side set-up of one diffusive domain from the argument list of c_diffnw
 * (/root/reference/src/kernel/diffusive/pydiffusive.f90:8-52; diffnw's own set-up is diffusive.f90:214-548).
 *
 * Builds every array of trtdw::Dom in ONE host pool of doubles and ONE of ints (inputs copied in Fortran order, state
 * zero-initialised, time axes and the mainstem list precomputed).  The CUDA library uploads the two pools and rebases the
 * pointers (diffusive.cu); the test-only host replica uses them in place.
 */
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "diffusive_device.cuh"

namespace trtdw {

/* the 42 arguments of c_diffnw, in order; every one a pointer (Fortran passes by reference) */
struct DiffnwArgs {
    const double* timestep_ar_g; const int *nts_ql_g, *nts_ub_g, *nts_db_g, *ntss_ev_g, *nts_qtrib_g, *nts_da_g, *mxncomp_g, *nrch_g;
    const double *z_ar_g, *bo_ar_g, *traps_ar_g, *tw_ar_g, *twcc_ar_g, *mann_ar_g, *manncc_ar_g; double* so_ar_g;
    const double *dx_ar_g, *iniq; const int *frnw_col, *frnw_ar_g; const double *qlat_g, *ubcd_g, *dbcd_g, *qtrib_g;
    const int* paradim; const double* para_ar_g; const int* mxnbathy_g; const double *x_bathy_g, *z_bathy_g, *mann_bathy_g;
    const int* size_bathy_g; const double* usgs_da_g; const int* usgs_da_reach_g; const double* rdx_ar_g;
    const int *cwnrow_g, *cwncol_g; const double *crosswalk_g, *z_thalweg_g; double *q_ev_g, *elv_ev_g, *depth_ev_g;
};

struct DomHost {
    Dom d;                          /* pointers into dpool / ipool / bpool (host addresses) */
    std::vector<double> dpool;
    std::vector<int> ipool;
    std::vector<unsigned char> bpool;
    size_t n_nodes = 0;             /* mx * nl */
    size_t n_out = 0;               /* nev * mx * nl */
    /* tab, tabmin, q_ev, elv_ev, depth_ev are NOT in the pools (large / output): the owner allocates them */
};

/* returns "" or an error message */
inline std::string dw_build_host(const DiffnwArgs& a, DomHost& H)
{
    Dom& D = H.d;
    std::memset(&D, 0, sizeof D);
    if (!a.mxncomp_g || !a.nrch_g || !a.timestep_ar_g || !a.para_ar_g || !a.frnw_ar_g || !a.frnw_col) return "NULL scalar argument";
    const int mx = *a.mxncomp_g, nl = *a.nrch_g;
    if (mx < 2 || nl < 1) return "mxncomp_g must be >= 2 and nrch_g >= 1";
    if (*a.mxnbathy_g < 0) return "mxnbathy_g < 0";
    D.mxnbathy = *a.mxnbathy_g;
    if (*a.cwnrow_g > 0) return "refactored-hydrofabric crosswalk (cwnrow_g > 0, diffusive.f90:837-903) is not supported";
    if (*a.paradim < 11) return "para_ar_g needs 11 entries";
    D.mx = mx; D.nl = nl; D.nev = *a.ntss_ev_g; D.nql = *a.nts_ql_g; D.nqt = *a.nts_qtrib_g; D.ndb = *a.nts_db_g;
    D.frnw_col = *a.frnw_col;
    if (D.nev < 1 || D.nql < 1 || D.nqt < 2 || D.ndb < 0) return "bad time-series lengths";
    const double* ts = a.timestep_ar_g;                                   /* :214-226 */
    D.dtini_given = ts[0]; D.t0 = ts[1]; D.tfin = ts[2]; D.saveInterval = ts[3]; D.dt_ql = ts[4]; D.dt_db = ts[6];
    D.dt_qtrib = ts[7]; D.dtini_min = ts[0] / ts[9];
    const double* pa = a.para_ar_g;                                       /* :235-245 */
    D.cfl = pa[0]; D.C_llm = pa[1]; D.D_llm = pa[2]; D.D_ulm = pa[3]; D.q_llm = pa[7]; D.so_llm = pa[8]; D.theta = pa[9];
    D.dsbc_option = (int)pa[10];
    if (D.dsbc_option == 1 && D.ndb < 2) return "dsbc_option 1 needs a tailwater depth series";
    if (!(D.dtini_given > 0.0) || !(D.saveInterval > 0.0) || !(ts[9] > 0.0)) return "bad timestep_ar_g";

    const size_t n2 = (size_t)mx * nl;
    H.n_nodes = n2; H.n_out = (size_t)D.nev * n2;
    /* ---- ints: frnw (as given), mstem, hint_q; bytes: is_main */
    const size_t nfr = (size_t)nl * D.frnw_col;
    H.ipool.assign(nfr + (size_t)nl + n2 + 1 + (size_t)nl + 2 + (size_t)nl + n2, 0);
    int* ip = H.ipool.data();
    std::memcpy(ip, a.frnw_ar_g, nfr * sizeof(int));
    D.frnw = ip; ip += nfr;
    int* mstem = ip; ip += nl;
    int* hint = ip; ip += n2;
    D.status = ip; ip += 1;
    int* lvl_ptr = ip; ip += (size_t)nl + 2;
    int* lvl_reach = ip; ip += nl;
    int* size_bathy = ip; ip += n2;
    if (D.mxnbathy > 0) std::memcpy(size_bathy, a.size_bathy_g, n2 * sizeof(int));
    D.size_bathy = size_bathy;
    for (size_t k = 0; k < n2; ++k) hint[k] = NEL / 2;
    H.bpool.assign((size_t)nl + 1, 0);
    int nm = 0;
    for (int j = 1; j <= nl; ++j) {                                       /* :403-417 */
        const int ncomp = DW_FRNW(j, 1), nus = DW_FRNW(j, 3);
        if (ncomp < 1 || ncomp > mx) return "frnw_g: node count of a reach out of range";
        if (nus < 0 || 3 + nus + 1 > D.frnw_col) return "frnw_g: too many upstream reaches for frnw_col";
        for (int k = 1; k <= nus; ++k)
            if (DW_FRNW(j, 3 + k) < 1 || DW_FRNW(j, 3 + k) > nl) return "frnw_g: upstream reach index out of range";
        if (DW_FRNW(j, 2) > nl) return "frnw_g: downstream reach index out of range";
        if (DW_FRNW(j, 3 + nus + 1) == 555) {
            if (ncomp < 2) return "a mainstem reach needs at least 2 nodes";
            mstem[nm++] = j; H.bpool[(size_t)j] = 1;
        }
    }
    if (nm == 0) return "no mainstem reach (flag 555) in frnw_g";
    if (D.mxnbathy > 0)
        for (int jm = 0; jm < nm; ++jm)
            for (int i = 1; i <= DW_FRNW(mstem[jm], 1); ++i) {
                const int nb = size_bathy[(i - 1) + (size_t)(mstem[jm] - 1) * mx];
                if (nb < 2 || nb > D.mxnbathy) return "size_bathy_g: a mainstem node needs 2 .. mxnbathy_g cross-section vertices";
            }
    D.nm = nm; D.mstem = mstem; D.hint_q = hint; D.is_main = H.bpool.data();
    {
        /* distance of every mainstem reach from the tailwater, counted in mainstem reaches; a reach drains into a reach
         * with a larger index (fp_network_map numbers reaches upstream to downstream, diffusive_utils_v02.py:90-101) */
        std::vector<int> lev((size_t)nl + 1, -1);
        int nlev = 0;
        for (int jm = nm - 1; jm >= 0; --jm) {
            const int j = mstem[jm], ds = DW_FRNW(j, 2);
            if (ds >= 1) {
                if (ds <= j) return "frnw_g: a reach must drain into a reach with a larger index";
                if (!H.bpool[(size_t)ds]) return "frnw_g: a mainstem reach drains into a tributary reach";
                lev[(size_t)j] = lev[(size_t)ds] + 1;
            } else lev[(size_t)j] = 0;
            if (lev[(size_t)j] + 1 > nlev) nlev = lev[(size_t)j] + 1;
        }
        for (int jm = 0; jm < nm; ++jm) lvl_ptr[lev[(size_t)mstem[jm]] + 1]++;
        for (int l = 0; l < nlev; ++l) lvl_ptr[l + 1] += lvl_ptr[l];
        std::vector<int> fill(lvl_ptr, lvl_ptr + nlev);
        for (int jm = nm - 1; jm >= 0; --jm) lvl_reach[fill[(size_t)lev[(size_t)mstem[jm]]]++] = mstem[jm];
        D.nlev = nlev; D.lvl_ptr = lvl_ptr; D.lvl_reach = lvl_reach;
    }

    /* ---- doubles */
    const size_t nq = (size_t)D.nql * n2, nt = (size_t)D.nqt * nl;
    const size_t nbathy = (size_t)D.mxnbathy * n2;
    const size_t total = 4 * nbathy /*vertices + mann15*/ + 8 * n2 /*geometry in*/ + nq + nt + (size_t)D.ndb + n2 /*iniq*/ + (size_t)(D.nql + 1) + D.nqt + D.ndb /*time axes*/
                         + (size_t)nl /*rmax*/ + 33 * n2 /*state*/ + (size_t)D.ndb /*varr_db*/ + 8 /*scal*/;
    H.dpool.assign(total, 0.0);
    double* p = H.dpool.data();
    auto take_in = [&](const double* src, size_t cnt) { double* q = p; if (cnt) std::memcpy(q, src, cnt * sizeof(double)); p += cnt; return (const double*)q; };
    auto take = [&](size_t cnt) { double* q = p; p += cnt; return q; };
    D.z_in = take_in(a.z_ar_g, n2); D.bo_in = take_in(a.bo_ar_g, n2); D.traps_in = take_in(a.traps_ar_g, n2);
    D.tw_in = take_in(a.tw_ar_g, n2); D.twcc_in = take_in(a.twcc_ar_g, n2); D.mann_in = take_in(a.mann_ar_g, n2);
    D.manncc_in = take_in(a.manncc_ar_g, n2); D.dx_in = take_in(a.dx_ar_g, n2);
    D.qlat = take_in(a.qlat_g, nq); D.qtrib = take_in(a.qtrib_g, nt); D.dbcd = take_in(a.dbcd_g, (size_t)D.ndb);
    D.iniq = take_in(a.iniq, n2);
    D.x_bathy = take_in(a.x_bathy_g, nbathy); D.z_bathy = take_in(a.z_bathy_g, nbathy); D.mann_bathy = take_in(a.mann_bathy_g, nbathy);
    D.mann15 = take(nbathy);
    double* tql = take((size_t)D.nql + 1); double* tqt = take((size_t)D.nqt); double* tdb = take((size_t)D.ndb);
    for (int n = 1; n <= D.nql; ++n) tql[n] = D.t0 * 60.0 + D.dt_ql * (double)n / 60.0;          /* :512-516 */
    tql[0] = D.t0 * 60;
    for (int n = 1; n <= D.nqt; ++n) tqt[n - 1] = D.t0 * 60.0 + D.dt_qtrib * (double)(n - 1) / 60.0;
    for (int n = 1; n <= D.ndb; ++n) tdb[n - 1] = D.t0 * 60.0 + D.dt_db * (double)(n - 1) / 60.0;
    D.tarr_ql = tql; D.tarr_qtrib = tqt; D.tarr_db = tdb;
    D.rmax = take((size_t)nl);
    double** state[] = {&D.z, &D.dx, &D.bo, &D.pere, &D.qp, &D.qpx, &D.sk, &D.co, &D.oldQ, &D.newQ, &D.oldArea, &D.newArea,
                        &D.oldY, &D.newY, &D.lateralFlow, &D.celerity, &D.diffusivity, &D.celerity2, &D.diffusivity2, &D.eei,
                        &D.ffi, &D.exi, &D.fxi, &D.c_ppi, &D.c_qqi, &D.c_rri, &D.c_ssi, &D.c_sxi, &D.b_ynorm, &D.b_x1, &D.b_x2,
                        &D.b_sf1, &D.b_sf2};
    for (double** s : state) *s = take(n2);
    D.varr_db = take((size_t)D.ndb);
    D.scal = take(8);
    /* initial state :372-385 */
    std::memcpy(D.z, a.z_ar_g, n2 * sizeof(double));
    std::memcpy(D.oldQ, a.iniq, n2 * sizeof(double));
    std::memcpy(D.newQ, a.iniq, n2 * sizeof(double));
    std::memcpy(D.qp, a.iniq, n2 * sizeof(double));
    for (size_t k = 0; k < n2; ++k) D.newY[k] = -999;
    double minDx = 1e10;                                                  /* :419-430 */
    for (int jm = 0; jm < nm; ++jm) {
        const int j = mstem[jm], ncomp = DW_FRNW(j, 1);
        for (int i = 1; i <= ncomp - 1; ++i) {
            DW_A2(D.dx, i, j) = DW_A2(a.dx_ar_g, i, j);
            if (!(DW_A2(D.dx, i, j) > 0.0)) return "dx_ar_g of a mainstem segment is not positive";
            if (DW_A2(D.dx, i, j) < minDx) minDx = DW_A2(D.dx, i, j);
        }
    }
    D.scal[0] = 1.0 / minDx;                                              /* maxCelerity / minDx :638-639 */
    return "";
}

/* Re-point every pool pointer of `D` (built by dw_build_host over H's pools) at copies of the pools that start at
 * dd / di / db: what the CUDA library does after uploading the pools (diffusive.cu), testable on the host. */
template <class T>
inline void dw_rebase_ptr(T*& p, const void* hbase, size_t hbytes, void* dbase)
{
    if (!p) return;
    const char* c = (const char*)p;
    if (c >= (const char*)hbase && c < (const char*)hbase + hbytes) p = (T*)((char*)dbase + (c - (const char*)hbase));
}

inline Dom dw_rebase(const DomHost& H, double* dd, int* di, unsigned char* db)
{
    Dom D = H.d;
    const void* hd = H.dpool.data(); const size_t nd = H.dpool.size() * sizeof(double);
    const void* hi = H.ipool.data(); const size_t ni = H.ipool.size() * sizeof(int);
    const void* hb = H.bpool.data(); const size_t nb = H.bpool.size();
    const double** cdp[] = {&D.z_in, &D.bo_in, &D.traps_in, &D.tw_in, &D.twcc_in, &D.mann_in, &D.manncc_in, &D.dx_in, &D.qlat,
                            &D.qtrib, &D.dbcd, &D.iniq, &D.tarr_ql, &D.tarr_qtrib, &D.tarr_db, &D.x_bathy, &D.z_bathy, &D.mann_bathy};
    for (const double** q : cdp) dw_rebase_ptr(*q, hd, nd, dd);
    dw_rebase_ptr(D.mann15, hd, nd, dd);
    dw_rebase_ptr(D.size_bathy, hi, ni, di);
    double** dp[] = {&D.rmax, &D.z, &D.dx, &D.bo, &D.pere, &D.qp, &D.qpx, &D.sk, &D.co, &D.oldQ, &D.newQ, &D.oldArea, &D.newArea,
                     &D.oldY, &D.newY, &D.lateralFlow, &D.celerity, &D.diffusivity, &D.celerity2, &D.diffusivity2, &D.eei,
                     &D.ffi, &D.exi, &D.fxi, &D.c_ppi, &D.c_qqi, &D.c_rri, &D.c_ssi, &D.c_sxi, &D.b_ynorm, &D.b_x1, &D.b_x2, &D.b_sf1,
                     &D.b_sf2, &D.varr_db, &D.scal};
    for (double** q : dp) dw_rebase_ptr(*q, hd, nd, dd);
    dw_rebase_ptr(D.frnw, hi, ni, di);
    dw_rebase_ptr(D.mstem, hi, ni, di);
    dw_rebase_ptr(D.hint_q, hi, ni, di);
    dw_rebase_ptr(D.status, hi, ni, di);
    dw_rebase_ptr(D.lvl_ptr, hi, ni, di);
    dw_rebase_ptr(D.lvl_reach, hi, ni, di);
    dw_rebase_ptr(D.is_main, hb, nb, db);
    return D;
}

}  // namespace trtdw
