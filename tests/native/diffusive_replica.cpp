/*
 * diffusive_replica.cpp -- TEST INFRASTRUCTURE.  Compiles the product's solver source (t-route_b200/csrc/diffusive_device.cuh,
 * TRT_HD functions) for the HOST with g++ and runs it single-threaded, so that the solver logic -- table construction,
 * arithmetic row look-ups, the re-ordered sweeps of dw_time_loop -- can be compared with oracle/diffusive_oracle.c without
 * a GPU (tests/test_diffusive_replica.py).  Never linked into libtroute_b200.so; the product path has no CPU fallback.
 *
 * Build: g++ -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (tests/helpers_diffusive.py).
 */
#include <cstdlib>
#include <string>
#include <vector>

#include "../../t-route_b200/csrc/diffusive_setup.h"

using namespace trtdw;

extern "C" int trt_replica_diffnw(
    const double* timestep_ar_g, const int* nts_ql_g, const int* nts_ub_g, const int* nts_db_g, const int* ntss_ev_g,
    const int* nts_qtrib_g, const int* nts_da_g, const int* mxncomp_g, const int* nrch_g, const double* z_ar_g,
    const double* bo_ar_g, const double* traps_ar_g, const double* tw_ar_g, const double* twcc_ar_g, const double* mann_ar_g,
    const double* manncc_ar_g, double* so_ar_g, const double* dx_ar_g, const double* iniq, const int* frnw_col,
    const int* frnw_ar_g, const double* qlat_g, const double* ubcd_g, const double* dbcd_g, const double* qtrib_g,
    const int* paradim, const double* para_ar_g, const int* mxnbathy_g, const double* x_bathy_g, const double* z_bathy_g,
    const double* mann_bathy_g, const int* size_bathy_g, const double* usgs_da_g, const int* usgs_da_reach_g,
    const double* rdx_ar_g, const int* cwnrow_g, const int* cwncol_g, const double* crosswalk_g, const double* z_thalweg_g,
    double* q_ev_g, double* elv_ev_g, double* depth_ev_g)
{
    DiffnwArgs a = {timestep_ar_g, nts_ql_g, nts_ub_g, nts_db_g, ntss_ev_g, nts_qtrib_g, nts_da_g, mxncomp_g, nrch_g, z_ar_g,
                    bo_ar_g, traps_ar_g, tw_ar_g, twcc_ar_g, mann_ar_g, manncc_ar_g, so_ar_g, dx_ar_g, iniq, frnw_col, frnw_ar_g,
                    qlat_g, ubcd_g, dbcd_g, qtrib_g, paradim, para_ar_g, mxnbathy_g, x_bathy_g, z_bathy_g, mann_bathy_g,
                    size_bathy_g, usgs_da_g, usgs_da_reach_g, rdx_ar_g, cwnrow_g, cwncol_g, crosswalk_g, z_thalweg_g, q_ev_g,
                    elv_ev_g, depth_ev_g};
    DomHost H;
    const std::string err = dw_build_host(a, H);
    if (!err.empty()) return -1;
    /* Same steps as the CUDA library's upload (diffusive.cu upload_domain): the pools are COPIED, the pointers re-pointed
     * at the copies with dw_rebase, and the originals poisoned -- a pointer that dw_rebase forgets reads NaN / garbage here. */
    std::vector<double> dd(H.dpool);
    std::vector<int> di(H.ipool);
    std::vector<unsigned char> db(H.bpool);
    Dom D = dw_rebase(H, dd.data(), di.data(), db.data());
    for (double& v : H.dpool) v = trt64_from_bits(0x7ff8dead0000beefULL);
    for (int& v : H.ipool) v = 0x40000000;
    for (unsigned char& v : H.bpool) v = 0xff;
    std::vector<double> tab(H.n_nodes * NCOL * LD, 0.0), tabmin(H.n_nodes * NCOL, 0.0);
    D.tab = tab.data(); D.tabmin = tabmin.data();
    D.q_ev = q_ev_g; D.elv_ev = elv_ev_g; D.depth_ev = depth_ev_g;
    for (size_t k = 0; k < H.n_out; ++k) { q_ev_g[k] = 0.0; elv_ev_g[k] = 0.0; depth_ev_g[k] = 0.0; }
    for (int jm = 0; jm < D.nm; ++jm) {
        const int j = D.mstem[jm];
        for (int i = 1; i <= DW_FRNW(j, 1); ++i) {
            if (D.mxnbathy == 0) {
                for (int row = 1; row <= NEL; ++row) dw_table_pass1(D, i, j, row);
            } else {
                for (int ic = 1; ic <= D.size_bathy[(i - 1) + (size_t)(j - 1) * D.mx]; ++ic) dw_nat_prep(D, i, j, ic);
                for (int row = 1; row <= NEL; ++row) dw_nat_pass1(D, i, j, row);
                dw_nat_smooth(D, i, j);
            }
        }
    }
    for (int jm = 0; jm < D.nm; ++jm) {
        const int j = D.mstem[jm];
        for (int i = 1; i <= DW_FRNW(j, 1); ++i) {
            for (int row = 1; row <= NEL; ++row) dw_table_pass2(D, i, j, row);
            for (int col = 0; col < NCOL; ++col) dw_table_min(D, i, j, col);
        }
    }
    dw_time_loop(D);
    return *D.status;
}

/* column `col` (0..7, trtdw::C_*) of the look-up table of node (i, j), 1-based, for the table tests */
extern "C" int trt_replica_table(const double* z_ar_g, const double* bo_ar_g, const double* traps_ar_g, const double* tw_ar_g,
                                 const double* twcc_ar_g, const double* mann_ar_g, const double* manncc_ar_g, double* out6x501)
{
    Xs S;
    dw_xs_setup(S, *z_ar_g, *bo_ar_g, *traps_ar_g, *tw_ar_g, *twcc_ar_g, 1.0 / (1.0 / *mann_ar_g), 1.0 / (1.0 / *manncc_ar_g));
    for (int row = 1; row <= NEL; ++row) {
        const XsRow r = dw_xs_row(S, row);
        out6x501[0 * NEL + row - 1] = r.el; out6x501[1 * NEL + row - 1] = r.area; out6x501[2 * NEL + row - 1] = r.peri;
        out6x501[3 * NEL + row - 1] = r.conv; out6x501[4 * NEL + row - 1] = r.topw; out6x501[5 * NEL + row - 1] = r.skk;
    }
    return 0;
}

/* probe of dw_locate_hint for tests/test_diffusive_replica.py */
extern "C" int trt_replica_locate(const double* xx, int n, double x, int hint) { return dw_locate_hint(xx, n, x, hint); }

/* look-up table of ONE surveyed cross section (nb vertices): out = [NCOL][501] columns C_ELEV, C_AREA, C_PERI, C_CONV, C_TOPW,
 * C_DKDA, (unused), C_SKK after the monotone smoothing; returns the bed elevation in *z_out */
extern "C" int trt_replica_table_natural(int nb, const double* x, const double* z, const double* mann, double* out8x501, double* z_out)
{
    Dom D;
    std::memset(&D, 0, sizeof D);
    D.mx = 1; D.nl = 1; D.nm = 1; D.mxnbathy = nb;
    std::vector<double> tab((size_t)NCOL * LD, 0.0), m15((size_t)nb, 0.0), zz(1, 0.0);
    int size = nb;
    D.x_bathy = x; D.z_bathy = z; D.mann_bathy = mann; D.size_bathy = &size; D.mann15 = m15.data();
    D.tab = tab.data(); D.z = zz.data();
    for (int ic = 1; ic <= nb; ++ic) dw_nat_prep(D, 1, 1, ic);
    for (int row = 1; row <= NEL; ++row) dw_nat_pass1(D, 1, 1, row);
    dw_nat_smooth(D, 1, 1);
    for (int c = 0; c < NCOL; ++c)
        for (int r = 0; r < NEL; ++r) out8x501[(size_t)c * NEL + r] = tab[(size_t)c * LD + r];
    *z_out = zz[0];
    return 0;
}
