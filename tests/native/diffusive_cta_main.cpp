/*
 * diffusive_cta_main.cpp -- TEST INFRASTRUCTURE.  Runs the product's dw_time_loop (t-route_b200/csrc/diffusive_device.cuh) as
 * an EMULATED CTA: N host threads (N a multiple of 32, thread id = threadIdx.x) that meet at a pthread barrier wherever the
 * device code calls __syncthreads().  Built with -fsanitize=thread, a phase of the time loop that reads what another
 * thread of the same phase writes -- i.e. a missing barrier, which on the GPU would be a silent race -- is reported by
 * ThreadSanitizer; the outputs are compared with the oracle by tests/test_diffusive_cta_emulation.py.
 *
 *   diffusive_cta_main <nthreads> <input.bin> <output.bin>
 * input.bin : the 39 input arguments of c_diffnw in order, each as  int64 count, then the values (int32 or float64)
 * output.bin: q_ev_g, elv_ev_g, depth_ev_g (float64, Fortran order)
 */
#include <pthread.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../t-route_b200/csrc/diffusive_setup.h"

thread_local int dw_emu_tid = 0;
int dw_emu_nt = 1;
static pthread_barrier_t g_barrier;
void dw_emu_sync() { pthread_barrier_wait(&g_barrier); }

using namespace trtdw;

struct Job { Dom* D; int tid; };
static void* run(void* p)
{
    Job* j = (Job*)p;
    dw_emu_tid = j->tid;
    dw_time_loop(*j->D);
    return nullptr;
}

int main(int argc, char** argv)
{
    if (argc != 4) return 2;
    const int nt = atoi(argv[1]);
    if (nt < 32 || nt % 32) return 2;
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 3;
    /* kinds of the 39 inputs: D double array, I int array / scalar */
    const char* kinds = "DIIIIIIIIDDDDDDDDDDIIDDDDIDIDDDIDIDIIDD";
    std::vector<std::vector<double>> dv(39);
    std::vector<std::vector<int>> iv(39);
    const void* ptr[42];
    for (int k = 0; k < 39; ++k) {
        int64_t n = 0;
        if (fread(&n, sizeof n, 1, f) != 1) return 4;
        if (kinds[k] == 'D') { dv[k].resize((size_t)(n > 0 ? n : 1)); if (n && fread(dv[k].data(), 8, (size_t)n, f) != (size_t)n) return 4; ptr[k] = dv[k].data(); }
        else { iv[k].resize((size_t)(n > 0 ? n : 1)); if (n && fread(iv[k].data(), 4, (size_t)n, f) != (size_t)n) return 4; ptr[k] = iv[k].data(); }
    }
    fclose(f);
    const int nev = iv[4][0], mx = iv[7][0], nl = iv[8][0];
    const size_t nout = (size_t)nev * mx * nl;
    std::vector<double> q(nout, 0.0), e(nout, 0.0), d(nout, 0.0);
    ptr[39] = q.data(); ptr[40] = e.data(); ptr[41] = d.data();
    const void* const* a = ptr;
    DiffnwArgs A{(const double*)a[0], (const int*)a[1], (const int*)a[2], (const int*)a[3], (const int*)a[4], (const int*)a[5],
                 (const int*)a[6], (const int*)a[7], (const int*)a[8], (const double*)a[9], (const double*)a[10],
                 (const double*)a[11], (const double*)a[12], (const double*)a[13], (const double*)a[14], (const double*)a[15],
                 (double*)a[16], (const double*)a[17], (const double*)a[18], (const int*)a[19], (const int*)a[20],
                 (const double*)a[21], (const double*)a[22], (const double*)a[23], (const double*)a[24], (const int*)a[25],
                 (const double*)a[26], (const int*)a[27], (const double*)a[28], (const double*)a[29], (const double*)a[30],
                 (const int*)a[31], (const double*)a[32], (const int*)a[33], (const double*)a[34], (const int*)a[35],
                 (const int*)a[36], (const double*)a[37], (const double*)a[38], q.data(), e.data(), d.data()};
    DomHost H;
    const std::string err = dw_build_host(A, H);
    if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 5; }
    Dom& D = H.d;
    std::vector<double> tab(H.n_nodes * NCOL * LD, 0.0), tabmin(H.n_nodes * NCOL, 0.0);
    D.tab = tab.data(); D.tabmin = tabmin.data();
    D.q_ev = q.data(); D.elv_ev = e.data(); D.depth_ev = d.data();
    /* tables: every (node, row) is an independent thread on the device; built serially here */
    for (int jm = 0; jm < D.nm; ++jm) {
        const int j = D.mstem[jm];
        for (int i = 1; i <= DW_FRNW(j, 1); ++i) {
            if (D.mxnbathy == 0) { for (int row = 1; row <= NEL; ++row) dw_table_pass1(D, i, j, row); }
            else {
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
    dw_emu_nt = nt;
    pthread_barrier_init(&g_barrier, nullptr, (unsigned)nt);
    std::vector<pthread_t> th((size_t)nt);
    std::vector<Job> jobs((size_t)nt);
    for (int t = 0; t < nt; ++t) { jobs[(size_t)t] = Job{&D, t}; pthread_create(&th[(size_t)t], nullptr, run, &jobs[(size_t)t]); }
    for (int t = 0; t < nt; ++t) pthread_join(th[(size_t)t], nullptr);
    pthread_barrier_destroy(&g_barrier);
    if (*D.status != 0) return 6;
    FILE* o = fopen(argv[3], "wb");
    if (!o) return 3;
    fwrite(q.data(), 8, nout, o); fwrite(e.data(), 8, nout, o); fwrite(d.data(), 8, nout, o);
    fclose(o);
    return 0;
}
