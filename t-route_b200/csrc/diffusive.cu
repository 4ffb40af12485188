/*
 * diffusive.cu -- kernels and C ABI of the diffusive-wave mainstem solver (include/troute_b200.h: trt_c_diffnw,
 * trt_diffnw_batch).  Drop-in for c_diffnw (/root/reference/src/kernel/diffusive/pydiffusive.f90:8-52), the Fortran entry
 * point the reference binds in fast_reach/fortran_wrappers.pxd and calls once per tailwater domain from
 * compute_diffusive_routing (compute.py:1740-1884).
 *
 * Device work per call:
 *   table1_kernel   thread per (node, table row): cross-section geometry -> elevation, area, perimeter, conveyance, top
 *                   width, 1/n columns (readXsection :2272-2426); 501 rows x nodes x domains independent threads
 *   table2_kernel   thread per (node, row): dK/dA and the uniform-flow column (:2395-2402, :487-506); thread per
 *                   (node, column) for the column minima
 *   time_loop_kernel  ONE CTA per domain runs the whole simulation (dw_time_loop, diffusive_device.cuh): no launch, no
 *                   host round trip per time step; the CFL-adaptive step size is a CTA-local reduction
 * The solver arithmetic lives in diffusive_device.cuh; see its header for what runs in parallel and what is a chain.
 */
#include <cstdio>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/troute_b200.h"
#include "diffusive_setup.h"
#include "internal.h"

using namespace trtdw;

namespace {

constexpr int kLoopThreads = 256;
int g_device = 0;
double g_table_ms = 0.0, g_loop_ms = 0.0;
long long g_launches = 0;

__global__ void __launch_bounds__(256) table1_kernel(const Dom* __restrict__ doms)
{
    Dom D = doms[blockIdx.y];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)D.nm * D.mx * NEL) return;
    const int row = (int)(idx % NEL) + 1;
    const int node = (int)(idx / NEL);
    const int j = D.mstem[node / D.mx], i = node % D.mx + 1;
    if (i <= DW_FRNW(j, 1)) {
        if (D.mxnbathy == 0) dw_table_pass1(D, i, j, row);
        else dw_nat_pass1(D, i, j, row);
    }
}

/* surveyed cross sections: thread per (node, vertex) before table1_kernel, thread per node after it */
__global__ void __launch_bounds__(256) natprep_kernel(const Dom* __restrict__ doms)
{
    Dom D = doms[blockIdx.y];
    if (D.mxnbathy == 0) return;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)D.nm * D.mx * D.mxnbathy) return;
    const int ic = (int)(idx % D.mxnbathy) + 1;
    const int node = (int)(idx / D.mxnbathy);
    const int j = D.mstem[node / D.mx], i = node % D.mx + 1;
    if (i <= DW_FRNW(j, 1) && ic <= D.size_bathy[(i - 1) + (size_t)(j - 1) * D.mx]) dw_nat_prep(D, i, j, ic);
}

__global__ void __launch_bounds__(128) natsmooth_kernel(const Dom* __restrict__ doms)
{
    Dom D = doms[blockIdx.y];
    if (D.mxnbathy == 0) return;
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= D.nm * D.mx) return;
    const int j = D.mstem[node / D.mx], i = node % D.mx + 1;
    if (i <= DW_FRNW(j, 1)) dw_nat_smooth(D, i, j);
}

__global__ void __launch_bounds__(256) table2_kernel(const Dom* __restrict__ doms)
{
    Dom D = doms[blockIdx.y];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)D.nm * D.mx * NEL) return;
    const int row = (int)(idx % NEL) + 1;
    const int node = (int)(idx / NEL);
    const int j = D.mstem[node / D.mx], i = node % D.mx + 1;
    if (i <= DW_FRNW(j, 1)) dw_table_pass2(D, i, j, row);
}

__global__ void __launch_bounds__(256) tablemin_kernel(const Dom* __restrict__ doms)
{
    Dom D = doms[blockIdx.y];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)D.nm * D.mx * NCOL) return;
    const int col = (int)(idx % NCOL);
    const int node = (int)(idx / NCOL);
    const int j = D.mstem[node / D.mx], i = node % D.mx + 1;
    if (i <= DW_FRNW(j, 1)) dw_table_min(D, i, j, col);
}

__global__ void __launch_bounds__(kLoopThreads) time_loop_kernel(const Dom* __restrict__ doms)
{
    __shared__ Dom D;
    if (threadIdx.x == 0) D = doms[blockIdx.x];
    __syncthreads();
    dw_time_loop(D);
}

/* One device allocation per call for every domain: pools, look-up tables and outputs are carved out of it at 256-byte
 * alignment (a cudaMalloc / cudaFree pair per array and domain would cost more than the kernels for small domains). */
struct Arena {
    char* base = nullptr;
    size_t size = 0, off = 0;
    ~Arena() { cudaFree(base); }
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    void* take(size_t bytes)
    {
        void* p = base + off;
        off += pad(bytes);
        return p;
    }
};

struct DevDomain {
    DomHost H;
    double *d_pool = nullptr, *d_tab = nullptr, *d_tabmin = nullptr, *d_out = nullptr;
    int* i_pool = nullptr;
    unsigned char* b_pool = nullptr;
    size_t bytes() const
    {
        return Arena::pad(H.dpool.size() * sizeof(double)) + Arena::pad(H.ipool.size() * sizeof(int)) + Arena::pad(H.bpool.size()) +
               Arena::pad(H.n_nodes * NCOL * LD * sizeof(double)) + Arena::pad(H.n_nodes * NCOL * sizeof(double)) +
               Arena::pad(3 * H.n_out * sizeof(double));
    }
};

#define CUD(call)                                                                                                 \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) {                                                                                 \
            char b__[512];                                                                                        \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return trt_internal_fail(TRT_ERR_CUDA, b__);                                                          \
        }                                                                                                         \
    } while (0)

/* carve the arrays of one domain out of the (zero-filled) arena and copy its host pools in */
int upload_domain(DevDomain& V, Arena& A, Dom& out)
{
    DomHost& H = V.H;
    const size_t db = H.dpool.size() * sizeof(double), ib = H.ipool.size() * sizeof(int), bb = H.bpool.size();
    V.d_pool = (double*)A.take(db);
    V.i_pool = (int*)A.take(ib);
    V.b_pool = (unsigned char*)A.take(bb);
    V.d_tab = (double*)A.take(H.n_nodes * NCOL * LD * sizeof(double));
    V.d_tabmin = (double*)A.take(H.n_nodes * NCOL * sizeof(double));
    V.d_out = (double*)A.take(3 * H.n_out * sizeof(double));          /* q_ev_g = elv_ev_g = depth_ev_g = 0 (:391-393) */
    if (A.off > A.size) return trt_internal_fail(TRT_ERR_STATE, "diffusive arena overflow");
    CUD(cudaMemcpy(V.d_pool, H.dpool.data(), db, cudaMemcpyHostToDevice));
    CUD(cudaMemcpy(V.i_pool, H.ipool.data(), ib, cudaMemcpyHostToDevice));
    CUD(cudaMemcpy(V.b_pool, H.bpool.data(), bb, cudaMemcpyHostToDevice));
    Dom D = dw_rebase(H, V.d_pool, V.i_pool, V.b_pool);
    D.tab = V.d_tab; D.tabmin = V.d_tabmin;
    D.q_ev = V.d_out; D.elv_ev = V.d_out + H.n_out; D.depth_ev = V.d_out + 2 * H.n_out;
    out = D;
    return TRT_OK;
}

}  // namespace

extern "C" {

int trt_diffusive_set_device(int device)
{
    if (device < 0) return trt_internal_fail(TRT_ERR_INVALID, "device must be >= 0");
    g_device = device;
    return TRT_OK;
}

int trt_diffusive_last_run(double* table_ms, double* loop_ms, long long* launches)
{
    if (table_ms) *table_ms = g_table_ms;
    if (loop_ms) *loop_ms = g_loop_ms;
    if (launches) *launches = g_launches;
    return TRT_OK;
}

int trt_diffnw_batch(int n_domains, const void* const* argv)
{
    if (n_domains < 1 || !argv) return trt_internal_fail(TRT_ERR_INVALID, "n_domains < 1 or NULL argument list");
    std::vector<DevDomain> doms((size_t)n_domains);
    std::vector<DiffnwArgs> args((size_t)n_domains);
    for (int d = 0; d < n_domains; ++d) {
        const void* const* a = argv + (size_t)d * 42;
        for (int k = 0; k < 42; ++k)
            if (!a[k]) {
                char b[96];
                snprintf(b, sizeof b, "domain %d: argument %d of c_diffnw is NULL", d, k);
                return trt_internal_fail(TRT_ERR_INVALID, b);
            }
        DiffnwArgs& A = args[(size_t)d];
        A = DiffnwArgs{(const double*)a[0], (const int*)a[1], (const int*)a[2], (const int*)a[3], (const int*)a[4], (const int*)a[5],
                       (const int*)a[6], (const int*)a[7], (const int*)a[8], (const double*)a[9], (const double*)a[10],
                       (const double*)a[11], (const double*)a[12], (const double*)a[13], (const double*)a[14],
                       (const double*)a[15], (double*)a[16], (const double*)a[17], (const double*)a[18], (const int*)a[19],
                       (const int*)a[20], (const double*)a[21], (const double*)a[22], (const double*)a[23],
                       (const double*)a[24], (const int*)a[25], (const double*)a[26], (const int*)a[27], (const double*)a[28],
                       (const double*)a[29], (const double*)a[30], (const int*)a[31], (const double*)a[32], (const int*)a[33],
                       (const double*)a[34], (const int*)a[35], (const int*)a[36], (const double*)a[37], (const double*)a[38],
                       (double*)a[39], (double*)a[40], (double*)a[41]};
        const std::string err = dw_build_host(A, doms[(size_t)d].H);
        if (!err.empty()) {
            const std::string m = "domain " + std::to_string(d) + ": " + err;
            return trt_internal_fail(TRT_ERR_INVALID, m.c_str());
        }
    }
    CUD(cudaSetDevice(g_device));
    std::vector<Dom> hdoms((size_t)n_domains);
    long long max_rows = 0, max_cols = 0, max_verts = 0, max_nodes = 0;
    Arena arena;
    for (int d = 0; d < n_domains; ++d) arena.size += doms[(size_t)d].bytes();
    arena.size += Arena::pad(sizeof(Dom) * (size_t)n_domains);
    CUD(cudaMalloc((void**)&arena.base, arena.size));
    CUD(cudaMemset(arena.base, 0, arena.size));
    for (int d = 0; d < n_domains; ++d) {
        const int rc = upload_domain(doms[(size_t)d], arena, hdoms[(size_t)d]);
        if (rc != TRT_OK) return rc;
        const long long nodes = (long long)hdoms[(size_t)d].nm * hdoms[(size_t)d].mx;
        max_rows = std::max(max_rows, nodes * NEL);
        max_cols = std::max(max_cols, nodes * NCOL);
        max_nodes = std::max(max_nodes, nodes);
        max_verts = std::max(max_verts, nodes * hdoms[(size_t)d].mxnbathy);
    }
    Dom* d_doms = (Dom*)arena.take(sizeof(Dom) * (size_t)n_domains);
    if (arena.off > arena.size) return trt_internal_fail(TRT_ERR_STATE, "diffusive arena overflow");
    CUD(cudaMemcpy(d_doms, hdoms.data(), sizeof(Dom) * (size_t)n_domains, cudaMemcpyHostToDevice));
    struct Events {
        cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } ev;
    for (cudaEvent_t& x : ev.e) CUD(cudaEventCreate(&x));
    cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2];
    CUD(cudaEventRecord(e0));
    const dim3 g1((unsigned)((max_rows + 255) / 256), (unsigned)n_domains), gm((unsigned)((max_cols + 255) / 256), (unsigned)n_domains);
    if (max_verts > 0) {
        const dim3 gv((unsigned)((max_verts + 255) / 256), (unsigned)n_domains);
        natprep_kernel<<<gv, 256>>>(d_doms);
    }
    table1_kernel<<<g1, 256>>>(d_doms);
    if (max_verts > 0) {
        const dim3 gn((unsigned)((max_nodes + 127) / 128), (unsigned)n_domains);
        natsmooth_kernel<<<gn, 128>>>(d_doms);
    }
    table2_kernel<<<g1, 256>>>(d_doms);
    tablemin_kernel<<<gm, 256>>>(d_doms);
    CUD(cudaGetLastError());
    CUD(cudaEventRecord(e1));
    time_loop_kernel<<<n_domains, kLoopThreads>>>(d_doms);
    CUD(cudaGetLastError());
    CUD(cudaEventRecord(e2));
    CUD(cudaDeviceSynchronize());
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) g_table_ms = ms;
    if (cudaEventElapsedTime(&ms, e1, e2) == cudaSuccess) g_loop_ms = ms;
    g_launches = max_verts > 0 ? 6 : 4;
    for (int d = 0; d < n_domains; ++d) {
        DevDomain& V = doms[(size_t)d];
        const size_t nb = V.H.n_out * sizeof(double);
        CUD(cudaMemcpy(args[(size_t)d].q_ev_g, V.d_out, nb, cudaMemcpyDeviceToHost));
        CUD(cudaMemcpy(args[(size_t)d].elv_ev_g, V.d_out + V.H.n_out, nb, cudaMemcpyDeviceToHost));
        CUD(cudaMemcpy(args[(size_t)d].depth_ev_g, V.d_out + 2 * V.H.n_out, nb, cudaMemcpyDeviceToHost));
        int status = 0;
        CUD(cudaMemcpy(&status, hdoms[(size_t)d].status, sizeof(int), cudaMemcpyDeviceToHost));
        if (status != 0) {
            const std::string m = "domain " + std::to_string(d) + ": the adaptive time step collapsed to <= 0 (the reference would "
                                  "never leave its time loop, diffusive.f90:655)";
            return trt_internal_fail(TRT_ERR_STATE, m.c_str());
        }
    }
    return TRT_OK;
}

int trt_c_diffnw(const double* timestep_ar_g, const int* nts_ql_g, const int* nts_ub_g, const int* nts_db_g, const int* ntss_ev_g,
                 const int* nts_qtrib_g, const int* nts_da_g, const int* mxncomp_g, const int* nrch_g, const double* z_ar_g,
                 const double* bo_ar_g, const double* traps_ar_g, const double* tw_ar_g, const double* twcc_ar_g,
                 const double* mann_ar_g, const double* manncc_ar_g, double* so_ar_g, const double* dx_ar_g, const double* iniq,
                 const int* frnw_col, const int* frnw_ar_g, const double* qlat_g, const double* ubcd_g, const double* dbcd_g,
                 const double* qtrib_g, const int* paradim, const double* para_ar_g, const int* mxnbathy_g,
                 const double* x_bathy_g, const double* z_bathy_g, const double* mann_bathy_g, const int* size_bathy_g,
                 const double* usgs_da_g, const int* usgs_da_reach_g, const double* rdx_ar_g, const int* cwnrow_g,
                 const int* cwncol_g, const double* crosswalk_g, const double* z_thalweg_g, double* q_ev_g, double* elv_ev_g,
                 double* depth_ev_g)
{
    const void* argv[42] = {timestep_ar_g, nts_ql_g, nts_ub_g, nts_db_g, ntss_ev_g, nts_qtrib_g, nts_da_g, mxncomp_g, nrch_g,
                            z_ar_g, bo_ar_g, traps_ar_g, tw_ar_g, twcc_ar_g, mann_ar_g, manncc_ar_g, so_ar_g, dx_ar_g, iniq,
                            frnw_col, frnw_ar_g, qlat_g, ubcd_g, dbcd_g, qtrib_g, paradim, para_ar_g, mxnbathy_g, x_bathy_g,
                            z_bathy_g, mann_bathy_g, size_bathy_g, usgs_da_g, usgs_da_reach_g, rdx_ar_g, cwnrow_g, cwncol_g,
                            crosswalk_g, z_thalweg_g, q_ev_g, elv_ev_g, depth_ev_g};
    return trt_diffnw_batch(1, argv);
}

}  // extern "C"
