// nk_kernels.cu -- the translation unit of libnk_b200.so: sm_100a kernels (included below) + the host side of the C ABI
// of the Nano-kappa particle loop (include/nk_b200.h).
//
// One timestep = two launches on one stream (three with the per-(mode, subvolume) tables), no host synchronisation:
//
//   k_step / k_step_tab   (nk_stream.cuh) streaming pass over the particle SoA, 84 algorithmic bytes per particle:
//               [lifetime relaxation of the previous step] -> drift -> nearest-subvolume -> block-privatised per-SV
//               energy / count / flux bins.  Particles whose next collision falls inside this step are only appended to
//               a hit list; the prologue scans the reservoir tables and fills the emission list.
//   k_rare      (nk_rare.cuh) one thread per hit-list / emission-list entry: absorb / periodic / rough event loop, new
//               particles into recycled slots.  Its last block closes the step: per-SV sums (exchanged with the peer
//               GPUs over NVLink when sharded) -> energy density -> temperature, heat flux and kappa on convergence
//               steps, reservoir balances, reset.
//   k_mode_tables  (nk_stream.cuh) next step's {n0, exp(-dt/tau)} per (mode, subvolume) when the table variant is active.
//
// lifetime_scattering(k) needs T_sv(k), a grid-wide dependency; instead of a second pass over the particles it is applied
// at the head of k_step(k+1) (nothing reads `occ` in between except outputs, which call nk_flush_relaxation).  Energies
// use the pre-relaxation occupation and the previous step's T_sv exactly like the reference (Population.py:704-713,
// :1754-1756).
//
// nk_types.cuh    parameter block (NkP), device-resident step state (NkDyn), accumulator / result layouts
// nk_device.cuh   exact-arithmetic helpers, Philox, table look-ups, classification, ray-triangle search, boundary events
// nk_tiles.cuh    TMA-staged triangle tiles shared by the ray kernels
// nk_ops.cuh ... nk_hostpipe.cuh   kernels, in dependency order
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#include <thread>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <algorithm>

#include "../../include/nk_b200.h"
#include "nk_device.cuh"

#include "nk_tiles.cuh"
#include "nk_ops.cuh"
#include "nk_stream.cuh"
#include "nk_rare.cuh"
#include "nk_sort.cuh"
#include "nk_hostpipe.cuh"

// =================================================================================================
// host side
// =================================================================================================
struct nk_ctx {
    int device = 0;
    NkP P;
    cudaStream_t stream = 0;
    std::string err;
    std::vector<void*> owned;      // device allocations freed in nk_destroy
    int n_sm = 148;
    int has_rough = 0;
    bool particles_bound = false;
    // host mirrors needed to rebuild derived tables
    std::vector<double> h_tau, h_Tg, h_mode;
    double hot_lo = 0, hot_hi = 0;
    int step_blocks = 0;
    int step_blocks_variant = -1;
    bool force_tiled = false;      // NK_RARE_TILED=1: use the tiled rare-path kernel also for small meshes (tests)
    bool rare_attr_set = false;
    int rare_blocks_per_sm = 16;   // NK_RARE_BLOCKS_PER_SM: grid cap of the rare-path kernel (items beyond it are taken by the grid-stride loop)
    double* snap_host[4] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t snap_ev[4] = {nullptr, nullptr, nullptr, nullptr}; int snap_len = 0; unsigned snap_next = 0;
    bool last_rare_tiled = false;
    int* sort_count = nullptr; int* sort_cursor = nullptr; int* mode_first_dev = nullptr; long long* sort_totals = nullptr;
    bool use_tab = true;           // NK_STEP_TAB=0 disables the per-(mode, subvolume) table variant
    bool force_tab = false;        // NK_STEP_TAB=force: use it regardless of the particle count (tests)
    bool tab_dirty = true;         // T_sv changed since k_mode_tables last ran
    long long h_slots_hint = 0;
    int last_variant = 0;
    // chunked host-buffer pipeline (nk_advance_host)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    void* patch_dev = nullptr; void* patch_host = nullptr; long long patch_cap = 0;
    unsigned int* patch_count_dev = nullptr;
    void* cold_dev = nullptr; void* cold_host = nullptr; long long cold_cap = 0;
    cudaEvent_t ev_cold = nullptr;
    bool sparse_cold = true;       // NK_HOST_SPARSE=0: upload the cold arrays densely
    bool l2_persist = false, l2_window_set = false;      // NK_L2_PERSIST=1 (experiment)
    bool use_pipeline = true;      // NK_HOST_PIPELINE=0 disables it
    long long xfer_h2d = 0, xfer_d2h = 0;   // bytes of the last nk_advance_host call
    void* comm_block = nullptr;    // flags + mailboxes of the fused exchange
    unsigned int comm_imported = 0;
    std::vector<void*> ipc_open;   // peer mailboxes mapped with cudaIpcOpenMemHandle    // slot count at the last nk_set_slot_count
    bool profiling = false;
    long long h_step = 0;          // host mirror of NkDyn::step
    bool h_relax_pending = false;  // host mirror of NkDyn::relax_pending
    std::vector<cudaEvent_t> ev;   // 5 events per profiled step
    size_t ev_used = 0;
};

static void nk_prof_mark(nk_ctx* ctx) {
    if (!ctx->profiling) return;
    if (ctx->ev_used == ctx->ev.size()) {
        cudaEvent_t e; cudaEventCreate(&e); ctx->ev.push_back(e);
    }
    cudaEventRecord(ctx->ev[ctx->ev_used++], ctx->stream);
}

static std::string g_create_err;
static bool getenv_is_one(const char* name) { const char* e = getenv(name); return e && !strcmp(e, "1"); }

#define NK_CK(call)                                                                          \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

template <class T>
static T* nk_upload(nk_ctx* ctx, const T* h, size_t n) {
    T* d = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    ctx->owned.push_back(d);
    if (n && h) {
        if (cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    } else {
        cudaMemset(d, 0, bytes);
    }
    return d;
}
#define NK_UP(dst, T, h, n)                                                     \
    do {                                                                        \
        dst = nk_upload<T>(ctx, h, n);                                          \
        if (!dst) { ctx->err = "device allocation/upload failed: " #dst; return -1; } \
    } while (0)

extern "C" {

int nk_version(void) { return 100; }

const char* nk_last_error(nk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int nk_create(int device, nk_ctx** out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return -1;
    }
    if (device < 0 || device >= n) { g_create_err = "device index out of range"; return -1; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return -1; }
    nk_ctx* ctx = new nk_ctx();
    memset(&ctx->P, 0, sizeof(NkP));
    ctx->device = device;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->n_sm = prop.multiProcessorCount;
    ctx->P.world = 1;
    ctx->P.slot_lo = 0; ctx->P.slot_hi = 0x7fffffffffffffffLL; ctx->P.scan_emit = 1;
    if (const char* e = getenv("NK_HOST_PIPELINE")) ctx->use_pipeline = strcmp(e, "0") != 0;
    if (const char* e = getenv("NK_HOST_SPARSE")) ctx->sparse_cold = strcmp(e, "0") != 0;
    if (const char* e = getenv("NK_L2_PERSIST")) ctx->l2_persist = strcmp(e, "0") != 0;
    ctx->P.trace = nullptr;
    if (const char* e = getenv("NK_TRACE")) {
        if (strcmp(e, "0") != 0 && cudaMalloc(&ctx->P.trace, 8 * sizeof(unsigned long long)) == cudaSuccess)
            cudaMemset(ctx->P.trace, 0, 8 * sizeof(unsigned long long));
    }
    if (const char* e = getenv("NK_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));   // experiment: 32 / 64 / 128
    if (const char* e = getenv("NK_STEP_TAB")) { ctx->use_tab = strcmp(e, "0") != 0; ctx->force_tab = !strcmp(e, "force"); }
    ctx->force_tiled = getenv_is_one("NK_RARE_TILED");
    if (const char* e = getenv("NK_RARE_BLOCKS_PER_SM")) ctx->rare_blocks_per_sm = std::max(1, atoi(e));
    NkDyn z; memset(&z, 0, sizeof(z));
    ctx->P.dyn = nk_upload<NkDyn>(ctx, &z, 1);
    *out = ctx;
    return 0;
}

void nk_destroy(nk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void* p : ctx->ipc_open) cudaIpcCloseMemHandle(p);
    for (void* p : ctx->owned) cudaFree(p);
    for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_k) cudaEventDestroy(e);
    if (ctx->s_in) { cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_out); }
    if (ctx->patch_dev) cudaFree(ctx->patch_dev);
    if (ctx->patch_host) cudaFreeHost(ctx->patch_host);
    if (ctx->cold_dev) cudaFree(ctx->cold_dev);
    if (ctx->cold_host) cudaFreeHost(ctx->cold_host);
    if (ctx->ev_cold) cudaEventDestroy(ctx->ev_cold);
    if (ctx->patch_count_dev) cudaFree(ctx->patch_count_dev);
    for (int k = 0; k < 4; ++k) { if (ctx->snap_host[k]) cudaFreeHost(ctx->snap_host[k]); if (ctx->snap_ev[k]) cudaEventDestroy(ctx->snap_ev[k]); }
    delete ctx;
}

int nk_set_stream(nk_ctx* ctx, void* s) { ctx->stream = (cudaStream_t)s; return 0; }
int nk_synchronize(nk_ctx* ctx) { NK_CK(cudaStreamSynchronize(ctx->stream)); return 0; }

int nk_set_mesh(nk_ctx* ctx, int F, const double* fn, const double* fk, const double* flo, const double* fhi,
                const double* fo, const double* fb, const int32_t* ff, const double* fverts, const double* fareas,
                int nf, const int32_t* bc, const int32_t* partner, const int32_t* fres, const int32_t* frough,
                const double* fnormal, const double* fcentroid, const double* farea,
                const int32_t* ffptr, const int32_t* ffaces, const double* bounds) {
    cudaSetDevice(ctx->device);
    std::vector<NkFace> faces(F);
    // scale of the mesh for the error bounds of the ray pre-filter: largest |coordinate| of the bounding box / |plane offset|
    double mesh_scale = 1e-300;
    for (int k = 0; k < 6; ++k) mesh_scale = std::max(mesh_scale, std::fabs(bounds[k]));
    for (int f = 0; f < F; ++f) mesh_scale = std::max(mesh_scale, std::fabs(fk[f]));
    ctx->P.mesh_scale = mesh_scale;
    for (int f = 0; f < F; ++f) {
        NkFace& T = faces[f];
        T.nx = fn[3 * f]; T.ny = fn[3 * f + 1]; T.nz = fn[3 * f + 2]; T.k = fk[f];
        T.lox = flo[3 * f] - NK_TOL; T.loy = flo[3 * f + 1] - NK_TOL; T.loz = flo[3 * f + 2] - NK_TOL;
        T.hix = fhi[3 * f] + NK_TOL; T.hiy = fhi[3 * f + 1] + NK_TOL; T.hiz = fhi[3 * f + 2] + NK_TOL;
        T.mx = 0.5 * (T.lox + T.hix); T.my = 0.5 * (T.loy + T.hiy); T.mz = 0.5 * (T.loz + T.hiz);
        T.ex = 0.5 * (T.hix - T.lox) + 1e-6 * mesh_scale; T.ey = 0.5 * (T.hiy - T.loy) + 1e-6 * mesh_scale; T.ez = 0.5 * (T.hiz - T.loz) + 1e-6 * mesh_scale;
        T.ox = fo[3 * f]; T.oy = fo[3 * f + 1]; T.oz = fo[3 * f + 2];
        // inverse of A = face_basis_matrix (columns b1, b2, n), rows 0 and 1, by cofactors
        const double* A = fb + 9 * f;
        double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], g = A[5], h = A[6], i = A[7], j = A[8];
        double det = a * (e * j - g * i) - b * (d * j - g * h) + c * (d * i - e * h);
        T.ia0 = (e * j - g * i) / det; T.ia1 = (c * i - b * j) / det; T.ia2 = (b * g - c * e) / det;
        T.ib0 = (g * h - d * j) / det; T.ib1 = (a * j - c * h) / det; T.ib2 = (c * d - a * g) / det;
        T.facet = (double)ff[f];
    }
    NkP& P = ctx->P;
    P.F = F; P.nf = nf; ctx->rare_attr_set = false;
    NkFace* dfaces; NK_UP(dfaces, NkFace, faces.data(), (size_t)F); P.faces = dfaces;
    int* di; double* dd;
    NK_UP(di, int, bc, nf); P.facet_bc = di;
    NK_UP(di, int, partner, nf); P.facet_partner = di;
    NK_UP(di, int, fres, nf); P.facet_res = di;
    NK_UP(di, int, frough, nf); P.facet_rough = di;
    NK_UP(dd, double, fnormal, 3 * (size_t)nf); P.facet_normal = dd;
    NK_UP(dd, double, fcentroid, 3 * (size_t)nf); P.facet_centroid = dd;
    NK_UP(dd, double, farea, nf); P.facet_area = dd;
    NK_UP(dd, double, fverts, 9 * (size_t)F); P.face_vertices = dd;
    for (int k = 0; k < 3; ++k) { P.blo[k] = bounds[k]; P.bhi[k] = bounds[3 + k]; }
    // reservoir sampling tables: facets with res >= 0, in reservoir order
    int R = 0;
    for (int f = 0; f < nf; ++f) if (fres[f] >= 0) R = std::max(R, fres[f] + 1);
    std::vector<int> rptr(R + 1, 0), rfaces; std::vector<double> rcdf;
    for (int r = 0; r < R; ++r) {
        int facet = -1;
        for (int f = 0; f < nf; ++f) if (fres[f] == r) facet = f;
        rptr[r] = (int)rfaces.size();
        if (facet >= 0) {
            // np.random.choice(faces, p = areas/areas.sum()): cdf = cumsum(p); cdf /= cdf[-1]
            double tot = 0.0;
            for (int q = ffptr[facet]; q < ffptr[facet + 1]; ++q) tot += fareas[ffaces[q]];
            std::vector<double> cdf; double run = 0.0;
            for (int q = ffptr[facet]; q < ffptr[facet + 1]; ++q) { run += fareas[ffaces[q]] / tot; cdf.push_back(run); rfaces.push_back(ffaces[q]); }
            for (double& cval : cdf) cval /= run;
            rcdf.insert(rcdf.end(), cdf.begin(), cdf.end());
        }
    }
    rptr[R] = (int)rfaces.size();
    NK_UP(di, int, rptr.data(), rptr.size()); P.res_face_ptr = di;
    NK_UP(di, int, rfaces.data(), rfaces.size()); P.res_faces = di;
    NK_UP(dd, double, rcdf.data(), rcdf.size()); P.res_face_cdf = dd;
    return 0;
}

int nk_set_subvols(nk_ctx* ctx, int S, const double* centres, const double* volumes, int is_slice, int axis, int interp) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (S < 1) { ctx->err = "n_subvols must be >= 1"; return -1; }
    if (S > 1400) { ctx->err = "n_subvols above 1400 is not supported: the per-block subvolume bins no longer fit the 227 KB of shared memory"; return -1; }
    if (interp == NK_INTERP_LINEAR && !is_slice) { ctx->err = "linear temperature interpolation needs slice subvolumes; the reference falls back to NK_INTERP_RADIAL there"; return -1; }
    if (interp == NK_INTERP_RADIAL && is_slice) { ctx->err = "radial temperature interpolation on slice subvolumes is singular upstream (collinear centres)"; return -1; }
    if (interp < NK_INTERP_NEAREST || interp > NK_INTERP_RADIAL) { ctx->err = "unknown temp_interp"; return -1; }
    P.S = S; P.is_slice = is_slice; P.axis = axis; P.interp = interp;
    P.rbf_nd = 0; P.rbf_w = nullptr;
    {
        std::vector<double> z(S + 4, 0.0);
        NK_UP(P.rbf_coef, double, z.data(), z.size());
    }
    double* dd;
    NK_UP(dd, double, centres, 3 * (size_t)S); P.svc = dd;
    NK_UP(dd, double, volumes, S); P.sv_volume = dd;
    std::vector<double> ax(S), mid(S > 1 ? S - 1 : 1, 0.0);
    for (int s = 0; s < S; ++s) ax[s] = centres[3 * s + axis];
    for (int s = 0; s + 1 < S; ++s) mid[s] = ax[s + 1] / 2.0 + ax[s] / 2.0;   // x_bds = x/2; x_bds[1:] + x_bds[:-1]
    NK_UP(dd, double, ax.data(), S); P.sv_axis = dd;
    NK_UP(dd, double, mid.data(), mid.size()); P.sv_mid = dd;
    P.sv_inv_dx = (S > 1 && ax[S - 1] != ax[0]) ? (S - 1) / (ax[S - 1] - ax[0]) : 0.0;
    P.sv_x0 = (S > 1 && P.sv_inv_dx > 0) ? ax[0] - 0.5 / P.sv_inv_dx : 0.0;
    std::vector<double> T0(S, 0.0);
    NK_UP(P.T_sv, double, T0.data(), S);
    return 0;
}

static int nk_build_tau4(nk_ctx* ctx) {
    NkP& P = ctx->P;
    if (ctx->h_tau.empty()) return 0;
    int NT = P.NT, M = P.M;
    // slabs i0..i0+3 cover [Tg[i0], Tg[i0+3]]: centre the window on the expected temperature range
    int ilo = 0;
    while (ilo + 1 < NT - 1 && ctx->h_Tg[ilo + 1] <= ctx->hot_lo) ++ilo;
    int ihi = ilo;
    while (ihi + 1 < NT - 1 && ctx->h_Tg[ihi + 1] <= ctx->hot_hi) ++ihi;
    int i0 = ilo - std::max(0, (2 - (ihi - ilo)) / 2);
    i0 = std::max(0, std::min(i0, NT - 4));
    if (NT < 4) i0 = 0;
    std::vector<NkTau4> t4(M);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < 4; ++k) t4[m].t[k] = (i0 + k < NT) ? ctx->h_tau[(size_t)(i0 + k) * M + m] : 0.0;
    NkTau4* d; NK_UP(d, NkTau4, t4.data(), (size_t)M);
    P.tau4 = d; P.tau_i0 = (NT >= 4) ? i0 : -1000000;
    // 64 B hot record per mode for the streaming kernel: {omega, v_g} + the same four tau slabs
    std::vector<NkModeHot> hot(M);
    for (int m = 0; m < M; ++m) {
        hot[m].omega = ctx->h_mode[4 * (size_t)m]; hot[m].vx = ctx->h_mode[4 * (size_t)m + 1];
        hot[m].vy = ctx->h_mode[4 * (size_t)m + 2]; hot[m].vz = ctx->h_mode[4 * (size_t)m + 3];
        for (int k = 0; k < 4; ++k) hot[m].t[k] = t4[m].t[k];
    }
    NkModeHot* dh; NK_UP(dh, NkModeHot, hot.data(), (size_t)M);
    P.mhot = dh;
    ctx->step_blocks = 0; ctx->tab_dirty = true;
    return 0;
}

int nk_set_phonon(nk_ctx* ctx, int Q, int J, int NT, const double* Tg, const double* omega, const double* vg,
                  const double* tau, double hbar, double kb, double V_uc, int64_t n_active,
                  int nE, const double* Ea, const double* Ta) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (NT < 2) { ctx->err = "need at least two temperatures"; return -1; }
    P.Q = Q; P.J = J; P.M = Q * J; P.NT = NT;
    int M = P.M;
    std::vector<NkMode> mp(M);
    for (int m = 0; m < M; ++m) { mp[m].omega = omega[m]; mp[m].vx = vg[3 * m]; mp[m].vy = vg[3 * m + 1]; mp[m].vz = vg[3 * m + 2]; }
    NkMode* dm; NK_UP(dm, NkMode, mp.data(), (size_t)M); P.mprop = dm;
    double* dd;
    NK_UP(dd, double, Tg, NT); P.Tg = dd;
    NK_UP(dd, double, tau, (size_t)NT * M); P.tau = dd;
    NK_UP(dd, double, Ea, nE); P.Ea = dd;
    NK_UP(dd, double, Ta, nE); P.Ta = dd;
    P.nE = nE; P.hbar = hbar; P.kb = kb; P.V_uc = V_uc; P.n_active = (double)n_active;
    P.dens_norm = (double)Q * V_uc;
    P.Tg_inv_d = 1.0 / (Tg[1] - Tg[0]);
    P.Tg0 = Tg[0]; P.Tg_d = Tg[1] - Tg[0]; P.Tg_uniform = P.Tg_d > 0.0 ? 1 : 0;
    for (int i = 0; i < NT && P.Tg_uniform; ++i) if (Tg[i] != P.Tg0 + (double)i * P.Tg_d) P.Tg_uniform = 0;
    P.Ta_inv_d = (nE > 1 && Ta[nE - 1] != Ta[0]) ? (nE - 1) / (Ta[nE - 1] - Ta[0]) : 0.0;
    ctx->h_mode.resize(4 * (size_t)M);
    for (int m = 0; m < M; ++m) { ctx->h_mode[4 * (size_t)m] = mp[m].omega; ctx->h_mode[4 * (size_t)m + 1] = mp[m].vx; ctx->h_mode[4 * (size_t)m + 2] = mp[m].vy; ctx->h_mode[4 * (size_t)m + 3] = mp[m].vz; }
    ctx->h_tau.assign(tau, tau + (size_t)NT * M);
    ctx->h_Tg.assign(Tg, Tg + NT);
    if (ctx->hot_hi == 0) { ctx->hot_lo = 295.0; ctx->hot_hi = 305.0; }
    return nk_build_tau4(ctx);
}

int nk_set_population(nk_ctx* ctx, double dt, int norm_mean, double density, int n_dt_to_conv, uint64_t seed,
                      double unit_flux, double a_in_m, double hot_lo, double hot_hi) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    P.dt = dt; P.norm_mean = norm_mean; P.particle_density = density; P.n_dt_to_conv = n_dt_to_conv;
    P.seed_lo = (unsigned int)(seed & 0xFFFFFFFFull); P.seed_hi = (unsigned int)(seed >> 32);
    P.eVpsa2_in_Wm2 = unit_flux; P.a_in_m = a_in_m;
    ctx->hot_lo = hot_lo; ctx->hot_hi = hot_hi;
    return nk_build_tau4(ctx);
}

static int nk_alloc_scratch(nk_ctx* ctx) {
    NkP& P = ctx->P;
    double* dd;
    NK_UP(dd, double, (const double*)nullptr, (size_t)nk_acc_len(P.S, P.R)); P.acc = dd;
    { unsigned long long* dq; NK_UP(dq, unsigned long long, (const unsigned long long*)nullptr, (size_t)2 * nk_acc_len(P.S, P.R)); P.acc_q = dq; }
    NK_UP(dd, double, (const double*)nullptr, (size_t)4 * std::max(P.R, 1)); P.res_acc = dd;
    NK_UP(dd, double, (const double*)nullptr, (size_t)nk_out_len(P.S, P.R)); P.out = dd;
    P.hot_tab = nullptr;
    if (P.is_slice && P.interp == NK_INTERP_NEAREST && ctx->use_tab) {
        double2* dt2 = nullptr;
        if (cudaMalloc(&dt2, (size_t)P.M * P.S * sizeof(double2)) == cudaSuccess) { ctx->owned.push_back(dt2); P.hot_tab = dt2; }
        else cudaGetLastError();          // not enough memory: the direct variant is used
    }
    ctx->tab_dirty = true; ctx->step_blocks = 0; ctx->rare_attr_set = false;
    return 0;
}

int nk_set_reservoirs(nk_ctx* ctx, int R, const int32_t* res_facet, const double* res_T, const double* enter_prob,
                      const double* res_counter) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (P.M == 0 || P.S == 0) { ctx->err = "call nk_set_phonon and nk_set_subvols before nk_set_reservoirs"; return -1; }
    P.R = R;
    int* di; double* dd;
    NK_UP(di, int, res_facet, R); P.res_facet = di;
    NK_UP(dd, double, res_T, R); P.res_T = dd;
    NK_UP(dd, double, enter_prob, (size_t)R * P.M); P.enter_prob = dd;
    NK_UP(dd, double, res_counter, (size_t)R * P.M); P.res_counter = dd;
    int2* de; NK_UP(de, int2, (const int2*)nullptr, (size_t)std::max(R, 1) * P.M); P.emitlist = de;
    P.newslots_cap = std::max<long long>((long long)std::max(R, 1) * P.M * 2, 1 << 20);
    int* dn; NK_UP(dn, int, (const int*)nullptr, (size_t)P.newslots_cap); P.newslots = dn;
    // emission-mode extras (nk_set_reservoir_mode): initial N_leaving (Population.py:344), roulette of one_to_one (:465-466)
    P.res_gen = NK_RESGEN_CONSTANT;
    std::vector<double> nl(std::max(R, 1), 0.0), rou((size_t)std::max(R, 1) * P.M, 0.0);
    for (int r = 0; r < R; ++r) {
        const double* p = enter_prob + (size_t)r * P.M;
        double sum = 0.0;                                    // np.sum is pairwise, np.cumsum sequential
        for (int m = 0; m < P.M; ++m) { sum += p[m]; rou[(size_t)r * P.M + m] = sum; }
        double mx = 0.0;
        for (int m = 0; m < P.M; ++m) mx = std::max(mx, rou[(size_t)r * P.M + m]);
        for (int m = 0; m < P.M; ++m) rou[(size_t)r * P.M + m] /= mx;
        nl[r] = std::nearbyint(sum);
    }
    NK_UP(dd, double, rou.data(), rou.size()); P.res_roulette = dd;
    NK_UP(dd, double, nl.data(), nl.size()); P.res_nleave = dd;
    NK_UP(dd, double, (const double*)nullptr, (size_t)std::max(R, 1) * P.M); P.emit_u = dd;
    { unsigned char* du; NK_UP(du, unsigned char, (const unsigned char*)nullptr, (size_t)std::max(R, 1) * P.M); P.res_fire = du; }
    return nk_alloc_scratch(ctx);
}

int nk_set_reservoir_mode(nk_ctx* ctx, int mode, const double* n_leaving) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (!P.res_nleave) { ctx->err = "nk_set_reservoirs first"; return -1; }
    if (mode < NK_RESGEN_CONSTANT || mode > NK_RESGEN_ONE_TO_ONE) { ctx->err = "unknown reservoir generation mode"; return -1; }
    P.res_gen = mode;
    if (n_leaving && P.R > 0) {
        NK_CK(cudaStreamSynchronize(ctx->stream));
        NK_CK(cudaMemcpy(P.res_nleave, n_leaving, P.R * sizeof(double), cudaMemcpyHostToDevice));
    }
    return 0;
}

int nk_get_res_counter(nk_ctx* ctx, double* h) {
    cudaSetDevice(ctx->device);
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(h, ctx->P.res_counter, (size_t)ctx->P.R * ctx->P.M * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int nk_set_boundary_luts(nk_ctx* ctx, int Fr, const double* spec, const uint8_t* ts, const int32_t* so, const double* rou) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    P.Fr = Fr;
    size_t n = (size_t)Fr * P.M;
    double* dd; unsigned char* du; int* di;
    NK_UP(dd, double, spec, n); P.specularity = dd;
    NK_UP(du, unsigned char, ts, n); P.true_spec = du;
    NK_UP(di, int, so, n); P.spec_out = di;
    NK_UP(dd, double, rou, n); P.roulette = dd;
    ctx->has_rough = Fr > 0; P.has_rough = Fr > 0;
    // guide table for the diffuse pick (nk_event_advance): only valid when every row is non-decreasing -- the reference's
    // creation rates can go negative (Population.py:906-939), and on such a row the plain bisection's answer depends on its
    // probe sequence, so those set-ups keep the full bisection
    P.rou_guide = nullptr; P.rou_guide_k = 0;
    bool monotonic = Fr > 0 && P.M > 0;
    for (size_t f = 0; monotonic && f < (size_t)Fr; ++f)
        for (int m = 1; m < P.M; ++m)
            if (!(rou[f * P.M + m] >= rou[f * P.M + m - 1])) { monotonic = false; break; }
    if (monotonic) {
        int K = 16384;
        while (K > 256 && (size_t)Fr * (K + 1) * sizeof(int) > ((size_t)64 << 20)) K >>= 1;
        std::vector<int> guide((size_t)Fr * (K + 1));
        for (size_t f = 0; f < (size_t)Fr; ++f) {
            const double* row = rou + f * P.M;
            const double total = row[P.M - 1];
            for (int b = 0; b <= K; ++b) {
                const double thr = ((double)b / (double)K) * total;
                guide[f * (K + 1) + b] = (int)(std::lower_bound(row, row + P.M, thr) - row);
            }
        }
        NK_UP(di, int, guide.data(), guide.size());
        P.rou_guide = di; P.rou_guide_k = K;
    }
    return 0;
}

int nk_bind_particles(nk_ctx* ctx, int64_t cap, double* px, double* py, double* pz, double* tc, double* occ,
                      int32_t* mode, int32_t* omode, int32_t* cfacet, double* cx, double* cy, double* cz, int64_t* pid) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (cap < 2 || (cap & 1)) { ctx->err = "capacity must be even and >= 2"; return -1; }
    if (cap > 2147483646LL) { ctx->err = "capacity above 2^31-2 slots per GPU is not supported"; return -1; }
    if (((uintptr_t)px | (uintptr_t)py | (uintptr_t)pz | (uintptr_t)tc | (uintptr_t)occ) & 15) { ctx->err = "particle arrays must be 16-byte aligned"; return -1; }
    if (((uintptr_t)mode | (uintptr_t)omode) & 7) { ctx->err = "mode arrays must be 8-byte aligned"; return -1; }
    P.cap = cap; P.px = px; P.py = py; P.pz = pz; P.tc = tc; P.occ = occ; P.mode = mode; P.omode = omode; P.cfacet = cfacet;
    P.cx = cx; P.cy = cy; P.cz = cz; P.pid = (long long*)pid;
    if (P.M <= 0) { ctx->err = "nk_set_phonon before nk_bind_particles (the free-slot rings are per mode)"; return -1; }
    int* di;
    NK_UP(di, int, (const int*)nullptr, (size_t)cap); P.hitlist = di;
    // dense hit records for up to 1/16 of the slots per step (more hits fall back to the particle arrays)
    P.hitrec_cap = std::min<long long>(cap, std::max<long long>(65536, cap / 16));
    { NkHitRec* dr; NK_UP(dr, NkHitRec, (const NkHitRec*)nullptr, (size_t)P.hitrec_cap); P.hitrec = dr; }
    // free-slot rings (NkP::fr_*): ring 0 = global ring (freelist[cap, 2 cap)), ring 1 + m = region of mode m (freelist[0, cap))
    NK_UP(di, int, (const int*)nullptr, 2 * (size_t)cap); P.freelist = di;
    { long long* dl; NK_UP(dl, long long, (const long long*)nullptr, 3 * ((size_t)P.M + 1)); P.fr_ctr = dl; }
    NK_UP(ctx->mode_first_dev, int, (const int*)nullptr, (size_t)P.M + 1);
    NK_UP(ctx->sort_count, int, (const int*)nullptr, (size_t)P.M);
    NK_UP(ctx->sort_cursor, int, (const int*)nullptr, (size_t)P.M);
    NK_UP(ctx->sort_totals, long long, (const long long*)nullptr, 2);
    P.mode_first = ctx->mode_first_dev; P.n_rings = 1; P.fr_sorted = 0;      // until nk_sort_by_mode: every slot recycles through the global ring
    ctx->particles_bound = true;
    return 0;
}

static int nk_read_dyn(nk_ctx* ctx, NkDyn* d) {
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(d, ctx->P.dyn, sizeof(NkDyn), cudaMemcpyDeviceToHost));
    return 0;
}
static int nk_write_dyn(nk_ctx* ctx, const NkDyn* d) {
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(ctx->P.dyn, d, sizeof(NkDyn), cudaMemcpyHostToDevice));
    return 0;
}

int nk_set_slot_count(nk_ctx* ctx, int64_t n_slots) {
    cudaSetDevice(ctx->device);
    if (!ctx->particles_bound) { ctx->err = "nk_bind_particles first"; return -1; }
    if (n_slots < 0 || n_slots > ctx->P.cap) { ctx->err = "n_slots out of range"; return -1; }
    NkDyn d; if (nk_read_dyn(ctx, &d)) return -1;
    ctx->h_slots_hint = n_slots;
    d.n_slots = n_slots; d.fr_head = d.fr_tail = d.fr_snap = 0; d.n_hits = 0; d.n_emit = 0; d.n_new = 0; d.last_hits = 0; d.last_new = 0; d.blocks_done = 0;
    if (nk_write_dyn(ctx, &d)) return -1;
    // a new slot layout: all rings empty, no mode regions (nk_sort_by_mode publishes them again); the census below puts
    // every free slot (mode < 0) of [0, n_slots) on the global ring, so holes in the caller's arrays are recycled
    NK_CK(cudaMemsetAsync(ctx->P.fr_ctr, 0, 3 * ((size_t)ctx->P.M + 1) * sizeof(long long), ctx->stream));
    ctx->P.n_rings = 1; ctx->P.fr_sorted = 0;
    unsigned long long* dc = (unsigned long long*)ctx->sort_totals;       // persistent scratch: no allocation (and no implicit device sync) per call
    NK_CK(cudaMemsetAsync(dc, 0, 8, ctx->stream));
    k_census<<<ctx->n_sm * 4, 256, 0, ctx->stream>>>(ctx->P, dc);
    k_census_publish<<<1, 1, 0, ctx->stream>>>(ctx->P);
    NK_CK(cudaGetLastError());
    unsigned long long hc = 0;
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost));
    if (nk_read_dyn(ctx, &d)) return -1;
    d.n_alive = (long long)hc;
    return nk_write_dyn(ctx, &d);
}

// Maintenance pass between timesteps (nk_sort.cuh): live particles ordered by mode into the BACK buffers, free slots
// squeezed out; with pool_frac / pool_fixed > 0 every mode region gets spare slots and its own free-slot ring.  On return the
// back buffers are the bound particle arrays (the caller swaps its handles) and the old front buffers are scratch.
int nk_sort_by_mode(nk_ctx* ctx, double* px, double* py, double* pz, double* tc, double* occ, int32_t* mode, int32_t* omode,
                    int32_t* cfacet, double* cx, double* cy, double* cz, int64_t* pid, double pool_frac, int pool_fixed,
                    int64_t* n_slots_out, int64_t* n_alive_out) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (!ctx->particles_bound) { ctx->err = "nk_bind_particles first"; return -1; }
    if (((uintptr_t)px | (uintptr_t)py | (uintptr_t)pz | (uintptr_t)tc | (uintptr_t)occ) & 15) { ctx->err = "particle arrays must be 16-byte aligned"; return -1; }
    if (((uintptr_t)mode | (uintptr_t)omode) & 7) { ctx->err = "mode arrays must be 8-byte aligned"; return -1; }
    if (px == P.px || mode == P.mode) { ctx->err = "nk_sort_by_mode needs a second set of arrays (back buffers)"; return -1; }
    cudaStream_t st = ctx->stream;
    const int M = P.M;
    bool pools = (pool_frac > 0.0 || pool_fixed > 0) && !ctx->has_rough;      // rough walls change modes in place: regions would not hold
    for (int attempt = 0; attempt < 2; ++attempt) {
        NK_CK(cudaMemsetAsync(ctx->sort_count, 0, (size_t)M * sizeof(int), st));
        k_sort_hist<<<ctx->n_sm * 8, 256, 0, st>>>(P, ctx->sort_count);
        k_sort_scan<<<1, 1024, 0, st>>>(M, ctx->sort_count, ctx->mode_first_dev, ctx->sort_cursor, pools ? pool_frac : 0.0, pools ? pool_fixed : 0,
                                        ctx->sort_totals);
        NK_CK(cudaGetLastError());
        long long tot[2] = {0, 0};
        NK_CK(cudaMemcpyAsync(tot, ctx->sort_totals, sizeof(tot), cudaMemcpyDeviceToHost, st));
        NK_CK(cudaStreamSynchronize(st));
        if (tot[0] > P.cap) {
            if (pools) { pools = false; continue; }               // no room for the spare slots: plain compaction
            ctx->err = "nk_sort_by_mode: more live particles than capacity"; return -1;
        }
        const long long n_dst = tot[0], n_live = tot[1];
        int* perm = P.hitlist;                                    // cap ints, free between timesteps
        if (n_dst > 0) NK_CK(cudaMemsetAsync(perm, 0xFF, (size_t)n_dst * sizeof(int), st));
        k_sort_rank<<<ctx->n_sm * 8, 256, 0, st>>>(P, ctx->sort_cursor, perm);
        NkSoA src{P.px, P.py, P.pz, P.tc, P.occ, P.cx, P.cy, P.cz, P.mode, P.omode, P.cfacet, P.pid};
        NkSoA dst{px, py, pz, tc, occ, cx, cy, cz, mode, omode, cfacet, (long long*)pid};
        if (n_dst > 0) k_sort_permute<<<ctx->n_sm * 8, 256, 0, st>>>(src, dst, perm, n_dst);
        if (P.cap > n_dst) NK_CK(cudaMemsetAsync(mode + n_dst, 0xFF, (size_t)(P.cap - n_dst) * sizeof(int), st));    // free slots beyond
        k_sort_pools<<<ctx->n_sm, 256, 0, st>>>(M, ctx->sort_count, ctx->mode_first_dev, P.fr_ctr, P.freelist, pools ? 1 : 0);
        NK_CK(cudaGetLastError());
        P.px = px; P.py = py; P.pz = pz; P.tc = tc; P.occ = occ; P.mode = mode; P.omode = omode; P.cfacet = cfacet;
        P.cx = cx; P.cy = cy; P.cz = cz; P.pid = (long long*)pid;
        P.fr_sorted = pools ? n_dst : 0;
        P.n_rings = pools ? M + 1 : 1;
        NkDyn d; if (nk_read_dyn(ctx, &d)) return -1;
        d.n_slots = n_dst; d.n_alive = n_live; d.fr_head = d.fr_tail = d.fr_snap = 0;
        d.n_hits = 0; d.n_emit = 0; d.n_new = 0; d.last_hits = 0; d.last_new = 0; d.blocks_done = 0;
        if (nk_write_dyn(ctx, &d)) return -1;
        ctx->h_slots_hint = n_dst;
        if (n_slots_out) *n_slots_out = n_dst;
        if (n_alive_out) *n_alive_out = n_live;
        return 0;
    }
    return -1;
}

int nk_get_slot_count(nk_ctx* ctx, int64_t* n_slots, int64_t* n_alive) {
    cudaSetDevice(ctx->device);
    NkDyn d; if (nk_read_dyn(ctx, &d)) return -1;
    if (d.error) {
        ctx->err = std::string("device error bits: ") + ((d.error & NK_ERR_CAPACITY) ? "[particle capacity exhausted] " : "") +
                   ((d.error & NK_ERR_EVENTS) ? "[boundary event cap / broken periodic pair] " : "") +
                   ((d.error & NK_ERR_CMAX) ? "[more than 64 copies of one mode emitted in a step] " : "") +
                   ((d.error & NK_ERR_COMM) ? "[a peer rank did not deliver its sums within the time-out] " : "");
        return -2;
    }
    if (n_slots) *n_slots = d.n_slots;
    if (n_alive) *n_alive = d.n_alive;
    return 0;
}

// coefficients of the cubic RBF temperature field for the current T_sv (outside a step: set-up, restart)
__global__ void k_rbf_coef(NkP P) {
    nk_rbf_refresh(P, P.T_sv);
}

int nk_set_rbf(nk_ctx* ctx, int n_dims, const int32_t* dims, const double* shift, const double* scale, const double* weights) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (!P.svc) { ctx->err = "nk_set_subvols first"; return -1; }
    if (P.interp != NK_INTERP_RADIAL) { ctx->err = "nk_set_rbf needs temp_interp = NK_INTERP_RADIAL"; return -1; }
    if (n_dims < 1 || n_dims > 3) { ctx->err = "n_dims must be 1..3"; return -1; }
    for (int k = 0; k < n_dims; ++k) {
        if (dims[k] < 0 || dims[k] > 2) { ctx->err = "dims entries must be 0..2"; return -1; }
        if (!(scale[k] != 0.0)) { ctx->err = "scale must be non-zero"; return -1; }
        P.rbf_dim[k] = dims[k]; P.rbf_shift[k] = shift[k]; P.rbf_scale[k] = scale[k];
    }
    P.rbf_nd = n_dims;
    double* dd;
    NK_UP(dd, double, weights, (size_t)(P.S + n_dims + 1) * P.S); P.rbf_w = dd;
    k_rbf_coef<<<1, 128, 0, ctx->stream>>>(P);
    NK_CK(cudaGetLastError());
    return 0;
}

int nk_set_sv_temperature(nk_ctx* ctx, const double* T) {
    cudaSetDevice(ctx->device);
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(ctx->P.T_sv, T, ctx->P.S * sizeof(double), cudaMemcpyHostToDevice));
    ctx->tab_dirty = true;
    if (ctx->P.interp == NK_INTERP_RADIAL && ctx->P.rbf_w) {
        k_rbf_coef<<<1, 128, 0, ctx->stream>>>(ctx->P);
        NK_CK(cudaGetLastError());
    }
    return 0;
}
int nk_get_sv_temperature(nk_ctx* ctx, double* T) {
    cudaSetDevice(ctx->device);
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(T, ctx->P.T_sv, ctx->P.S * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
int nk_set_timestep(nk_ctx* ctx, int64_t k) {
    cudaSetDevice(ctx->device);
    NkDyn d; if (nk_read_dyn(ctx, &d)) return -1;
    d.step = k; d.relax_pending = 0;
    ctx->h_step = k; ctx->h_relax_pending = false;
    if (ctx->comm_block) NK_CK(cudaMemset(ctx->comm_block, 0, 256));
    return nk_write_dyn(ctx, &d);
}
int nk_get_timestep(nk_ctx* ctx, int64_t* k) {
    cudaSetDevice(ctx->device);
    NkDyn d; if (nk_read_dyn(ctx, &d)) return -1;
    *k = d.step; return 0;
}

// ---- operator seams ----------------------------------------------------------------------------------------
static inline int nk_grid(long long n, int threads, int cap_blocks) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (cap_blocks > 0 && b > cap_blocks) b = cap_blocks;
    return (int)b;
}

int nk_find_boundary(nk_ctx* ctx, int64_t n, const double* x, const double* v, double* xc, double* tc, int32_t* fc) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return 0;
    k_find_boundary<<<nk_grid(n, NK_RAY_THREADS, ctx->n_sm * 8), NK_RAY_THREADS, NK_TILE_SMEM_BYTES, ctx->stream>>>(ctx->P, n, x, v, xc, tc, fc);
    NK_CK(cudaGetLastError());
    return 0;
}
int nk_contains(nk_ctx* ctx, int64_t n, const double* x, uint8_t* inside) {
    cudaSetDevice(ctx->device);
    if (!ctx->P.faces) { ctx->err = "nk_set_mesh first"; return -1; }
    if (n <= 0) return 0;
    k_contains<<<nk_grid(n, NK_RAY_THREADS, ctx->n_sm * 4), NK_RAY_THREADS, NK_TILE_SMEM_BYTES, ctx->stream>>>(ctx->P, n, x, inside);
    NK_CK(cudaGetLastError());
    return 0;
}
int nk_classify(nk_ctx* ctx, int64_t n, const double* x, int32_t* sv, int64_t* counts) {
    cudaSetDevice(ctx->device);
    if (counts) NK_CK(cudaMemsetAsync(counts, 0, ctx->P.S * sizeof(int64_t), ctx->stream));
    if (n <= 0) return 0;
    size_t smem = nk_sv_smem_doubles(ctx->P.S) * 8 + ctx->P.S * 4;
    k_classify<<<nk_grid(n, 256, ctx->n_sm * 8), 256, smem, ctx->stream>>>(ctx->P, n, x, sv, (unsigned long long*)counts);
    NK_CK(cudaGetLastError());
    return 0;
}
int nk_occupation(nk_ctx* ctx, int64_t n, const double* T, const double* omega, double* occ) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return 0;
    k_occupation<<<nk_grid(n, 256, ctx->n_sm * 8), 256, 0, ctx->stream>>>(ctx->P, n, T, omega, occ);
    NK_CK(cudaGetLastError());
    return 0;
}
int nk_lifetime(nk_ctx* ctx, int64_t n, const double* T, const int32_t* mode, double* tau) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return 0;
    k_lifetime<<<nk_grid(n, 256, ctx->n_sm * 8), 256, 0, ctx->stream>>>(ctx->P, n, T, mode, tau);
    NK_CK(cudaGetLastError());
    return 0;
}
int nk_temperature_of_energy(nk_ctx* ctx, int64_t n, const double* E, double* T) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return 0;
    k_table<<<nk_grid(n, 256, ctx->n_sm * 8), 256, 0, ctx->stream>>>(ctx->P, n, E, T, 1);
    NK_CK(cudaGetLastError());
    return 0;
}
int nk_energy_of_temperature(nk_ctx* ctx, int64_t n, const double* T, double* E) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return 0;
    k_table<<<nk_grid(n, 256, ctx->n_sm * 8), 256, 0, ctx->stream>>>(ctx->P, n, T, E, 0);
    NK_CK(cudaGetLastError());
    return 0;
}
int nk_particle_temperature(nk_ctx* ctx, int64_t n, const double* x, double* T) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return 0;
    k_particle_T<<<nk_grid(n, 256, ctx->n_sm * 8), 256, 0, ctx->stream>>>(ctx->P, n, x, T);
    NK_CK(cudaGetLastError());
    return 0;
}

// ---- the timestep ---------------------------------------------------------------------------------------------
static int nk_check_ready(nk_ctx* ctx) {
    const NkP& P = ctx->P;
    if (!ctx->particles_bound) { ctx->err = "particles not bound"; return -1; }
    if (!P.faces || !P.svc || !P.mprop || !P.mhot || !P.acc) { ctx->err = "tables missing: call nk_set_mesh, nk_set_subvols, nk_set_phonon, nk_set_population, nk_set_reservoirs first"; return -1; }
    if (P.dt <= 0) { ctx->err = "nk_set_population not called"; return -1; }
    if (P.interp == NK_INTERP_RADIAL && !P.rbf_w) { ctx->err = "temp_interp radial: nk_set_rbf not called"; return -1; }
    return 0;
}

int nk_init_collisions(nk_ctx* ctx) {
    cudaSetDevice(ctx->device);
    if (nk_check_ready(ctx)) return -1;
    k_init_collisions<<<ctx->n_sm * 4, NK_RAY_THREADS, NK_TILE_SMEM_BYTES, ctx->stream>>>(ctx->P);
    NK_CK(cudaGetLastError());
    return 0;
}

#define NK_TAB_MAX_ENTRIES (6LL << 20)      // 96 MB of {n0, decay} pairs

// per-(mode, subvolume) tables {n0, decay} of the table variant for the step that comes next (T_sv changed)
static int nk_refresh_tables(nk_ctx* ctx, int variant) {
    const NkP& P = ctx->P;
    if (variant == 4) {
        k_mode_tables<<<ctx->n_sm * 8, 256, nk_hot_smem_bytes(P.S), ctx->stream>>>(P);
        NK_CK(cudaGetLastError());
        ctx->tab_dirty = false;
    }
    return 0;
}
static size_t nk_step_smem(const NkP& P) { return (nk_sv_smem_doubles(P.S) + 8 * (size_t)P.S) * 8 + ((size_t)P.S + 2) * 4 + nk_hot_smem_bytes(P.S) + 32; }

// chunked launch of the streaming kernel for the host-buffer pipeline: chunk c waits for its upload event and
// signals its own completion event
struct NkChunkPlan { int n; long long chunk, total; cudaEvent_t* ev_in; cudaEvent_t* ev_k; };

// kernels of one step; fuse_finalize: the last block of k_rare closes the step (no collective in between)
// phase 0: whole step; 1: streaming part only (chunks of the host pipeline); 2: rare path + closing of a step started with 1
static int nk_step_kernels(nk_ctx* ctx, bool fuse_finalize, const NkChunkPlan* plan = nullptr, int phase = 0) {
    cudaSetDevice(ctx->device);
    if (nk_check_ready(ctx)) return -1;
    const NkP& P = ctx->P;
    size_t smem = nk_step_smem(P);
    const bool fast = P.is_slice && P.interp == NK_INTERP_NEAREST;
    const int kind = fast ? NK_KIND_FAST : ((P.is_slice && P.interp == NK_INTERP_LINEAR) ? NK_KIND_SLICE_LINEAR : NK_KIND_GENERAL);
    // the step counter and the relaxation flag are mirrored on the host (every mutation goes through this
    // library), so the launch-uniform RELAX / FLUX variants can be chosen without a device read-back
    const bool relax = ctx->h_relax_pending;
    const bool flux = ((ctx->h_step + 1) % P.n_dt_to_conv) == 0;
    int variant = 0;                                          // 0: direct arithmetic, 4: per-(mode, subvolume) tables
    if (phase == 2) variant = ctx->last_variant;
    else {
    // per-(mode, subvolume) tables pay off once there are a few particles per table entry
    // ... and while the table (16 B per entry, rebuilt every step) stays L2-sized: at S = 100 x 1.8e5 modes (286 MB) the
    // rebuild costs what the leaner inner loop saves (profiles/README.md)
    if (ctx->use_tab && fast && P.hot_tab &&
        (ctx->force_tab || (ctx->h_slots_hint >= 2 * (long long)P.M * P.S && (long long)P.M * P.S <= NK_TAB_MAX_ENTRIES))) variant = 4;
    if (ctx->tab_dirty && nk_refresh_tables(ctx, variant)) return -1;
    // experiment (NK_L2_PERSIST=1): keep the (mode, subvolume) table in the persisting part of the L2 while 8 GB of particle
    // state stream through it
    if (variant == 4 && ctx->l2_persist && !ctx->l2_window_set) {
        cudaDeviceProp prop; cudaGetDeviceProperties(&prop, ctx->device);
        const size_t bytes = (size_t)P.M * P.S * sizeof(double2);
        const size_t win = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(win, (size_t)prop.persistingL2CacheMaxSize));
        cudaStreamAttrValue attr; memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = P.hot_tab;
        attr.accessPolicyWindow.num_bytes = win;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
        ctx->l2_window_set = true;
    }
    nk_step_fn kern = nk_pick_step(variant, ctx->has_rough, kind, relax, flux);
    if (!ctx->step_blocks || ctx->step_blocks_variant != variant) {
        int per_sm = 0;
        for (int r = 0; r < 2; ++r) for (int f = 0; f < 2; ++f) {       // every variant may need the opt-in shared memory size
            nk_step_fn k = nk_pick_step(variant, ctx->has_rough, kind, r, f);
            if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nk_pick_step(variant, ctx->has_rough, kind, true, true), NK_STEP_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        ctx->step_blocks = per_sm * ctx->n_sm;
        ctx->step_blocks_variant = variant;
    }
    nk_prof_mark(ctx);
    if (!plan) {
        kern<<<ctx->step_blocks, NK_STEP_THREADS, smem, ctx->stream>>>(P);
    } else {
        for (int c = 0; c < plan->n; ++c) {
            NkP Pc = P;
            Pc.slot_lo = (long long)c * plan->chunk;
            Pc.slot_hi = std::min(Pc.slot_lo + plan->chunk, plan->total);
            Pc.scan_emit = c == 0;
            NK_CK(cudaStreamWaitEvent(ctx->stream, plan->ev_in[c], 0));
            kern<<<ctx->step_blocks, NK_STEP_THREADS, smem, ctx->stream>>>(Pc);
            NK_CK(cudaEventRecord(plan->ev_k[c], ctx->stream));
        }
    }
    NK_CK(cudaGetLastError());
    nk_prof_mark(ctx);
    ctx->last_variant = variant;
    }
    if (phase == 1) return 0;
    const size_t fin_smem = nk_rare_fin_doubles(P.S, P.R) * 8;
    const bool tiled = P.F > NK_RARE_FACES || ctx->force_tiled;
    ctx->last_rare_tiled = tiled;
    const size_t rare_smem = tiled ? fin_smem + NK_TILE_SMEM_BYTES : nk_rare_smem_bytes(P.S, P.R, P.F, P.nf);
    if (!ctx->rare_attr_set) {
        const int want = (int)std::max(fin_smem + NK_TILE_SMEM_BYTES, nk_rare_smem_bytes(P.S, P.R, P.F, P.nf));
        if (want > 48 * 1024) {
            cudaFuncSetAttribute(k_rare<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            cudaFuncSetAttribute(k_rare<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            cudaFuncSetAttribute(k_rare_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            cudaFuncSetAttribute(k_rare_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
        }
        ctx->rare_attr_set = true;
    }
    // one item per thread; the number of items is known on the device only, so size the grid for ~5 % of the slots
    // (hits + emission are 0.2-2 % of the particles per step; more items are covered by the grid-stride loop)
    const long long want_blocks = (ctx->h_slots_hint / 20 + NK_RARE_THREADS - 1) / NK_RARE_THREADS;
    const int rare_blocks = (int)std::min<long long>((long long)ctx->n_sm * ctx->rare_blocks_per_sm, std::max<long long>(ctx->n_sm, want_blocks));
    const int tiled_blocks = (int)std::min<long long>((long long)ctx->n_sm * 16, std::max<long long>(ctx->n_sm, (ctx->h_slots_hint / 20 + NK_RARE_TILED_THREADS - 1) / NK_RARE_TILED_THREADS));
    if (tiled) {
        if (fuse_finalize) k_rare_tiled<true><<<tiled_blocks, NK_RARE_TILED_THREADS, rare_smem, ctx->stream>>>(P);
        else k_rare_tiled<false><<<tiled_blocks, NK_RARE_TILED_THREADS, rare_smem, ctx->stream>>>(P);
    } else {
        if (fuse_finalize) k_rare<true><<<rare_blocks, NK_RARE_THREADS, rare_smem, ctx->stream>>>(P);
        else k_rare<false><<<rare_blocks, NK_RARE_THREADS, rare_smem, ctx->stream>>>(P);
    }
    NK_CK(cudaGetLastError());
    nk_prof_mark(ctx);
    if (fuse_finalize) {
        ctx->h_step += 1; ctx->h_relax_pending = true; ctx->tab_dirty = true;
        if (nk_refresh_tables(ctx, variant)) return -1;      // next step's tables right away (T_sv is final once k_rare has finished)
        nk_prof_mark(ctx);
    }
    ctx->last_variant = variant;
    return 0;
}

int nk_step_local(nk_ctx* ctx) { return nk_step_kernels(ctx, false); }

int nk_last_step_variant(nk_ctx* ctx) { return ctx->last_variant + (ctx->last_rare_tiled ? 8 : 0); }

int nk_step_finalize(nk_ctx* ctx) {
    cudaSetDevice(ctx->device);
    const NkP& P = ctx->P;
    int threads = 32;
    while (threads < P.S && threads < 1024) threads <<= 1;
    k_finalize<<<1, threads, 3 * (size_t)P.S * 8, ctx->stream>>>(P);
    NK_CK(cudaGetLastError());
    ctx->h_step += 1; ctx->h_relax_pending = true; ctx->tab_dirty = true;
    if (nk_refresh_tables(ctx, ctx->last_variant)) return -1;
    nk_prof_mark(ctx);
    return 0;
}

int nk_profile_begin(nk_ctx* ctx) {
    cudaSetDevice(ctx->device);
    ctx->profiling = true; ctx->ev_used = 0;
    return 0;
}
int nk_profile_end(nk_ctx* ctx, double* ms, int64_t* n_steps) {
    cudaSetDevice(ctx->device);
    ctx->profiling = false;
    NK_CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 4; ++k) ms[k] = 0.0;
    size_t steps = ctx->ev_used / 4;                 // marks: | k_step | k_rare | finalize (0 when fused) |
    for (size_t s = 0; s < steps; ++s)
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
            NK_CK(cudaEventElapsedTime(&t, ctx->ev[4 * s + k], ctx->ev[4 * s + k + 1]));
            ms[k] += t;
        }
    if (n_steps) *n_steps = (int64_t)steps;
    ctx->ev_used = 0;
    return 0;
}

int nk_step(nk_ctx* ctx, int n_steps) {
    const bool fused = ctx->P.world == 1 || ctx->P.comm_on;      // otherwise the caller must all-reduce between the halves
    if (!fused) { ctx->err = "nk_step with world > 1 needs nk_comm_enable; or use nk_step_local / all-reduce / nk_step_finalize"; return -1; }
    for (int k = 0; k < n_steps; ++k)
        if (nk_step_kernels(ctx, true)) return -1;
    return 0;
}

int nk_flush_relaxation(nk_ctx* ctx) {
    cudaSetDevice(ctx->device);
    if (nk_check_ready(ctx)) return -1;
    size_t smem = nk_sv_smem_doubles(ctx->P.S) * 8 + nk_hot_smem_bytes(ctx->P.S);
    if (ctx->P.is_slice && ctx->P.interp == NK_INTERP_NEAREST) k_flush_relax<NK_KIND_FAST><<<ctx->n_sm * 8, 256, smem, ctx->stream>>>(ctx->P);
    else if (ctx->P.is_slice && ctx->P.interp == NK_INTERP_LINEAR) k_flush_relax<NK_KIND_SLICE_LINEAR><<<ctx->n_sm * 8, 256, smem, ctx->stream>>>(ctx->P);
    else k_flush_relax<NK_KIND_GENERAL><<<ctx->n_sm * 8, 256, smem, ctx->stream>>>(ctx->P);
    NK_CK(cudaGetLastError());
    k_clear_relax<<<1, 1, 0, ctx->stream>>>(ctx->P);
    NK_CK(cudaGetLastError());
    ctx->h_relax_pending = false;
    return 0;
}

static void nk_unpack_results(const NkP& P, const double* h, double* T_sv, double* E_sv, int64_t* N_sv, double* flux, double* kappa_sv, double* kappa,
                              double* res_E_bal, double* res_flux, int64_t* N_leaving, double* total_energy) {
    const int S = P.S, R = P.R;
    if (T_sv) memcpy(T_sv, &h[NK_OUT_T(S, R)], S * 8);
    if (E_sv) memcpy(E_sv, &h[NK_OUT_E(S, R)], S * 8);
    if (N_sv) for (int s = 0; s < S; ++s) N_sv[s] = (int64_t)h[NK_OUT_N(S, R) + s];
    if (flux) memcpy(flux, &h[NK_OUT_FLUX(S, R)], 3 * S * 8);
    if (kappa_sv) memcpy(kappa_sv, &h[NK_OUT_KSV(S, R)], S * 8);
    if (kappa) *kappa = h[NK_OUT_KAPPA(S, R)];
    if (res_E_bal) memcpy(res_E_bal, &h[NK_OUT_REBAL(S, R)], R * 8);
    if (res_flux) memcpy(res_flux, &h[NK_OUT_RFLUX(S, R)], 3 * R * 8);
    if (N_leaving) for (int r = 0; r < R; ++r) N_leaving[r] = (int64_t)h[NK_OUT_NLEAVE(S, R) + r];
    if (total_energy) *total_energy = h[NK_OUT_ETOT(S, R)];
}

int nk_get_results(nk_ctx* ctx, double* T_sv, double* E_sv, int64_t* N_sv, double* flux, double* kappa_sv, double* kappa,
                   double* res_E_bal, double* res_flux, int64_t* N_leaving, double* total_energy) {
    cudaSetDevice(ctx->device);
    const NkP& P = ctx->P;
    std::vector<double> h(nk_out_len(P.S, P.R));
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(h.data(), P.out, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    nk_unpack_results(P, h.data(), T_sv, E_sv, N_sv, flux, kappa_sv, kappa, res_E_bal, res_flux, N_leaving, total_energy);
    return 0;
}

// Asynchronous variant for callers that keep the device busy: nk_snapshot_results enqueues a copy of the results block (as it
// is after the steps enqueued so far) into one of four pinned host buffers and returns its ticket; more steps may be enqueued
// right away.  nk_get_snapshot waits for that copy only and unpacks it like nk_get_results.
int nk_snapshot_results(nk_ctx* ctx, int* ticket) {
    cudaSetDevice(ctx->device);
    const NkP& P = ctx->P;
    if (!P.out) { ctx->err = "results block not allocated yet"; return -1; }
    const int len = nk_out_len(P.S, P.R);
    if (ctx->snap_len != len) {
        for (int k = 0; k < 4; ++k) {
            if (ctx->snap_host[k]) cudaFreeHost(ctx->snap_host[k]);
            NK_CK(cudaMallocHost(&ctx->snap_host[k], (size_t)len * sizeof(double)));
            if (!ctx->snap_ev[k]) NK_CK(cudaEventCreateWithFlags(&ctx->snap_ev[k], cudaEventDisableTiming));
        }
        ctx->snap_len = len;
    }
    const int k = (int)(ctx->snap_next++ & 3u);
    NK_CK(cudaMemcpyAsync(ctx->snap_host[k], P.out, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NK_CK(cudaEventRecord(ctx->snap_ev[k], ctx->stream));
    *ticket = k;
    return 0;
}
int nk_get_snapshot(nk_ctx* ctx, int ticket, double* T_sv, double* E_sv, int64_t* N_sv, double* flux, double* kappa_sv, double* kappa,
                    double* res_E_bal, double* res_flux, int64_t* N_leaving, double* total_energy) {
    cudaSetDevice(ctx->device);
    if (ticket < 0 || ticket > 3 || !ctx->snap_host[ticket]) { ctx->err = "no such results snapshot"; return -1; }
    NK_CK(cudaEventSynchronize(ctx->snap_ev[ticket]));
    nk_unpack_results(ctx->P, ctx->snap_host[ticket], T_sv, E_sv, N_sv, flux, kappa_sv, kappa, res_E_bal, res_flux, N_leaving, total_energy);
    return 0;
}

// Plain version: upload everything, step, download everything.
static int nk_advance_host_simple(nk_ctx* ctx, int64_t n_in, int n_steps, double* px, double* py, double* pz, double* tc, double* occ,
                                  int32_t* mode, int32_t* omode, int32_t* cfacet, double* cx, double* cy, double* cz, int64_t* pid,
                                  int64_t* n_out) {
    NkP& P = ctx->P;
    cudaStream_t st = ctx->stream;
    size_t n = (size_t)n_in;
    NK_CK(cudaMemcpyAsync(P.px, px, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.py, py, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.pz, pz, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.tc, tc, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.occ, occ, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.mode, mode, n * 4, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.omode, omode, n * 4, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.cfacet, cfacet, n * 4, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.cx, cx, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.cy, cy, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.cz, cz, n * 8, cudaMemcpyHostToDevice, st));
    NK_CK(cudaMemcpyAsync(P.pid, pid, n * 8, cudaMemcpyHostToDevice, st));
    if (nk_set_slot_count(ctx, n_in)) return -1;
    if (nk_step(ctx, n_steps)) return -1;
    if (nk_flush_relaxation(ctx)) return -1;
    int64_t ns = 0, na = 0;
    if (nk_get_slot_count(ctx, &ns, &na)) return -1;
    size_t m = (size_t)ns;
    NK_CK(cudaMemcpyAsync(px, P.px, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(py, P.py, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(pz, P.pz, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(tc, P.tc, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(occ, P.occ, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(mode, P.mode, m * 4, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(omode, P.omode, m * 4, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(cfacet, P.cfacet, m * 4, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(cx, P.cx, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(cy, P.cy, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(cz, P.cz, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaMemcpyAsync(pid, P.pid, m * 8, cudaMemcpyDeviceToHost, st));
    NK_CK(cudaStreamSynchronize(st));
    ctx->xfer_h2d = (long long)n * 84; ctx->xfer_d2h = (long long)m * 84;
    if (n_out) *n_out = ns;
    return 0;
}

// Pipelined single-step version: the slots are cut into chunks; chunk c+1 is uploaded (stream s_in) while chunk c
// runs through the streaming kernel (ctx stream) and the positions / collision clocks of chunk c-1 are downloaded
// (stream s_out), so the PCIe link works in both directions at once.  The arrays k_step never writes (modes,
// collision data, ids) are not downloaded again: the few slots the rare path touched come back as a packed patch
// that is applied to the host arrays.  Occupations are downloaded after the deferred relaxation has been flushed.
#define NK_PIPE_CHUNKS 16
static int nk_advance_host_pipelined(nk_ctx* ctx, int64_t n_in, double* px, double* py, double* pz, double* tc, double* occ,
                                     int32_t* mode, int32_t* omode, int32_t* cfacet, double* cx, double* cy, double* cz, int64_t* pid,
                                     int64_t* n_out) {
    NkP& P = ctx->P;
    cudaStream_t st = ctx->stream;
    if (!ctx->s_in) {
        NK_CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        NK_CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        ctx->ev_in.resize(NK_PIPE_CHUNKS); ctx->ev_k.resize(NK_PIPE_CHUNKS);
        for (int c = 0; c < NK_PIPE_CHUNKS; ++c) {
            NK_CK(cudaEventCreateWithFlags(&ctx->ev_in[c], cudaEventDisableTiming));
            NK_CK(cudaEventCreateWithFlags(&ctx->ev_k[c], cudaEventDisableTiming));
        }
        NK_CK(cudaMalloc(&ctx->patch_count_dev, 2 * sizeof(unsigned int)));
    }
    long long want = std::max<long long>(1 << 16, P.cap / 64);
    if (const char* e = getenv("NK_PIPE_PATCH_CAP")) want = std::max<long long>(1, atoll(e));     // tests: force the fallback
    if (ctx->patch_cap != want) {
        if (ctx->patch_dev) cudaFree(ctx->patch_dev);
        if (ctx->patch_host) cudaFreeHost(ctx->patch_host);
        NK_CK(cudaMalloc(&ctx->patch_dev, (size_t)want * 80));
        NK_CK(cudaMallocHost(&ctx->patch_host, (size_t)want * 80));
        ctx->patch_cap = want;
    }
    const long long cap = ctx->patch_cap;
    auto carve = [&](void* base) {
        NkPatch q; char* b = (char*)base;
        q.x = (double*)b; q.y = q.x + cap; q.z = q.y + cap; q.tc = q.z + cap; q.cx = q.tc + cap; q.cy = q.cx + cap; q.cz = q.cy + cap;
        q.pid = (long long*)(q.cz + cap); q.slot = (int*)(q.pid + cap); q.mode = q.slot + cap; q.omode = q.mode + cap; q.cfacet = q.omode + cap;
        return q;
    };
    const NkPatch pd = carve(ctx->patch_dev), ph = carve(ctx->patch_host);

    const size_t n = (size_t)n_in;
    NK_CK(cudaStreamSynchronize(st));
    static const bool trace = getenv("NK_PIPE_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (trace) fprintf(stderr, "[nk pipe] %-10s %8.2f ms\n", what,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    // 1. the slot census needs the modes
    NK_CK(cudaMemcpyAsync(P.mode, mode, n * 4, cudaMemcpyHostToDevice, ctx->s_in));
    NK_CK(cudaStreamSynchronize(ctx->s_in));
    if (nk_set_slot_count(ctx, n_in)) return -1;
    lap("census");
    // 2. uploads, chunk by chunk: only what the streaming kernel reads (48 of the 84 bytes per particle)
    long long chunk = ((long long)n_in + NK_PIPE_CHUNKS - 1) / NK_PIPE_CHUNKS;
    chunk = (chunk + 511) / 512 * 512;
    const int nc = (int)(((long long)n_in + chunk - 1) / chunk);
    const bool sparse = ctx->sparse_cold;
    // without rough facets nothing separates the omega-carrying mode from the mode: the kernels do not read `omode` then,
    // except the rare path for the particles it handles (it travels with the cold record)
    const bool dense_omode = ctx->has_rough || !sparse;
    for (int c = 0; c < nc; ++c) {
        const size_t lo = (size_t)c * chunk, len = std::min<size_t>(chunk, n - lo);
        NK_CK(cudaMemcpyAsync(P.px + lo, px + lo, len * 8, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.py + lo, py + lo, len * 8, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.pz + lo, pz + lo, len * 8, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.tc + lo, tc + lo, len * 8, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.occ + lo, occ + lo, len * 8, cudaMemcpyHostToDevice, ctx->s_in));
        if (dense_omode) NK_CK(cudaMemcpyAsync(P.omode + lo, omode + lo, len * 4, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
    }
    // 3. the streaming kernel per chunk on the ctx stream, as soon as the chunk has arrived
    NkChunkPlan plan{nc, chunk, (long long)n_in, ctx->ev_in.data(), ctx->ev_k.data()};
    if (nk_step_kernels(ctx, true, &plan, 1)) return -1;
    // 4. positions and clocks of every chunk go back as soon as its kernel is done (stale for the few slots the
    //    rare path rewrites afterwards: the patch below overrides them)
    for (int c = 0; c < nc; ++c) {
        const size_t lo = (size_t)c * chunk, len = std::min<size_t>(chunk, n - lo);
        NK_CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[c], 0));
        NK_CK(cudaMemcpyAsync(px + lo, P.px + lo, len * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        NK_CK(cudaMemcpyAsync(py + lo, P.py + lo, len * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        NK_CK(cudaMemcpyAsync(pz + lo, P.pz + lo, len * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        NK_CK(cudaMemcpyAsync(tc + lo, P.tc + lo, len * 8, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    lap("enq hot");
    // 4b. cold fields (collision facet / position, id): only the rare path reads them, and only for particles whose
    //     collision falls inside this step (tc - 1 < 0 <=> tc < 1, exact in IEEE).  While the DMA engines are busy with the
    //     chunks, host threads scan `tc` for those particles and pack their cold fields; NaN clocks are included (harmless).
    long long n_cold = -1;
    if (sparse) {
        const long long want_cold = std::max<long long>(1 << 16, P.cap / 32);
        if (ctx->cold_cap != want_cold) {
            if (ctx->cold_dev) cudaFree(ctx->cold_dev);
            if (ctx->cold_host) cudaFreeHost(ctx->cold_host);
            ctx->cold_dev = ctx->cold_host = nullptr; ctx->cold_cap = 0;
            NK_CK(cudaMalloc(&ctx->cold_dev, (size_t)want_cold * 44));
            NK_CK(cudaMallocHost(&ctx->cold_host, (size_t)want_cold * 44));
            ctx->cold_cap = want_cold;
        }
        if (!ctx->ev_cold) NK_CK(cudaEventCreateWithFlags(&ctx->ev_cold, cudaEventDisableTiming));
        const long long ccap = ctx->cold_cap;
        auto carve_cold = [&](void* base) {
            NkCold q; char* b = (char*)base;
            q.cx = (double*)b; q.cy = q.cx + ccap; q.cz = q.cy + ccap; q.pid = (long long*)(q.cz + ccap);
            q.slot = (int*)(q.pid + ccap); q.cfacet = q.slot + ccap; q.omode = q.cfacet + ccap;
            return q;
        };
        const NkCold cd = carve_cold(ctx->cold_dev), ch = carve_cold(ctx->cold_host);
        const unsigned hw = std::thread::hardware_concurrency();
        const unsigned share = (hw ? hw : 1u) / (unsigned)std::max(1, P.world);          // the ranks of a box share its cores
        const int nt = (int)std::max(2u, std::min(16u, share));
        std::vector<std::vector<int>> found(nt);
        {
            std::vector<std::thread> th;
            for (int k = 0; k < nt; ++k)
                th.emplace_back([&, k]() {
                    const long long i0 = (long long)n_in * k / nt, i1 = (long long)n_in * (k + 1) / nt;
                    std::vector<int>& v = found[k];
                    for (long long i = i0; i < i1; ++i)
                        if (!(tc[i] >= 1.0) && mode[i] >= 0) v.push_back((int)i);
                });
            for (auto& t : th) t.join();
        }
        std::vector<long long> off(nt + 1, 0);
        for (int k = 0; k < nt; ++k) off[k + 1] = off[k] + (long long)found[k].size();
        if (off[nt] <= ccap) {
            n_cold = off[nt];
            std::vector<std::thread> th;
            for (int k = 0; k < nt; ++k)
                th.emplace_back([&, k]() {
                    long long o = off[k];
                    for (int i : found[k]) {
                        ch.slot[o] = i; ch.cfacet[o] = cfacet[i]; ch.omode[o] = omode[i]; ch.pid[o] = pid[i];
                        ch.cx[o] = cx[i]; ch.cy[o] = cy[i]; ch.cz[o] = cz[i];
                        ++o;
                    }
                });
            for (auto& t : th) t.join();
            const size_t k = (size_t)n_cold;
            if (k) {
                NK_CK(cudaMemcpyAsync(cd.cx, ch.cx, k * 8, cudaMemcpyHostToDevice, ctx->s_in));
                NK_CK(cudaMemcpyAsync(cd.cy, ch.cy, k * 8, cudaMemcpyHostToDevice, ctx->s_in));
                NK_CK(cudaMemcpyAsync(cd.cz, ch.cz, k * 8, cudaMemcpyHostToDevice, ctx->s_in));
                NK_CK(cudaMemcpyAsync(cd.pid, ch.pid, k * 8, cudaMemcpyHostToDevice, ctx->s_in));
                NK_CK(cudaMemcpyAsync(cd.slot, ch.slot, k * 4, cudaMemcpyHostToDevice, ctx->s_in));
                NK_CK(cudaMemcpyAsync(cd.cfacet, ch.cfacet, k * 4, cudaMemcpyHostToDevice, ctx->s_in));
                NK_CK(cudaMemcpyAsync(cd.omode, ch.omode, k * 4, cudaMemcpyHostToDevice, ctx->s_in));
            }
            NK_CK(cudaEventRecord(ctx->ev_cold, ctx->s_in));
            NK_CK(cudaStreamWaitEvent(st, ctx->ev_cold, 0));
            if (k) {
                k_unpack_cold<<<ctx->n_sm * 2, 256, 0, st>>>(P, cd, n_cold);
                NK_CK(cudaGetLastError());
            }
        }
        lap("cold scan");
    }
    if (n_cold < 0) {             // dense upload of the cold arrays (NK_HOST_SPARSE=0, or more candidates than the staging holds)
        NK_CK(cudaMemcpyAsync(P.cfacet, cfacet, n * 4, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.cx, cx, n * 8, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.cy, cy, n * 8, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.cz, cz, n * 8, cudaMemcpyHostToDevice, ctx->s_in));
        NK_CK(cudaMemcpyAsync(P.pid, pid, n * 8, cudaMemcpyHostToDevice, ctx->s_in));
        if (!dense_omode) NK_CK(cudaMemcpyAsync(P.omode, omode, n * 4, cudaMemcpyHostToDevice, ctx->s_in));
        if (!ctx->ev_cold) NK_CK(cudaEventCreateWithFlags(&ctx->ev_cold, cudaEventDisableTiming));
        NK_CK(cudaEventRecord(ctx->ev_cold, ctx->s_in));
        NK_CK(cudaStreamWaitEvent(st, ctx->ev_cold, 0));
    }
    // 4c. rare path + closing of the step
    if (nk_step_kernels(ctx, true, nullptr, 2)) return -1;
    // 5. tail: flush the deferred relaxation, pack the dirty slots
    if (nk_flush_relaxation(ctx)) return -1;
    k_pack_dirty<<<ctx->n_sm * 4, 256, 0, st>>>(P, pd, cap, 0, ctx->patch_count_dev);
    NK_CK(cudaGetLastError());
    unsigned int cnt[2] = {0, 0};
    NK_CK(cudaMemcpyAsync(cnt, ctx->patch_count_dev, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    lap("enqueued");
    NK_CK(cudaStreamSynchronize(st));
    lap("kernels");
    int64_t ns = 0, na = 0;
    if (nk_get_slot_count(ctx, &ns, &na)) return -1;
    const size_t m = (size_t)ns;
    const long long nd = (long long)cnt[0] + cnt[1];
    const bool patch_ok = (long long)cnt[1] <= P.newslots_cap;
    if (!patch_ok && n_cold >= 0) {
        // the new-slot list overflowed AND the device holds the cold arrays only for the touched slots: neither a patch
        // nor a full download can rebuild the host arrays
        ctx->err = "nk_advance_host: more emitted particles than the new-slot list holds; rerun with NK_HOST_SPARSE=0 or NK_HOST_PIPELINE=0";
        return -1;
    }
    auto fetch_patch = [&](size_t k, cudaStream_t q) -> int {
        double* const dsrc[7] = {pd.x, pd.y, pd.z, pd.tc, pd.cx, pd.cy, pd.cz};
        double* const ddst[7] = {ph.x, ph.y, ph.z, ph.tc, ph.cx, ph.cy, ph.cz};
        for (int a = 0; a < 7; ++a) NK_CK(cudaMemcpyAsync(ddst[a], dsrc[a], k * 8, cudaMemcpyDeviceToHost, q));
        NK_CK(cudaMemcpyAsync(ph.pid, pd.pid, k * 8, cudaMemcpyDeviceToHost, q));
        int* const isrc[4] = {pd.slot, pd.mode, pd.omode, pd.cfacet};
        int* const idst[4] = {ph.slot, ph.mode, ph.omode, ph.cfacet};
        for (int a = 0; a < 4; ++a) NK_CK(cudaMemcpyAsync(idst[a], isrc[a], k * 4, cudaMemcpyDeviceToHost, q));
        return 0;
    };
    // scattered writes, latency-bound on one core: a few host threads (a slot listed twice carries the same values)
    auto apply_patch = [&](long long cntp) {
        auto apply = [&](long long i0, long long i1) {
            for (long long i = i0; i < i1; ++i) {
                const int s = ph.slot[i];
                px[s] = ph.x[i]; py[s] = ph.y[i]; pz[s] = ph.z[i]; tc[s] = ph.tc[i];
                cx[s] = ph.cx[i]; cy[s] = ph.cy[i]; cz[s] = ph.cz[i];
                mode[s] = ph.mode[i]; omode[s] = ph.omode[i]; cfacet[s] = ph.cfacet[i]; pid[s] = ph.pid[i];
            }
        };
        const unsigned hw = std::thread::hardware_concurrency();
        const int nt = cntp < 8192 ? 1 : (int)std::min<unsigned>(8, hw ? hw : 1);
        if (nt <= 1) { apply(0, cntp); return; }
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(apply, cntp * k / nt, cntp * (k + 1) / nt);
        for (auto& t : th) t.join();
    };
    // the first round of the (small) patch on s_out; the occupations follow on the ctx stream while the host applies it
    const long long first = std::min(nd, cap);
    if (patch_ok && fetch_patch((size_t)first, ctx->s_out)) return -1;
    NK_CK(cudaMemcpyAsync(occ, P.occ, m * 8, cudaMemcpyDeviceToHost, st));
    if (!patch_ok) {
        NK_CK(cudaStreamSynchronize(ctx->s_out));    // the chunk downloads write the same host arrays
        NK_CK(cudaMemcpyAsync(mode, P.mode, m * 4, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(omode, P.omode, m * 4, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(cfacet, P.cfacet, m * 4, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(cx, P.cx, m * 8, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(cy, P.cy, m * 8, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(cz, P.cz, m * 8, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(pid, P.pid, m * 8, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(px, P.px, m * 8, cudaMemcpyDeviceToHost, st));      // stale for the rewritten slots
        NK_CK(cudaMemcpyAsync(py, P.py, m * 8, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(pz, P.pz, m * 8, cudaMemcpyDeviceToHost, st));
        NK_CK(cudaMemcpyAsync(tc, P.tc, m * 8, cudaMemcpyDeviceToHost, st));
    }
    NK_CK(cudaStreamSynchronize(ctx->s_out));    // all chunk downloads (enqueued before the patch) and the patch
    lap("d2h patch");
    if (patch_ok) {
        apply_patch(first);
        for (long long off = cap; off < nd; off += cap) {       // more dirty slots than one staging buffer holds: further rounds
            const long long k = std::min(cap, nd - off);
            k_pack_dirty<<<ctx->n_sm * 4, 256, 0, ctx->s_out>>>(P, pd, cap, off, ctx->patch_count_dev);
            NK_CK(cudaGetLastError());
            if (fetch_patch((size_t)k, ctx->s_out)) return -1;
            NK_CK(cudaStreamSynchronize(ctx->s_out));
            apply_patch(k);
        }
    }
    lap("patched");
    NK_CK(cudaStreamSynchronize(st));
    lap("d2h occ");
    ctx->xfer_h2d = (long long)n * (dense_omode ? 48 : 44) + (n_cold >= 0 ? n_cold * 44 : (long long)n * (dense_omode ? 36 : 40));
    ctx->xfer_d2h = (long long)n * 32 + (long long)m * 8 + 8 + (patch_ok ? nd * 80 : (long long)m * 76);
    if (trace) fprintf(stderr, "[nk pipe] patch entries %lld (hits %u, new %u), cold records uploaded %lld\n", nd, cnt[0], cnt[1], n_cold);
    if (n_out) *n_out = ns;
    return 0;
}

int nk_advance_host(nk_ctx* ctx, int64_t n_in, int n_steps, double* px, double* py, double* pz, double* tc, double* occ,
                    int32_t* mode, int32_t* omode, int32_t* cfacet, double* cx, double* cy, double* cz, int64_t* pid,
                    int64_t* n_out, double* T_sv_out, double* E_sv_out, int64_t* N_sv_out) {
    cudaSetDevice(ctx->device);
    if (nk_check_ready(ctx)) return -1;
    NkP& P = ctx->P;
    if (n_in > P.cap) { ctx->err = "n_in exceeds bound capacity"; return -1; }
    const bool pipe = ctx->use_pipeline && n_steps == 1 && (P.world == 1 || P.comm_on) && n_in >= (1 << 20) && !ctx->profiling;
    int rc = pipe ? nk_advance_host_pipelined(ctx, n_in, px, py, pz, tc, occ, mode, omode, cfacet, cx, cy, cz, pid, n_out)
                  : nk_advance_host_simple(ctx, n_in, n_steps, px, py, pz, tc, occ, mode, omode, cfacet, cx, cy, cz, pid, n_out);
    if (rc) return rc;
    return nk_get_results(ctx, T_sv_out, E_sv_out, N_sv_out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int nk_outside_slots(nk_ctx* ctx, double tol, int32_t* slots_dev, int64_t cap, int64_t* n_found) {
    cudaSetDevice(ctx->device);
    if (nk_check_ready(ctx)) return -1;
    if (!ctx->patch_count_dev) NK_CK(cudaMalloc(&ctx->patch_count_dev, 2 * sizeof(unsigned int)));
    NK_CK(cudaMemsetAsync(ctx->patch_count_dev, 0, 2 * sizeof(unsigned int), ctx->stream));
    k_outside_slots<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(ctx->P, tol, slots_dev, cap, ctx->patch_count_dev);
    NK_CK(cudaGetLastError());
    unsigned int c = 0;
    NK_CK(cudaMemcpyAsync(&c, ctx->patch_count_dev, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    NK_CK(cudaStreamSynchronize(ctx->stream));
    if (n_found) *n_found = (int64_t)c;
    return 0;
}

int nk_debug_trace(nk_ctx* ctx, uint64_t* out8) {
    cudaSetDevice(ctx->device);
    if (!ctx->P.trace) { ctx->err = "tracing is off (NK_TRACE=1 at nk_create)"; return -1; }
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(out8, ctx->P.trace, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    NK_CK(cudaMemset(ctx->P.trace, 0, 8 * sizeof(uint64_t)));
    return 0;
}

int nk_last_transfer_bytes(nk_ctx* ctx, int64_t* h2d, int64_t* d2h) {
    if (h2d) *h2d = ctx->xfer_h2d;
    if (d2h) *d2h = ctx->xfer_d2h;
    return 0;
}

// ---- multi-GPU plumbing ---------------------------------------------------------------------------------------
int nk_set_rank(nk_ctx* ctx, int rank, int world) {
    NkP& P = ctx->P;
    if (world < 1 || rank < 0 || rank >= world) { ctx->err = "bad rank/world"; return -1; }
    if (world > 255) { ctx->err = "at most 255 ranks (emission copies are dealt with an 8-bit counter)"; return -1; }
    // every rank advances the whole reservoir table; copies are dealt round-robin per table entry (nk_emit_owner)
    P.rank = rank; P.world = world;
    return 0;
}

// ---- run state that is neither a table nor a particle: reservoir counters, emission deal counters, the accumulators of the
//      current convergence window, N_leaving of the previous step, the results block.  Checkpoints use it to continue
//      bit-exactly WITHOUT rebuilding the tables (nk_set_reservoirs would also reset the accumulators and the rank).
int nk_results_len(nk_ctx* ctx) { return nk_out_len(ctx->P.S, ctx->P.R); }
static int nk_run_state_io(nk_ctx* ctx, bool write, double* res_counter, uint8_t* res_fire, double* res_acc, double* n_leaving, double* results) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (!P.res_counter || !P.out) { ctx->err = "nk_set_reservoirs first"; return -1; }
    NK_CK(cudaStreamSynchronize(ctx->stream));
    const size_t rm = (size_t)P.R * P.M;
    auto io = [&](void* dev, void* host, size_t bytes) -> cudaError_t {
        if (!host || !bytes) return cudaSuccess;
        return write ? cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice) : cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost);
    };
    NK_CK(io(P.res_counter, res_counter, rm * 8));
    NK_CK(io(P.res_fire, res_fire, rm));
    NK_CK(io(P.res_acc, res_acc, 4 * (size_t)P.R * 8));
    NK_CK(io(P.res_nleave, n_leaving, (size_t)P.R * 8));
    NK_CK(io(P.out, results, (size_t)nk_out_len(P.S, P.R) * 8));
    return 0;
}
int nk_get_run_state(nk_ctx* ctx, double* res_counter, uint8_t* res_fire, double* res_acc, double* n_leaving, double* results) {
    return nk_run_state_io(ctx, false, res_counter, res_fire, res_acc, n_leaving, results);
}
int nk_set_run_state(nk_ctx* ctx, const double* res_counter, const uint8_t* res_fire, const double* res_acc, const double* n_leaving,
                     const double* results) {
    return nk_run_state_io(ctx, true, (double*)res_counter, (uint8_t*)res_fire, (double*)res_acc, (double*)n_leaving, (double*)results);
}
int nk_acc_buffer(nk_ctx* ctx, double** p, int64_t* n) {
    if (!ctx->P.acc) { ctx->err = "accumulators not allocated yet"; return -1; }
    *p = ctx->P.acc; *n = nk_acc_len(ctx->P.S, ctx->P.R);
    return 0;
}
// ---- set-up helper: E(T) table of Phonon.initialise_temperature_function on the device --------------------------
// crystal_energy(T) = sum_active hbar*omega*n0(T, omega) / (Q V_uc) + zero_point   (Phonon.py:352-362).  One block per
// temperature, f64 tree reduction (the order of the sum differs from NumPy's pairwise sum: ~1e-16 relative).
__global__ void __launch_bounds__(256) k_energy_table(int M, const double* __restrict__ omega, const unsigned char* __restrict__ active,
                                                      int nT, const double* __restrict__ T, double hbar, double kb, double dens_norm,
                                                      double zero_point, double* __restrict__ out) {
    __shared__ double red[256];
    for (int it = blockIdx.x; it < nT; it += gridDim.x) {
        const double Tk = T[it];
        double acc = 0.0;
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            const double w = omega[m];
            if (active[m] && Tk > 0.0 && w > 0.0) {
                const double x = nk_div(nk_mul(w, hbar), nk_mul(Tk, kb));
                acc += nk_mul(nk_mul(hbar, w), nk_div(1.0, nk_sub(exp(x), 1.0)));
            }
        }
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[it] = nk_add(nk_div(red[0], dens_norm), zero_point);
        __syncthreads();
    }
}

int nk_energy_table(int device, int M, const double* omega, const uint8_t* active, int nT, const double* T, double hbar, double kb,
                    double dens_norm, double zero_point, double* out) {
    if (cudaSetDevice(device) != cudaSuccess) { g_create_err = "nk_energy_table: no such CUDA device"; return -1; }
    double *d_w = nullptr, *d_T = nullptr, *d_o = nullptr; unsigned char* d_a = nullptr;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t c) { if (e == cudaSuccess) e = c; return c == cudaSuccess; };
    ok(cudaMalloc(&d_w, (size_t)M * 8)); ok(cudaMalloc(&d_a, (size_t)M)); ok(cudaMalloc(&d_T, (size_t)nT * 8)); ok(cudaMalloc(&d_o, (size_t)nT * 8));
    if (e == cudaSuccess) {
        ok(cudaMemcpy(d_w, omega, (size_t)M * 8, cudaMemcpyHostToDevice));
        ok(cudaMemcpy(d_a, active, (size_t)M, cudaMemcpyHostToDevice));
        ok(cudaMemcpy(d_T, T, (size_t)nT * 8, cudaMemcpyHostToDevice));
        k_energy_table<<<std::min(nT, 148 * 8), 256>>>(M, d_w, d_a, nT, d_T, hbar, kb, dens_norm, zero_point, d_o);
        ok(cudaGetLastError());
        ok(cudaMemcpy(out, d_o, (size_t)nT * 8, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_w); cudaFree(d_a); cudaFree(d_T); cudaFree(d_o);
    if (e != cudaSuccess) { g_create_err = std::string("nk_energy_table: ") + cudaGetErrorString(e); return -1; }
    return 0;
}

// ---- fused exchange over NVLink peer memory ---------------------------------------------------------------------
int nk_comm_export(nk_ctx* ctx, void* handle_out) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (!P.acc) { ctx->err = "set the tables first (accumulator length unknown)"; return -1; }
    if (P.world < 1 || P.world > 8) { ctx->err = "fused exchange supports 1..8 ranks"; return -1; }
    if (!ctx->comm_block) {
        const size_t flag_bytes = 256;
        const size_t bytes = flag_bytes + 2 * (size_t)P.world * 3 * nk_acc_len(P.S, P.R) * sizeof(double);   // {lo, hi, side} per entry
        // an allocation of its own (>= 2 MB): IPC handles name the underlying allocation, and small blocks of several
        // contexts of one process would share one -- a peer cannot map the same allocation twice
        NK_CK(cudaMalloc(&ctx->comm_block, std::max<size_t>(bytes, (size_t)2 << 20)));
        NK_CK(cudaMemset(ctx->comm_block, 0, bytes));
        ctx->owned.push_back(ctx->comm_block);
        P.flags_local = reinterpret_cast<unsigned long long*>(ctx->comm_block);
        P.mbox_local = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->comm_block) + flag_bytes);
    }
    cudaIpcMemHandle_t h;
    NK_CK(cudaIpcGetMemHandle(&h, ctx->comm_block));
    static_assert(sizeof(h) == 64, "IPC handle is 64 bytes");
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int nk_comm_import(nk_ctx* ctx, int peer, const void* handle) {
    cudaSetDevice(ctx->device);
    NkP& P = ctx->P;
    if (!ctx->comm_block) { ctx->err = "call nk_comm_export first"; return -1; }
    if (peer < 0 || peer >= P.world) { ctx->err = "peer rank out of range"; return -1; }
    void* base = ctx->comm_block;
    if (peer != P.rank) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof(h));
        NK_CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_open.push_back(base);
    }
    P.peer_flags[peer] = reinterpret_cast<unsigned long long*>(base);
    P.peer_mbox[peer] = reinterpret_cast<double*>(reinterpret_cast<char*>(base) + 256);
    ctx->comm_imported |= 1u << peer;
    return 0;
}

int nk_comm_enable(nk_ctx* ctx, int enable) {
    NkP& P = ctx->P;
    if (enable) {
        if (P.world > 1 && ctx->comm_imported != (1u << P.world) - 1u) { ctx->err = "not every peer mailbox has been imported"; return -1; }
        P.comm_on = P.world > 1 ? 1 : 0;
    } else {
        // also unmap the peers' mailboxes, so that every rank can release them before any rank frees its own block
        P.comm_on = 0;
        cudaSetDevice(ctx->device);
        NK_CK(cudaStreamSynchronize(ctx->stream));
        for (void* p : ctx->ipc_open) cudaIpcCloseMemHandle(p);
        ctx->ipc_open.clear();
        ctx->comm_imported = 0;
    }
    return 0;
}

}  // extern "C"
