// nk_kernels.cu -- sm_100a kernels + C ABI of the Nano-kappa particle loop (see include/nk_b200.h).
//
// One timestep = four launches on one stream, no host synchronisation, replayable as a CUDA graph:
//
//   k_step      streaming pass over the particle SoA (84 algorithmic bytes per particle):
//               [lifetime relaxation of the previous step] -> drift -> nearest-subvolume ->
//               block-privatised per-SV energy/count/flux bins.  Particles whose next collision
//               falls inside this step are only appended to a hit list.
//   k_emit      reservoir emission over the (R, Q*J) entry table (fill_reservoirs +
//               add_reservoir_particles), recycling free slots.
//   k_boundary  one thread per hit-list entry: absorb / periodic / rough event loop.
//   k_finalize  per-SV sums -> energy density -> temperature (table inversion), heat flux and
//               kappa on convergence steps, reservoir balances; resets the accumulators.
//
// lifetime_scattering(k) needs T_sv(k), a grid-wide dependency; instead of a second pass over the
// particles it is applied at the head of k_step(k+1) (nothing reads `occ` in between except
// outputs, which call nk_flush_relaxation).  Energies use the pre-relaxation occupation and the
// previous step's T_sv exactly like the reference (Population.py:704-713, :1754-1756).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#include <thread>
#include <vector>
#include <cmath>
#include <cstdlib>

#include "../../include/nk_b200.h"
#include "nk_device.cuh"

// =================================================================================================
// kernels
// =================================================================================================

// ---- shared-memory copy of the subvolume tables used by the streaming kernels --------------------
struct NkSvSmem {
    double* svc; double* sv_axis; double* sv_mid; double* T_sv;
};
__device__ __forceinline__ NkSvSmem nk_load_sv(const NkP& P, double* sm) {
    NkSvSmem s;
    s.svc = sm; s.sv_axis = sm + 3 * P.S; s.sv_mid = s.sv_axis + P.S; s.T_sv = s.sv_mid + P.S;
    for (int i = threadIdx.x; i < 3 * P.S; i += blockDim.x) s.svc[i] = P.svc[i];
    for (int i = threadIdx.x; i < P.S; i += blockDim.x) {
        s.sv_axis[i] = P.sv_axis[i];
        s.T_sv[i] = P.T_sv[i];
        if (i < P.S - 1) s.sv_mid[i] = P.sv_mid[i];
    }
    return s;
}
__host__ __device__ static inline size_t nk_sv_smem_doubles(int S) { return (size_t)6 * S; }

// Mesh.find_boundary operator seam: one ray per thread, triangles staged through shared memory.
#define NK_FACE_TILE 256
__global__ void __launch_bounds__(256) k_find_boundary(NkP P, long long n, const double* __restrict__ x,
                                                        const double* __restrict__ v, double* __restrict__ xc,
                                                        double* __restrict__ tc, int* __restrict__ fc) {
    __shared__ NkFace sf[NK_FACE_TILE];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0;
    if (i < n) { px = x[3 * i]; py = x[3 * i + 1]; pz = x[3 * i + 2]; vx = v[3 * i]; vy = v[3 * i + 1]; vz = v[3 * i + 2]; }
    double tbest = CUDART_INF; int fbest = -1;
    for (int f0 = 0; f0 < P.F; f0 += NK_FACE_TILE) {
        int nt = min(NK_FACE_TILE, P.F - f0);
        __syncthreads();
        const double* src = reinterpret_cast<const double*>(P.faces + f0);
        double* dst = reinterpret_cast<double*>(sf);
        for (int k = threadIdx.x; k < nt * (int)(sizeof(NkFace) / 8); k += blockDim.x) dst[k] = src[k];
        __syncthreads();
        if (i < n) nk_ray_faces(sf, nt, px, py, pz, vx, vy, vz, tbest, fbest);
    }
    if (i < n) {
        tc[i] = tbest; fc[i] = fbest;
        xc[3 * i] = nk_add(px, nk_mul(tbest, vx)); xc[3 * i + 1] = nk_add(py, nk_mul(tbest, vy)); xc[3 * i + 2] = nk_add(pz, nk_mul(tbest, vz));
    }
}

// first collision of every live slot (Population.py:308-316)
__global__ void __launch_bounds__(256) k_init_collisions(NkP P) {
    __shared__ NkFace sf[NK_FACE_TILE];
    const long long n = P.dyn->n_slots;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        long long i = base + threadIdx.x;
        bool live = i < n && P.mode[i] >= 0;
        double px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0;
        if (live) { NkMode m = P.mprop[P.mode[i]]; px = P.px[i]; py = P.py[i]; pz = P.pz[i]; vx = m.vx; vy = m.vy; vz = m.vz; }
        double tbest = CUDART_INF; int fbest = -1;
        for (int f0 = 0; f0 < P.F; f0 += NK_FACE_TILE) {
            int nt = min(NK_FACE_TILE, P.F - f0);
            __syncthreads();
            const double* src = reinterpret_cast<const double*>(P.faces + f0);
            double* dst = reinterpret_cast<double*>(sf);
            for (int k = threadIdx.x; k < nt * (int)(sizeof(NkFace) / 8); k += blockDim.x) dst[k] = src[k];
            __syncthreads();
            if (live) nk_ray_faces(sf, nt, px, py, pz, vx, vy, vz, tbest, fbest);
        }
        if (live) {
            P.tc[i] = nk_div(tbest, P.dt); P.cfacet[i] = fbest;
            P.cx[i] = nk_add(px, nk_mul(tbest, vx)); P.cy[i] = nk_add(py, nk_mul(tbest, vy)); P.cz[i] = nk_add(pz, nk_mul(tbest, vz));
        }
    }
}

__global__ void __launch_bounds__(256) k_classify(NkP P, long long n, const double* __restrict__ x, int* __restrict__ sv,
                                                   unsigned long long* __restrict__ counts) {
    extern __shared__ double sm[];
    NkSvSmem s = nk_load_sv(P, sm);
    unsigned int* hist = reinterpret_cast<unsigned int*>(sm + nk_sv_smem_doubles(P.S));
    for (int i = threadIdx.x; i < P.S; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int r = nk_classify(P, s.svc, s.sv_mid, x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        sv[i] = r;
        if (counts) atomicAdd(hist + r, 1u);
    }
    __syncthreads();
    if (counts) for (int i = threadIdx.x; i < P.S; i += blockDim.x) if (hist[i]) atomicAdd(counts + i, (unsigned long long)hist[i]);
}

__global__ void k_occupation(NkP P, long long n, const double* T, const double* omega, double* occ) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        occ[i] = nk_bose(P, T[i], omega[i]);
}
__global__ void k_lifetime(NkP P, long long n, const double* T, const int* mode, double* tau) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        tau[i] = nk_tau(P, T[i], mode[i]);
}
__global__ void k_table(NkP P, long long n, const double* in, double* out, int e_to_t) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = e_to_t ? nk_interp_table(P.Ea, P.Ta, P.nE, in[i], P.Ta[0], P.Ta[P.nE - 1])
                        : nk_interp_table(P.Ta, P.Ea, P.nE, in[i], P.Ea[0], P.Ea[P.nE - 1]);
}
__global__ void k_particle_T(NkP P, long long n, const double* x, double* T) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        T[i] = nk_particle_T(P, P.svc, P.sv_axis, P.sv_mid, P.T_sv, x[3 * i], x[3 * i + 1], x[3 * i + 2], -1);
}

// ---- lifetime_scattering of one particle (Population.py:1701-1710) ---------------------------------
__device__ __forceinline__ double nk_relax(const NkP& P, double T, int mode, double omega, double occ) {
    double tau = nk_tau(P, T, mode);
    double n0 = nk_bose(P, T, omega);
    if (tau > 0.0) return nk_add(n0, nk_mul(nk_sub(occ, n0), exp(nk_div(-P.dt, tau))));
    return n0;
}

// ---- the streaming kernel ----------------------------------------------------------------------------
//
// One pass over the particle SoA per timestep: 5 x 16 B + 8 B vector loads, 5 x 16 B vector stores per
// particle PAIR, one 64 B gather of the mode record {omega, v_g, tau slabs}.  Everything that depends only
// on the subvolume (1/(k_B T_sv), tau interpolation weight and slab) is hoisted into a per-block
// shared-memory table, so the per-particle arithmetic is: one Bose-Einstein evaluation shared by the
// relaxation of the previous step and the energy of this one (the particle usually stays in its
// subvolume), one decay exponential, the drift and a 1-D slice lookup.
//
// FAST = slice subvolumes + nearest temperature rule (the Si/Ge thin-film configurations).  The general
// variant (linear interpolation along the slices, or grid/voronoi subvolumes) evaluates the per-particle
// temperature and tau explicitly.
//
// Occupation / energy arithmetic uses a Newton-refined reciprocal instead of IEEE division (<= 2 ulp);
// positions, collision times and every integer result keep the reference's exact operation order.
#define NK_STEP_THREADS 256
#ifndef NK_STEP_MIN_BLOCKS
#define NK_STEP_MIN_BLOCKS 4
#endif

// Newton-refined reciprocal of a positive normal double (<= 2 ulp): MUFU.RCP64H + 4 DFMA, no branch
__device__ __forceinline__ double nk_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

// Branch-free exp for the occupation arithmetic: argument clamped to [-708, 709] (results there are
// ~1e-308 / ~1e308, i.e. 0 / inf for every use below), Cody-Waite reduction, degree-13 Taylor polynomial on
// |r| <= ln2/2 (truncation 4e-18), exponent added with integer arithmetic.  Coefficients live in constant
// memory so that they are DFMA operands instead of 64-bit immediates.
__constant__ double NK_EXP_C[12] = {
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
    1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
__device__ __forceinline__ double nk_exp(double x) {
    x = fmin(fmax(x, -708.0), 709.0);
    const double magic = 6755399441055744.0;                      // 2^52 + 2^51: rounds to nearest integer
    const double t = fma(x, 1.4426950408889634, magic);
    const int k = __double2loint(t);
    const double kf = t - magic;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double p = NK_EXP_C[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, r, NK_EXP_C[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// Bose-Einstein with the hoisted 1/(k_B T): a = hbar*omega.  exp(x)-1 == 0 only for x == 0 -> inf like 1/0.
__device__ __forceinline__ double nk_bose_fast(double a, double omega, double invb) {
    const double d = nk_exp(a * invb) - 1.0;
    const double v = d > 0.0 ? nk_rcp(d) : CUDART_INF;
    return (invb > 0.0 && omega > 0.0) ? v : 0.0;
}
__device__ __forceinline__ double nk_decay(double dt, double tau) {       // exp(-dt/tau), tau > 0
    return nk_exp(-dt * nk_rcp(tau));
}

// slice index of a coordinate = searchsorted(mid, xa, 'left') for uniformly spaced slices: arithmetic guess
// verified against the padded boundary table midp[0..S] (midp[0] = -inf, midp[S] = +inf); the exact
// bisection runs only when the guess is off (never for in-range coordinates, kept for safety).
__device__ __forceinline__ int nk_slice_lookup(const NkP& P, const double* midp, const double* mid, double xa, double& lo, double& hi) {
    int g = __double2int_rd((xa - P.sv_x0) * P.sv_inv_dx);
    g = max(0, min(g, P.S - 1));
    lo = midp[g]; hi = midp[g + 1];
    if (!((lo < xa) && (xa <= hi))) {
        g = P.S > 1 ? nk_searchsorted_left(mid, P.S - 1, xa, P.sv_inv_dx) : 0;
        lo = midp[g]; hi = midp[g + 1];
    }
    return g;
}

struct NkSvHot {            // per-subvolume values hoisted out of the particle loop (shared memory)
    double* invb;           // 1 / (k_B T_sv)   (0 when T_sv <= 0 -> occupation 0)
    double* tw;             // tau interpolation weight w
    double* midp;           // (S+1) slice boundaries padded with -inf / +inf
    int* tr;                // slab offset into the mode record (0..2) or -1 -> full table
    int* ti;                // absolute slab index
};
__host__ __device__ static inline size_t nk_hot_smem_bytes(int S) { return (size_t)S * (3 * 8 + 2 * 4) + 16; }
// carve + fill the table; caller syncs
__device__ __forceinline__ NkSvHot nk_load_hot(const NkP& P, void* mem) {
    NkSvHot h;
    const int S = P.S;
    h.invb = reinterpret_cast<double*>(mem); h.tw = h.invb + S; h.midp = h.tw + S;
    h.tr = reinterpret_cast<int*>(h.midp + S + 1); h.ti = h.tr + S;
    for (int i = threadIdx.x; i <= S; i += blockDim.x)
        h.midp[i] = i == 0 ? -CUDART_INF : (i == S ? CUDART_INF : P.sv_mid[i - 1]);
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        double T = P.T_sv[i];
        h.invb[i] = T > 0.0 ? nk_div(1.0, nk_mul(T, P.kb)) : 0.0;
        int it = nk_T_index(P, T);
        double t0 = P.Tg[it], t1 = P.Tg[it + 1];
        h.tw[i] = nk_div(nk_sub(T, t0), nk_sub(t1, t0));
        if (P.is_slice && P.interp == NK_INTERP_LINEAR)        // the linear rule has no per-slice tau weight: reuse the slot
            h.tw[i] = i > 0 ? nk_div(1.0, nk_sub(P.sv_axis[i], P.sv_axis[i - 1])) : 0.0;
        int r = it - P.tau_i0;
        h.tr[i] = (r >= 0 && r <= 2) ? r : -1;
        h.ti[i] = it;
    }
    return h;
}

// 64 B mode record as two 256-bit non-coherent loads (LDG.E.256): {omega, v} and the tau slabs
__device__ __forceinline__ void nk_ld256(const double* p, double4& v) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
}


// lifetime_scattering of one particle (Population.py:1701-1710) at its position BEFORE the drift of the
// next step.  Returns the relaxed occupation; be0 / g0 = equilibrium occupation and slice used (FAST).
template <bool FAST>
__device__ __forceinline__ double nk_relax_particle(const NkP& P, const NkSvSmem& s, const NkSvHot& h, double x, double y, double z,
                                                    int mode, double omega, double a, const double4& mt, double occ,
                                                    double& be0, int& g0) {
    double tau;
    if (FAST) {
        const double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
        double lo_b, hi_b;
        g0 = nk_slice_lookup(P, h.midp, s.sv_mid, xa, lo_b, hi_b);                        // interp1d 'nearest'
        be0 = nk_bose_fast(a, omega, h.invb[g0]);
        const int r = h.tr[g0];
        const double w = h.tw[g0];
        double lo = r == 0 ? mt.x : (r == 1 ? mt.y : mt.z);
        double hi = r == 0 ? mt.y : (r == 1 ? mt.z : mt.w);
        if (r < 0) {                                                                       // temperature outside the packed slabs
            const int it = h.ti[g0];
            lo = __ldg(P.tau + (size_t)it * P.M + mode);
            hi = __ldg(P.tau + (size_t)(it + 1) * P.M + mode);
        }
        tau = nk_add(nk_mul(lo, nk_sub(1.0, w)), nk_mul(hi, w));
    } else {
        // general rule (linear interpolation between slices, nearest centre / RBF of grid and voronoi subvolumes).  The
        // temperature only feeds occupations, so reciprocals replace IEEE divisions (1e-16 relative), and the lifetime
        // comes from the tau slabs of the mode record already in registers whenever T lies inside them.
        double Ti;
        if (P.is_slice && P.interp == NK_INTERP_LINEAR && P.S > 1) {
            const double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
            int idx = nk_searchsorted_left(s.sv_axis, P.S, xa, P.sv_inv_dx);
            idx = max(1, min(idx, P.S - 1));
            const double xl = s.sv_axis[idx - 1], xh = s.sv_axis[idx];
            const double inv = h.tw[idx];                                              // 1 / (xh - xl), see nk_load_hot
            Ti = ((xa - xl) * inv) * s.T_sv[idx] + ((xh - xa) * inv) * s.T_sv[idx - 1];
        } else {
            Ti = nk_particle_T(P, s.svc, s.sv_axis, s.sv_mid, s.T_sv, x, y, z, -1);
        }
        const int it = nk_T_index(P, Ti);
        const double t0 = __ldg(P.Tg + it), t1 = __ldg(P.Tg + it + 1);
        const double w = (Ti - t0) * nk_rcp(t1 - t0);
        const int r = it - P.tau_i0;
        double lo = r == 0 ? mt.x : (r == 1 ? mt.y : mt.z);
        double hi = r == 0 ? mt.y : (r == 1 ? mt.z : mt.w);
        if (r < 0 || r > 2) {
            lo = __ldg(P.tau + (size_t)it * P.M + mode);
            hi = __ldg(P.tau + (size_t)(it + 1) * P.M + mode);
        }
        tau = nk_add(nk_mul(lo, nk_sub(1.0, w)), nk_mul(hi, w));
        be0 = nk_bose_fast(a, omega, Ti > 0.0 ? nk_rcp(nk_mul(Ti, P.kb)) : 0.0);
        g0 = -1;
    }
    const double relaxed = be0 + (occ - be0) * nk_decay(P.dt, tau > 0.0 ? tau : 1.0);
    return tau > 0.0 ? relaxed : be0;
}

// Reservoir counters (Population.fill_reservoirs 'constant', Population.py:358-370): every entry of the
// (R, Q*J) table advances its fractional counter; entries that emit this step are appended to the
// emission list that k_rare consumes.  Runs as the prologue of the streaming kernel (grid-stride over all
// its blocks): it does not depend on the particles at all.
__device__ __forceinline__ long long nk_one_to_one_share(const NkP& P, int r) {
    const long long n = (long long)P.res_nleave[r];
    return n > P.rank ? (n - P.rank + P.world - 1) / P.world : 0;
}

__device__ __forceinline__ void nk_emit_scan(const NkP& P) {
    if (P.res_gen == NK_RESGEN_ONE_TO_ONE) {
        // one_to_one (Population.py:457-489): as many particles as the reservoir absorbed in the previous step; the
        // k-th of them belongs to rank k % world.  No table scan: k_rare decodes (reservoir, k) from the item index.
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned int total = 0;
            for (int r = 0; r < P.R; ++r) total += (unsigned int)nk_one_to_one_share(P, r);
            P.dyn->n_emit = total;
        }
        return;
    }
    const int mspan = P.emit_m_hi - P.emit_m_lo;
    const long long total = (long long)P.R * mspan;
    const long long step = P.dyn->step;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / mspan);
        const int m = P.emit_m_lo + (int)(e % mspan);
        const size_t idx = (size_t)r * P.M + m;
        const double prob = P.enter_prob[idx];
        const double fixed = floor(prob);
        int extra;
        if (P.res_gen == NK_RESGEN_FIXED_RATE) {
            // fixed_rate (Population.py:408-417): a fresh dice per (reservoir, mode) and step instead of the counter
            double dice, unused;
            nk_uniforms(P, NK_EMIT_ID_BASE + (((step * P.R + r) * (long long)P.M + m) * NK_EMIT_CMAX), step, NK_STREAM_EMIT_C, dice, unused);
            extra = dice <= nk_sub(prob, fixed) ? 1 : 0;
            P.emit_u[idx] = dice;
        } else {
            double cnt = nk_add(P.res_counter[idx], nk_sub(prob, fixed));
            extra = cnt >= 1.0 ? 1 : 0;
            cnt = nk_sub(cnt, (double)extra);
            P.res_counter[idx] = cnt;
        }
        int n_new = (int)fixed + extra;
        if (n_new == 0) continue;
        if (n_new > NK_EMIT_CMAX) { atomicOr(&P.dyn->error, NK_ERR_CMAX); n_new = NK_EMIT_CMAX; }
        const unsigned int k = atomicAdd(&P.dyn->n_emit, 1u);
        P.emitlist[k] = make_int2((r << 8) | n_new, m);
    }
}

// (block-private fixed-point bins: nk_bin_add in nk_device.cuh)

// one live particle: deferred relaxation -> drift -> (if no collision this step) subvolume + energy bins.
// Returns true when the particle's collision falls inside this step (it then goes to the hit list).
template <bool HAS_ROUGH, bool FAST, bool RELAX, bool FLUX>
__device__ __forceinline__ bool nk_step_particle(const NkP& P, const NkSvSmem& s, const NkSvHot& h, long long* binE, long long* binF,
                                                 double* binX, unsigned int* binC, int md, int om, double& x, double& y, double& z,
                                                 double& tc, double& occ) {
    const NkModeHot* __restrict__ mhot = P.mhot;
    double4 ma, mt;
    nk_ld256(&mhot[md].omega, ma);        // omega, v_g
    nk_ld256(&mhot[md].t[0], mt);         // tau slabs
    double omega = ma.x;
    if (HAS_ROUGH && om != md) omega = mhot[om].omega;
    const double a = nk_mul(P.hbar, omega);
    const double dt = P.dt;
    double be0 = 0.0; int g0 = -1;
    if (RELAX) occ = nk_relax_particle<FAST>(P, s, h, x, y, z, md, omega, a, mt, occ, be0, g0);
    x = nk_add(x, nk_mul(ma.y, dt)); y = nk_add(y, nk_mul(ma.z, dt)); z = nk_add(z, nk_mul(ma.w, dt));
    tc = nk_sub(tc, 1.0);
    if (tc < 0.0) return true;
    int sv;
    if (FAST) {
        // nearest centre of a slice stack = 1-D lookup; inside 1e-6 A of a slice boundary the full
        // squared-distance comparison decides, so the index equals the reference's
        const double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
        double lo_b, hi_b;
        sv = nk_slice_lookup(P, h.midp, s.sv_mid, xa, lo_b, hi_b);
        if ((xa - lo_b < 1e-6) || (hi_b - xa < 1e-6)) sv = nk_classify(P, s.svc, s.sv_mid, x, y, z);
    } else {
        sv = nk_classify(P, s.svc, s.sv_mid, x, y, z);
    }
    double be1 = be0;
    if (!(FAST && RELAX && sv == g0)) be1 = nk_bose_fast(a, omega, h.invb[sv]);
    const double e = a * (occ - be1);
    nk_bin_add(binE + sv, binX + sv, e, NK_QE);
    atomicAdd(binC + sv, 1u);
    if (FLUX) {
        nk_bin_add(binF + 3 * sv, binX + P.S + 3 * sv, ma.y * e, NK_QF);
        nk_bin_add(binF + 3 * sv + 1, binX + P.S + 3 * sv + 1, ma.z * e, NK_QF);
        nk_bin_add(binF + 3 * sv + 2, binX + P.S + 3 * sv + 2, ma.w * e, NK_QF);
    }
    return false;
}

// warp-aggregated append of up to two slots per lane to the hit list (full-mask votes: call converged)
__device__ __forceinline__ void nk_push_hits(const NkP& P, unsigned int lane, bool h0, bool h1, long long base) {
    const unsigned int m0 = __ballot_sync(0xffffffffu, h0);
    const unsigned int m1 = __ballot_sync(0xffffffffu, h1);
    if (m0 | m1) {
        unsigned int pos = 0;
        if (lane == 0) pos = atomicAdd(&P.dyn->n_hits, __popc(m0) + __popc(m1));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        const unsigned int below = (1u << lane) - 1u;
        if (h0) P.hitlist[pos + __popc(m0 & below)] = (int)base;
        if (h1) P.hitlist[pos + __popc(m0) + __popc(m1 & below)] = (int)(base + 1);
    }
}

template <bool FLUX>
__device__ __forceinline__ void nk_flush_bins(const NkP& P, const long long* binE, const long long* binF, const double* binX,
                                              const unsigned int* binC) {
    const int S = P.S;
    double* acc = P.acc;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        if (binC[i]) {
            atomicAdd(acc + NK_ACC_E(S, P.R) + i, (double)binE[i] * (1.0 / NK_QE) + binX[i]);
            atomicAdd(acc + NK_ACC_CNT(S, P.R) + i, (double)binC[i]);
            if (FLUX) {
                for (int k = 0; k < 3; ++k)
                    atomicAdd(acc + NK_ACC_FLUX(S, P.R) + 3 * i + k, (double)binF[3 * i + k] * (1.0 / NK_QF) + binX[S + 3 * i + k]);
            }
        }
    }
}

// ---- variant A: direct 128-bit global loads/stores (any capacity) ------------------------------------------
template <bool HAS_ROUGH, bool FAST, bool RELAX, bool FLUX>
__global__ void __launch_bounds__(NK_STEP_THREADS, NK_STEP_MIN_BLOCKS) k_step(NkP P) {
    extern __shared__ double sm[];
    NkSvSmem s = nk_load_sv(P, sm);
    const int S = P.S;
    long long* binE = reinterpret_cast<long long*>(sm + nk_sv_smem_doubles(S));   // S   fixed-point energy sums
    long long* binF = binE + S;                                                      // 3S  fixed-point flux sums
    double* binX = reinterpret_cast<double*>(binF + 3 * S);                         // 4S  f64 side bins
    unsigned int* binC = reinterpret_cast<unsigned int*>(binX + 4 * S);             // S (+ pad to 8 B)
    NkSvHot h = nk_load_hot(P, binC + S + (S & 1));
    for (int i = threadIdx.x; i < S; i += blockDim.x) { binE[i] = 0; binC[i] = 0u; for (int k = 0; k < 3; ++k) binF[3 * i + k] = 0; for (int k = 0; k < 4; ++k) binX[4 * i + k] = 0.0; }
    __syncthreads();

    if (P.scan_emit) nk_emit_scan(P);
    const long long n = min((long long)P.dyn->n_slots, P.slot_hi);
    const unsigned int lane = threadIdx.x & 31u;

    // the loop bound is WARP-uniform (lane 0's index) because the hit-list append uses full-mask warp
    // votes; lanes past the end carry dead slots
    for (long long wbase = P.slot_lo + 2 * ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)); wbase < n;
         wbase += 2 * (long long)gridDim.x * blockDim.x) {
        const long long base = wbase + 2 * lane;
        const bool inb = base < n;
        double2 X = make_double2(0, 0), Y = X, Z = X, TC = X, OC = X;
        int2 MD = make_int2(-1, -1), OM = MD;
        if (inb) {
            X = *reinterpret_cast<const double2*>(P.px + base);
            Y = *reinterpret_cast<const double2*>(P.py + base);
            Z = *reinterpret_cast<const double2*>(P.pz + base);
            TC = *reinterpret_cast<const double2*>(P.tc + base);
            OC = *reinterpret_cast<const double2*>(P.occ + base);
            MD = *reinterpret_cast<const int2*>(P.mode + base);
            OM = MD;
            if (HAS_ROUGH) OM = *reinterpret_cast<const int2*>(P.omode + base);
        }
        bool h0 = false, h1 = false;
        if (base < n && MD.x >= 0) h0 = nk_step_particle<HAS_ROUGH, FAST, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.x, OM.x, X.x, Y.x, Z.x, TC.x, OC.x);
        if (base + 1 < n && MD.y >= 0) h1 = nk_step_particle<HAS_ROUGH, FAST, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.y, OM.y, X.y, Y.y, Z.y, TC.y, OC.y);
        nk_push_hits(P, lane, h0, h1, base);
        if (inb) {
            *reinterpret_cast<double2*>(P.px + base) = X;
            *reinterpret_cast<double2*>(P.py + base) = Y;
            *reinterpret_cast<double2*>(P.pz + base) = Z;
            *reinterpret_cast<double2*>(P.tc + base) = TC;
            *reinterpret_cast<double2*>(P.occ + base) = OC;
        }
    }
    __syncthreads();
    nk_flush_bins<FLUX>(P, binE, binF, binX, binC);
}

__device__ __forceinline__ unsigned int nk_smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
// ---- variant A2: variant A + per-thread software prefetch through shared memory (cp.async / LDGSTS) -----------
// Every thread copies the 88 bytes of ITS next two particles into a private shared-memory slot with cp.async
// while it works on the current pair, so two tiles of loads are in flight per warp without holding them in
// registers (occupancy stays at 4 blocks / SM).  A thread only reads back what it copied itself: no barrier.
struct NkPfStage {
    double2 x[NK_STEP_THREADS], y[NK_STEP_THREADS], z[NK_STEP_THREADS], tc[NK_STEP_THREADS], oc[NK_STEP_THREADS];
    int2 md[NK_STEP_THREADS], om[NK_STEP_THREADS];
};
__device__ __forceinline__ void nk_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(nk_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void nk_cp8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(nk_smem_u32(dst)), "l"(src) : "memory");
}
template <bool HAS_ROUGH, bool FAST, bool RELAX, bool FLUX>
__global__ void __launch_bounds__(NK_STEP_THREADS, NK_STEP_MIN_BLOCKS) k_step_pf(NkP P) {
    extern __shared__ __align__(128) unsigned char smraw[];
    NkPfStage* stage = reinterpret_cast<NkPfStage*>(smraw);
    double* sm = reinterpret_cast<double*>(smraw + 2 * sizeof(NkPfStage));
    NkSvSmem s = nk_load_sv(P, sm);
    const int S = P.S;
    long long* binE = reinterpret_cast<long long*>(sm + nk_sv_smem_doubles(S));
    long long* binF = binE + S;
    double* binX = reinterpret_cast<double*>(binF + 3 * S);
    unsigned int* binC = reinterpret_cast<unsigned int*>(binX + 4 * S);
    NkSvHot h = nk_load_hot(P, binC + S + (S & 1));
    for (int i = threadIdx.x; i < S; i += blockDim.x) { binE[i] = 0; binC[i] = 0u; for (int k = 0; k < 3; ++k) binF[3 * i + k] = 0; for (int k = 0; k < 4; ++k) binX[4 * i + k] = 0.0; }
    __syncthreads();

    nk_emit_scan(P);
    const long long n = P.dyn->n_slots;
    const unsigned int lane = threadIdx.x & 31u;
    const int t = threadIdx.x;
    const long long stride = 2 * (long long)gridDim.x * blockDim.x;

    auto prefetch = [&](int st, long long base) {
        if (base < n) {
            NkPfStage& T = stage[st];
            nk_cp16(&T.x[t], P.px + base); nk_cp16(&T.y[t], P.py + base); nk_cp16(&T.z[t], P.pz + base);
            nk_cp16(&T.tc[t], P.tc + base); nk_cp16(&T.oc[t], P.occ + base);
            nk_cp8(&T.md[t], P.mode + base);
            if (HAS_ROUGH) nk_cp8(&T.om[t], P.omode + base);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    long long wbase = 2 * ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u));
    int st = 0;
    if (wbase < n) prefetch(0, wbase + 2 * lane);
    for (; wbase < n; wbase += stride, st ^= 1) {
        const long long base = wbase + 2 * lane;
        prefetch(st ^ 1, base + stride);                        // next pair (an empty group past the end)
        asm volatile("cp.async.wait_group 1;" ::: "memory");    // everything but the newest group has landed
        const bool inb = base < n;
        double2 X = make_double2(0, 0), Y = X, Z = X, TC = X, OC = X;
        int2 MD = make_int2(-1, -1), OM = MD;
        if (inb) {
            const NkPfStage& T = stage[st];
            X = T.x[t]; Y = T.y[t]; Z = T.z[t]; TC = T.tc[t]; OC = T.oc[t]; MD = T.md[t];
            OM = MD;
            if (HAS_ROUGH) OM = T.om[t];
        }
        bool h0 = false, h1 = false;
        if (base < n && MD.x >= 0) h0 = nk_step_particle<HAS_ROUGH, FAST, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.x, OM.x, X.x, Y.x, Z.x, TC.x, OC.x);
        if (base + 1 < n && MD.y >= 0) h1 = nk_step_particle<HAS_ROUGH, FAST, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.y, OM.y, X.y, Y.y, Z.y, TC.y, OC.y);
        nk_push_hits(P, lane, h0, h1, base);
        if (inb) {
            *reinterpret_cast<double2*>(P.px + base) = X;
            *reinterpret_cast<double2*>(P.py + base) = Y;
            *reinterpret_cast<double2*>(P.pz + base) = Z;
            *reinterpret_cast<double2*>(P.tc + base) = TC;
            *reinterpret_cast<double2*>(P.occ + base) = OC;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    nk_flush_bins<FLUX>(P, binE, binF, binX, binC);
}

// ---- variant T: per-(mode, subvolume) tables ------------------------------------------------------------------
// With the nearest-temperature rule both the equilibrium occupation and the relaxation factor of a particle are
// functions of (mode, subvolume) only.  When there are many particles per (mode, subvolume) pair it is cheaper
// to tabulate {n0, exp(-dt/tau)} once per step (k_mode_tables, M x S entries, the arithmetic of nk_bose_fast /
// nk_decay, so results are bit-identical to the direct variants) than to evaluate two exponentials and three
// reciprocals per particle.  Particles are ordered by mode, so a warp gathers from a handful of table rows.
__global__ void __launch_bounds__(256) k_mode_tables(NkP P) {
    extern __shared__ double sm[];
    NkSvHot h = nk_load_hot(P, sm);
    __syncthreads();
    const int S = P.S;
    const long long total = (long long)P.M * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / S), sv = (int)(i % S);
        double4 ma, mt;
        nk_ld256(&P.mhot[m].omega, ma);
        nk_ld256(&P.mhot[m].t[0], mt);
        const double a = nk_mul(P.hbar, ma.x);
        const double be = nk_bose_fast(a, ma.x, h.invb[sv]);
        const int r = h.tr[sv];
        const double w = h.tw[sv];
        double lo = r == 0 ? mt.x : (r == 1 ? mt.y : mt.z);
        double hi = r == 0 ? mt.y : (r == 1 ? mt.z : mt.w);
        if (r < 0) {
            const int it = h.ti[sv];
            lo = __ldg(P.tau + (size_t)it * P.M + m);
            hi = __ldg(P.tau + (size_t)(it + 1) * P.M + m);
        }
        const double tau = nk_add(nk_mul(lo, nk_sub(1.0, w)), nk_mul(hi, w));
        const double dec = tau > 0.0 ? nk_decay(P.dt, tau) : 0.0;       // tau <= 0: relax straight to n0
        P.hot_tab[i] = make_double2(be, dec);
    }
}

template <bool HAS_ROUGH, bool RELAX, bool FLUX>
__device__ __forceinline__ bool nk_step_particle_tab(const NkP& P, const NkSvSmem& s, const NkSvHot& h, long long* binE, long long* binF,
                                                     double* binX, unsigned int* binC, int md, int om, double& x, double& y, double& z,
                                                     double& tc, double& occ) {
    double4 ma;
    nk_ld256(&P.mhot[md].omega, ma);        // omega, v_g
    double omega = ma.x;
    if (HAS_ROUGH && om != md) omega = P.mhot[om].omega;
    const double a = nk_mul(P.hbar, omega);
    const double dt = P.dt;
    const int S = P.S;
    const double2* __restrict__ row = P.hot_tab + (size_t)md * S;
    const double2* __restrict__ orow = (HAS_ROUGH && om != md) ? P.hot_tab + (size_t)om * S : row;
    double be0 = 0.0; int g0 = -1;
    if (RELAX) {
        const double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
        double lo_b, hi_b;
        g0 = nk_slice_lookup(P, h.midp, s.sv_mid, xa, lo_b, hi_b);          // interp1d 'nearest'
        const double2 t0 = __ldg(row + g0);
        be0 = (HAS_ROUGH && om != md) ? __ldg(orow + g0).x : t0.x;
        const double relaxed = be0 + (occ - be0) * t0.y;
        occ = t0.y > 0.0 ? relaxed : be0;
    }
    x = nk_add(x, nk_mul(ma.y, dt)); y = nk_add(y, nk_mul(ma.z, dt)); z = nk_add(z, nk_mul(ma.w, dt));
    tc = nk_sub(tc, 1.0);
    if (tc < 0.0) return true;
    const double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
    double lo_b, hi_b;
    int sv = nk_slice_lookup(P, h.midp, s.sv_mid, xa, lo_b, hi_b);
    if ((xa - lo_b < 1e-6) || (hi_b - xa < 1e-6)) sv = nk_classify(P, s.svc, s.sv_mid, x, y, z);
    double be1 = be0;
    if (!(RELAX && sv == g0)) be1 = __ldg(orow + sv).x;
    const double e = a * (occ - be1);
    nk_bin_add(binE + sv, binX + sv, e, NK_QE);
    atomicAdd(binC + sv, 1u);
    if (FLUX) {
        nk_bin_add(binF + 3 * sv, binX + P.S + 3 * sv, ma.y * e, NK_QF);
        nk_bin_add(binF + 3 * sv + 1, binX + P.S + 3 * sv + 1, ma.z * e, NK_QF);
        nk_bin_add(binF + 3 * sv + 2, binX + P.S + 3 * sv + 2, ma.w * e, NK_QF);
    }
    return false;
}

template <bool HAS_ROUGH, bool FAST, bool RELAX, bool FLUX>
__global__ void __launch_bounds__(NK_STEP_THREADS, NK_STEP_MIN_BLOCKS) k_step_tab(NkP P) {
    extern __shared__ double sm[];
    NkSvSmem s = nk_load_sv(P, sm);
    const int S = P.S;
    long long* binE = reinterpret_cast<long long*>(sm + nk_sv_smem_doubles(S));
    long long* binF = binE + S;
    double* binX = reinterpret_cast<double*>(binF + 3 * S);
    unsigned int* binC = reinterpret_cast<unsigned int*>(binX + 4 * S);
    NkSvHot h = nk_load_hot(P, binC + S + (S & 1));
    NK_TRACE_MARK_FIRST(P, 0);
    for (int i = threadIdx.x; i < S; i += blockDim.x) { binE[i] = 0; binC[i] = 0u; for (int k = 0; k < 3; ++k) binF[3 * i + k] = 0; for (int k = 0; k < 4; ++k) binX[4 * i + k] = 0.0; }
    __syncthreads();

    if (P.scan_emit) nk_emit_scan(P);
    const long long n = min((long long)P.dyn->n_slots, P.slot_hi);
    const unsigned int lane = threadIdx.x & 31u;
    for (long long wbase = P.slot_lo + 2 * ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)); wbase < n;
         wbase += 2 * (long long)gridDim.x * blockDim.x) {
        const long long base = wbase + 2 * lane;
        const bool inb = base < n;
        double2 X = make_double2(0, 0), Y = X, Z = X, TC = X, OC = X;
        int2 MD = make_int2(-1, -1), OM = MD;
        if (inb) {
            X = *reinterpret_cast<const double2*>(P.px + base);
            Y = *reinterpret_cast<const double2*>(P.py + base);
            Z = *reinterpret_cast<const double2*>(P.pz + base);
            TC = *reinterpret_cast<const double2*>(P.tc + base);
            OC = *reinterpret_cast<const double2*>(P.occ + base);
            MD = *reinterpret_cast<const int2*>(P.mode + base);
            OM = MD;
            if (HAS_ROUGH) OM = *reinterpret_cast<const int2*>(P.omode + base);
        }
        bool h0 = false, h1 = false;
        if (base < n && MD.x >= 0) h0 = nk_step_particle_tab<HAS_ROUGH, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.x, OM.x, X.x, Y.x, Z.x, TC.x, OC.x);
        if (base + 1 < n && MD.y >= 0) h1 = nk_step_particle_tab<HAS_ROUGH, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.y, OM.y, X.y, Y.y, Z.y, TC.y, OC.y);
        nk_push_hits(P, lane, h0, h1, base);
        if (inb) {
            *reinterpret_cast<double2*>(P.px + base) = X;
            *reinterpret_cast<double2*>(P.py + base) = Y;
            *reinterpret_cast<double2*>(P.pz + base) = Z;
            *reinterpret_cast<double2*>(P.tc + base) = TC;
            *reinterpret_cast<double2*>(P.occ + base) = OC;
        }
    }
    __syncthreads();
    nk_flush_bins<FLUX>(P, binE, binF, binX, binC);
    NK_TRACE_MARK_MAX(P, 1);
}

// ---- variant A1: one particle per thread, 64-bit accesses (fewer live registers -> more resident warps) ------
#ifndef NK_STEP1_MIN_BLOCKS
#define NK_STEP1_MIN_BLOCKS 5
#endif
template <bool HAS_ROUGH, bool FAST, bool RELAX, bool FLUX>
__global__ void __launch_bounds__(NK_STEP_THREADS, NK_STEP1_MIN_BLOCKS) k_step1(NkP P) {
    extern __shared__ double sm[];
    NkSvSmem s = nk_load_sv(P, sm);
    const int S = P.S;
    long long* binE = reinterpret_cast<long long*>(sm + nk_sv_smem_doubles(S));
    long long* binF = binE + S;
    double* binX = reinterpret_cast<double*>(binF + 3 * S);
    unsigned int* binC = reinterpret_cast<unsigned int*>(binX + 4 * S);
    NkSvHot h = nk_load_hot(P, binC + S + (S & 1));
    for (int i = threadIdx.x; i < S; i += blockDim.x) { binE[i] = 0; binC[i] = 0u; for (int k = 0; k < 3; ++k) binF[3 * i + k] = 0; for (int k = 0; k < 4; ++k) binX[4 * i + k] = 0.0; }
    __syncthreads();
    nk_emit_scan(P);
    const long long n = P.dyn->n_slots;
    const unsigned int lane = threadIdx.x & 31u;
    for (long long wbase = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wbase < n;
         wbase += (long long)gridDim.x * blockDim.x) {
        const long long i = wbase + lane;
        bool hit = false;
        if (i < n) {
            const int md = P.mode[i];
            if (md >= 0) {
                const int om = HAS_ROUGH ? P.omode[i] : md;
                double x = P.px[i], y = P.py[i], z = P.pz[i], tc = P.tc[i], occ = P.occ[i];
                hit = nk_step_particle<HAS_ROUGH, FAST, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, md, om, x, y, z, tc, occ);
                P.px[i] = x; P.py[i] = y; P.pz[i] = z; P.tc[i] = tc; P.occ[i] = occ;
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            unsigned int pos = 0;
            if (lane == 0) pos = atomicAdd(&P.dyn->n_hits, __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (hit) P.hitlist[pos + __popc(m & ((1u << lane) - 1u))] = (int)i;
        }
    }
    __syncthreads();
    nk_flush_bins<FLUX>(P, binE, binF, binX, binC);
}

// ---- variant B: TMA bulk-copy pipeline -------------------------------------------------------------------------
// The particle SoA is streamed through shared memory in tiles of NK_TILE slots by the bulk async-copy engine
// (cp.async.bulk, SASS UBLKCP) with an mbarrier per stage: NK_STAGES tiles are in flight per block regardless
// of register pressure, consumer warps read/write the tile in shared memory, and the updated tile goes back
// with a bulk store.  Needs capacity % NK_TILE == 0 (slots past n_slots are dead: mode = -1).
#define NK_TILE 512
#define NK_STAGES 3
__device__ __forceinline__ void nk_mbar_init(void* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nk_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void nk_mbar_expect_tx(void* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nk_mbar_wait(void* bar, unsigned int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "NK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NK_DONE_%=;\n\t"
        "bra NK_WAIT_%=;\n\t"
        "NK_DONE_%=:\n\t}" ::"r"(nk_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void nk_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(nk_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(nk_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nk_bulk_s2g(void* dst_gmem, const void* src_smem, unsigned int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(nk_smem_u32(src_smem)), "r"(bytes) : "memory");
}

struct NkTileSmem {                         // one pipeline stage (24.5 KB)
    double x[NK_TILE], y[NK_TILE], z[NK_TILE], tc[NK_TILE], occ[NK_TILE];
    int mode[NK_TILE], omode[NK_TILE];
};

template <bool HAS_ROUGH, bool FAST, bool RELAX, bool FLUX>
__global__ void __launch_bounds__(NK_STEP_THREADS, 2) k_step_tma(NkP P) {
    extern __shared__ __align__(128) unsigned char smraw[];
    NkTileSmem* stage = reinterpret_cast<NkTileSmem*>(smraw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smraw + NK_STAGES * sizeof(NkTileSmem));
    double* sm = reinterpret_cast<double*>(full + NK_STAGES + 1);
    NkSvSmem s = nk_load_sv(P, sm);
    const int S = P.S;
    long long* binE = reinterpret_cast<long long*>(sm + nk_sv_smem_doubles(S));
    long long* binF = binE + S;
    double* binX = reinterpret_cast<double*>(binF + 3 * S);
    unsigned int* binC = reinterpret_cast<unsigned int*>(binX + 4 * S);
    NkSvHot h = nk_load_hot(P, binC + S + (S & 1));
    for (int i = threadIdx.x; i < S; i += blockDim.x) { binE[i] = 0; binC[i] = 0u; for (int k = 0; k < 3; ++k) binF[3 * i + k] = 0; for (int k = 0; k < 4; ++k) binX[4 * i + k] = 0.0; }

    nk_emit_scan(P);
    const long long n = P.dyn->n_slots;
    const long long n_tiles = (n + NK_TILE - 1) / NK_TILE;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int tile_bytes = NK_TILE * (5 * 8 + (HAS_ROUGH ? 8 : 4));

    auto issue_load = [&](int st, long long tile) {
        NkTileSmem& T = stage[st];
        const long long o = tile * NK_TILE;
        nk_mbar_expect_tx(&full[st], tile_bytes);
        nk_bulk_g2s(T.x, P.px + o, NK_TILE * 8, &full[st]);
        nk_bulk_g2s(T.y, P.py + o, NK_TILE * 8, &full[st]);
        nk_bulk_g2s(T.z, P.pz + o, NK_TILE * 8, &full[st]);
        nk_bulk_g2s(T.tc, P.tc + o, NK_TILE * 8, &full[st]);
        nk_bulk_g2s(T.occ, P.occ + o, NK_TILE * 8, &full[st]);
        nk_bulk_g2s(T.mode, P.mode + o, NK_TILE * 4, &full[st]);
        if (HAS_ROUGH) nk_bulk_g2s(T.omode, P.omode + o, NK_TILE * 4, &full[st]);
    };

    if (threadIdx.x == 0) {
        for (int st = 0; st < NK_STAGES; ++st) nk_mbar_init(&full[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int st = 0; st < NK_STAGES; ++st) {
            const long long tile = (long long)blockIdx.x + (long long)st * gridDim.x;
            if (tile < n_tiles) issue_load(st, tile);
        }
    }

    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int st = (int)(it % NK_STAGES);
        const unsigned int parity = (unsigned int)((it / NK_STAGES) & 1);
        nk_mbar_wait(&full[st], parity);
        NkTileSmem& T = stage[st];
        const int j = 2 * threadIdx.x;
        double2 X = *reinterpret_cast<double2*>(T.x + j), Y = *reinterpret_cast<double2*>(T.y + j), Z = *reinterpret_cast<double2*>(T.z + j);
        double2 TC = *reinterpret_cast<double2*>(T.tc + j), OC = *reinterpret_cast<double2*>(T.occ + j);
        int2 MD = *reinterpret_cast<int2*>(T.mode + j), OM = MD;
        if (HAS_ROUGH) OM = *reinterpret_cast<int2*>(T.omode + j);
        const long long base = tile * NK_TILE + j;
        bool h0 = false, h1 = false;
        if (base < n && MD.x >= 0) h0 = nk_step_particle<HAS_ROUGH, FAST, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.x, OM.x, X.x, Y.x, Z.x, TC.x, OC.x);
        if (base + 1 < n && MD.y >= 0) h1 = nk_step_particle<HAS_ROUGH, FAST, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.y, OM.y, X.y, Y.y, Z.y, TC.y, OC.y);
        nk_push_hits(P, lane, h0, h1, base);
        *reinterpret_cast<double2*>(T.x + j) = X; *reinterpret_cast<double2*>(T.y + j) = Y; *reinterpret_cast<double2*>(T.z + j) = Z;
        *reinterpret_cast<double2*>(T.tc + j) = TC; *reinterpret_cast<double2*>(T.occ + j) = OC;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long o = tile * NK_TILE;
            nk_bulk_s2g(P.px + o, T.x, NK_TILE * 8);
            nk_bulk_s2g(P.py + o, T.y, NK_TILE * 8);
            nk_bulk_s2g(P.pz + o, T.z, NK_TILE * 8);
            nk_bulk_s2g(P.tc + o, T.tc, NK_TILE * 8);
            nk_bulk_s2g(P.occ + o, T.occ, NK_TILE * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the stage written back one iteration ago has been read by now: refill it
            if (it >= 1) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                const long long nt = tile + (long long)(NK_STAGES - 1) * gridDim.x;
                if (nt < n_tiles) issue_load((int)((it - 1) % NK_STAGES), nt);
            }
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    nk_flush_bins<FLUX>(P, binE, binF, binX, binC);
}

typedef void (*nk_step_fn)(NkP);
template <int V, bool A, bool B, bool C, bool D>
static nk_step_fn nk_pick5() {
    return V == 1 ? (nk_step_fn)k_step_tma<A, B, C, D> : (V == 2 ? (nk_step_fn)k_step1<A, B, C, D> : (V == 3 ? (nk_step_fn)k_step_pf<A, B, C, D> : (V == 4 ? (nk_step_fn)k_step_tab<A, B, C, D> : (nk_step_fn)k_step<A, B, C, D>)));
}
template <int V, bool A, bool B, bool C>
static nk_step_fn nk_pick4(bool d) { return d ? nk_pick5<V, A, B, C, true>() : nk_pick5<V, A, B, C, false>(); }
template <int V, bool A, bool B>
static nk_step_fn nk_pick3(bool c, bool d) { return c ? nk_pick4<V, A, B, true>(d) : nk_pick4<V, A, B, false>(d); }
template <int V, bool A>
static nk_step_fn nk_pick2(bool b, bool c, bool d) { return b ? nk_pick3<V, A, true>(c, d) : nk_pick3<V, A, false>(c, d); }
template <int V>
static nk_step_fn nk_pick1(bool a, bool b, bool c, bool d) { return a ? nk_pick2<V, true>(b, c, d) : nk_pick2<V, false>(b, c, d); }
static nk_step_fn nk_pick_step(int variant, bool rough, bool fast, bool relax, bool flux) {
    return variant == 1 ? nk_pick1<1>(rough, fast, relax, flux) : (variant == 2 ? nk_pick1<2>(rough, fast, relax, flux) : (variant == 3 ? nk_pick1<3>(rough, fast, relax, flux) : (variant == 4 ? nk_pick1<4>(rough, fast, relax, flux) : nk_pick1<0>(rough, fast, relax, flux))));
}

// ---- helpers shared by the rare-path code ------------------------------------------------------------------
__device__ __forceinline__ void nk_store_particle(const NkP& P, long long i, const NkParticle& p) {
    P.px[i] = p.x; P.py[i] = p.y; P.pz[i] = p.z; P.tc[i] = p.tc; P.occ[i] = p.occ;
    P.mode[i] = p.mode; P.omode[i] = p.omode; P.cfacet[i] = p.cf; P.cx[i] = p.cx; P.cy[i] = p.cy; P.cz[i] = p.cz;
}
// refresh_temperatures contribution of one particle handled outside k_step; `acc` is the block-private
// (shared memory) copy of the accumulator vector
__device__ __forceinline__ void nk_accumulate(const NkP& P, double* acc, const NkParticle& p, bool with_flux) {
    int sv = nk_classify(P, P.svc, P.sv_mid, p.x, p.y, p.z);
    double e = nk_mul(nk_mul(P.hbar, p.omega), nk_sub(p.occ, nk_bose(P, P.T_sv[sv], p.omega)));
    NK_RACC_E(P, acc, NK_ACC_E(P.S, P.R) + sv, e);
    NK_RACC_N(P, acc, NK_ACC_CNT(P.S, P.R) + sv);
    if (with_flux) {
        NK_RACC_F(P, acc, NK_ACC_FLUX(P.S, P.R) + 3 * sv, nk_mul(p.vx, e));
        NK_RACC_F(P, acc, NK_ACC_FLUX(P.S, P.R) + 3 * sv + 1, nk_mul(p.vy, e));
        NK_RACC_F(P, acc, NK_ACC_FLUX(P.S, P.R) + 3 * sv + 2, nk_mul(p.vz, e));
    }
}

// Free slots live in a ring: absorbed particles push at `fr_tail`, emission pops at `fr_head` but only
// entries pushed in EARLIER steps (below `fr_snap`, advanced by the finalize), so that pushes and pops
// of the same launch never touch the same entry.
__device__ __forceinline__ void nk_kill(const NkP& P, double* acc, long long i) {
    P.mode[i] = -1;
    unsigned long long k = nk_agg_inc((unsigned long long*)&P.dyn->fr_tail);
    P.freelist[k % (unsigned long long)P.cap] = (int)i;
    NK_RACC_N(P, acc, NK_ACC_NABS(P.S, P.R));
}
__device__ __forceinline__ long long nk_take_slot(const NkP& P) {
    // claims beyond fr_snap are not returned: the finalize clamps fr_head back to fr_snap
    long long old = (long long)nk_agg_inc((unsigned long long*)&P.dyn->fr_head);
    if (old < P.dyn->fr_snap) return P.freelist[old % P.cap];
    long long slot = (long long)nk_agg_inc((unsigned long long*)&P.dyn->n_slots);      // nothing recyclable: append
    if (slot >= P.cap) {
        atomicAdd((unsigned long long*)&P.dyn->n_slots, (unsigned long long)(-1LL));
        atomicOr(&P.dyn->error, NK_ERR_CAPACITY);
        return -1;
    }
    return slot;
}

// One emission-list entry: n_new copies of mode m entering through reservoir r (Population.py:385-406,
// :491-508, add_reservoir_particles :525-552, Mesh.sample_surface Mesh.py:923-951).
// One new particle of reservoir r in mode m entering the domain dt_in before the end of the step
// (Population.fill_reservoirs :491-508 + add_reservoir_particles :525-552 + Mesh.sample_surface :923-951).
__device__ __forceinline__ void nk_emit_particle(const NkP& P, const NkGeo& G, double* acc, int r, int m, long long id, double dt_in,
                                                 double uface, double us, double ur, long long step, bool with_flux) {
    const double dt = P.dt;
    const NkMode mp = P.mprop[m];
    NkParticle p;
    p.id = id;
    // face ~ area: searchsorted(cdf, u, side='right') as np.random.choice does
    const int f0 = P.res_face_ptr[r], f1 = P.res_face_ptr[r + 1];
    int lo = f0, hi = f1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (P.res_face_cdf[mid] <= uface) lo = mid + 1; else hi = mid; }
    const int face = P.res_faces[min(lo, f1 - 1)];
    const double* V = P.face_vertices + 9 * (size_t)face;
    const double rs = sqrt(us);
    const double a0 = nk_sub(1.0, rs), a1 = nk_mul(nk_sub(1.0, ur), rs), a2 = nk_mul(ur, rs);
    const double x0 = nk_add(nk_add(nk_mul(a0, V[0]), nk_mul(a1, V[3])), nk_mul(a2, V[6]));
    const double y0 = nk_add(nk_add(nk_mul(a0, V[1]), nk_mul(a1, V[4])), nk_mul(a2, V[7]));
    const double z0 = nk_add(nk_add(nk_mul(a0, V[2]), nk_mul(a1, V[5])), nk_mul(a2, V[8]));
    p.mode = m; p.omode = m; p.omega = mp.omega; p.vx = mp.vx; p.vy = mp.vy; p.vz = mp.vz;
    double t;
    nk_find_boundary_1(P, G.faces, x0, y0, z0, p.vx, p.vy, p.vz, p.cx, p.cy, p.cz, t, p.cf);
    p.tc = nk_sub(nk_div(t, dt), nk_div(dt_in, dt));
    p.x = nk_add(x0, nk_mul(p.vx, dt_in)); p.y = nk_add(y0, nk_mul(p.vy, dt_in)); p.z = nk_add(z0, nk_mul(p.vz, dt_in));
    p.occ = nk_bose(P, P.res_T[r], p.omega);
    p.alive = true;
    NK_RACC_N(P, acc, NK_ACC_NEMIT(P.S, P.R));
    if (p.tc < 0.0) nk_boundary_events(P, G, p, step, acc);
    if (!p.alive) { NK_RACC_N(P, acc, NK_ACC_NABS(P.S, P.R)); return; }   // crossed the whole domain within the step
    const long long slot = nk_take_slot(P);
    if (slot < 0) return;
    nk_store_particle(P, slot, p);
    P.pid[slot] = p.id;
    {
        const unsigned int k = nk_agg_inc(&P.dyn->n_new);
        if ((long long)k < P.newslots_cap) P.newslots[k] = (int)slot;
    }
    nk_accumulate(P, acc, p, with_flux);
}

// One emission-list entry (constant / fixed_rate): n_new copies of mode m from reservoir r.
__device__ __forceinline__ void nk_emit_entry(const NkP& P, const NkGeo& G, double* acc, int r, int m, int n_new, long long step, bool with_flux) {
    const double dt = P.dt;
    const size_t idx = (size_t)r * P.M + m;
    const double prob = P.enter_prob[idx];
    // numerator of the first copy's entry time: the counter after this step's update, or this step's dice
    const double lead = P.res_gen == NK_RESGEN_FIXED_RATE ? P.emit_u[idx] : P.res_counter[idx];
    for (int c = n_new; c >= 1; --c) {
        const long long id = NK_EMIT_ID_BASE + (((step * P.R + r) * (long long)P.M + m) * NK_EMIT_CMAX + (c - 1));
        double ua, uface, us, ur;
        nk_uniforms(P, id, step, NK_STREAM_EMIT_A, ua, uface);
        nk_uniforms(P, id, step, NK_STREAM_EMIT_B, us, ur);
        const double dt_in = (c == 1) ? nk_mul(dt, nk_sub(1.0, nk_div(lead, prob)))
                                      : nk_mul(dt, nk_sub(1.0, nk_div(nk_add((double)(c - 1), ua), prob)));
        nk_emit_particle(P, G, acc, r, m, id, dt_in, uface, us, ur, step, with_flux);
    }
}

// One re-emitted particle of the one_to_one mode: k-th particle of reservoir r (Population.py:457-489).
__device__ __forceinline__ void nk_emit_one_to_one(const NkP& P, const NkGeo& G, double* acc, long long e, long long step, bool with_flux) {
    int r = 0;
    for (; r < P.R; ++r) {
        const long long share = nk_one_to_one_share(P, r);
        if (e < share) break;
        e -= share;
    }
    if (r >= P.R) return;
    const long long k = P.rank + e * P.world;
    const long long id = NK_EMIT_ID_BASE + (step * P.R + r) * ((long long)P.M * NK_EMIT_CMAX) + k;
    double ua, uface, us, ur, umode, udt;
    nk_uniforms(P, id, step, NK_STREAM_EMIT_A, ua, uface);
    nk_uniforms(P, id, step, NK_STREAM_EMIT_B, us, ur);
    nk_uniforms(P, id, step, NK_STREAM_EMIT_C, umode, udt);
    const double* rou = P.res_roulette + (size_t)r * P.M;
    int lo = 0, hi = P.M;                                    // searchsorted left
    while (lo < hi) { int mid = (lo + hi) >> 1; if (rou[mid] < umode) lo = mid + 1; else hi = mid; }
    nk_emit_particle(P, G, acc, r, min(lo, P.M - 1), id, nk_mul(P.dt, udt), uface, us, ur, step, with_flux);
}

// One hit-list entry: the boundary event loop of an existing particle.
__device__ __forceinline__ void nk_hit_entry(const NkP& P, const NkGeo& G, double* acc, long long i, long long step, bool with_flux) {
    NkParticle p;
    p.x = P.px[i]; p.y = P.py[i]; p.z = P.pz[i]; p.tc = P.tc[i]; p.occ = P.occ[i];
    p.mode = P.mode[i]; p.omode = P.omode[i];
    const NkMode m = P.mprop[p.mode];
    p.vx = m.vx; p.vy = m.vy; p.vz = m.vz;
    p.omega = (p.omode == p.mode) ? m.omega : P.mprop[p.omode].omega;
    p.cf = P.cfacet[i]; p.cx = P.cx[i]; p.cy = P.cy[i]; p.cz = P.cz[i];
    p.id = P.pid[i]; p.alive = true;
    nk_boundary_events(P, G, p, step, acc);
    if (p.alive) {
        nk_store_particle(P, i, p);
        nk_accumulate(P, acc, p, with_flux);
    } else {
        nk_kill(P, acc, i);
    }
}

// coef = rbf_w . T (T may live in shared memory); all threads of the block take rows
__device__ __forceinline__ void nk_rbf_refresh(const NkP& P, const double* T) {
    const int rows = P.S + P.rbf_nd + 1;
    for (int j = threadIdx.x; j < rows; j += blockDim.x) {
        const double* w = P.rbf_w + (size_t)j * P.S;
        double a = 0.0;
        for (int s = 0; s < P.S; ++s) a += w[s] * T[s];
        P.rbf_coef[j] = a;
    }
}

// ---- close the step: calculate_energy normalisation, temperature_function, heat flux, kappa,
//      reservoir balances (Population.py:704-728, :692, :730-788, :1685-1699).  One block. -----------------------
__device__ void nk_finalize_block(const NkP& P, double* sm) {
    const int S = P.S, R = P.R;
    double* sT = sm;             // new T_sv
    double* sPhi = sm + S;       // flux along the slice axis
    double* sN = sm + 2 * S;     // counts
    double* acc = P.acc; double* out = P.out;
    const long long step_done = P.dyn->step + 1;
    const bool conv = (step_done % P.n_dt_to_conv) == 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        double cnt = __ldcg(acc + NK_ACC_CNT(S, R) + s);
        double esum = __ldcg(acc + NK_ACC_E(S, R) + s);
        double norm;
        if (P.norm_mean) { norm = nk_div(P.n_active, cnt); if (norm != norm) norm = 0.0; }
        else norm = nk_div(P.n_active, nk_mul(P.particle_density, P.sv_volume[s]));
        double Tprev = P.T_sv[s];
        // both tables share the index of the (uniform) temperature grid, and T moves little per step: start the bracket
        // searches at the previous temperature's index
        const int ig = P.nE > 1 ? (int)((Tprev - P.Ta[0]) * P.Ta_inv_d) : 0;
        double ref = nk_interp_table_from(P.Ta, P.Ea, P.nE, Tprev, P.Ea[0], P.Ea[P.nE - 1], ig);
        double E = nk_add(nk_div(nk_mul(esum, norm), P.dens_norm), ref);
        double Tn = nk_interp_table_from(P.Ea, P.Ta, P.nE, E, P.Ta[0], P.Ta[P.nE - 1], ig);
        sT[s] = Tn; sN[s] = cnt;
        out[NK_OUT_T(S, R) + s] = Tn;
        out[NK_OUT_E(S, R) + s] = E;
        out[NK_OUT_N(S, R) + s] = cnt;
        if (conv) {
            double f[3];
            for (int k = 0; k < 3; ++k) {
                f[k] = nk_mul(nk_div(nk_mul(__ldcg(acc + NK_ACC_FLUX(S, R) + 3 * s + k), norm), P.dens_norm), P.eVpsa2_in_Wm2);
                out[NK_OUT_FLUX(S, R) + 3 * s + k] = f[k];
            }
            sPhi[s] = f[P.axis];
        }
    }
    __syncthreads();
    // reservoirs: accumulate this step, normalise on convergence steps
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        out[NK_OUT_NLEAVE(S, R) + r] = __ldcg(acc + NK_ACC_NLEAVE(S, R) + r);
        P.res_nleave[r] = __ldcg(acc + NK_ACC_NLEAVE(S, R) + r);
        double eb = nk_add(P.res_acc[r], __ldcg(acc + NK_ACC_EBAL(S, R) + r));
        double fx[3];
        for (int k = 0; k < 3; ++k) fx[k] = nk_add(P.res_acc[R + 3 * r + k], __ldcg(acc + NK_ACC_RFLUX(S, R) + 3 * r + k));
        if (conv) {
            double area = P.facet_area[P.res_facet[r]];
            double den = nk_mul(nk_mul(nk_mul(P.particle_density, P.dt), (double)P.n_dt_to_conv), area);
            double cf = nk_div(P.n_active, den);
            for (int k = 0; k < 3; ++k) out[NK_OUT_RFLUX(S, R) + 3 * r + k] = nk_mul(nk_div(nk_mul(fx[k], cf), P.dens_norm), P.eVpsa2_in_Wm2);
            double ce = nk_div(P.n_active, nk_mul(nk_mul(P.particle_density, P.dt), (double)P.n_dt_to_conv));
            out[NK_OUT_REBAL(S, R) + r] = nk_div(nk_mul(eb, ce), P.dens_norm);
            eb = 0.0; fx[0] = fx[1] = fx[2] = 0.0;
        }
        P.res_acc[r] = eb;
        for (int k = 0; k < 3; ++k) P.res_acc[R + 3 * r + k] = fx[k];
    }
    if (threadIdx.x == 0) {
        double np = 0.0, et = 0.0;
        for (int s = 0; s < S; ++s) { np += sN[s]; et += __ldcg(acc + NK_ACC_E(S, R) + s); }
        out[NK_OUT_NP(S, R)] = np;
        out[NK_OUT_ETOT(S, R)] = et;
        if (conv && P.is_slice && R == 2) {
            // calculate_kappa, slice subvolumes (Population.py:750-771)
            double L = nk_sub(P.bhi[P.axis], P.blo[P.axis]);
            double dx = nk_div(nk_mul(nk_mul(2.0, L), P.a_in_m), (double)S);
            double DX = nk_div(nk_mul(nk_mul(L, P.a_in_m), (double)(1 + S)), (double)S);
            double T0 = P.res_T[0], T1 = P.res_T[1];
            double sum = 0.0;
            for (int s = 0; s < S; ++s) {
                double Tm = s == 0 ? T0 : sT[s - 1];
                double Tp = s == S - 1 ? T1 : sT[s + 1];
                double k = nk_div(nk_mul(-sPhi[s], dx), nk_sub(Tp, Tm));
                if (isinf(k)) k = 0.0;
                out[NK_OUT_KSV(S, R) + s] = k;
                sum += nk_mul(sPhi[s], sN[s]);
            }
            out[NK_OUT_KAPPA(S, R)] = nk_div(nk_mul(-sum, nk_div(DX, nk_sub(T1, T0))), np);
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) P.T_sv[s] = sT[s];
    if (P.interp == NK_INTERP_RADIAL) nk_rbf_refresh(P, sT);
    for (int i = threadIdx.x; i < nk_acc_len(S, R); i += blockDim.x) acc[i] = 0.0;
    if (threadIdx.x == 0) {
        NkDyn* d = P.dyn;
        d->step = step_done;
        d->relax_pending = 1;
        d->last_hits = d->n_hits; d->last_new = d->n_new;
        d->n_hits = 0;
        d->n_emit = 0;
        d->n_new = 0;
        if (d->fr_head > d->fr_snap) d->fr_head = d->fr_snap;      // over-claims of an exhausted free list (nk_take_slot)
        d->fr_snap = d->fr_tail;           // slots freed in this step become recyclable from the next one
        d->blocks_done = 0;
    }
}

// All-reduce (sum) of the accumulator vector across the ranks of one box, done by the block that closes the
// step: every rank stores its vector straight into every peer's mailbox over NVLink (peer-mapped memory),
// publishes a sequence number, waits for the peers' numbers and adds the world's vectors in rank order, so
// all ranks get bit-identical sums without a separate collective launch.  Two mailbox parities: a rank can be
// at most one step ahead of the slowest one.  The wait is bounded (~20 s): a missing peer raises NK_ERR_COMM
// instead of hanging the GPU.
__device__ void nk_exchange_sums(const NkP& P) {
    const int len = nk_acc_len(P.S, P.R);
    const int W = P.world;
    const unsigned long long seq = (unsigned long long)(P.dyn->step + 1);
    const int par = (int)(seq & 1ull);
    for (int r = 0; r < W; ++r) {
        double* dst = P.peer_mbox[r] + ((size_t)par * W + P.rank) * len;
        for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = __ldcg(P.acc + i);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < W) {
        volatile unsigned long long* f = P.peer_flags[threadIdx.x] + (size_t)par * W + P.rank;
        *f = seq;
        __threadfence_system();
        volatile unsigned long long* mine = P.flags_local + (size_t)par * W + threadIdx.x;
        const long long t0 = clock64();
        while (*mine != seq) {
            if (clock64() - t0 > 40000000000LL) { atomicOr(&P.dyn->error, NK_ERR_COMM); break; }   // ~20 s
        }
    }
    __syncthreads();
    __threadfence_system();
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        double sum = 0.0;
        for (int r = 0; r < W; ++r) sum += *((volatile double*)(P.mbox_local + ((size_t)par * W + r) * len + i));
        P.acc[i] = sum;
    }
    __threadfence();
    __syncthreads();
}

__global__ void __launch_bounds__(1024) k_finalize(NkP P) {
    extern __shared__ double sm[];
    nk_finalize_block(P, sm);
}

// ---- the rare path of a step: boundary events of the hit list + reservoir emission -----------------------------
// Work items [0, n_hits) are existing particles whose collision falls inside the step, [n_hits, n_hits + n_emit)
// are emission-list entries.  One thread per item; triangles staged in shared memory when they fit.  With
// FUSE the last block to finish closes the step (single-GPU path: no collective between the two halves).
#define NK_RARE_THREADS 128
#define NK_RARE_FACES 128
#define NK_RARE_FACETS 64
template <bool FUSE>
#ifndef NK_RARE_MIN_BLOCKS
#define NK_RARE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(NK_RARE_THREADS, NK_RARE_MIN_BLOCKS) k_rare(NkP P) {
    __shared__ NkFace sfaces[NK_RARE_FACES];
    __shared__ int sfi[4 * NK_RARE_FACETS];
    __shared__ double sfd[6 * NK_RARE_FACETS];
    extern __shared__ double sm_fin[];
    __shared__ int s_last;
    NkGeo G;
    G.faces = P.faces; G.bc = P.facet_bc; G.partner = P.facet_partner; G.res = P.facet_res; G.rough = P.facet_rough;
    G.normal = P.facet_normal; G.centroid = P.facet_centroid;
    NK_TRACE_MARK_FIRST(P, 2);
    // blocks beyond the work list (most of them when few particles hit a wall) go straight to the closing protocol
    const bool has_work = (unsigned long long)blockIdx.x * blockDim.x < (unsigned long long)P.dyn->n_hits + P.dyn->n_emit;
    if (has_work) {
        if (P.F <= NK_RARE_FACES) {
            const double* src = reinterpret_cast<const double*>(P.faces);
            double* dst = reinterpret_cast<double*>(sfaces);
            for (int k = threadIdx.x; k < P.F * (int)(sizeof(NkFace) / 8); k += blockDim.x) dst[k] = src[k];
            G.faces = sfaces;
        }
        if (P.nf <= NK_RARE_FACETS) {
            for (int k = threadIdx.x; k < P.nf; k += blockDim.x) {
                sfi[k] = P.facet_bc[k]; sfi[NK_RARE_FACETS + k] = P.facet_partner[k];
                sfi[2 * NK_RARE_FACETS + k] = P.facet_res[k]; sfi[3 * NK_RARE_FACETS + k] = P.facet_rough[k];
            }
            for (int k = threadIdx.x; k < 3 * P.nf; k += blockDim.x) { sfd[k] = P.facet_normal[k]; sfd[3 * NK_RARE_FACETS + k] = P.facet_centroid[k]; }
            G.bc = sfi; G.partner = sfi + NK_RARE_FACETS; G.res = sfi + 2 * NK_RARE_FACETS; G.rough = sfi + 3 * NK_RARE_FACETS;
            G.normal = sfd; G.centroid = sfd + 3 * NK_RARE_FACETS;
        }
        // block-private accumulators: thousands of items would otherwise hammer the same ~40 global addresses
        double* racc = sm_fin + 3 * P.S;
        const int nacc = nk_acc_len(P.S, P.R);
        long long* rq = reinterpret_cast<long long*>(racc + nacc);      // fixed-point halves of the same entries (nk_racc_*)
        for (int k = threadIdx.x; k < nacc; k += blockDim.x) { racc[k] = 0.0; rq[k] = 0; }
        __syncthreads();
        const unsigned int nh = P.dyn->n_hits, ne = P.dyn->n_emit;
        const long long step = P.dyn->step;
        const bool with_flux = ((step + 1) % P.n_dt_to_conv) == 0;
        for (unsigned int w = blockIdx.x * blockDim.x + threadIdx.x; w < nh + ne; w += gridDim.x * blockDim.x) {
            if (w < nh) {
                nk_hit_entry(P, G, racc, P.hitlist[w], step, with_flux);
            } else {
                if (P.res_gen == NK_RESGEN_ONE_TO_ONE) {
                    nk_emit_one_to_one(P, G, racc, (long long)(w - nh), step, with_flux);
                } else {
                    const int2 e = P.emitlist[w - nh];
                    nk_emit_entry(P, G, racc, e.x >> 8, e.y, e.x & 0xff, step, with_flux);
                }
            }
        }
        __syncthreads();
        NK_TRACE_MARK_MAX(P, 3);
        for (int k = threadIdx.x; k < nacc; k += blockDim.x)
        {
            const double v = (double)rq[k] * nk_racc_inv_scale(P.S, P.R, k) + racc[k];
            if (v != 0.0) atomicAdd(P.acc + k, v);
        }
        if (threadIdx.x == 0) {
            // live count: + particles that got a slot (emitted - absorbed on arrival) - absorbed
            const double d = (double)(rq[NK_ACC_NEMIT(P.S, P.R)] - rq[NK_ACC_NABS(P.S, P.R)]);
            if (d != 0.0) atomicAdd((unsigned long long*)&P.dyn->n_alive, (unsigned long long)(long long)d);
        }
    }
    if (FUSE) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(&P.dyn->blocks_done, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (P.comm_on) nk_exchange_sums(P);
            if (P.trace && threadIdx.x == 0) P.trace[4] = nk_globaltimer();
            nk_finalize_block(P, sm_fin);
            __syncthreads();
            if (P.trace && threadIdx.x == 0) P.trace[5] = nk_globaltimer();
        }
    }
}

// apply the deferred lifetime_scattering so that `occ` is what the reference holds after run_timestep
// (same arithmetic as the head of k_step, so flushing between steps is bit-neutral)
template <bool FAST>
__global__ void __launch_bounds__(256) k_flush_relax(NkP P) {
    extern __shared__ double sm[];
    NkSvSmem s = nk_load_sv(P, sm);
    NkSvHot h = nk_load_hot(P, sm + nk_sv_smem_doubles(P.S));
    __syncthreads();
    if (!P.dyn->relax_pending) return;
    const long long n = P.dyn->n_slots;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int md = P.mode[i];
        if (md < 0) continue;
        const int om = P.has_rough ? P.omode[i] : md;      // as the streaming kernel: without rough facets omode == mode
        double4 ma, mt;
        nk_ld256(&P.mhot[md].omega, ma);
        nk_ld256(&P.mhot[md].t[0], mt);
        double omega = om == md ? ma.x : P.mhot[om].omega;
        double be0; int g0;
        P.occ[i] = nk_relax_particle<FAST>(P, s, h, P.px[i], P.py[i], P.pz[i], md, omega, nk_mul(P.hbar, omega), mt, P.occ[i], be0, g0);
    }
}
__global__ void k_clear_relax(NkP P) { P.dyn->relax_pending = 0; }

// Host-buffer pipeline: the slots the rare path touched in the step just closed (hit list + emitted slots) are
// packed into a small patch so that the host does not have to download the cold arrays of all particles again.
struct NkPatch {                  // structure of arrays, `cap` records each
    int *slot, *mode, *omode, *cfacet; long long* pid;
    double *x, *y, *z, *tc, *cx, *cy, *cz;
};
__global__ void __launch_bounds__(256) k_pack_dirty(NkP P, NkPatch out, long long cap, long long offset, unsigned int* count) {
    const unsigned int nh = P.dyn->last_hits, nn = P.dyn->last_new;
    const long long total = (long long)nh + nn;
    if (blockIdx.x == 0 && threadIdx.x == 0) { count[0] = nh; count[1] = nn; }
    if ((long long)nn > P.newslots_cap) return;              // the list of new slots is incomplete: the host decides
    const long long hi = min(total, offset + cap);           // this round packs entries [offset, hi) of hits ++ new slots
    for (long long i = offset + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        const int s = i < nh ? P.hitlist[i] : P.newslots[i - nh];
        const long long o = i - offset;
        out.slot[o] = s; out.mode[o] = P.mode[s]; out.omode[o] = P.omode[s]; out.cfacet[o] = P.cfacet[s]; out.pid[o] = P.pid[s];
        out.x[o] = P.px[s]; out.y[o] = P.py[s]; out.z[o] = P.pz[s]; out.tc[o] = P.tc[s];
        out.cx[o] = P.cx[s]; out.cy[o] = P.cy[s]; out.cz[o] = P.cz[s];
    }
}

// Population.contains_check (Population.py:1712-1722): live particles outside the bounding box +- tol
__global__ void __launch_bounds__(256) k_outside_slots(NkP P, double tol, int* out, long long cap, unsigned int* count) {
    const long long n = P.dyn->n_slots;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (P.mode[i] < 0) continue;
        const double x = P.px[i], y = P.py[i], z = P.pz[i];
        const bool outside = x < P.blo[0] - tol || y < P.blo[1] - tol || z < P.blo[2] - tol ||
                             x > P.bhi[0] + tol || y > P.bhi[1] + tol || z > P.bhi[2] + tol;
        if (outside) {
            const unsigned int k = nk_agg_inc(count);
            if ((long long)k < cap) out[k] = (int)i;
        }
    }
}

// Host-buffer pipeline, upload side: the streaming kernel never reads collision facet / position or the particle id,
// the rare path reads them only for particles whose collision falls inside the step (tc < 1 on entry).  The host
// finds those (a scan of `tc`), packs their cold fields and this kernel scatters them into the device arrays.
struct NkCold {                   // structure of arrays, `cap` records each
    int *slot, *cfacet, *omode; long long* pid; double *cx, *cy, *cz;
};
__global__ void __launch_bounds__(256) k_unpack_cold(NkP P, NkCold in, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = in.slot[i];
        P.cfacet[s] = in.cfacet[i]; P.omode[s] = in.omode[i]; P.pid[s] = in.pid[i];
        P.cx[s] = in.cx[i]; P.cy[s] = in.cy[i]; P.cz[s] = in.cz[i];
    }
}

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
    int step_variant = 0;
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
    ctx->P.trace = nullptr;
    if (const char* e = getenv("NK_TRACE")) {
        if (strcmp(e, "0") != 0 && cudaMalloc(&ctx->P.trace, 8 * sizeof(unsigned long long)) == cudaSuccess)
            cudaMemset(ctx->P.trace, 0, 8 * sizeof(unsigned long long));
    }
    if (const char* e = getenv("NK_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));   // experiment: 32 / 64 / 128
    if (const char* e = getenv("NK_STEP_TAB")) { ctx->use_tab = strcmp(e, "0") != 0; ctx->force_tab = !strcmp(e, "force"); }
    if (const char* e = getenv("NK_STEP_IMPL")) ctx->step_variant = !strcmp(e, "tma") ? 1 : (!strcmp(e, "ldg1") ? 2 : (!strcmp(e, "pf") ? 3 : 0));
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
    for (int f = 0; f < F; ++f) {
        NkFace& T = faces[f];
        T.nx = fn[3 * f]; T.ny = fn[3 * f + 1]; T.nz = fn[3 * f + 2]; T.k = fk[f];
        T.lox = flo[3 * f] - NK_TOL; T.loy = flo[3 * f + 1] - NK_TOL; T.loz = flo[3 * f + 2] - NK_TOL;
        T.hix = fhi[3 * f] + NK_TOL; T.hiy = fhi[3 * f + 1] + NK_TOL; T.hiz = fhi[3 * f + 2] + NK_TOL;
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
    P.F = F; P.nf = nf;
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
    P.Ta_inv_d = (nE > 1 && Ta[nE - 1] != Ta[0]) ? (nE - 1) / (Ta[nE - 1] - Ta[0]) : 0.0;
    ctx->h_mode.resize(4 * (size_t)M);
    for (int m = 0; m < M; ++m) { ctx->h_mode[4 * (size_t)m] = mp[m].omega; ctx->h_mode[4 * (size_t)m + 1] = mp[m].vx; ctx->h_mode[4 * (size_t)m + 2] = mp[m].vy; ctx->h_mode[4 * (size_t)m + 3] = mp[m].vz; }
    ctx->h_tau.assign(tau, tau + (size_t)NT * M);
    ctx->h_Tg.assign(Tg, Tg + NT);
    if (ctx->hot_hi == 0) { ctx->hot_lo = 295.0; ctx->hot_hi = 305.0; }
    P.emit_m_lo = 0; P.emit_m_hi = M;
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
    NK_UP(dd, double, (const double*)nullptr, (size_t)4 * std::max(P.R, 1)); P.res_acc = dd;
    NK_UP(dd, double, (const double*)nullptr, (size_t)nk_out_len(P.S, P.R)); P.out = dd;
    P.hot_tab = nullptr;
    if (P.is_slice && P.interp == NK_INTERP_NEAREST && ctx->use_tab) {
        double2* dt2 = nullptr;
        if (cudaMalloc(&dt2, (size_t)P.M * P.S * sizeof(double2)) == cudaSuccess) { ctx->owned.push_back(dt2); P.hot_tab = dt2; }
        else cudaGetLastError();          // not enough memory: the direct variant is used
    }
    ctx->tab_dirty = true; ctx->step_blocks = 0;
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
    int* di;
    NK_UP(di, int, (const int*)nullptr, (size_t)cap); P.hitlist = di;
    NK_UP(di, int, (const int*)nullptr, (size_t)cap); P.freelist = di;
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

// count live slots (mode >= 0) on the host side of a tiny kernel-free path: done with a reduction kernel
__global__ void k_count_alive(NkP P, unsigned long long* out) {
    const long long n = P.dyn->n_slots;
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += P.mode[i] >= 0;
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int nk_set_slot_count(nk_ctx* ctx, int64_t n_slots) {
    cudaSetDevice(ctx->device);
    if (!ctx->particles_bound) { ctx->err = "nk_bind_particles first"; return -1; }
    if (n_slots < 0 || n_slots > ctx->P.cap) { ctx->err = "n_slots out of range"; return -1; }
    NkDyn d; if (nk_read_dyn(ctx, &d)) return -1;
    ctx->h_slots_hint = n_slots;
    d.n_slots = n_slots; d.fr_head = d.fr_tail = d.fr_snap = 0; d.n_hits = 0; d.n_emit = 0; d.n_new = 0; d.last_hits = 0; d.last_new = 0; d.blocks_done = 0;
    if (nk_write_dyn(ctx, &d)) return -1;
    unsigned long long* dc; NK_CK(cudaMalloc(&dc, 8)); NK_CK(cudaMemset(dc, 0, 8));
    k_count_alive<<<ctx->n_sm * 4, 256, 0, ctx->stream>>>(ctx->P, dc);
    unsigned long long hc = 0;
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost));
    cudaFree(dc);
    if (nk_read_dyn(ctx, &d)) return -1;
    d.n_alive = (long long)hc;
    return nk_write_dyn(ctx, &d);
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
    k_find_boundary<<<nk_grid(n, 256, 0), 256, 0, ctx->stream>>>(ctx->P, n, x, v, xc, tc, fc);
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
    k_init_collisions<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(ctx->P);
    NK_CK(cudaGetLastError());
    return 0;
}

#define NK_TAB_MAX_ENTRIES (6LL << 20)      // 96 MB of {n0, decay} pairs
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
    // the step counter and the relaxation flag are mirrored on the host (every mutation goes through this
    // library), so the launch-uniform RELAX / FLUX variants can be chosen without a device read-back
    const bool relax = ctx->h_relax_pending;
    const bool flux = ((ctx->h_step + 1) % P.n_dt_to_conv) == 0;
    int variant = ctx->step_variant;                          // 0: 2 particles/thread LDG.128, 1: TMA pipeline, 2: 1 particle/thread
    if (phase == 2) variant = ctx->last_variant;
    else {
    if (variant == 1 && (P.cap % NK_TILE) != 0) variant = 0;
    // per-(mode, subvolume) tables pay off once there are a few particles per table entry
    // ... and while the table (16 B per entry, rebuilt every step) stays L2-sized: at S = 100 x 1.8e5 modes (286 MB) the
    // rebuild costs what the leaner inner loop saves (profiles/README.md)
    if (variant == 0 && ctx->use_tab && fast && P.hot_tab &&
        (ctx->force_tab || (ctx->h_slots_hint >= 2 * (long long)P.M * P.S && (long long)P.M * P.S <= NK_TAB_MAX_ENTRIES))) variant = 4;
    if (variant == 4 && ctx->tab_dirty) {
        k_mode_tables<<<ctx->n_sm * 8, 256, nk_hot_smem_bytes(P.S), ctx->stream>>>(P);
        NK_CK(cudaGetLastError());
        ctx->tab_dirty = false;
    }
    if (variant == 1) smem = NK_STAGES * sizeof(NkTileSmem) + (NK_STAGES + 1) * 8 + nk_step_smem(P);
    if (variant == 3) smem = 2 * sizeof(NkPfStage) + nk_step_smem(P);
    nk_step_fn kern = nk_pick_step(variant, ctx->has_rough, fast, relax, flux);
    if (!ctx->step_blocks || ctx->step_blocks_variant != variant) {
        int per_sm = 0;
        for (int r = 0; r < 2; ++r) for (int f = 0; f < 2; ++f) {       // every variant may need the opt-in shared memory size
            nk_step_fn k = nk_pick_step(variant, ctx->has_rough, fast, r, f);
            if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nk_pick_step(variant, ctx->has_rough, fast, true, true), NK_STEP_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        ctx->step_blocks = per_sm * ctx->n_sm;
        ctx->step_blocks_variant = variant;
    }
    nk_prof_mark(ctx);
    if (!plan) {
        kern<<<ctx->step_blocks, NK_STEP_THREADS, smem, ctx->stream>>>(P);
    } else {
        if (variant != 0 && variant != 4) { ctx->err = "chunked launch needs the direct or table step kernel"; return -1; }
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
    const size_t fin_smem = (3 * (size_t)P.S + 2 * (size_t)nk_acc_len(P.S, P.R)) * 8;
    // one item per thread; the number of items is known on the device only, so size the grid for ~5 % of the slots
    // (hits + emission are 0.2-2 % of the particles per step; more items are covered by the grid-stride loop)
    const long long want_blocks = (ctx->h_slots_hint / 20 + NK_RARE_THREADS - 1) / NK_RARE_THREADS;
    const int rare_blocks = (int)std::min<long long>((long long)ctx->n_sm * 16, std::max<long long>(ctx->n_sm, want_blocks));
    if (fuse_finalize) k_rare<true><<<rare_blocks, NK_RARE_THREADS, fin_smem, ctx->stream>>>(P);
    else k_rare<false><<<rare_blocks, NK_RARE_THREADS, fin_smem, ctx->stream>>>(P);
    NK_CK(cudaGetLastError());
    nk_prof_mark(ctx);
    if (fuse_finalize) {
        ctx->h_step += 1; ctx->h_relax_pending = true; ctx->tab_dirty = true;
        if (variant == 4) {            // next step's tables right away (T_sv is final once k_rare has finished)
            k_mode_tables<<<ctx->n_sm * 8, 256, nk_hot_smem_bytes(P.S), ctx->stream>>>(P);
            NK_CK(cudaGetLastError());
            ctx->tab_dirty = false;
        }
        nk_prof_mark(ctx);
    }
    ctx->last_variant = variant;
    return 0;
}

int nk_step_local(nk_ctx* ctx) { return nk_step_kernels(ctx, false); }

int nk_step_finalize(nk_ctx* ctx) {
    cudaSetDevice(ctx->device);
    const NkP& P = ctx->P;
    int threads = 32;
    while (threads < P.S && threads < 1024) threads <<= 1;
    k_finalize<<<1, threads, 3 * (size_t)P.S * 8, ctx->stream>>>(P);
    NK_CK(cudaGetLastError());
    ctx->h_step += 1; ctx->h_relax_pending = true; ctx->tab_dirty = true;
    if (ctx->last_variant == 4) {
        k_mode_tables<<<ctx->n_sm * 8, 256, nk_hot_smem_bytes(P.S), ctx->stream>>>(P);
        NK_CK(cudaGetLastError());
        ctx->tab_dirty = false;
    }
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
    if (ctx->P.is_slice && ctx->P.interp == NK_INTERP_NEAREST) k_flush_relax<true><<<ctx->n_sm * 8, 256, smem, ctx->stream>>>(ctx->P);
    else k_flush_relax<false><<<ctx->n_sm * 8, 256, smem, ctx->stream>>>(ctx->P);
    NK_CK(cudaGetLastError());
    k_clear_relax<<<1, 1, 0, ctx->stream>>>(ctx->P);
    NK_CK(cudaGetLastError());
    ctx->h_relax_pending = false;
    return 0;
}

int nk_get_results(nk_ctx* ctx, double* T_sv, double* E_sv, int64_t* N_sv, double* flux, double* kappa_sv, double* kappa,
                   double* res_E_bal, double* res_flux, int64_t* N_leaving, double* total_energy) {
    cudaSetDevice(ctx->device);
    const NkP& P = ctx->P;
    const int S = P.S, R = P.R;
    std::vector<double> h(nk_out_len(S, R));
    NK_CK(cudaStreamSynchronize(ctx->stream));
    NK_CK(cudaMemcpy(h.data(), P.out, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
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
    const bool pipe = ctx->use_pipeline && n_steps == 1 && (P.world == 1 || P.comm_on) && n_in >= (1 << 20) && !ctx->profiling &&
                      (ctx->step_variant == 0);
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
    if (P.M == 0) { ctx->err = "nk_set_phonon first"; return -1; }
    P.rank = rank; P.world = world;
    long long M = P.M;
    P.emit_m_lo = (int)(M * rank / world);
    P.emit_m_hi = (int)(M * (rank + 1) / world);
    return 0;
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
        const size_t bytes = flag_bytes + 2 * (size_t)P.world * nk_acc_len(P.S, P.R) * sizeof(double);
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
