// nk_device.cuh -- device functions of the particle loop.  All arithmetic is IEEE f64 with explicit
// round-to-nearest mul/add (no FMA contraction) wherever the reference's NumPy expression order is
// what decides an integer result (facet, subvolume, mode index).
#pragma once
#include "nk_types.cuh"
#include <math_constants.h>

#define NK_DEVI __device__ __forceinline__

NK_DEVI double nk_mul(double a, double b) { return __dmul_rn(a, b); }
NK_DEVI double nk_add(double a, double b) { return __dadd_rn(a, b); }
NK_DEVI double nk_sub(double a, double b) { return __dsub_rn(a, b); }
NK_DEVI double nk_div(double a, double b) { return __ddiv_rn(a, b); }
// np.sum over a length-3 axis is left-associated: (a+b)+c
NK_DEVI double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return nk_add(nk_add(nk_mul(ax, bx), nk_mul(ay, by)), nk_mul(az, bz));
}
NK_DEVI double norm3(double x, double y, double z) { return sqrt(dot3(x, y, z, x, y, z)); }

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

// Phonon.calculate_occupation for the rare path (absorbed / emitted / diffusely scattered particles and their energy
// terms): the same lean arithmetic as the streaming kernel (<= 2 ulp; never feeds an integer result) instead of libm's exp
// and two IEEE divisions, which sat on every item's dependency chain
__device__ __forceinline__ double nk_bose_lean(double hbar, double kb, double T, double omega) {
    return nk_bose_fast(hbar * omega, omega, T > 0.0 ? nk_rcp(T * kb) : 0.0);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 keyed by (particle id, step, stream); see oracle/philox.py for the contract
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long nk_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// trace slots: 0 k_step first block in, 1 k_step last block out, 2 k_rare first block in, 3 k_rare last item done,
//              4 finalize in, 5 finalize out
#define NK_TRACE_MARK_FIRST(P, slot) do { if ((P).trace && threadIdx.x == 0 && blockIdx.x == 0) (P).trace[slot] = nk_globaltimer(); } while (0)
#define NK_TRACE_MARK_MAX(P, slot)   do { if ((P).trace && threadIdx.x == 0) atomicMax((P).trace + (slot), nk_globaltimer()); } while (0)

// Block-private per-subvolume sums.  On sm_100 shared-memory atomicAdd is native only for 32-bit integers
// (ATOMS.ADD / ATOMS.POPC.INC); the f64 and the 64-bit integer versions are compare-and-swap loops
// (ATOMS.CAST.SPIN.64) and were a quarter of the streaming kernel's stall samples (and 9-12 us of the rare path).  Terms
// are therefore accumulated in 64-bit FIXED POINT built from two native 32-bit adds: the low word's returned old value
// tells this add whether it wrapped, and the carry rides on the high word's add (two's complement, so signed terms
// just work, and the result does not depend on the order of the adds).  Quantum: 2^-46 eV for energies
// (1.4e-14 eV, the size of the f64 rounding noise of the reference's own sum), 2^-38 for flux terms
// (|v e| is at most ~0.03 eV A/ps for a few K of temperature difference, so 1e-8 of a single typical term).  A
// term outside the fixed-point range (|q| >= 2^40; never for physical occupations) or non-finite goes to
// an f64 side bin.  A block adds at most a few million terms: |sum| < 2^62.
#define NK_QE 70368744177664.0          // 2^46
#define NK_QF 274877906944.0            // 2^38
__device__ __forceinline__ void nk_bin_add(long long* q, double* side, double v, double scale) {
    const double t = v * scale;
    if (fabs(t) < 1099511627776.0) {
        const long long i = __double2ll_rn(t);
        const unsigned int lo = (unsigned int)i, hi = (unsigned int)(i >> 32);
        unsigned int* w = reinterpret_cast<unsigned int*>(q);            // little endian: w[0] low, w[1] high
        const unsigned int old = atomicAdd(w, lo);
        atomicAdd(w + 1, hi + ((old + lo) < old ? 1u : 0u));
    } else {
        atomicAdd(side, v);
    }
}

// Global merge of the block sums: the same fixed point one level up, 128 bits wide (two native 64-bit atomics with the
// carry riding on the high word), so that neither the order in which blocks finish nor the number of particles can change
// or overflow the sums: T_sv is a pure function of the particle set.
__device__ __forceinline__ void nk_gacc_add(unsigned long long* q, long long v) {
    if (v == 0) return;
    const unsigned long long lo = (unsigned long long)v;
    const unsigned long long hi = v < 0 ? ~0ull : 0ull;
    const unsigned long long old = atomicAdd(q, lo);
    atomicAdd(q + 1, hi + ((old + lo) < old ? 1ull : 0ull));
}
__device__ __forceinline__ double nk_gacc_value(const unsigned long long* q, double inv_scale) {
    const unsigned long long lo = __ldcg(q);
    const long long hi = (long long)__ldcg(q + 1);
    if ((hi == 0 && (long long)lo >= 0) || (hi == -1 && (long long)lo < 0)) return (double)(long long)lo * inv_scale;
    return ((double)hi * 18446744073709551616.0 + (double)lo) * inv_scale;
}

// The rare path's block-private copy of the accumulator vector: `acc` (double, the side bins) is followed by the
// fixed-point halves of the same entries.  Energies / reservoir balances use NK_QE, fluxes NK_QF, counters are integers.
#define NK_RACC_Q(P, acc, idx) (reinterpret_cast<long long*>((acc) + nk_acc_len((P).S, (P).R)) + (idx))
#define NK_RACC_E(P, acc, idx, v) nk_bin_add(NK_RACC_Q(P, acc, idx), (acc) + (idx), (v), NK_QE)
#define NK_RACC_F(P, acc, idx, v) nk_bin_add(NK_RACC_Q(P, acc, idx), (acc) + (idx), (v), NK_QF)
#define NK_RACC_N(P, acc, idx) nk_bin_add(NK_RACC_Q(P, acc, idx), (acc) + (idx), 1.0, 1.0)
__device__ __forceinline__ double nk_racc_inv_scale(int S, int R, int k) {
    if (k < S) return 1.0 / NK_QE;                                        // sum e
    if (k < 2 * S) return 1.0;                                            // count
    if (k < 5 * S) return 1.0 / NK_QF;                                    // sum v e
    if (k < 5 * S + R) return 1.0;                                        // N_leaving
    if (k < 5 * S + 2 * R) return 1.0 / NK_QE;                            // reservoir energy balance
    if (k < 5 * S + 5 * R) return 1.0 / NK_QF;                            // reservoir flux
    return 1.0;                                                           // emitted, absorbed
}
// One block, after every other block of the step has merged its sums: fixed point -> the f64 accumulator vector that the
// closing block (and, between ranks, the exchange) reads; the side bins already sit in `acc`.
__device__ __forceinline__ void nk_gacc_to_f64(const NkP& P) {
    const int n = nk_acc_len(P.S, P.R);
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double v = nk_gacc_value(P.acc_q + 2 * k, nk_racc_inv_scale(P.S, P.R, k));
        P.acc[k] = v + __ldcg(P.acc + k);
        P.acc_q[2 * k] = 0ull; P.acc_q[2 * k + 1] = 0ull;
    }
}

#define NK_STREAM_EMIT_A 0u
#define NK_STREAM_EMIT_B 1u
#define NK_STREAM_ROUGH0 2u          // + index of the boundary event within the step
#define NK_STREAM_EMIT_C 65536u      // fixed_rate dice / one_to_one (mode, entry time)

NK_DEVI void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                           unsigned int k0, unsigned int k1, unsigned int out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two doubles in [0,1) built like NumPy's random_sample: ((a>>5)*2^26 + (b>>6)) / 2^53
NK_DEVI void nk_uniforms(const NkP& P, long long id, long long step, unsigned int stream, double& u0, double& u1) {
    unsigned int r[4];
    unsigned long long uid = (unsigned long long)id;
    philox4x32_10((unsigned int)(uid & 0xFFFFFFFFull), (unsigned int)(uid >> 32), (unsigned int)step, stream,
                  P.seed_lo, P.seed_hi, r);
    u0 = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6)) / 9007199254740992.0;
    u1 = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6)) / 9007199254740992.0;
}

// ------------------------------------------------------------------------------------------------
// mode-table functions
// ------------------------------------------------------------------------------------------------
// Phonon.calculate_occupation (Phonon.py:338-345): 1/(exp(omega*hbar/(T*kb)) - 1), 0 if T<=0 or omega<=0
NK_DEVI double nk_bose(const NkP& P, double T, double omega) {
    if (!(T > 0.0) || !(omega > 0.0)) return 0.0;
    double x = nk_div(nk_mul(omega, P.hbar), nk_mul(T, P.kb));
    return nk_div(1.0, nk_sub(exp(x), 1.0));
}

// index i with Tg[i] <= T < Tg[i+1], clipped to [0, NT-2]  (searchsorted(right)-1)
NK_DEVI int nk_T_index(const NkP& P, double T) {
    int i = (int)floor((T - __ldg(P.Tg)) * P.Tg_inv_d);
    i = max(0, min(i, P.NT - 2));
    while (i > 0 && T < __ldg(P.Tg + i)) --i;
    while (i < P.NT - 2 && T >= __ldg(P.Tg + i + 1)) ++i;
    return i;
}

// Phonon.lifetime_function (Phonon.py:326-336) at integer (q,j): linear in T between two slabs
NK_DEVI double nk_tau(const NkP& P, double T, int m) {
    int i = nk_T_index(P, T);
    double t0 = __ldg(P.Tg + i), t1 = __ldg(P.Tg + i + 1);
    double w = nk_div(nk_sub(T, t0), nk_sub(t1, t0));
    double lo, hi;
    int r = i - P.tau_i0;
    if (r >= 0 && r <= 2) {
        const double4 q = *reinterpret_cast<const double4*>(P.tau4 + m);     // one 32 B sector
        lo = r == 0 ? q.x : (r == 1 ? q.y : q.z);
        hi = r == 0 ? q.y : (r == 1 ? q.z : q.w);
    } else {
        lo = __ldg(P.tau + (size_t)i * P.M + m);
        hi = __ldg(P.tau + (size_t)(i + 1) * P.M + m);
    }
    return nk_add(nk_mul(lo, nk_sub(1.0, w)), nk_mul(hi, w));
}

// np.interp(x, xp, fp) with interp1d's (below, above) fill values (Phonon.py:387-390)
NK_DEVI double nk_interp_table(const double* xp, const double* fp, int n, double x, double below, double above) {
    if (x != x) return x;
    if (x < xp[0]) return below;
    if (x > xp[n - 1]) return above;
    int lo = 0, hi = n - 1;            // invariant xp[lo] <= x, lo < n
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (x >= xp[n - 1]) return fp[n - 1];
    if (xp[lo] == x) return fp[lo];
    double slope = nk_div(nk_sub(fp[lo + 1], fp[lo]), nk_sub(xp[lo + 1], xp[lo]));
    return nk_add(nk_mul(slope, nk_sub(x, xp[lo])), fp[lo]);
}

// Same result as nk_interp_table, but the bracket search starts at index `guess` and gallops outwards: a couple of
// probes instead of log2(n) dependent loads when the guess is close (uniform grids, slowly changing temperatures).
NK_DEVI double nk_interp_table_from(const double* xp, const double* fp, int n, double x, double below, double above, int guess) {
    if (x != x) return x;
    if (x < xp[0]) return below;
    if (x > xp[n - 1]) return above;
    int lo = max(0, min(guess, n - 2)), hi;
    if (xp[lo] <= x) {
        int stepw = 1;
        hi = lo + 1;
        while (hi < n - 1 && xp[hi] <= x) { lo = hi; hi = min(n - 1, hi + stepw); stepw <<= 1; }
    } else {
        int stepw = 1;
        hi = lo; lo = max(0, hi - 1);
        while (lo > 0 && xp[lo] > x) { hi = lo; lo = max(0, lo - stepw); stepw <<= 1; }
    }
    while (hi - lo > 1) {              // invariant xp[lo] <= x < xp[hi] (or hi == n-1)
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (x >= xp[n - 1]) return fp[n - 1];
    if (xp[lo] == x) return fp[lo];
    double slope = nk_div(nk_sub(fp[lo + 1], fp[lo]), nk_sub(xp[lo + 1], xp[lo]));
    return nk_add(nk_mul(slope, nk_sub(x, xp[lo])), fp[lo]);
}

// Warp-aggregated counter increment: the lanes that arrive together issue ONE atomic and share the range.
NK_DEVI unsigned long long nk_agg_inc(unsigned long long* ctr) {
    const unsigned int m = __activemask();
    const unsigned int lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}
NK_DEVI unsigned int nk_agg_inc(unsigned int* ctr) {
    const unsigned int m = __activemask();
    const unsigned int lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if ((int)lane == leader) base = atomicAdd(ctr, (unsigned int)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------------
// subvolumes
// ------------------------------------------------------------------------------------------------
NK_DEVI double nk_dist2(const double* c, double x, double y, double z) {
    double dx = nk_sub(x, c[0]), dy = nk_sub(y, c[1]), dz = nk_sub(z, c[2]);
    return dot3(dx, dy, dz, dx, dy, dz);
}

// searchsorted(a, v, side='left') on an almost uniform ascending array with spacing 1/inv_d
NK_DEVI int nk_searchsorted_left(const double* a, int n, double v, double inv_d) {
    int g = (int)floor((v - a[0]) * inv_d) + 1;
    g = max(0, min(g, n));
    while (g > 0 && a[g - 1] >= v) --g;
    while (g < n && a[g] < v) ++g;
    return g;
}

// SubvolClassifier.predict (Geometry.py:1198-1213): nearest centre, squared distances summed in
// x,y,z order; exact ties (undefined upstream) resolve to the lowest index.
// `svc`, `sv_axis`, `sv_mid` may point to shared memory copies.
NK_DEVI int nk_classify(const NkP& P, const double* svc, const double* sv_mid, double x, double y, double z) {
    if (P.is_slice) {
        double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
        int g = P.S > 1 ? nk_searchsorted_left(sv_mid, P.S - 1, xa, P.sv_inv_dx) : 0;
        int best = g; double dbest = nk_dist2(svc + 3 * g, x, y, z);
        if (g > 0) { double d = nk_dist2(svc + 3 * (g - 1), x, y, z); if (d <= dbest) { dbest = d; best = g - 1; } }
        if (g + 1 < P.S) { double d = nk_dist2(svc + 3 * (g + 1), x, y, z); if (d < dbest) { dbest = d; best = g + 1; } }
        return best;
    }
    int best = 0; double dbest = nk_dist2(svc, x, y, z);
    for (int s = 1; s < P.S; ++s) {
        double d = nk_dist2(svc + 3 * s, x, y, z);
        if (d < dbest) { dbest = d; best = s; }
    }
    return best;
}

// scipy RBFInterpolator(kernel='cubic') evaluation (Population.py:588, :697-702) with the step's coefficients
NK_DEVI double nk_rbf_T(const NkP& P, const double* svc, const double* coef, double x, double y, double z) {
    const int nd = P.rbf_nd, S = P.S;
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double r2 = 0.0;
        for (int k = 0; k < nd; ++k) {
            const int d = P.rbf_dim[k];
            const double dx = (d == 0 ? x : (d == 1 ? y : z)) - svc[3 * s + d];
            r2 += dx * dx;
        }
        acc += coef[s] * (r2 * sqrt(r2));
    }
    acc += coef[S];
    for (int k = 0; k < nd; ++k) {
        const int d = P.rbf_dim[k];
        acc += coef[S + 1 + k] * (((d == 0 ? x : (d == 1 ? y : z)) - P.rbf_shift[k]) / P.rbf_scale[k]);
    }
    return acc;
}

// temperature_interpolator(x) (Population.py:570-590, :694-702; scipy interp1d formulas restated in
// oracle/nk_oracle.py:particle_temperature).  `sv` is the particle's subvolume if already known (-1
// otherwise); it is only used by the non-slice nearest rule.
NK_DEVI double nk_particle_T(const NkP& P, const double* svc, const double* sv_axis, const double* sv_mid,
                             const double* T_sv, double x, double y, double z, int sv) {
    if (P.is_slice) {
        double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
        if (P.interp == NK_INTERP_LINEAR && P.S > 1) {
            int idx = nk_searchsorted_left(sv_axis, P.S, xa, P.sv_inv_dx);
            idx = max(1, min(idx, P.S - 1));
            double xl = sv_axis[idx - 1], xh = sv_axis[idx];
            double den = nk_sub(xh, xl);
            return nk_add(nk_mul(nk_div(nk_sub(xa, xl), den), T_sv[idx]), nk_mul(nk_div(nk_sub(xh, xa), den), T_sv[idx - 1]));
        }
        int idx = P.S > 1 ? nk_searchsorted_left(sv_mid, P.S - 1, xa, P.sv_inv_dx) : 0;
        return T_sv[min(idx, P.S - 1)];
    }
    if (P.interp == NK_INTERP_RADIAL) return nk_rbf_T(P, svc, P.rbf_coef, x, y, z);
    if (sv < 0) sv = nk_classify(P, svc, sv_mid, x, y, z);
    return T_sv[sv];
}

// ------------------------------------------------------------------------------------------------
// Mesh.find_boundary (Mesh.py:806-856) for one ray.  `faces` may point to shared memory.
//
// Two stages per triangle.  Stage 1 (every lane, no divergence, ~23 FP64 instructions): an FMA-contracted copy of the
// plane test gives an approximate hit parameter ta = |num| / |den| through a Newton reciprocal and an approximate hit point;
// the triangle is skipped when that point lies outside its (margin-widened) bounding box, when the ray moves away from the
// plane, or when ta cannot beat the running minimum.  Every skip is CONSERVATIVE: the approximations are accurate to
// 1e-10 x mesh scale wherever |num| and |den| are not tiny, the margins are 1e-6 x mesh scale and 1 %, and whatever is tiny,
// non-finite or far outside the mesh is "ambiguous" and goes to stage 2.  Stage 2 (the few surviving candidates) is the
// reference's arithmetic, operation by operation: IEEE division, unfused products, the 1e-10 tolerances.  So the result
// -- facet, t, and the first-minimum tie rule -- is exactly what the unfiltered loop gives.
// ------------------------------------------------------------------------------------------------
// stage 2: the reference's test of one triangle, operation by operation (Mesh.py:818-856)
NK_DEVI void nk_ray_face_exact(const NkFace& T, double x, double y, double z, double vx, double vy, double vz, double& tbest, int& fbest) {
    double num = nk_add(dot3(x, y, z, T.nx, T.ny, T.nz), T.k);
    double den = dot3(vx, vy, vz, T.nx, T.ny, T.nz);
    // t = -num/den can only reach the tolerance when num and den have opposite signs (0, inf and NaN quotients are rejected
    // below anyway): skip the IEEE division for the planes the ray moves away from
    if (!((num < 0.0 && den > 0.0) || (num > 0.0 && den < 0.0))) return;
    // |num| * rcp(|den|) is within a few ulp of t: a plane whose approximate t exceeds the best one by more than 1e-12
    // relative can never satisfy t < tbest below, so the IEEE division is skipped for it
    if (tbest < CUDART_INF && fabs(num) * nk_rcp(fabs(den)) > tbest * (1.0 + 1e-12)) return;
    double t = -nk_div(num, den);
    if (!(t >= NK_TOL) || isinf(t)) return;                   // also rejects NaN
    if (!(t < tbest)) return;                                 // cannot become the first minimum
    double cx = nk_add(x, nk_mul(t, vx)), cy = nk_add(y, nk_mul(t, vy)), cz = nk_add(z, nk_mul(t, vz));
    if (!(cx >= T.lox && cy >= T.loy && cz >= T.loz && cx <= T.hix && cy <= T.hiy && cz <= T.hiz)) return;
    double dx = nk_sub(cx, T.ox), dy = nk_sub(cy, T.oy), dz = nk_sub(cz, T.oz);
    double a = T.ia0 * dx + T.ia1 * dy + T.ia2 * dz;
    double b = T.ib0 * dx + T.ib1 * dy + T.ib2 * dz;
    double w = nk_sub(1.0, nk_add(a, b));
    const double lo = -NK_TOL, hi = 1.0 + NK_TOL;
    if (!(a >= lo && a <= hi && b >= lo && b <= hi && w >= lo && w <= hi)) return;
    tbest = t; fbest = (int)T.facet;
}
// Small meshes (a box is 12 triangles) staged per block: every thread walks all faces with the exact test; the pre-filter
// below costs more than it saves there.
NK_DEVI void nk_ray_faces(const NkFace* faces, int F, double mesh_scale, double x, double y, double z, double vx, double vy, double vz,
                          double& tbest, int& fbest);
NK_DEVI void nk_ray_faces_small(const NkFace* faces, int F, double mesh_scale, double x, double y, double z, double vx, double vy, double vz,
                                double& tbest, int& fbest) {
#ifdef NK_SMALL_FILTER
    nk_ray_faces(faces, F, mesh_scale, x, y, z, vx, vy, vz, tbest, fbest);
#else
    for (int f = 0; f < F; ++f) nk_ray_face_exact(faces[f], x, y, z, vx, vy, vz, tbest, fbest);
#endif
}

struct NkRayPre {               // per-ray constants of the pre-filter
    double eps_n, den_thr, graze_lim;
};
NK_DEVI NkRayPre nk_ray_pre(double mesh_scale, double x, double y, double z, double vx, double vy, double vz) {
    NkRayPre r;
    const double vn1 = fabs(vx) + fabs(vy) + fabs(vz);
    const double xs = fabs(x) + fabs(y) + fabs(z);
    // |num| below eps_n (rounding noise x 1e3) -> ambiguous; an origin far outside the mesh (or NaN) makes everything ambiguous
    r.eps_n = (xs <= 10.0 * mesh_scale) ? 1e-11 * (xs + mesh_scale) : CUDART_INF;
    r.den_thr = 1e-3 * vn1;
    // a grazing ray (|den| <= den_thr, including the exact zeros of symmetric mode tables) meets the plane at t >= |num| / den_thr,
    // i.e. at an L1 distance >= 1e3 |num| from its origin: beyond every point of the mesh when that exceeds 2 (|x|_1 + 3 scale)
    r.graze_lim = (xs <= 10.0 * mesh_scale) ? 2.0 * (xs + 3.0 * mesh_scale) / 999.0 : CUDART_INF;
    return r;
}
// stage 1: true when the triangle certainly cannot be the first hit (all comparisons are false for NaN -> not skipped)
NK_DEVI bool nk_ray_face_skip(const NkFace& T, const NkRayPre& r, double x, double y, double z, double vx, double vy, double vz, double tb_hi) {
    const double nf = fma(x, T.nx, fma(y, T.ny, fma(z, T.nz, T.k)));
    const double df = fma(vx, T.nx, fma(vy, T.ny, vz * T.nz));
    const double an = fabs(nf), ad = fabs(df);
    const bool opposite = (__double2hiint(nf) ^ __double2hiint(df)) < 0;
    const double ta = an * nk_rcp(ad);
    const bool outside = (fabs(fma(ta, vx, x) - T.mx) > T.ex) | (fabs(fma(ta, vy, y) - T.my) > T.ey) | (fabs(fma(ta, vz, z) - T.mz) > T.ez);
    return (ad > r.den_thr) ? ((an > r.eps_n) & (!opposite | outside | (ta > tb_hi))) : (an > r.graze_lim);
}
NK_DEVI void nk_ray_faces(const NkFace* faces, int F, double mesh_scale, double x, double y, double z, double vx, double vy, double vz,
                          double& tbest, int& fbest) {
    const NkRayPre r = nk_ray_pre(mesh_scale, x, y, z, vx, vy, vz);
    double tb_hi = tbest * 1.01;                                  // inf stays inf
    int f = 0;
    // two triangles per round: their pre-filters are independent chains of FP64 instructions that overlap, and one branch
    // covers both (a stale tb_hi for the second one is only less strict)
    for (; f + 1 < F; f += 2) {
        const bool s0 = nk_ray_face_skip(faces[f], r, x, y, z, vx, vy, vz, tb_hi);
        const bool s1 = nk_ray_face_skip(faces[f + 1], r, x, y, z, vx, vy, vz, tb_hi);
        if (s0 & s1) continue;
        if (!s0) nk_ray_face_exact(faces[f], x, y, z, vx, vy, vz, tbest, fbest);
        if (!s1) nk_ray_face_exact(faces[f + 1], x, y, z, vx, vy, vz, tbest, fbest);
        tb_hi = tbest * 1.01;
    }
    if (f < F && !nk_ray_face_skip(faces[f], r, x, y, z, vx, vy, vz, tb_hi)) nk_ray_face_exact(faces[f], x, y, z, vx, vy, vz, tbest, fbest);
}

NK_DEVI void nk_find_boundary_1(const NkP& P, const NkFace* faces, double x, double y, double z,
                                double vx, double vy, double vz,
                                double& xc, double& yc, double& zc, double& tc, int& fc) {
    double tbest = CUDART_INF; int fbest = -1;
    nk_ray_faces_small(faces, P.F, P.mesh_scale, x, y, z, vx, vy, vz, tbest, fbest);
    tc = tbest; fc = fbest;
    xc = nk_add(x, nk_mul(tbest, vx)); yc = nk_add(y, nk_mul(tbest, vy)); zc = nk_add(z, nk_mul(tbest, vz));   // inf*0 = NaN like NumPy
}

// ------------------------------------------------------------------------------------------------
// one particle in registers
// ------------------------------------------------------------------------------------------------
struct NkParticle {
    double x, y, z, tc, occ;
    int mode, omode;
    double omega, vx, vy, vz;
    int cf; double cx, cy, cz;
    long long id;                // < 0: not loaded yet (read from pid[slot] on first use)
    long long slot;              // slot of an existing particle, -1 for a particle being emitted
    bool alive;
    bool mode_changed, occ_changed;   // what a rough-wall event modified (write-back of the rare path)
};

// Boundary loop of one particle whose next collision lies inside the current step
// (Population.boundary_scattering :1546-1683 with periodic_boundary_condition :1463-1489 and
// roughness_boundary_condition :1491-1544 / select_reflected_modes :941-988 /
// pick_diffuse_modes :990-1015), executed as a sequence of events.  `acc` receives the reservoir
// statistics (:1585-1602).  Returns with p.alive == false when the particle was absorbed.
// geometry tables the event loop reads; the pointers go to shared-memory copies when the mesh is small
struct NkGeo {
    const NkFace* faces;
    const int* bc; const int* partner; const int* res; const int* rough;
    const double* normal; const double* centroid;
};

struct NkEvState {
    double done;        // fraction of the step already simulated (calculated_ts)
    double ts;          // time to the next collision in units of dt (n_timesteps being rebuilt)
    unsigned int ev;    // rough-wall events so far in this step (Philox stream index)
    int it;             // event counter (cap)
};
__device__ __forceinline__ void nk_event_begin(const NkParticle& p, NkEvState& st) { st.done = 0.0; st.ts = p.tc; st.ev = 0; st.it = 0; }

// Advances the event loop of one particle until it either needs a new ray (returns true: origin p.x/y/z, direction p.v;
// feed the answer to nk_event_ray_done and call again) or is finished (returns false: absorbed -> p.alive == false,
// otherwise p.tc holds the new clock).
__device__ __forceinline__ bool nk_event_advance(const NkP& P, const NkGeo& G, NkParticle& p, NkEvState& st, long long step, double* acc) {
    const double dt = P.dt;
    if (st.it++ >= 4096) { atomicOr(&P.dyn->error, NK_ERR_EVENTS); p.tc = st.ts; return false; }
    int cfi = p.cf < 0 ? P.nf - 1 : p.cf;                    // bound_cond[-1] for escaped rays (:667, :1487)
    int cond = G.bc[cfi];
    double rem = nk_sub(1.0, st.done);
    if (rem > st.ts) {
        if (cond == NK_BC_T || cond == NK_BC_F) {
            // I. absorbed by a reservoir (:1565-1608)
            int r = G.res[cfi];
            if (r >= 0) {
                double e = nk_mul(nk_mul(P.hbar, p.omega), nk_sub(p.occ, nk_bose_lean(P.hbar, P.kb, P.res_T[r], p.omega)));
                const double* n = G.normal + 3 * cfi;
                double vn = dot3(p.vx, p.vy, p.vz, n[0], n[1], n[2]);
                NK_RACC_N(P, acc, NK_ACC_NLEAVE(P.S, P.R) + r);
                NK_RACC_E(P, acc, NK_ACC_EBAL(P.S, P.R) + r, -e);
                NK_RACC_F(P, acc, NK_ACC_RFLUX(P.S, P.R) + 3 * r + 0, nk_div(nk_mul(e, p.vx), vn));
                NK_RACC_F(P, acc, NK_ACC_RFLUX(P.S, P.R) + 3 * r + 1, nk_div(nk_mul(e, p.vy), vn));
                NK_RACC_F(P, acc, NK_ACC_RFLUX(P.S, P.R) + 3 * r + 2, nk_div(nk_mul(e, p.vz), vn));
            }
            p.alive = false;
            return false;
        }
        // start of the path segment that ends at the collision point (:1472-1474, :1504-1508)
        double qx = p.x, qy = p.y, qz = p.z;
        if (st.done == 0.0) { qx = nk_sub(qx, nk_mul(p.vx, dt)); qy = nk_sub(qy, nk_mul(p.vy, dt)); qz = nk_sub(qz, nk_mul(p.vz, dt)); }
        double dist = norm3(nk_sub(p.cx, qx), nk_sub(p.cy, qy), nk_sub(p.cz, qz));
        if (cond == NK_BC_P) {
            // II. periodic wrap (:1463-1489)
            int g = p.cf >= 0 ? G.partner[p.cf] : -1;
            if (g < 0) { atomicOr(&P.dyn->error, NK_ERR_EVENTS); p.tc = st.ts; return false; }
            const double* cg = G.centroid + 3 * g; const double* ch = G.centroid + 3 * p.cf;
            double nx = nk_add(p.cx, nk_sub(cg[0], ch[0])), ny = nk_add(p.cy, nk_sub(cg[1], ch[1])), nz = nk_add(p.cz, nk_sub(cg[2], ch[2]));
            st.done = nk_add(st.done, nk_div(dist, norm3(nk_mul(p.vx, dt), nk_mul(p.vy, dt), nk_mul(p.vz, dt))));
            p.x = nx; p.y = ny; p.z = nz;
        } else {
            // III. rough facet: specular or diffuse (:1491-1544, :941-1015)
            st.done = nk_add(st.done, nk_div(dist, nk_mul(norm3(p.vx, p.vy, p.vz), dt)));
            p.x = p.cx; p.y = p.cy; p.z = p.cz;
            int fr = G.rough[cfi];
            double u_dice, u_pick;
            if (p.id < 0) p.id = P.pid[p.slot];
            nk_uniforms(P, p.id, step, NK_STREAM_ROUGH0 + st.ev, u_dice, u_pick);
            ++st.ev;
            size_t li = (size_t)fr * P.M + p.mode;
            bool spec = P.true_spec[li] && (u_dice <= P.specularity[li]);
            p.mode_changed = true;
            if (spec) {
                p.mode = P.spec_out[li];                         // omega and occupation are kept (:955-971)
            } else {
                const double* rou = P.roulette + (size_t)fr * P.M;
                double target = nk_mul(u_pick, rou[P.M - 1]);
                int lo = 0, hi = P.M;                            // searchsorted left
                if (P.rou_guide) {
                    // monotonic table: u_pick in [b/K, (b+1)/K) brackets the answer between two guide entries (exact: the
                    // product with the total is monotonic in u), which cuts 18 dependent loads to 4-5
                    const int K = P.rou_guide_k;
                    const int b = min((int)(u_pick * (double)K), K - 1);
                    const int* g = P.rou_guide + (size_t)fr * (K + 1);
                    lo = g[b]; hi = g[b + 1];
                }
                while (lo < hi) { int mid = (lo + hi) >> 1; if (rou[mid] < target) lo = mid + 1; else hi = mid; }
                p.mode = min(lo, P.M - 1);
                p.omode = p.mode;
                p.omega = P.mprop[p.mode].omega;
                double Tc = nk_particle_T(P, P.svc, P.sv_axis, P.sv_mid, P.T_sv, p.cx, p.cy, p.cz, -1);
                p.occ = nk_bose_lean(P.hbar, P.kb, Tc, p.omega);
                p.occ_changed = true;
            }
            NkMode m = P.mprop[p.mode];
            p.vx = m.vx; p.vy = m.vy; p.vz = m.vz;
        }
        return true;
    }
    // IV. no further collision in this step (:1670-1681).  rem == ts never terminates upstream;
    // here it takes this branch.
    p.x = nk_add(p.x, nk_mul(nk_mul(p.vx, dt), rem));
    p.y = nk_add(p.y, nk_mul(nk_mul(p.vy, dt), rem));
    p.z = nk_add(p.z, nk_mul(nk_mul(p.vz, dt), rem));
    st.ts = nk_sub(st.ts, rem);
    p.tc = st.ts;
    return false;
}
// answer of the ray query an event asked for: the next collision (find_boundary of the new ray, :1480, :1526-1532)
__device__ __forceinline__ void nk_event_ray_done(const NkP& P, NkParticle& p, NkEvState& st, double t, int cf) {
    p.cf = cf;
    p.cx = nk_add(p.x, nk_mul(t, p.vx)); p.cy = nk_add(p.y, nk_mul(t, p.vy)); p.cz = nk_add(p.z, nk_mul(t, p.vz));   // inf*0 = NaN like NumPy
    st.ts = nk_div(t, P.dt);
}

// the whole loop for one thread that sweeps the triangles by itself (small meshes staged in shared memory)
__device__ __forceinline__ void nk_boundary_events(const NkP& P, const NkGeo& G, NkParticle& p, long long step, double* acc) {
    NkEvState st;
    nk_event_begin(p, st);
    while (nk_event_advance(P, G, p, st, step, acc)) {
        double tbest = CUDART_INF; int fbest = -1;
        nk_ray_faces_small(G.faces, P.F, P.mesh_scale, p.x, p.y, p.z, p.vx, p.vy, p.vz, tbest, fbest);
        nk_event_ray_done(P, p, st, tbest, fbest);
    }
}
