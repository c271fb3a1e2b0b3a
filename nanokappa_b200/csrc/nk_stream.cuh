// nk_stream.cuh -- the streaming kernel of a timestep: relaxation, drift, binning (direct and table variants)
// Part of the single translation unit nk_kernels.cu (included in this order: nk_tiles.cuh, nk_ops.cuh, nk_stream.cuh,
// nk_rare.cuh, nk_sort.cuh, nk_hostpipe.cuh); see DESIGN.md section 4.
#pragma once

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


// Kernel kinds (compile-time): how subvolume and particle temperature are found
#define NK_KIND_GENERAL 0        // grid / voronoi subvolumes (nearest centre by squared distance), nearest or RBF temperature
#define NK_KIND_FAST 1           // slice subvolumes + nearest temperature rule (the Si/Ge thin-film configurations)
#define NK_KIND_SLICE_LINEAR 2   // slice subvolumes + linear interpolation between the slice centres (parameters_test.txt)

// interp1d(kind='linear', fill_value='extrapolate') on the slice axis, given the slice g that holds xa (Population.py:570-573):
// idx = searchsorted(centres, xa) clipped to [1, S-1] follows from g and one comparison with the centre of slice g
__device__ __forceinline__ double nk_linear_T(const NkP& P, const NkSvSmem& s, const NkSvHot& h, double xa, int g) {
    int idx = xa > s.sv_axis[g] ? g + 1 : g;
    idx = max(1, min(idx, P.S - 1));
    const double xl = s.sv_axis[idx - 1], xh = s.sv_axis[idx];
    const double inv = h.tw[idx];                                              // 1 / (xh - xl), see nk_load_hot
    return ((xa - xl) * inv) * s.T_sv[idx] + ((xh - xa) * inv) * s.T_sv[idx - 1];
}

// lifetime_scattering of one particle (Population.py:1701-1710) at its position BEFORE the drift of the
// next step.  Returns the relaxed occupation; be0 / g0 = equilibrium occupation and slice used (FAST).
template <int KIND>
__device__ __forceinline__ double nk_relax_particle(const NkP& P, const NkSvSmem& s, const NkSvHot& h, double x, double y, double z,
                                                    int mode, double omega, double a, const double4& mt, double occ,
                                                    double& be0, int& g0) {
    double tau;
    if (KIND == NK_KIND_FAST) {
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
        if (KIND == NK_KIND_SLICE_LINEAR) {
            const double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
            double lo_b, hi_b;
            const int g = nk_slice_lookup(P, h.midp, s.sv_mid, xa, lo_b, hi_b);
            Ti = P.S > 1 ? nk_linear_T(P, s, h, xa, g) : s.T_sv[0];
        } else {
            Ti = nk_particle_T(P, s.svc, s.sv_axis, s.sv_mid, s.T_sv, x, y, z, -1);
        }
        int it; double w;
        if (KIND == NK_KIND_SLICE_LINEAR && P.Tg_uniform) {
            // uniform temperature grid (every node is Tg0 + i d exactly): the bracket and the weight by arithmetic, one fix-up
            // for a quotient that rounded across a node; same index as the search, weight within an ulp or two (it only feeds
            // occupations)
            it = max(0, min(__double2int_rd((Ti - P.Tg0) * P.Tg_inv_d), P.NT - 2));
            double t0 = fma((double)it, P.Tg_d, P.Tg0);
            if (Ti < t0 && it > 0) { --it; t0 -= P.Tg_d; }
            else if (Ti >= t0 + P.Tg_d && it < P.NT - 2) { ++it; t0 += P.Tg_d; }
            w = (Ti - t0) * P.Tg_inv_d;
        } else {
            it = nk_T_index(P, Ti);
            const double t0 = __ldg(P.Tg + it), t1 = __ldg(P.Tg + it + 1);
            w = (Ti - t0) * nk_rcp(t1 - t0);
        }
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

// which rank injects copy `k` (0-based) of a table entry that has emitted `fire` particles before this step
__device__ __forceinline__ int nk_emit_owner(const NkP& P, int m, unsigned int fire, int k) {
    return (int)((fire + (unsigned int)k + (unsigned int)m) % (unsigned int)P.world);
}

__device__ __forceinline__ void nk_emit_scan(const NkP& P) {
    // free-slot rings: slots freed in the previous step become recyclable, over-claims of exhausted rings are dropped
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < P.n_rings; b += (long long)gridDim.x * blockDim.x) {
        long long* c = P.fr_ctr + 3 * b;
        const long long tail = c[1];
        if (c[0] > c[2]) c[0] = c[2];
        if (c[2] != tail) c[2] = tail;
    }
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
    // Every rank advances the WHOLE table (identical on all ranks) and keeps the entries of which it owns at least one copy:
    // copies are dealt round-robin per entry (nk_emit_owner), so each rank injects 1/world of every mode and its per-mode
    // particle numbers stay in balance with what it absorbs.
    const unsigned int total = (unsigned int)P.R * (unsigned int)P.M;       // (R, Q*J) entries: far below 2^32; 32-bit index arithmetic
    const long long step = P.dyn->step;
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int r = (int)(e / (unsigned int)P.M);
        const int m = (int)(e % (unsigned int)P.M);
        const size_t idx = (size_t)e;
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
        const unsigned int fire = P.res_fire[idx];
        P.res_fire[idx] = (unsigned char)(fire + (unsigned int)n_new);
        bool mine = P.world == 1 || n_new >= P.world;
        for (int k = 0; !mine && k < n_new; ++k) mine = nk_emit_owner(P, m, fire, k) == P.rank;
        if (!mine) continue;
        const unsigned int k = atomicAdd(&P.dyn->n_emit, 1u);
        P.emitlist[k] = make_int2((r << 16) | ((int)fire << 8) | n_new, m);
    }
}

// (block-private fixed-point bins: nk_bin_add in nk_device.cuh)

// one live particle: deferred relaxation -> drift -> (if no collision this step) subvolume + energy bins.
// Returns true when the particle's collision falls inside this step (it then goes to the hit list).
template <bool HAS_ROUGH, int KIND, bool RELAX, bool FLUX>
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
    if (RELAX) occ = nk_relax_particle<KIND>(P, s, h, x, y, z, md, omega, a, mt, occ, be0, g0);
    x = nk_add(x, nk_mul(ma.y, dt)); y = nk_add(y, nk_mul(ma.z, dt)); z = nk_add(z, nk_mul(ma.w, dt));
    tc = nk_sub(tc, 1.0);
    if (tc < 0.0) return true;
    int sv;
    if (KIND != NK_KIND_GENERAL) {
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
    // (a per-(mode, subvolume) table of this occupation for the general kinds was measured: no gain -- the gather costs what
    //  the exponential saves, profiles/README.md)
    if (!(KIND == NK_KIND_FAST && RELAX && sv == g0)) be1 = nk_bose_fast(a, omega, h.invb[sv]);
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

// warp-aggregated append of up to two slots per lane to the hit list (full-mask votes: call converged).  The first
// hitrec_cap entries of a step also get a dense 64-byte record of what this thread holds in registers, so that the rare
// path reads its input coalesced; entries beyond that are re-read from the particle arrays (written below as usual).
__device__ __forceinline__ void nk_write_hitrec(const NkP& P, unsigned int pos, int slot, double x, double y, double z, double tc, double occ,
                                                int mode, int omode) {
    if ((long long)pos >= P.hitrec_cap) return;
    double4* r = reinterpret_cast<double4*>(P.hitrec + pos);
    r[0] = make_double4(x, y, z, tc);
    r[1] = make_double4(occ, __hiloint2double(mode, slot), __hiloint2double(0, omode), 0.0);
}
__device__ __forceinline__ void nk_push_hits(const NkP& P, unsigned int lane, bool h0, bool h1, long long base, const double2& X, const double2& Y,
                                             const double2& Z, const double2& TC, const double2& OC, const int2& MD, const int2& OM) {
    const unsigned int m0 = __ballot_sync(0xffffffffu, h0);
    const unsigned int m1 = __ballot_sync(0xffffffffu, h1);
    if (m0 | m1) {
        unsigned int pos = 0;
        if (lane == 0) pos = atomicAdd(&P.dyn->n_hits, __popc(m0) + __popc(m1));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        const unsigned int below = (1u << lane) - 1u;
        if (h0) {
            const unsigned int k = pos + __popc(m0 & below);
            P.hitlist[k] = (int)base;
            nk_write_hitrec(P, k, (int)base, X.x, Y.x, Z.x, TC.x, OC.x, MD.x, OM.x);
        }
        if (h1) {
            const unsigned int k = pos + __popc(m0) + __popc(m1 & below);
            P.hitlist[k] = (int)(base + 1);
            nk_write_hitrec(P, k, (int)(base + 1), X.y, Y.y, Z.y, TC.y, OC.y, MD.y, OM.y);
        }
    }
}

template <bool FLUX>
__device__ __forceinline__ void nk_flush_bins(const NkP& P, const long long* binE, const long long* binF, const double* binX,
                                              const unsigned int* binC) {
    const int S = P.S;
    double* acc = P.acc;
    unsigned long long* q = P.acc_q;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        if (binC[i]) {
            nk_gacc_add(q + 2 * (NK_ACC_E(S, P.R) + i), binE[i]);
            if (binX[i] != 0.0) atomicAdd(acc + NK_ACC_E(S, P.R) + i, binX[i]);
            nk_gacc_add(q + 2 * (NK_ACC_CNT(S, P.R) + i), (long long)binC[i]);
            if (FLUX) {
                for (int k = 0; k < 3; ++k) {
                    nk_gacc_add(q + 2 * (NK_ACC_FLUX(S, P.R) + 3 * i + k), binF[3 * i + k]);
                    if (binX[S + 3 * i + k] != 0.0) atomicAdd(acc + NK_ACC_FLUX(S, P.R) + 3 * i + k, binX[S + 3 * i + k]);
                }
            }
        }
    }
}

// ---- variant A: direct 128-bit global loads/stores (any capacity) ------------------------------------------
template <bool HAS_ROUGH, int KIND, bool RELAX, bool FLUX>
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
        if (base < n && MD.x >= 0) h0 = nk_step_particle<HAS_ROUGH, KIND, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.x, OM.x, X.x, Y.x, Z.x, TC.x, OC.x);
        if (base + 1 < n && MD.y >= 0) h1 = nk_step_particle<HAS_ROUGH, KIND, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.y, OM.y, X.y, Y.y, Z.y, TC.y, OC.y);
        nk_push_hits(P, lane, h0, h1, base, X, Y, Z, TC, OC, MD, OM);
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

// ---- variant T: per-(mode, subvolume) tables ------------------------------------------------------------------
// With the nearest-temperature rule both the equilibrium occupation and the relaxation factor of a particle are
// functions of (mode, subvolume) only.  When there are many particles per (mode, subvolume) pair it is cheaper
// to tabulate {n0, exp(-dt/tau)} once per step (k_mode_tables, M x S entries, the arithmetic of nk_bose_fast /
// nk_decay, so results are bit-identical to the direct variants) than to evaluate two exponentials and three
// reciprocals per particle.  Particles are ordered by mode, so a warp gathers from a handful of table rows.
__device__ __forceinline__ double2 nk_mode_table_entry(const NkP& P, const NkSvHot& h, int m, int sv) {
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
    return make_double2(be, dec);
}
__global__ void __launch_bounds__(256) k_mode_tables(NkP P) {
    extern __shared__ double sm[];
    NkSvHot h = nk_load_hot(P, sm);
    __syncthreads();
    // 32-bit index arithmetic: the table has at most NK_TAB_MAX_ENTRIES (6 Mi) entries, and a 64-bit division per entry cost
    // more instructions than the two exponentials
    const unsigned int S = (unsigned int)P.S;
    const unsigned int total = (unsigned int)P.M * S;
    const unsigned int stride = gridDim.x * blockDim.x;
    // two entries per round: their exp / reciprocal chains are independent and overlap in the FP64 pipe
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < total && i + stride < total; i += 2 * stride) {
        const unsigned int j = i + stride;
        const double2 e0 = nk_mode_table_entry(P, h, (int)(i / S), (int)(i % S));
        const double2 e1 = nk_mode_table_entry(P, h, (int)(j / S), (int)(j % S));
        P.hot_tab[i] = e0; P.hot_tab[j] = e1;
    }
    if (i < total) P.hot_tab[i] = nk_mode_table_entry(P, h, (int)(i / S), (int)(i % S));
}

// Two phases per particle so that the gathers of BOTH particles of a thread are in flight before either is consumed (the
// stall samples of the one-phase version sat on the table-row gather: profiles/r2_film_kstep_tab_hotspots.txt).
// Phase A: mode record + table row of the slice the particle is in; no side effects, safe for dead slots (mode clamped).
struct NkTabPre {
    double4 ma;            // omega, v_g
    double2 t0;            // {n0, decay} of (mode, slice before the drift)
    double be0;            // n0 of the omega-carrying mode in that slice
    double omega;
    int g0;
};
template <bool HAS_ROUGH, bool RELAX>
__device__ __forceinline__ void nk_step_tab_gather(const NkP& P, const NkSvSmem& s, const NkSvHot& h, int md, int om, double x, double y, double z,
                                                   NkTabPre& q) {
    md = max(md, 0); om = max(om, 0);
    nk_ld256(&P.mhot[md].omega, q.ma);
    q.omega = q.ma.x;
    if (HAS_ROUGH && om != md) q.omega = P.mhot[om].omega;
    q.g0 = -1; q.t0 = make_double2(0.0, 0.0); q.be0 = 0.0;
    if (RELAX) {
        const double xa = P.axis == 0 ? x : (P.axis == 1 ? y : z);
        double lo_b, hi_b;
        q.g0 = nk_slice_lookup(P, h.midp, s.sv_mid, xa, lo_b, hi_b);          // interp1d 'nearest'
        q.t0 = __ldg(P.hot_tab + (size_t)md * P.S + q.g0);
        q.be0 = (HAS_ROUGH && om != md) ? __ldg(P.hot_tab + (size_t)om * P.S + q.g0).x : q.t0.x;
    }
}
// Phase B: relaxation, drift, binning.  Returns true when the particle's collision falls inside this step.
template <bool HAS_ROUGH, bool RELAX, bool FLUX>
__device__ __forceinline__ bool nk_step_particle_tab(const NkP& P, const NkSvSmem& s, const NkSvHot& h, long long* binE, long long* binF,
                                                     double* binX, unsigned int* binC, int md, int om, const NkTabPre& q, double& x, double& y, double& z,
                                                     double& tc, double& occ) {
    const double4 ma = q.ma;
    const double a = nk_mul(P.hbar, q.omega);
    const double dt = P.dt;
    const int S = P.S;
    const double2* __restrict__ orow = P.hot_tab + (size_t)((HAS_ROUGH && om != md) ? om : md) * S;
    const double be0 = q.be0; const int g0 = q.g0;
    if (RELAX) {
        const double relaxed = be0 + (occ - be0) * q.t0.y;
        occ = q.t0.y > 0.0 ? relaxed : be0;
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
        NkTabPre q0, q1;
        nk_step_tab_gather<HAS_ROUGH, RELAX>(P, s, h, MD.x, OM.x, X.x, Y.x, Z.x, q0);
        nk_step_tab_gather<HAS_ROUGH, RELAX>(P, s, h, MD.y, OM.y, X.y, Y.y, Z.y, q1);
        if (base < n && MD.x >= 0) h0 = nk_step_particle_tab<HAS_ROUGH, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.x, OM.x, q0, X.x, Y.x, Z.x, TC.x, OC.x);
        if (base + 1 < n && MD.y >= 0) h1 = nk_step_particle_tab<HAS_ROUGH, RELAX, FLUX>(P, s, h, binE, binF, binX, binC, MD.y, OM.y, q1, X.y, Y.y, Z.y, TC.y, OC.y);
        nk_push_hits(P, lane, h0, h1, base, X, Y, Z, TC, OC, MD, OM);
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


// ---- launch-uniform variants: RELAX (a deferred relaxation is pending), FLUX (convergence step), rough walls present ----
typedef void (*nk_step_fn)(NkP);
template <int KIND>
static nk_step_fn nk_pick_kind(bool a, bool c, bool d) {
    return a ? (c ? (d ? (nk_step_fn)k_step<true, KIND, true, true> : (nk_step_fn)k_step<true, KIND, true, false>)
                  : (d ? (nk_step_fn)k_step<true, KIND, false, true> : (nk_step_fn)k_step<true, KIND, false, false>))
             : (c ? (d ? (nk_step_fn)k_step<false, KIND, true, true> : (nk_step_fn)k_step<false, KIND, true, false>)
                  : (d ? (nk_step_fn)k_step<false, KIND, false, true> : (nk_step_fn)k_step<false, KIND, false, false>));
}
static nk_step_fn nk_pick_tab(bool a, bool c, bool d) {
    return a ? (c ? (d ? (nk_step_fn)k_step_tab<true, true, true, true> : (nk_step_fn)k_step_tab<true, true, true, false>)
                  : (d ? (nk_step_fn)k_step_tab<true, true, false, true> : (nk_step_fn)k_step_tab<true, true, false, false>))
             : (c ? (d ? (nk_step_fn)k_step_tab<false, true, true, true> : (nk_step_fn)k_step_tab<false, true, true, false>)
                  : (d ? (nk_step_fn)k_step_tab<false, true, false, true> : (nk_step_fn)k_step_tab<false, true, false, false>));
}
// variant 4 = per-(mode, subvolume) tables of {n0, decay} (needs NK_KIND_FAST), 0 = direct
static nk_step_fn nk_pick_step(int variant, bool rough, int kind, bool relax, bool flux) {
    if (variant == 4) return nk_pick_tab(rough, relax, flux);
    return kind == NK_KIND_FAST ? nk_pick_kind<NK_KIND_FAST>(rough, relax, flux)
                                : (kind == NK_KIND_SLICE_LINEAR ? nk_pick_kind<NK_KIND_SLICE_LINEAR>(rough, relax, flux)
                                                                : nk_pick_kind<NK_KIND_GENERAL>(rough, relax, flux));
}
