// nk_rare.cuh -- the rare path of a timestep (boundary events, emission), the closing block, the in-kernel exchange, the relaxation flush
// Part of the single translation unit nk_kernels.cu (included in this order: nk_tiles.cuh, nk_ops.cuh, nk_stream.cuh,
// nk_rare.cuh, nk_sort.cuh, nk_hostpipe.cuh); see DESIGN.md section 4.
#pragma once

// ---- helpers shared by the rare-path code ------------------------------------------------------------------
__device__ __forceinline__ void nk_store_particle(const NkP& P, long long i, const NkParticle& p) {
    P.px[i] = p.x; P.py[i] = p.y; P.pz[i] = p.z; P.tc[i] = p.tc; P.occ[i] = p.occ;
    P.mode[i] = p.mode; P.omode[i] = p.omode; P.cfacet[i] = p.cf; P.cx[i] = p.cx; P.cy[i] = p.cy; P.cz[i] = p.cz;
}
// write-back of a particle that went through the event loop: position, clock and the new collision always change, mode /
// omega-carrying mode / occupation only at a rough wall.  Unchanged fields are not rewritten (every scattered 8-byte
// store costs a 32-byte sector read-modify-write in DRAM).
__device__ __forceinline__ void nk_store_after_events(const NkP& P, long long i, const NkParticle& p) {
    P.px[i] = p.x; P.py[i] = p.y; P.pz[i] = p.z; P.tc[i] = p.tc;
    P.cfacet[i] = p.cf; P.cx[i] = p.cx; P.cy[i] = p.cy; P.cz[i] = p.cz;
    if (p.mode_changed) { P.mode[i] = p.mode; P.omode[i] = p.omode; }
    if (p.occ_changed) P.occ[i] = p.occ;
}
// refresh_temperatures contribution of one particle handled outside k_step; `acc` is the block-private
// (shared memory) copy of the accumulator vector
__device__ __forceinline__ void nk_accumulate(const NkP& P, double* acc, const NkParticle& p, bool with_flux) {
    int sv = nk_classify(P, P.svc, P.sv_mid, p.x, p.y, p.z);
    double e = nk_mul(nk_mul(P.hbar, p.omega), nk_sub(p.occ, nk_bose_lean(P.hbar, P.kb, P.T_sv[sv], p.omega)));
    NK_RACC_E(P, acc, NK_ACC_E(P.S, P.R) + sv, e);
    NK_RACC_N(P, acc, NK_ACC_CNT(P.S, P.R) + sv);
    if (with_flux) {
        NK_RACC_F(P, acc, NK_ACC_FLUX(P.S, P.R) + 3 * sv, nk_mul(p.vx, e));
        NK_RACC_F(P, acc, NK_ACC_FLUX(P.S, P.R) + 3 * sv + 1, nk_mul(p.vy, e));
        NK_RACC_F(P, acc, NK_ACC_FLUX(P.S, P.R) + 3 * sv + 2, nk_mul(p.vz, e));
    }
}

// Free slots live in rings (NkP::fr_*): an absorbed particle pushes its slot at the tail of the ring that owns the slot -- the
// ring of the mode region the slot lies in (ordered part of the array), the global ring otherwise -- and emission pops at
// the head of the new particle's own mode ring (then the global ring, then it appends), but only entries pushed in EARLIER
// steps (below the ring's `snap`, advanced in the prologue of the next streaming kernel), so that pushes and pops of the
// same launch never touch the same entry.  A mode ring only ever receives the slots of its own region and the global ring
// holds cap entries: no ring can overflow.
__device__ __forceinline__ int nk_region_of_slot(const NkP& P, long long i, int mode_hint) {
    const int* f = P.mode_first;
    if (mode_hint >= 0 && (long long)f[mode_hint] <= i && i < (long long)f[mode_hint + 1]) return mode_hint;
    int lo = 0, hi = P.M;                       // last m with first[m] <= i  (regions of size 0 share their start with the next one)
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((long long)f[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ void nk_kill(const NkP& P, double* acc, long long i, int old_mode) {
    P.mode[i] = -1;
    int ring = 0; long long base = P.cap; unsigned long long size = (unsigned long long)P.cap;
    if (i < P.fr_sorted) {
        const int m = nk_region_of_slot(P, i, old_mode);
        ring = 1 + m; base = P.mode_first[m]; size = (unsigned long long)(P.mode_first[m + 1] - P.mode_first[m]);
    }
    long long* c = P.fr_ctr + 3 * (size_t)ring;
    const unsigned long long k = atomicAdd((unsigned long long*)(c + 1), 1ull);
    P.freelist[base + (long long)(k % size)] = (int)i;
    NK_RACC_N(P, acc, NK_ACC_NABS(P.S, P.R));
}
__device__ __forceinline__ long long nk_pop_ring(const NkP& P, int ring, long long base, long long size) {
    long long* c = P.fr_ctr + 3 * (size_t)ring;
    if (*(volatile long long*)c >= c[2]) return -1;                     // nothing recyclable here (cheap look before the atomic)
    const long long old = (long long)atomicAdd((unsigned long long*)c, 1ull);
    // a claim beyond the snapshot is not returned: the next prologue clamps the head back
    return old < c[2] ? (long long)P.freelist[base + (old % size)] : -1;
}
__device__ __forceinline__ long long nk_pop_mode_ring(const NkP& P, int m) {
    const long long base = P.mode_first[m], size = P.mode_first[m + 1] - base;
    return size > 0 ? nk_pop_ring(P, 1 + m, base, size) : -1;
}
__device__ __forceinline__ long long nk_take_slot(const NkP& P, int mode) {
    if (P.fr_sorted > 0) {
        long long s = nk_pop_mode_ring(P, mode);
        if (s >= 0) return s;
        // the pool of this mode is dry (its population fluctuates by ~sqrt(count)): a spare slot of a neighbouring mode keeps
        // the particle next to its own table rows ...
        for (int d = 1; d <= 4; ++d) {
            if (mode - d >= 0 && (s = nk_pop_mode_ring(P, mode - d)) >= 0) return s;
            if (mode + d < P.M && (s = nk_pop_mode_ring(P, mode + d)) >= 0) return s;
        }
    }
    {
        const long long s = nk_pop_ring(P, 0, P.cap, P.cap);
        if (s >= 0) return s;
    }
    if (P.fr_sorted > 0) {
        // ... and before the slot range grows, any free slot will do: the slots absorbed particles left behind sit in the
        // rings of THEIR regions, and a population whose total is steady must not creep towards the capacity
        unsigned int h = (unsigned int)clock64() * 2654435761u + (unsigned int)mode * 40503u + threadIdx.x;
        for (int t = 0; t < 24; ++t) {
            h = h * 1664525u + 1013904223u;
            const long long s = nk_pop_mode_ring(P, (int)((h >> 4) % (unsigned int)P.M));
            if (s >= 0) return s;
        }
    }
    long long slot = (long long)nk_agg_inc((unsigned long long*)&P.dyn->n_slots);      // nothing recyclable: append
    if (slot >= P.cap) {
        atomicAdd((unsigned long long*)&P.dyn->n_slots, (unsigned long long)(-1LL));
        atomicOr(&P.dyn->error, NK_ERR_CAPACITY);
        return -1;
    }
    return slot;
}

// One new particle of reservoir r in mode m entering the domain dt_in before the end of the step
// (Population.fill_reservoirs :491-508 + add_reservoir_particles :525-552 + Mesh.sample_surface :923-951), in three
// parts around the ray query for its first collision.
__device__ __forceinline__ void nk_emit_setup(const NkP& P, int r, int m, long long id, double uface, double us, double ur,
                                              NkParticle& p, double& x0, double& y0, double& z0) {
    const NkMode mp = P.mprop[m];
    p.id = id; p.slot = -1; p.mode_changed = false; p.occ_changed = false;
    // face ~ area: searchsorted(cdf, u, side='right') as np.random.choice does
    const int f0 = P.res_face_ptr[r], f1 = P.res_face_ptr[r + 1];
    int lo = f0, hi = f1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (P.res_face_cdf[mid] <= uface) lo = mid + 1; else hi = mid; }
    const int face = P.res_faces[min(lo, f1 - 1)];
    const double* V = P.face_vertices + 9 * (size_t)face;
    const double rs = sqrt(us);
    const double a0 = nk_sub(1.0, rs), a1 = nk_mul(nk_sub(1.0, ur), rs), a2 = nk_mul(ur, rs);
    x0 = nk_add(nk_add(nk_mul(a0, V[0]), nk_mul(a1, V[3])), nk_mul(a2, V[6]));
    y0 = nk_add(nk_add(nk_mul(a0, V[1]), nk_mul(a1, V[4])), nk_mul(a2, V[7]));
    z0 = nk_add(nk_add(nk_mul(a0, V[2]), nk_mul(a1, V[5])), nk_mul(a2, V[8]));
    p.mode = m; p.omode = m; p.omega = mp.omega; p.vx = mp.vx; p.vy = mp.vy; p.vz = mp.vz;
    p.x = x0; p.y = y0; p.z = z0;                    // ray origin of the first-collision query
    p.alive = true;
}
// after the ray query (t, cf) from (x0, y0, z0): clocks, entry drift, occupation.  Returns true when the first collision
// falls inside this very step (the event loop must run).
__device__ __forceinline__ bool nk_emit_post(const NkP& P, double* acc, int r, double dt_in, double x0, double y0, double z0, double t, int cf,
                                             NkParticle& p) {
    const double dt = P.dt;
    p.cf = cf;
    p.cx = nk_add(x0, nk_mul(t, p.vx)); p.cy = nk_add(y0, nk_mul(t, p.vy)); p.cz = nk_add(z0, nk_mul(t, p.vz));
    p.tc = nk_sub(nk_div(t, dt), nk_div(dt_in, dt));
    p.x = nk_add(x0, nk_mul(p.vx, dt_in)); p.y = nk_add(y0, nk_mul(p.vy, dt_in)); p.z = nk_add(z0, nk_mul(p.vz, dt_in));
    p.occ = nk_bose_lean(P.hbar, P.kb, P.res_T[r], p.omega);
    NK_RACC_N(P, acc, NK_ACC_NEMIT(P.S, P.R));
    return p.tc < 0.0;
}
// the particle is final: give it a slot (unless it crossed the whole domain within the step) and bin it
__device__ __forceinline__ void nk_emit_finish(const NkP& P, double* acc, NkParticle& p, bool with_flux) {
    if (!p.alive) { NK_RACC_N(P, acc, NK_ACC_NABS(P.S, P.R)); return; }
    const long long slot = nk_take_slot(P, p.mode);
    if (slot < 0) return;
    nk_store_particle(P, slot, p);
    P.pid[slot] = p.id;
    {
        const unsigned int k = nk_agg_inc(&P.dyn->n_new);
        if ((long long)k < P.newslots_cap) P.newslots[k] = (int)slot;
    }
    nk_accumulate(P, acc, p, with_flux);
}
__device__ __forceinline__ void nk_emit_particle(const NkP& P, const NkGeo& G, double* acc, int r, int m, long long id, double dt_in,
                                                 double uface, double us, double ur, long long step, bool with_flux) {
    NkParticle p;
    double x0, y0, z0;
    nk_emit_setup(P, r, m, id, uface, us, ur, p, x0, y0, z0);
    double t = CUDART_INF; int cf = -1;
    nk_ray_faces_small(G.faces, P.F, P.mesh_scale, x0, y0, z0, p.vx, p.vy, p.vz, t, cf);
    if (nk_emit_post(P, acc, r, dt_in, x0, y0, z0, t, cf, p)) nk_boundary_events(P, G, p, step, acc);
    nk_emit_finish(P, acc, p, with_flux);
}

// id, entry time and surface draws of copy c (1-based) of an emission-list entry (constant / fixed_rate, Population.py:385-406)
__device__ __forceinline__ void nk_emit_copy_draws(const NkP& P, int r, int m, int c, long long step, long long& id, double& dt_in,
                                                   double& uface, double& us, double& ur) {
    const double dt = P.dt;
    const size_t idx = (size_t)r * P.M + m;
    const double prob = P.enter_prob[idx];
    // numerator of the first copy's entry time: the counter after this step's update, or this step's dice
    const double lead = P.res_gen == NK_RESGEN_FIXED_RATE ? P.emit_u[idx] : P.res_counter[idx];
    id = NK_EMIT_ID_BASE + (((step * P.R + r) * (long long)P.M + m) * NK_EMIT_CMAX + (c - 1));
    double ua;
    nk_uniforms(P, id, step, NK_STREAM_EMIT_A, ua, uface);
    nk_uniforms(P, id, step, NK_STREAM_EMIT_B, us, ur);
    dt_in = (c == 1) ? nk_mul(dt, nk_sub(1.0, nk_div(lead, prob)))
                     : nk_mul(dt, nk_sub(1.0, nk_div(nk_add((double)(c - 1), ua), prob)));
}
// One emission-list entry (constant / fixed_rate): the copies of mode m from reservoir r that belong to this rank.
__device__ __forceinline__ void nk_emit_entry(const NkP& P, const NkGeo& G, double* acc, int r, int m, int n_new, unsigned int fire,
                                              long long step, bool with_flux) {
    for (int c = n_new; c >= 1; --c) {
        if (P.world > 1 && nk_emit_owner(P, m, fire, c - 1) != P.rank) continue;
        long long id; double dt_in, uface, us, ur;
        nk_emit_copy_draws(P, r, m, c, step, id, dt_in, uface, us, ur);
        nk_emit_particle(P, G, acc, r, m, id, dt_in, uface, us, ur, step, with_flux);
    }
}

// draws of the k-th re-emitted particle of the one_to_one mode (Population.py:457-489); returns false past the end
__device__ __forceinline__ bool nk_one_to_one_draws(const NkP& P, long long e, long long step, int& r, int& m, long long& id, double& dt_in,
                                                    double& uface, double& us, double& ur) {
    r = 0;
    for (; r < P.R; ++r) {
        const long long share = nk_one_to_one_share(P, r);
        if (e < share) break;
        e -= share;
    }
    if (r >= P.R) return false;
    const long long k = P.rank + e * P.world;
    id = NK_EMIT_ID_BASE + (step * P.R + r) * ((long long)P.M * NK_EMIT_CMAX) + k;
    double ua, umode, udt;
    nk_uniforms(P, id, step, NK_STREAM_EMIT_A, ua, uface);
    nk_uniforms(P, id, step, NK_STREAM_EMIT_B, us, ur);
    nk_uniforms(P, id, step, NK_STREAM_EMIT_C, umode, udt);
    const double* rou = P.res_roulette + (size_t)r * P.M;
    int lo = 0, hi = P.M;                                    // searchsorted left
    while (lo < hi) { int mid = (lo + hi) >> 1; if (rou[mid] < umode) lo = mid + 1; else hi = mid; }
    m = min(lo, P.M - 1);
    dt_in = nk_mul(P.dt, udt);
    return true;
}
__device__ __forceinline__ void nk_emit_one_to_one(const NkP& P, const NkGeo& G, double* acc, long long e, long long step, bool with_flux) {
    int r, m; long long id; double dt_in, uface, us, ur;
    if (!nk_one_to_one_draws(P, e, step, r, m, id, dt_in, uface, us, ur)) return;
    nk_emit_particle(P, G, acc, r, m, id, dt_in, uface, us, ur, step, with_flux);
}

// load hit-list entry w: from its dense record when it has one, from the particle arrays otherwise; cold fields
// (collision facet / point) always come from the arrays, the id only when a rough wall needs it
__device__ __forceinline__ long long nk_load_hit(const NkP& P, unsigned int w, NkParticle& p) {
    // the slot comes from the hit list (4 B, coalesced), so the record and the four cold fields are independent loads in
    // flight together: one DRAM round trip less on the item's dependency chain
    const long long i = P.hitlist[w];
    const int cf = P.cfacet[i];
    const double cx = P.cx[i], cy = P.cy[i], cz = P.cz[i];
    if ((long long)w < P.hitrec_cap) {
        const double4* r = reinterpret_cast<const double4*>(P.hitrec + w);
        const double4 a = r[0], b = r[1];
        p.x = a.x; p.y = a.y; p.z = a.z; p.tc = a.w; p.occ = b.x;
        p.mode = __double2hiint(b.y); p.omode = __double2loint(b.z);
    } else {
        p.x = P.px[i]; p.y = P.py[i]; p.z = P.pz[i]; p.tc = P.tc[i]; p.occ = P.occ[i];
        p.mode = P.mode[i]; p.omode = P.omode[i];
    }
    const NkMode m = P.mprop[p.mode];
    p.vx = m.vx; p.vy = m.vy; p.vz = m.vz;
    p.omega = (p.omode == p.mode) ? m.omega : P.mprop[p.omode].omega;
    p.cf = cf; p.cx = cx; p.cy = cy; p.cz = cz;
    p.id = -1; p.slot = i; p.alive = true; p.mode_changed = false; p.occ_changed = false;
    return i;
}
__device__ __forceinline__ void nk_finish_hit(const NkP& P, double* acc, long long i, int mode_in, const NkParticle& p, bool with_flux) {
    if (p.alive) {
        nk_store_after_events(P, i, p);
        nk_accumulate(P, acc, p, with_flux);
    } else {
        nk_kill(P, acc, i, mode_in);
    }
}
// One hit-list entry: the boundary event loop of an existing particle.
__device__ __forceinline__ void nk_hit_entry(const NkP& P, const NkGeo& G, double* acc, unsigned int w, long long step, bool with_flux) {
    NkParticle p;
    const long long i = nk_load_hit(P, w, p);
    const int mode_in = p.mode;
    nk_boundary_events(P, G, p, step, acc);
    nk_finish_hit(P, acc, i, mode_in, p, with_flux);
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
        d->blocks_done = 0;
    }
}

// All-reduce (sum) of the accumulator vector across the ranks of one box, done by the block that closes the
// step: every rank stores its vector straight into every peer's mailbox over NVLink (peer-mapped memory),
// publishes a sequence number, waits for the peers' numbers and adds the world's vectors in rank order, so
// all ranks get bit-identical sums without a separate collective launch.  Two mailbox parities: a rank can be
// at most one step ahead of the slowest one.  The wait is bounded (~20 s): a missing peer raises NK_ERR_COMM
// instead of hanging the GPU.
__device__ bool nk_exchange_sums(const NkP& P) {
    // message of a rank: the 128-bit fixed-point sums {lo, hi} of every accumulator entry, then the f64 side bins.  The
    // integer parts are added exactly, so the world totals -- and with them T_sv -- do not depend on how the particles
    // are spread over the ranks: a sharded run repeats the single-GPU run bit for bit.
    const int len = nk_acc_len(P.S, P.R);
    const int W = P.world;
    const size_t stride = 3 * (size_t)len;                       // 8-byte words per (parity, sender)
    const unsigned long long seq = (unsigned long long)(P.dyn->step + 1);
    const int par = (int)(seq & 1ull);
    for (int r = 0; r < W; ++r) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.peer_mbox[r]) + ((size_t)par * W + P.rank) * stride;
        for (int i = threadIdx.x; i < 2 * len; i += blockDim.x) dst[i] = __ldcg(P.acc_q + i);
        for (int i = threadIdx.x; i < len; i += blockDim.x) dst[2 * len + i] = (unsigned long long)__double_as_longlong(__ldcg(P.acc + i));
    }
    __threadfence_system();
    __syncthreads();
    int timed_out = 0;
    if ((int)threadIdx.x < W) {
        volatile unsigned long long* f = P.peer_flags[threadIdx.x] + (size_t)par * W + P.rank;
        *f = seq;
        __threadfence_system();
        volatile unsigned long long* mine = P.flags_local + (size_t)par * W + threadIdx.x;
        const long long t0 = clock64();
        while (*mine != seq) {
            if (clock64() - t0 > 40000000000LL) { atomicOr(&P.dyn->error, NK_ERR_COMM); timed_out = 1; break; }   // ~20 s
        }
    }
    // a missing peer freezes the step (no temperatures from incomplete sums); the sticky error bit reaches the host
    if (__syncthreads_or(timed_out)) return false;
    __threadfence_system();
    const volatile unsigned long long* box = reinterpret_cast<const volatile unsigned long long*>(P.mbox_local) + (size_t)par * W * stride;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        unsigned long long lo = 0ull, hi = 0ull;
        double side = 0.0;
        for (int r = 0; r < W; ++r) {
            const volatile unsigned long long* m = box + (size_t)r * stride;
            const unsigned long long a = m[2 * i], b = m[2 * i + 1];
            const unsigned long long s = lo + a;
            hi += b + (s < lo ? 1ull : 0ull);
            lo = s;
            side += __longlong_as_double((long long)m[2 * len + i]);
        }
        P.acc_q[2 * i] = lo; P.acc_q[2 * i + 1] = hi;
        P.acc[i] = side;
    }
    __threadfence();
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(1024) k_finalize(NkP P) {
    extern __shared__ double sm[];
    nk_finalize_block(P, sm);
}

// ---- the rare path of a step: boundary events of the hit list + reservoir emission -----------------------------
// Work items [0, n_hits) are existing particles whose collision falls inside the step, [n_hits, n_hits + n_emit)
// are emission-list entries.  With FUSE the last block to finish closes the step (single-GPU path, or the in-kernel
// exchange: no collective launch between the two halves).
//   k_rare        meshes of at most NK_RARE_FACES triangles: they are staged in shared memory once per block and every
//                 thread runs the event loop of its item on its own.
//   k_rare_tiled  larger meshes: the threads of a block advance their items in lock step -- everybody runs until it
//                 needs a ray (or is done), then the block sweeps the triangle tiles together (nk_tiles.cuh) -- so a
//                 tile is fetched once per block and round instead of once per ray.
#ifndef NK_RARE_THREADS
#define NK_RARE_THREADS 128         // (64-thread blocks with 8 blocks/SM were measured: 0.134 vs 0.123 ms on the film, equal on the rough bar)
#endif
#define NK_RARE_FACES 128
#define NK_RARE_FACETS 64
#ifndef NK_RARE_MIN_BLOCKS
#define NK_RARE_MIN_BLOCKS 4
#endif
// dynamic shared memory of k_rare: [closing block scratch 3S | block accumulators 2 nacc | faces | facet tables]
__host__ __device__ inline size_t nk_rare_fin_doubles(int S, int R) { return 3 * (size_t)S + 2 * (size_t)nk_acc_len(S, R); }
__host__ __device__ inline size_t nk_rare_smem_bytes(int S, int R, int F, int nf) {
    size_t b = nk_rare_fin_doubles(S, R) * 8;
    if (F <= NK_RARE_FACES) b += (size_t)F * sizeof(NkFace);
    if (nf <= NK_RARE_FACETS) b += (size_t)nf * (6 * 8 + 4 * 4);
    return b + 16;
}

// merge of a block's private accumulators + the closing protocol shared by both kernels
__device__ __forceinline__ void nk_rare_merge(const NkP& P, const double* racc, const long long* rq) {
    const int nacc = nk_acc_len(P.S, P.R);
    for (int k = threadIdx.x; k < nacc; k += blockDim.x) {
        nk_gacc_add(P.acc_q + 2 * k, rq[k]);
        if (racc[k] != 0.0) atomicAdd(P.acc + k, racc[k]);
    }
    if (threadIdx.x == 0) {
        // live count: + particles that got a slot (emitted - absorbed on arrival) - absorbed
        const double d = (double)(rq[NK_ACC_NEMIT(P.S, P.R)] - rq[NK_ACC_NABS(P.S, P.R)]);
        if (d != 0.0) atomicAdd((unsigned long long*)&P.dyn->n_alive, (unsigned long long)(long long)d);
    }
}
template <bool FUSE>
__device__ __forceinline__ void nk_rare_close(const NkP& P, double* sm_fin, int* s_last) {
    // the last block to finish turns the fixed-point sums into the f64 accumulator vector; with FUSE it also closes the step
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&P.dyn->blocks_done, 1u);
        *s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (*s_last) {
        __threadfence();
        if (FUSE && P.comm_on) {
            if (!nk_exchange_sums(P)) { if (threadIdx.x == 0) P.dyn->blocks_done = 0; return; }     // world totals, still in fixed point
        }
        nk_gacc_to_f64(P);
        __threadfence();
        __syncthreads();
        if (!FUSE) { if (threadIdx.x == 0) P.dyn->blocks_done = 0; return; }
        if (P.trace && threadIdx.x == 0) P.trace[4] = nk_globaltimer();
        nk_finalize_block(P, sm_fin);
        __syncthreads();
        if (P.trace && threadIdx.x == 0) P.trace[5] = nk_globaltimer();
    }
}

template <bool FUSE>
__global__ void __launch_bounds__(NK_RARE_THREADS, NK_RARE_MIN_BLOCKS) k_rare(NkP P) {
    extern __shared__ double sm_fin[];
    __shared__ int s_last;
    NkGeo G;
    G.faces = P.faces; G.bc = P.facet_bc; G.partner = P.facet_partner; G.res = P.facet_res; G.rough = P.facet_rough;
    G.normal = P.facet_normal; G.centroid = P.facet_centroid;
    NK_TRACE_MARK_FIRST(P, 2);
    // blocks beyond the work list (most of them when few particles hit a wall) go straight to the closing protocol
    const bool has_work = (unsigned long long)blockIdx.x * blockDim.x < (unsigned long long)P.dyn->n_hits + P.dyn->n_emit;
    if (has_work) {
        double* extra = sm_fin + nk_rare_fin_doubles(P.S, P.R);
        if (P.F <= NK_RARE_FACES) {
            const double* src = reinterpret_cast<const double*>(P.faces);
            NkFace* sfaces = reinterpret_cast<NkFace*>(extra);
            for (int k = threadIdx.x; k < P.F * (int)(sizeof(NkFace) / 8); k += blockDim.x) extra[k] = src[k];
            G.faces = sfaces;
            extra += (size_t)P.F * (sizeof(NkFace) / 8);
        }
        if (P.nf <= NK_RARE_FACETS) {
            const int nf = P.nf;
            double* sfd = extra;                                  // normals (3 nf), centroids (3 nf)
            int* sfi = reinterpret_cast<int*>(extra + 6 * nf);    // bc, partner, res, rough (nf each)
            for (int k = threadIdx.x; k < nf; k += blockDim.x) {
                sfi[k] = P.facet_bc[k]; sfi[nf + k] = P.facet_partner[k];
                sfi[2 * nf + k] = P.facet_res[k]; sfi[3 * nf + k] = P.facet_rough[k];
            }
            for (int k = threadIdx.x; k < 3 * nf; k += blockDim.x) { sfd[k] = P.facet_normal[k]; sfd[3 * nf + k] = P.facet_centroid[k]; }
            G.bc = sfi; G.partner = sfi + nf; G.res = sfi + 2 * nf; G.rough = sfi + 3 * nf;
            G.normal = sfd; G.centroid = sfd + 3 * nf;
        }
        // block-private accumulators: thousands of items would otherwise hammer the same ~40 global addresses
        double* racc = sm_fin + 3 * P.S;
        const int nacc = nk_acc_len(P.S, P.R);
        long long* rq = reinterpret_cast<long long*>(racc + nacc);      // fixed-point halves of the same entries (nk_racc_*)
        for (int k = threadIdx.x; k < nacc; k += blockDim.x) { racc[k] = 0.0; rq[k] = 0; }
        __syncthreads();
        const unsigned int nh = P.dyn->n_hits, ne = P.dyn->n_emit;
        const long long step = P.dyn->step;
        const bool with_flux = ((unsigned int)(step + 1) % (unsigned int)P.n_dt_to_conv) == 0u;     // 32-bit: a 64-bit modulo per thread is ~100 instructions
        for (unsigned int w = blockIdx.x * blockDim.x + threadIdx.x; w < nh + ne; w += gridDim.x * blockDim.x) {
            if (w < nh) {
                nk_hit_entry(P, G, racc, w, step, with_flux);
            } else {
                if (P.res_gen == NK_RESGEN_ONE_TO_ONE) {
                    nk_emit_one_to_one(P, G, racc, (long long)(w - nh), step, with_flux);
                } else {
                    const int2 e = P.emitlist[w - nh];
                    nk_emit_entry(P, G, racc, e.x >> 16, e.y, e.x & 0xff, (unsigned int)((e.x >> 8) & 0xff), step, with_flux);
                }
            }
        }
        __syncthreads();
        NK_TRACE_MARK_MAX(P, 3);
        nk_rare_merge(P, racc, rq);
    }
    nk_rare_close<FUSE>(P, sm_fin, &s_last);
}

// The same work with block-cooperative triangle tiles.  Dynamic shared memory: NK_TILE_SMEM_BYTES for the tile pipeline,
// then the closing block's scratch + this block's private accumulators (as k_rare).
#define NK_ITEM_DONE 0
#define NK_ITEM_EMIT_NEXT 1      // emission entry: pick the next copy that belongs to this rank
#define NK_ITEM_EMIT_RAY 2       // waiting for the first-collision ray of a new particle
#define NK_ITEM_EVENTS 3         // inside the boundary event loop
#define NK_RARE_TILED_THREADS 128    // rays that share one fetch of a triangle tile
template <bool FUSE>
__global__ void __launch_bounds__(NK_RARE_TILED_THREADS, 2) k_rare_tiled(NkP P) {
    extern __shared__ __align__(128) unsigned char rare_smem[];
    __shared__ int s_last;
    double* sm_fin = reinterpret_cast<double*>(rare_smem + NK_TILE_SMEM_BYTES);
    NkGeo G;
    G.faces = P.faces; G.bc = P.facet_bc; G.partner = P.facet_partner; G.res = P.facet_res; G.rough = P.facet_rough;
    G.normal = P.facet_normal; G.centroid = P.facet_centroid;
    NK_TRACE_MARK_FIRST(P, 2);
    const bool has_work = (unsigned long long)blockIdx.x * blockDim.x < (unsigned long long)P.dyn->n_hits + P.dyn->n_emit;
    if (has_work) {
        NkTilePipe tp;
        nk_tiles_init(tp, rare_smem, P);
        double* racc = sm_fin + 3 * P.S;
        const int nacc = nk_acc_len(P.S, P.R);
        long long* rq = reinterpret_cast<long long*>(racc + nacc);
        for (int k = threadIdx.x; k < nacc; k += blockDim.x) { racc[k] = 0.0; rq[k] = 0; }
        __syncthreads();
        const unsigned int nh = P.dyn->n_hits, ne = P.dyn->n_emit;
        const long long step = P.dyn->step;
        const bool with_flux = ((unsigned int)(step + 1) % (unsigned int)P.n_dt_to_conv) == 0u;     // 32-bit: a 64-bit modulo per thread is ~100 instructions
        for (unsigned int w0 = blockIdx.x * blockDim.x; w0 < nh + ne; w0 += gridDim.x * blockDim.x) {     // block-uniform
            const unsigned int w = w0 + threadIdx.x;
            NkParticle p;
            NkEvState st;
            int phase = NK_ITEM_DONE, kind = 0;           // kind 1: hit, 2: emission entry, 3: one_to_one particle
            long long slot_i = -1; int mode_in = -1;
            int r = 0, m = 0, c = 0; unsigned int fire = 0;
            double dt_in = 0.0, x0 = 0.0, y0 = 0.0, z0 = 0.0;
            p.alive = false; p.x = p.y = p.z = p.vx = p.vy = p.vz = 0.0;
            if (w < nh) {
                kind = 1;
                slot_i = nk_load_hit(P, w, p);
                mode_in = p.mode;
                nk_event_begin(p, st);
                phase = NK_ITEM_EVENTS;
            } else if (w < nh + ne) {
                if (P.res_gen == NK_RESGEN_ONE_TO_ONE) {
                    kind = 3;
                    long long id; double uface, us, ur;
                    if (nk_one_to_one_draws(P, (long long)(w - nh), step, r, m, id, dt_in, uface, us, ur)) {
                        nk_emit_setup(P, r, m, id, uface, us, ur, p, x0, y0, z0);
                        phase = NK_ITEM_EMIT_RAY;
                    }
                } else {
                    kind = 2;
                    const int2 e = P.emitlist[w - nh];
                    r = e.x >> 16; m = e.y; c = e.x & 0xff; fire = (unsigned int)((e.x >> 8) & 0xff);
                    phase = NK_ITEM_EMIT_NEXT;
                }
            }
            for (;;) {
                // every thread runs until its item needs a ray or is finished
                bool need = false;
                while (phase != NK_ITEM_DONE && !need) {
                    if (phase == NK_ITEM_EMIT_NEXT) {
                        while (c >= 1 && P.world > 1 && nk_emit_owner(P, m, fire, c - 1) != P.rank) --c;
                        if (c < 1) { phase = NK_ITEM_DONE; break; }
                        long long id; double uface, us, ur;
                        nk_emit_copy_draws(P, r, m, c, step, id, dt_in, uface, us, ur);
                        nk_emit_setup(P, r, m, id, uface, us, ur, p, x0, y0, z0);
                        phase = NK_ITEM_EMIT_RAY;
                        need = true;
                    } else if (phase == NK_ITEM_EMIT_RAY) {
                        need = true;
                    } else {                                   // NK_ITEM_EVENTS
                        if (nk_event_advance(P, G, p, st, step, racc)) { need = true; break; }
                        if (kind == 1) { nk_finish_hit(P, racc, slot_i, mode_in, p, with_flux); phase = NK_ITEM_DONE; }
                        else {
                            nk_emit_finish(P, racc, p, with_flux);
                            if (kind == 2) { --c; phase = NK_ITEM_EMIT_NEXT; } else phase = NK_ITEM_DONE;
                        }
                    }
                }
                if (!__syncthreads_or(need ? 1 : 0)) break;
                double tb = CUDART_INF; int fb = -1;
                nk_tiles_sweep(tp, P, need, p.x, p.y, p.z, p.vx, p.vy, p.vz, tb, fb);
                if (need) {
                    if (phase == NK_ITEM_EMIT_RAY) {
                        if (nk_emit_post(P, racc, r, dt_in, x0, y0, z0, tb, fb, p)) { nk_event_begin(p, st); phase = NK_ITEM_EVENTS; }
                        else {
                            nk_emit_finish(P, racc, p, with_flux);
                            if (kind == 2) { --c; phase = NK_ITEM_EMIT_NEXT; } else phase = NK_ITEM_DONE;
                        }
                    } else {
                        nk_event_ray_done(P, p, st, tb, fb);
                    }
                }
            }
        }
        __syncthreads();
        NK_TRACE_MARK_MAX(P, 3);
        nk_rare_merge(P, racc, rq);
    }
    nk_rare_close<FUSE>(P, sm_fin, &s_last);
}

// apply the deferred lifetime_scattering so that `occ` is what the reference holds after run_timestep
// (same arithmetic as the head of k_step, so flushing between steps is bit-neutral)
template <int KIND>
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
        P.occ[i] = nk_relax_particle<KIND>(P, s, h, P.px[i], P.py[i], P.pz[i], md, omega, nk_mul(P.hbar, omega), mt, P.occ[i], be0, g0);
    }
}
__global__ void k_clear_relax(NkP P) { P.dyn->relax_pending = 0; }
