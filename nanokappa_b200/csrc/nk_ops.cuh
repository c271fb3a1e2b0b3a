// nk_ops.cuh -- shared-memory subvolume tables and the operator kernels behind the reference's method seams
// Part of the single translation unit nk_kernels.cu (included in this order: nk_tiles.cuh, nk_ops.cuh, nk_stream.cuh,
// nk_rare.cuh, nk_sort.cuh, nk_hostpipe.cuh); see DESIGN.md section 4.
#pragma once

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

// Mesh.find_boundary operator seam: one ray per thread, the triangles stream through shared memory in TMA-staged tiles
// (nk_tiles.cuh).  Launch with NK_TILE_SMEM_BYTES of dynamic shared memory.
#define NK_RAY_THREADS 256
__global__ void __launch_bounds__(NK_RAY_THREADS, 4) k_find_boundary(NkP P, long long n, const double* __restrict__ x,
                                                                   const double* __restrict__ v, double* __restrict__ xc,
                                                                   double* __restrict__ tc, int* __restrict__ fc) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    NkTilePipe tp;
    nk_tiles_init(tp, tile_smem, P);
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        const bool live = i < n;
        double px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0;
        if (live) { px = x[3 * i]; py = x[3 * i + 1]; pz = x[3 * i + 2]; vx = v[3 * i]; vy = v[3 * i + 1]; vz = v[3 * i + 2]; }
        double tbest = CUDART_INF; int fbest = -1;
        nk_tiles_sweep(tp, P, live, px, py, pz, vx, vy, vz, tbest, fbest);
        if (live) {
            tc[i] = tbest; fc[i] = fbest;
            xc[3 * i] = nk_add(px, nk_mul(tbest, vx)); xc[3 * i + 1] = nk_add(py, nk_mul(tbest, vy)); xc[3 * i + 2] = nk_add(pz, nk_mul(tbest, vz));
        }
    }
}

// first collision of every live slot (Population.py:308-316): P = N rays against all F triangles
__global__ void __launch_bounds__(NK_RAY_THREADS, 4) k_init_collisions(NkP P) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    NkTilePipe tp;
    nk_tiles_init(tp, tile_smem, P);
    const long long n = P.dyn->n_slots;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        long long i = base + threadIdx.x;
        bool live = i < n && P.mode[i] >= 0;
        double px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0;
        if (live) { NkMode m = P.mprop[P.mode[i]]; px = P.px[i]; py = P.py[i]; pz = P.pz[i]; vx = m.vx; vy = m.vy; vz = m.vz; }
        double tbest = CUDART_INF; int fbest = -1;
        nk_tiles_sweep(tp, P, live, px, py, pz, vx, vy, vz, tbest, fbest);
        if (live) {
            P.tc[i] = nk_div(tbest, P.dt); P.cfacet[i] = fbest;
            P.cx[i] = nk_add(px, nk_mul(tbest, vx)); P.cy[i] = nk_add(py, nk_mul(tbest, vy)); P.cz[i] = nk_add(pz, nk_mul(tbest, vz));
        }
    }
}

// Mesh.contains_naive (Mesh.py:785-804) for set-up at scale (rejection sampling of initial positions inside an arbitrary mesh,
// Population.py:209-246): crossing parity of a ray from the point through all triangles, streamed in tiles.  Two fixed,
// unrelated directions vote and a third decides when they disagree (a ray that grazes an edge shared by two triangles is
// counted twice or not at all), exactly as nanokappa_b200.classes.Mesh.contains does on the host.
__device__ __forceinline__ void nk_parity_faces(const NkFace* faces, int F, double x, double y, double z, double dx, double dy, double dz,
                                                unsigned int& crossings) {
    for (int f = 0; f < F; ++f) {
        const NkFace& T = faces[f];
        const double den = T.nx * dx + T.ny * dy + T.nz * dz;
        const double t = -(x * T.nx + y * T.ny + z * T.nz + T.k) / den;
        if (!(t > NK_TOL) || isinf(t)) continue;
        const double cx = x + t * dx - T.ox, cy = y + t * dy - T.oy, cz = z + t * dz - T.oz;
        const double a = T.ia0 * cx + T.ia1 * cy + T.ia2 * cz;
        const double b = T.ib0 * cx + T.ib1 * cy + T.ib2 * cz;
        if (a >= 0.0 && b >= 0.0 && a + b <= 1.0) ++crossings;
    }
}
__device__ __forceinline__ unsigned int nk_tiles_parity(NkTilePipe& tp, const NkP& P, bool need, double x, double y, double z,
                                                        double dx, double dy, double dz) {
    unsigned int c = 0;
    if (tp.resident) {
        if (need) nk_parity_faces(tp.buf, P.F, x, y, z, dx, dy, dz, c);
        return c;
    }
    const int F = P.F, nt = (F + NK_TILE_FACES - 1) / NK_TILE_FACES;
    nk_tiles_begin_sweep(tp, P, nt);
    for (int t = 0; t < nt; ++t) {
        const int b = t & 1;
        nk_mbar_wait(&tp.bar[b], b ? tp.parity1 : tp.parity0);
        if (b) tp.parity1 ^= 1u; else tp.parity0 ^= 1u;
        if (need) nk_parity_faces(tp.buf + (size_t)b * NK_TILE_FACES, min(NK_TILE_FACES, F - t * NK_TILE_FACES), x, y, z, dx, dy, dz, c);
        if (t + 2 < nt) nk_tile_release(tp, P, b, t + 2);
    }
    return c;
}
__global__ void __launch_bounds__(NK_RAY_THREADS, 4) k_contains(NkP P, long long n, const double* __restrict__ x, unsigned char* __restrict__ inside) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    NkTilePipe tp;
    nk_tiles_init(tp, tile_smem, P);
    // the three voting directions of Mesh.contains, normalised
    const double d0x = 0.68273165190000003, d0y = 0.53411270430000004, d0z = 0.49851632270000002;
    const double d1x = -0.37119428350000001, d1y = 0.82604375110000003, d1z = 0.42415598719999998;
    const double d2x = 0.29031784669999999, d2y = -0.44782912349999999, d2z = 0.84567710930000001;
    const double n0 = 1.0 / sqrt(d0x * d0x + d0y * d0y + d0z * d0z), n1 = 1.0 / sqrt(d1x * d1x + d1y * d1y + d1z * d1z),
                 n2 = 1.0 / sqrt(d2x * d2x + d2y * d2y + d2z * d2z);
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        bool live = i < n;
        double px = 0, py = 0, pz = 0;
        if (live) {
            px = x[3 * i]; py = x[3 * i + 1]; pz = x[3 * i + 2];
            live = px >= P.blo[0] - NK_TOL && py >= P.blo[1] - NK_TOL && pz >= P.blo[2] - NK_TOL &&
                   px <= P.bhi[0] + NK_TOL && py <= P.bhi[1] + NK_TOL && pz <= P.bhi[2] + NK_TOL;
            if (!live) inside[i] = 0;
        }
        const bool a = nk_tiles_parity(tp, P, live, px, py, pz, d0x * n0, d0y * n0, d0z * n0) & 1u;
        const bool b = nk_tiles_parity(tp, P, live, px, py, pz, d1x * n1, d1y * n1, d1z * n1) & 1u;
        const bool tie = live && (a != b);
        bool res = a && b;
        if (__syncthreads_or(tie ? 1 : 0)) {
            const bool c = nk_tiles_parity(tp, P, tie, px, py, pz, d2x * n2, d2y * n2, d2z * n2) & 1u;
            if (tie) res = c;
        }
        if (live) inside[i] = res ? 1 : 0;
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
