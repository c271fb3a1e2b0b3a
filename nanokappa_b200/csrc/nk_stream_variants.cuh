// nk_stream_variants.cuh -- A/B variants of the streaming kernel kept behind NK_STEP_IMPL (cp.async prefetch, one particle per thread, TMA pipeline) and the variant picker
// Part of the single translation unit nk_kernels.cu (included in this order: nk_ops.cuh, nk_stream.cuh,
// nk_stream_variants.cuh, nk_rare.cuh, nk_hostpipe.cuh); see DESIGN.md section 4.
#pragma once

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
