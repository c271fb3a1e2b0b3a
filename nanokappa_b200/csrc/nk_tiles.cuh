// nk_tiles.cuh -- block-cooperative triangle tiles for Mesh.find_boundary (Mesh.py:806-856) on large meshes
// Part of the single translation unit nk_kernels.cu (included in this order: nk_tiles.cuh, nk_ops.cuh, nk_stream.cuh,
// nk_rare.cuh, nk_sort.cuh, nk_hostpipe.cuh); see DESIGN.md section 4.
//
// The reference intersects P rays with all F triangles through dense (P, F) temporaries (Population.py:810 chunks P).
// Here the rays of a block keep their running minimum in registers while the triangles stream through shared memory in
// tiles of NK_TILE_FACES records (208 B each), staged by the bulk async-copy engine (cp.async.bulk, SASS UBLKCP) into two
// stages guarded by one mbarrier each: the copy of tile t+1 overlaps the sweep of tile t.  Every thread reads the same
// triangle at the same time (shared-memory broadcast), so the kernel is bound by the FP64 pipe: ~23 FP64 instructions per
// (ray, triangle) in the divergence-free pre-filter of nk_ray_faces, the reference's exact arithmetic only for the few
// candidates that survive it.  A mesh of at most one tile is loaded once per block and stays
// resident.  Tiles are swept in face order with a strict `<`, so ties resolve to the lowest face index like np.argmax.
#pragma once

#define NK_TILE_FACES 96
#define NK_TILE_STAGE_BYTES (NK_TILE_FACES * (int)sizeof(NkFace))           // 19.5 KB
#define NK_TILE_SMEM_BYTES (2 * NK_TILE_STAGE_BYTES + 128)                  // two stages + the mbarriers

__device__ __forceinline__ unsigned int nk_smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
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

struct NkTilePipe {
    NkFace* buf;                     // two stages of NK_TILE_FACES faces (128-byte aligned shared memory)
    unsigned long long* bar;         // one mbarrier per stage
    unsigned int parity0, parity1;   // phase parities of the two stages (same value in every thread of the block)
    unsigned int* done;              // per stage: warps that have finished the tile it holds (the last one refills it)
    bool resident;                   // the whole mesh fits stage 0 and has been loaded
};

__device__ __forceinline__ void nk_tile_issue(const NkTilePipe& tp, const NkP& P, int stage, int tile) {
    const int f0 = tile * NK_TILE_FACES;
    const unsigned int bytes = (unsigned int)(min(NK_TILE_FACES, P.F - f0) * (int)sizeof(NkFace));
    nk_mbar_expect_tx(&tp.bar[stage], bytes);
    nk_bulk_g2s(tp.buf + (size_t)stage * NK_TILE_FACES, P.faces + f0, bytes, &tp.bar[stage]);
}

// `smem` points to NK_TILE_SMEM_BYTES of 128-byte aligned shared memory.  Every thread of the block calls this once.
__device__ __forceinline__ void nk_tiles_init(NkTilePipe& tp, unsigned char* smem, const NkP& P) {
    tp.buf = reinterpret_cast<NkFace*>(smem);
    tp.bar = reinterpret_cast<unsigned long long*>(smem + 2 * NK_TILE_STAGE_BYTES);
    tp.done = reinterpret_cast<unsigned int*>(smem + 2 * NK_TILE_STAGE_BYTES + 16);
    tp.parity0 = tp.parity1 = 0u;
    tp.resident = false;
    if (threadIdx.x == 0) {
        nk_mbar_init(&tp.bar[0], 1);
        nk_mbar_init(&tp.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (P.F <= NK_TILE_FACES) {
        if (P.F > 0) {
            if (threadIdx.x == 0) nk_tile_issue(tp, P, 0, 0);
            nk_mbar_wait(&tp.bar[0], 0u);
            tp.parity0 = 1u;
        }
        tp.resident = true;
    }
}

// Steady state of a sweep without block-wide barriers: a warp that has finished the tile in stage b bumps done[b]; the LAST
// warp to do so refills the stage with tile t + 2 (nobody reads it any more) and everybody else moves on to the other stage, so
// fast warps run up to one tile ahead of slow ones instead of waiting for them at a __syncthreads per tile.  A warp can only
// wait for fill k + 1 of a stage after every warp has consumed fill k (the refill is issued by the last consumer), so the
// parity of the mbarrier never aliases.
__device__ __forceinline__ void nk_tile_release(const NkTilePipe& tp, const NkP& P, int stage, int next_tile) {
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) {
        const unsigned int nwarps = blockDim.x >> 5;
        if (atomicAdd(&tp.done[stage], 1u) == nwarps - 1u) {
            atomicExch(&tp.done[stage], 0u);
            nk_tile_issue(tp, P, stage, next_tile);
        }
    }
}
__device__ __forceinline__ void nk_tiles_begin_sweep(const NkTilePipe& tp, const NkP& P, int nt) {
    __syncthreads();                                 // the readers of the previous sweep are done with both stages
    if (threadIdx.x == 0) {
        tp.done[0] = 0u; tp.done[1] = 0u;
        nk_tile_issue(tp, P, 0, 0);
        if (nt > 1) nk_tile_issue(tp, P, 1, 1);
    }
    __syncthreads();                                 // counters reset before anybody releases a stage
}

// All threads of the block call this together; threads with need == false only help to keep the pipeline moving.
__device__ __forceinline__ void nk_tiles_sweep(NkTilePipe& tp, const NkP& P, bool need, double x, double y, double z,
                                               double vx, double vy, double vz, double& tbest, int& fbest) {
    if (tp.resident) {
        if (need) {
            if (P.F <= 32) nk_ray_faces_small(tp.buf, P.F, P.mesh_scale, x, y, z, vx, vy, vz, tbest, fbest);
            else nk_ray_faces(tp.buf, P.F, P.mesh_scale, x, y, z, vx, vy, vz, tbest, fbest);
        }
        return;
    }
    const int F = P.F, nt = (F + NK_TILE_FACES - 1) / NK_TILE_FACES;
    nk_tiles_begin_sweep(tp, P, nt);
    for (int t = 0; t < nt; ++t) {
        const int b = t & 1;
        nk_mbar_wait(&tp.bar[b], b ? tp.parity1 : tp.parity0);
        if (b) tp.parity1 ^= 1u; else tp.parity0 ^= 1u;
        if (need) nk_ray_faces(tp.buf + (size_t)b * NK_TILE_FACES, min(NK_TILE_FACES, F - t * NK_TILE_FACES), P.mesh_scale, x, y, z, vx, vy, vz, tbest, fbest);
        if (t + 2 < nt) nk_tile_release(tp, P, b, t + 2);
    }
}
