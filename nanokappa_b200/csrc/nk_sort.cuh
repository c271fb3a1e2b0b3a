// nk_sort.cuh -- maintenance pass: order the particle SoA by mode, compact it, and set up the per-mode slot pools
// Part of the single translation unit nk_kernels.cu (included in this order: nk_tiles.cuh, nk_ops.cuh, nk_stream.cuh,
// nk_rare.cuh, nk_sort.cuh, nk_hostpipe.cuh); see DESIGN.md section 4.
//
// The streaming kernel gathers a 64-byte mode record and a row of the per-(mode, subvolume) table per particle; when
// the particles of a warp share their mode these gathers are one cache line.  The reference has no such notion (its
// arrays are in creation order), sums are order independent and particle identity is carried by `pid`, so the order of
// the slots is ours to choose.  This is a counting sort with M buckets:
//   k_sort_hist     live particles per mode                                    (4 B read per slot)
//   k_sort_scan     region start of every mode = exclusive scan of (count + spare slots)          (one block, M entries)
//   k_sort_rank     destination of every live slot: region start + arrival rank -> perm[dest] = slot
//   k_sort_permute  ONE fused gather that moves all 84 bytes of a slot from the front buffers to the back buffers
//   k_sort_pools    free-slot ring of every mode region <- its spare slots
// 4 + 4 + 4 + 84 + 84 bytes per particle in total; the caller swaps front and back buffers afterwards.
#pragma once

// spare slots appended to a mode region that holds `count` live particles
__host__ __device__ inline int nk_pool_spare(int count, double frac, int fixed) {
    if (count <= 0 || (frac <= 0.0 && fixed <= 0)) return 0;
    return fixed + (int)ceil((double)count * frac);
}

__global__ void __launch_bounds__(256) k_sort_hist(NkP P, int* __restrict__ count) {
    const long long n = P.dyn->n_slots;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int md = P.mode[i];
        if (md < 0) continue;
        // neighbours mostly share their mode: one atomic per group of equal modes in the warp
        const unsigned int act = __activemask();
        const unsigned int same = __match_any_sync(act, md);
        if ((int)(threadIdx.x & 31u) == __ffs(same) - 1) atomicAdd(count + md, __popc(same));
    }
}

// first[m] = sum_{m' < m} (count[m'] + spare(count[m'])), first[M] = total; cursor = copy of first (k_sort_rank advances it)
__global__ void __launch_bounds__(1024) k_sort_scan(int M, const int* __restrict__ count, int* __restrict__ first, int* __restrict__ cursor,
                                                    double frac, int fixed, long long* __restrict__ totals) {
    __shared__ long long part[1024];
    const int per = (M + 1023) / 1024;
    const int lo = min(M, (int)threadIdx.x * per), hi = min(M, lo + per);
    long long sum = 0, live = 0;
    for (int m = lo; m < hi; ++m) { sum += count[m] + nk_pool_spare(count[m], frac, fixed); live += count[m]; }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                       // Hillis-Steele inclusive scan of the per-thread sums
        long long v = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    long long run = part[threadIdx.x] - sum;                   // exclusive prefix of this thread's chunk
    for (int m = lo; m < hi; ++m) {
        first[m] = (int)run; cursor[m] = (int)run;
        run += count[m] + nk_pool_spare(count[m], frac, fixed);
    }
    if (threadIdx.x == 1023) { first[M] = (int)part[1023]; totals[0] = part[1023]; }
    // live total: a second tiny reduction
    __syncthreads();
    part[threadIdx.x] = live;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) { if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) totals[1] = part[0];
}

__global__ void __launch_bounds__(256) k_sort_rank(NkP P, int* __restrict__ cursor, int* __restrict__ perm) {
    const long long n = P.dyn->n_slots;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int md = P.mode[i];
        if (md < 0) continue;
        const unsigned int act = __activemask();
        const unsigned int same = __match_any_sync(act, md);
        const unsigned int lane = threadIdx.x & 31u;
        const int leader = __ffs(same) - 1;
        int base = 0;
        if ((int)lane == leader) base = atomicAdd(cursor + md, __popc(same));
        base = __shfl_sync(same, base, leader);
        perm[base + __popc(same & ((1u << lane) - 1u))] = (int)i;
    }
}

struct NkSoA {                    // one set of particle arrays (front or back buffers)
    double *px, *py, *pz, *tc, *occ, *cx, *cy, *cz;
    int *mode, *omode, *cfacet;
    long long* pid;
};

// dst[d] = src[perm[d]] for all twelve fields in one pass; spare slots (perm < 0) become free slots (mode = -1)
__global__ void __launch_bounds__(256) k_sort_permute(NkSoA src, NkSoA dst, const int* __restrict__ perm, long long n_dst) {
    for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < n_dst; d += (long long)gridDim.x * blockDim.x) {
        const int j = perm[d];
        if (j < 0) {
            dst.mode[d] = -1; dst.omode[d] = -1; dst.cfacet[d] = -1;
            dst.px[d] = 0.0; dst.py[d] = 0.0; dst.pz[d] = 0.0; dst.tc[d] = 0.0; dst.occ[d] = 0.0;
            dst.cx[d] = 0.0; dst.cy[d] = 0.0; dst.cz[d] = 0.0; dst.pid[d] = -1;
            continue;
        }
        dst.px[d] = src.px[j]; dst.py[d] = src.py[j]; dst.pz[d] = src.pz[j]; dst.tc[d] = src.tc[j]; dst.occ[d] = src.occ[j];
        dst.mode[d] = src.mode[j]; dst.omode[d] = src.omode[j]; dst.cfacet[d] = src.cfacet[j];
        dst.cx[d] = src.cx[j]; dst.cy[d] = src.cy[j]; dst.cz[d] = src.cz[j]; dst.pid[d] = src.pid[j];
    }
}

// ring 1 + m <- the spare slots at the end of region m; ring 0 (global) empty
__global__ void __launch_bounds__(256) k_sort_pools(int M, const int* __restrict__ count, const int* __restrict__ first,
                                                    long long* __restrict__ fr_ctr, int* __restrict__ freelist, int with_pools) {
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        long long* c = fr_ctr + 3 * (size_t)(1 + m);
        const int lo = first[m], live = count[m], spare = first[m + 1] - lo - live;
        c[0] = 0; c[1] = with_pools ? spare : 0; c[2] = with_pools ? spare : 0;
        if (with_pools) for (int k = 0; k < spare; ++k) freelist[lo + k] = lo + live + k;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { fr_ctr[0] = 0; fr_ctr[1] = 0; fr_ctr[2] = 0; }
}

// census for nk_set_slot_count: live particles, and every hole (mode < 0) goes onto the global free ring so that a caller's
// compacted-or-not arrays keep recycling their free slots
__global__ void __launch_bounds__(256) k_census(NkP P, unsigned long long* out) {
    const long long n = P.dyn->n_slots;
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool live = P.mode[i] >= 0;
        c += live;
        if (!live) {
            const unsigned long long k = nk_agg_inc(reinterpret_cast<unsigned long long*>(P.fr_ctr + 1));
            P.freelist[P.cap + (long long)(k % (unsigned long long)P.cap)] = (int)i;
        }
    }
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
__global__ void k_census_publish(NkP P) { P.fr_ctr[2] = P.fr_ctr[1]; }      // the holes are recyclable right away
