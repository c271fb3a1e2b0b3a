// nk_hostpipe.cuh -- kernels of the host-buffer pipeline and of contains_check (dirty-slot patch, cold-field patch, escapees)
// Part of the single translation unit nk_kernels.cu (included in this order: nk_ops.cuh, nk_stream.cuh,
// nk_stream_variants.cuh, nk_rare.cuh, nk_hostpipe.cuh); see DESIGN.md section 4.
#pragma once

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
