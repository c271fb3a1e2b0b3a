// nk_types.cuh -- device-visible parameter blocks of the particle loop.
//
// NkP is passed BY VALUE to every kernel (constant bank, no pointer chasing); it only holds
// pointers and scalars that never change between nk_set_* calls, so a captured CUDA graph of a
// timestep stays valid.  Everything that changes from step to step lives in NkDyn (device memory).
#pragma once
#include <stdint.h>

#define NK_TOL 1e-10            // Mesh.py:24
#define NK_EMIT_CMAX 64         // copies of one mode a reservoir may emit per step (id packing)
#define NK_EMIT_ID_BASE (1LL << 62)

// one triangle, everything find_boundary needs (Mesh.py:806-856), 26 doubles = 208 B (a multiple of 16: bulk-copyable)
struct NkFace {
    double nx, ny, nz, k;        // plane: x.n + k = 0
    double mx, my, mz;           // centre of the widened AABB                     } conservative pre-filter of nk_ray_faces:
    double ex, ey, ez;           // its half extents + 1e-6 x mesh scale           } never decides a hit, only skips faces
    double lox, loy, loz;        // face AABB already widened by -tol / +tol (Mesh.py:828-829)
    double hix, hiy, hiz;
    double ox, oy, oz;           // vertex 0
    double ia0, ia1, ia2;        // row 0 of [b1 b2 n]^-1  -> barycentric a
    double ib0, ib1, ib2;        // row 1                  -> barycentric b
    double facet;                // facet id stored as double to keep the record homogeneous
};

// per-mode record gathered once per particle per step: exactly one 32 B sector
struct __align__(32) NkMode {
    double omega, vx, vy, vz;
};

// four consecutive tau(T) slabs of a mode starting at NkP::tau_i0: one 32 B sector
struct __align__(32) NkTau4 {
    double t[4];
};

// streaming-kernel record: one 64 B line segment per mode = {omega, v_g} + tau slabs tau_i0..tau_i0+3
struct __align__(64) NkModeHot {
    double omega, vx, vy, vz;
    double t[4];
};

// Dense hit record: what the streaming kernel holds in registers when a particle's collision falls inside the step.
// The rare path reads these 64 B coalesced instead of gathering seven scattered sectors per hit.
struct __align__(32) NkHitRec {
    double x, y, z, tc;          // position after the drift, clock after the decrement (< 0)
    double occ;                  // occupation after the deferred relaxation
    int slot, mode;
    int omode, pad0;
    double pad1;
};

struct alignas(128) NkDyn {      // device-resident, mutated by kernels
    // line 0: fields touched once per block or once per step
    long long fr_snap;           // free-slot ring: fr_tail at the end of the previous step (pop limit)
    long long n_alive;
    long long step;              // Population.current_timestep
    unsigned int n_emit;         // emission-list entries of this step
    unsigned int last_hits;      // n_hits / n_new of the step that was closed last (dirty-slot list for host patches)
    unsigned int last_new;
    int relax_pending;           // lifetime_scattering of step-1 still to be applied to `occ`
    int error;                   // sticky device-side error bits
    unsigned int blocks_done;    // last-block detection
    char pad0[128 - 3 * 8 - 6 * 4];
    // the counters the rare path hammers with atomics get a 128-byte line (= an L2 slice queue) each
    long long n_slots;           // slots [0, n_slots) are live or on the free list
    char pad1[120];
    long long fr_head;           // free-slot ring: next entry to recycle (may overshoot fr_snap inside a step; clamped by the finalize)
    char pad2[120];
    long long fr_tail;           //                 next entry to write (slots freed by absorption)
    char pad3[120];
    unsigned int n_hits;         // particles whose collision falls inside this step
    char pad4[124];
    unsigned int n_new;          // slots filled by emission in this step (newslots list)
    char pad5[124];
};

#define NK_ERR_CAPACITY 1        // emission ran out of slots
#define NK_ERR_EVENTS   2        // a particle exceeded the per-step event cap
#define NK_ERR_CMAX     4        // a mode emitted more than NK_EMIT_CMAX copies in one step
#define NK_ERR_COMM     8        // a peer did not deliver its sums in time (fused exchange)

struct NkP {
    // ---- mesh
    int F, nf;
    const NkFace* faces;
    const int* facet_bc; const int* facet_partner; const int* facet_res; const int* facet_rough;
    const double* facet_normal; const double* facet_centroid; const double* facet_area;
    // reservoir surface sampling (Mesh.sample_surface): per reservoir a CSR slice of faces + area cdf
    const int* res_face_ptr; const int* res_faces; const double* res_face_cdf;
    const double* face_vertices;      // (F,3,3)
    double blo[3], bhi[3];
    double mesh_scale;                // max |coordinate| / |plane offset| of the mesh (error bounds of the ray pre-filter)
    // ---- subvolumes
    int S, is_slice, axis, interp;
    const double* svc;                // (S,3)
    const double* sv_axis;            // (S)   centres on the slice axis
    const double* sv_mid;             // (S-1) x/2+x/2 midpoints (scipy interp1d 'nearest')
    const double* sv_volume;
    double sv_inv_dx;                 // 1 / slice spacing (guess only; exactness comes from fix-up)
    double sv_x0;                     // lower end of slice 0 on the slice axis (guess only)
    // cubic RBF temperature field (--temp_interp radial, nk_set_rbf): T(x) = sum_s coef[s] |x-c_s|^3 + coef[S] +
    // sum_k coef[S+1+k] (x_k - shift_k)/scale_k over the rbf_nd coordinates rbf_dim[]; coef = rbf_w . T_sv
    int rbf_nd; int rbf_dim[3];
    double rbf_shift[3], rbf_scale[3];
    const double* rbf_w;              // (S + rbf_nd + 1, S) row-major
    double* rbf_coef;                 // (S + rbf_nd + 1), refreshed whenever T_sv changes
    // ---- modes
    int Q, J, M, NT;
    const double* Tg;                 // (NT)
    const NkMode* mprop;              // (M)
    const double* tau;                // (NT, M)
    const NkTau4* tau4;               // (M) slabs tau_i0 .. tau_i0+3
    const NkModeHot* mhot;            // (M) hot record of the streaming kernel
    double2* hot_tab;                 // (M, S) {n0(T_sv), exp(-dt/tau(T_sv))} rebuilt every step (slice + nearest T), or null
    int tau_i0;
    double Tg_inv_d;                  // 1 / (Tg[1]-Tg[0]) guess
    int Tg_uniform; double Tg0, Tg_d; // Tg[i] == Tg0 + i * Tg_d exactly (phono3py's 0, 10, ... 1000 K): bracket by arithmetic
    int nE; const double* Ea; const double* Ta;
    double Ta_inv_d;                  // (nE-1)/(Ta[nE-1]-Ta[0]): index guess into the E(T) table
    double hbar, kb, V_uc, n_active, dens_norm;   // dens_norm = Q * V_uc
    // ---- population
    double dt, particle_density, eVpsa2_in_Wm2, a_in_m;
    int norm_mean, n_dt_to_conv;
    unsigned int seed_lo, seed_hi;
    // ---- reservoirs
    int R; const int* res_facet; const double* res_T; const double* enter_prob; double* res_counter;
    unsigned char* res_fire;          // (R, M) particles emitted so far by each table entry (mod 256): copy k of an entry
                                      // belongs to rank (fire + k + mode) % world, so every rank injects 1/world of EVERY mode
    int res_gen;                      // NK_RESGEN_* (--reservoir_gen)
    double* emit_u;                   // (R, M) this step's dice (fixed_rate)
    const double* res_roulette;       // (R, M) cumsum(enter_prob[r]) / max (one_to_one)
    double* res_nleave;               // (R) particles absorbed per reservoir in the previous step (one_to_one)
    // ---- rough-wall LUTs (Fr, M)
    int has_rough;                    // 0: no rough facet, the omega-carrying mode is always the mode itself (omode == mode)
    int Fr; const double* specularity; const unsigned char* true_spec; const int* spec_out; const double* roulette;
    // guide table of the diffuse roulette (only when every row is non-decreasing): rou_guide[f][b] = first index whose cumulative
    // rate reaches (b / rou_guide_k) x total, so the bisection starts inside a bracket of ~M / rou_guide_k entries
    const int* rou_guide; int rou_guide_k;
    // ---- particles (borrowed)
    long long cap;
    double *px, *py, *pz, *tc, *occ;
    int *mode, *omode, *cfacet;
    double *cx, *cy, *cz;
    long long* pid;
    // ---- scratch owned by the ctx
    int* hitlist; int* freelist;
    // Free slots are kept in rings with counters {head, tail, snap} in fr_ctr.  Ring 0 is the global ring (entries
    // freelist[cap, 2 cap)).  Once nk_sort_by_mode has ordered the particles by mode (fr_sorted > 0) every mode m owns the
    // slot region [mode_first[m], mode_first[m+1]) -- its live particles followed by a few spare slots -- and ring 1 + m
    // (entries freelist[mode_first[m], mode_first[m+1])) recycles exactly those slots: an absorbed particle frees a slot of
    // its region, the next emitted particle of that mode takes it, so the order by mode -- and with it the locality of the
    // mode-table gathers -- survives emission and absorption.  Slots beyond fr_sorted recycle through the global ring.
    long long* fr_ctr; int n_rings; const int* mode_first; long long fr_sorted;
    NkHitRec* hitrec; long long hitrec_cap;   // dense records of the first hitrec_cap hit-list entries of a step
    int2* emitlist;                   // (R*M) {reservoir << 8 | copies, mode} of the entries emitting this step
    int* newslots; long long newslots_cap;   // slots that received an emitted particle in this step
    long long slot_lo, slot_hi;       // slot range the streaming kernel covers in this launch (chunked host pipeline)
    int scan_emit;                    // 1: this launch advances the reservoir counters
    double* T_sv;                     // (S) current subvolume temperatures
    double* acc;                      // per-step accumulators, see layout below (f64: the exchange vector + side bins)
    unsigned long long* acc_q;        // the same entries in 128-bit fixed point {lo, hi}: order-independent block merges
    double* res_acc;                  // (R*4) E_bal + flux accumulated over the convergence window
    double* out;                      // results block, see NK_OUT_* offsets
    NkDyn* dyn;
    unsigned long long* trace;        // NK_TRACE=1: %globaltimer marks of the last step (see nk_debug_trace), else null
    int rank, world;
    // ---- fused exchange of the accumulator vector over NVLink peer memory (nk_comm_*)
    int comm_on;                      // 1: the last block of k_rare all-reduces P.acc through the mailboxes
    double* mbox_local;               // [2][world][acc_len] written by the peers (parity, sender)
    unsigned long long* flags_local;  // [2][world] sequence numbers written by the peers
    double* peer_mbox[8];             // every rank's mailbox (own entry = mbox_local)
    unsigned long long* peer_flags[8];
};

// accumulator layout (all double so that ONE f64 all-reduce covers it)
//   [0,S) sum e   [S,2S) count   [2S,5S) sum v*e   then per reservoir: N_leaving, E_bal, flux(3)
//   then 2 scalars: particles emitted, particles absorbed
__host__ __device__ inline int nk_acc_len(int S, int R) { return 5 * S + 5 * R + 2; }
#define NK_ACC_E(S, R)      0
#define NK_ACC_CNT(S, R)    (S)
#define NK_ACC_FLUX(S, R)   (2 * (S))
#define NK_ACC_NLEAVE(S, R) (5 * (S))
#define NK_ACC_EBAL(S, R)   (5 * (S) + (R))
#define NK_ACC_RFLUX(S, R)  (5 * (S) + 2 * (R))
#define NK_ACC_NEMIT(S, R)  (5 * (S) + 5 * (R))
#define NK_ACC_NABS(S, R)   (5 * (S) + 5 * (R) + 1)

// results block layout (doubles)
//   T_sv(S) E_sv(S) N_sv(S) flux(3S) kappa_sv(S) kappa(1) res_E_bal(R) res_flux(3R) N_leaving(R) total_energy(1) N_p(1)
__host__ __device__ inline int nk_out_len(int S, int R) { return 7 * S + 5 * R + 3; }
#define NK_OUT_T(S, R)      0
#define NK_OUT_E(S, R)      (S)
#define NK_OUT_N(S, R)      (2 * (S))
#define NK_OUT_FLUX(S, R)   (3 * (S))
#define NK_OUT_KSV(S, R)    (6 * (S))
#define NK_OUT_KAPPA(S, R)  (7 * (S))
#define NK_OUT_REBAL(S, R)  (7 * (S) + 1)
#define NK_OUT_RFLUX(S, R)  (7 * (S) + 1 + (R))
#define NK_OUT_NLEAVE(S, R) (7 * (S) + 1 + 4 * (R))
#define NK_OUT_ETOT(S, R)   (7 * (S) + 1 + 5 * (R))
#define NK_OUT_NP(S, R)     (7 * (S) + 2 + 5 * (R))
