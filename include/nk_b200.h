/*
 * nk_b200.h -- C ABI of the B200-native per-timestep particle loop of Nano-kappa.
 *
 * The reference (brunohs1993/Nanokappa) is pure Python and has no FFI; the seam this library
 * replaces is the body of the Python methods on the hot path.  Each entry point
 * below cites the reference method whose arithmetic it performs on the GPU.  Conventions:
 *
 *   - plain C types only; every function returns 0 on success, <0 on error
 *     (message: nk_last_error).  Nothing throws across the ABI.
 *   - "host" pointers are read during the call and copied; "dev" pointers are CUDA device
 *     pointers owned by the caller (torch tensors in the Python host layer), borrowed until
 *     nk_destroy / the next nk_bind_particles.
 *   - one nk_ctx per GPU per process; a ctx is not thread-safe.  All work is enqueued on the
 *     stream given to nk_set_stream (default: the legacy default stream); calls that return a host
 *     scalar synchronise that stream, the others do not.
 *   - flat mode index m = q * J + j everywhere (reference keeps (q, j) pairs, Population.py:133).
 *   - boundary-condition codes: 0 = T (isothermal reservoir), 1 = P (periodic), 2 = R (rough),
 *     3 = F (flux reservoir, treated like T as upstream does, Population.py:1569).
 */
#ifndef NK_B200_H
#define NK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nk_ctx nk_ctx;

#define NK_BC_T 0
#define NK_BC_P 1
#define NK_BC_R 2
#define NK_BC_F 3

#define NK_INTERP_NEAREST 0   /* --temp_interp nearest (Population.py:570-573) */
#define NK_INTERP_LINEAR  1   /* --temp_interp linear, slice subvolumes only (Population.py:570-571) */
#define NK_INTERP_RADIAL  2   /* --temp_interp radial, and linear on non-slice subvolumes: scipy RBFInterpolator(kernel='cubic')
                                 (Population.py:574-588, :697-702); needs nk_set_rbf */

/* ---- lifetime ------------------------------------------------------------------------------- */
int         nk_create(int device, nk_ctx** out);
void        nk_destroy(nk_ctx* ctx);
const char* nk_last_error(nk_ctx* ctx);          /* ctx may be NULL: last error of nk_create */
int         nk_version(void);
int         nk_set_stream(nk_ctx* ctx, void* cuda_stream);   /* cudaStream_t */
int         nk_synchronize(nk_ctx* ctx);

/* ---- static tables (host pointers, copied) ---------------------------------------------------- */

/* Triangle planes and facets: Mesh.get_faces_properties / get_facets_properties / get_face_k
 * (Mesh.py:205-243, :244-308, :323-327) and the BC assignment of Geometry.get_bound_facets /
 * check_facet_connections (Geometry.py:652-726).
 *   face_*      : F triangles; basis = face_basis_matrix (F,3,3) row-major (columns b1, b2, n)
 *   face_vertices (F,3,3), face_areas (F): used by Mesh.sample_surface (Mesh.py:923-951)
 *   facet_*     : n_facets coplanar groups; partner = periodic partner or -1; res = reservoir index
 *                 or -1; rough = index into the rough-facet LUTs or -1
 *   facet_faces_ptr (n_facets+1), facet_faces: CSR list of the triangles of each facet
 *   bounds (2,3): mesh bounding box                                                            */
int nk_set_mesh(nk_ctx* ctx, int n_faces,
                const double* face_normals, const double* face_k,
                const double* face_lo, const double* face_hi,
                const double* face_origins, const double* face_basis,
                const int32_t* face_facets, const double* face_vertices, const double* face_areas,
                int n_facets, const int32_t* facet_bc, const int32_t* facet_partner,
                const int32_t* facet_res, const int32_t* facet_rough,
                const double* facet_normal, const double* facet_centroid, const double* facet_area,
                const int32_t* facet_faces_ptr, const int32_t* facet_faces,
                const double* bounds);

/* Subvolume centres and the particle-temperature rule: Geometry.set_subvolumes (Geometry.py:446-544),
 * SubvolClassifier (Geometry.py:1198-1213), Population.assign_temperatures (Population.py:570-590). */
int nk_set_subvols(nk_ctx* ctx, int n_subvols, const double* centres, const double* volumes,
                   int is_slice, int slice_axis, int temp_interp);
/* Cubic RBF temperature field of NK_INTERP_RADIAL (call after nk_set_subvols).  scipy's system
 * [[|c_i-c_j|^3, P], [P^T, 0]] depends on the subvolume centres only, so the host set-up inverts it once:
 * weights (S+n_dims+1, S) row-major = lhs^-1[:, :S]; every step the library forms coeffs = weights . T_sv and evaluates
 * T(x) = sum_s coeffs[s] |x-c_s|^3 + coeffs[S] + sum_k coeffs[S+1+k] (x[dims[k]] - shift[k]) / scale[k].
 * dims (n_dims <= 3): the coordinates the interpolator sees (a grid with a collapsed direction drops it,
 * Population.py:697-699); shift, scale (n_dims): scipy's polynomial normalisation. */
int nk_set_rbf(nk_ctx* ctx, int n_dims, const int32_t* dims, const double* shift, const double* scale,
               const double* weights);

/* Mode tables: Phonon.load_base_properties / calculate_lifetime / initialise_temperature_function
 * (Phonon.py:66-151, :326-336, :372-390).  omega (Q,J) rad THz, group_vel (Q,J,3) A THz,
 * tau (NT,Q,J) ps (0 where gamma <= 0), T_grid (NT) K, energy_array/T_array (nE) the E(T) table. */
int nk_set_phonon(nk_ctx* ctx, int Q, int J, int NT, const double* T_grid,
                  const double* omega, const double* group_vel, const double* tau,
                  double hbar, double kb, double volume_unitcell, int64_t n_active_modes,
                  int nE, const double* energy_array, const double* T_array);

/* Population constants (Population.py:35-125) and physical unit factors (Constants.py:7-12).
 * hot_T_lo/hi: expected temperature window; only used to choose which tau slabs are packed next to
 * each mode (any T outside still takes the exact full-table path). */
int nk_set_population(nk_ctx* ctx, double dt, int norm_mean, double particle_density,
                      int n_dt_to_conv, uint64_t seed, double eVpsa2_in_Wm2, double a_in_m,
                      double hot_T_lo, double hot_T_hi);

/* Reservoirs: Population.initialise_reservoirs / enter_probability (Population.py:146-161, :323-354).
 * enter_prob, res_counter (R,Q,J). */
int nk_set_reservoirs(nk_ctx* ctx, int R, const int32_t* res_facet, const double* res_T,
                      const double* enter_prob, const double* res_counter);
int nk_get_res_counter(nk_ctx* ctx, double* res_counter_host);
/* --reservoir_gen (Population.fill_reservoirs, Population.py:356-489); call after nk_set_reservoirs, default constant.
 *   constant   : fractional counters per (reservoir, mode)                                   (:358-406)
 *   fixed_rate : a fresh uniform per (reservoir, mode) and step instead of the counter       (:408-455)
 *   one_to_one : every reservoir re-emits what it absorbed in the previous step, modes from the roulette
 *                cumsum(enter_prob[r]) / max, entry time uniform in the step                 (:457-489)
 * n_leaving (R doubles, may be NULL): particles "absorbed" before the first step; NULL keeps the reference's initial
 * value round(sum(enter_prob[r])) (Population.py:344). */
#define NK_RESGEN_CONSTANT   0
#define NK_RESGEN_FIXED_RATE 1
#define NK_RESGEN_ONE_TO_ONE 2
int nk_set_reservoir_mode(nk_ctx* ctx, int mode, const double* n_leaving);

/* Rough-wall tables: calculate_fbz_specularity / find_specular_correspondences /
 * diffuse_scat_probability (Population.py:852-939, :1042-1461).  All (Fr, Q*J); spec_out is the
 * flat outgoing mode of specular_function(-n_f, q, j) (Population.py:959-961) or -1. */
int nk_set_boundary_luts(nk_ctx* ctx, int n_rough, const double* specularity,
                         const uint8_t* true_specular, const int32_t* spec_out, const double* roulette);

/* ---- particle state (device pointers, borrowed) ----------------------------------------------- */

/* The 13 per-particle arrays of Population.delete_particles (Population.py:838-850) reduced to the
 * ones that are state; omega / group_vel / wavevectors are gathered from the mode tables.
 *   px,py,pz   positions [A]            tc   n_timesteps (time to next collision, in dt)
 *   occ        occupation               mode flat mode (v_g, tau); -1 marks a free slot
 *   omode      flat mode whose omega the particle carries (differs from mode after a specular
 *              reflection, Population.py:955-971)
 *   cfacet,cx,cy,cz  collision_facets / collision_positions      pid  stable particle id        */
int nk_bind_particles(nk_ctx* ctx, int64_t capacity,
                      double* px, double* py, double* pz, double* tc, double* occ,
                      int32_t* mode, int32_t* omode, int32_t* cfacet,
                      double* cx, double* cy, double* cz, int64_t* pid);
/* Slots [0, n_slots) are in use or free (mode = -1).  Runs a census: counts the live particles and puts every free
 * slot of the range on the free-slot ring, so a caller that hands over arrays with holes (np.delete never ran) gets them
 * recycled by the emission.  Forgets the mode regions of nk_sort_by_mode. */
int nk_set_slot_count(nk_ctx* ctx, int64_t n_slots);
int nk_get_slot_count(nk_ctx* ctx, int64_t* n_slots, int64_t* n_alive);
/* Maintenance pass between timesteps (replaces Population.delete_particles' compaction, Population.py:832-850, which the
 * reference runs on every step): counting sort of the live particles by flat mode into a second set of arrays ("back
 * buffers", same capacity and alignment as nk_bind_particles), one fused pass over all twelve fields.  On return the back
 * buffers ARE the bound particle arrays and the previous ones are free for the next call.  Why by mode: the streaming
 * kernel's mode-table gathers of neighbouring particles then hit one cache line.  With pool_frac > 0 or pool_fixed > 0
 * every mode region gets pool_fixed + ceil(pool_frac * count) spare slots and its own free-slot ring: an emitted particle
 * takes a slot inside the region of its own mode, so the order survives emission and absorption (ignored when the
 * geometry has rough facets -- particles change mode in place there -- or when the spare slots do not fit the capacity).
 * Results are unchanged by the pass (sums are order independent, identity is `pid`). */
int nk_sort_by_mode(nk_ctx* ctx, double* px, double* py, double* pz, double* tc, double* occ,
                    int32_t* mode, int32_t* omode, int32_t* cfacet, double* cx, double* cy, double* cz, int64_t* pid,
                    double pool_frac, int pool_fixed, int64_t* n_slots_out, int64_t* n_alive_out);
int nk_set_sv_temperature(nk_ctx* ctx, const double* T_sv_host);   /* Population.subvol_temperature */
int nk_get_sv_temperature(nk_ctx* ctx, double* T_sv_host);
int nk_set_timestep(nk_ctx* ctx, int64_t current_timestep);
int nk_get_timestep(nk_ctx* ctx, int64_t* current_timestep);

/* ---- operator seams (device pointers; (n,3) arrays are row-major like the reference's) ---------- */

/* Mesh.find_boundary(x, v) -> (xc, tc, fc)   Mesh.py:806-856.  fc int32, -1 = no hit (tc = inf). */
int nk_find_boundary(nk_ctx* ctx, int64_t n, const double* x, const double* v,
                     double* xc, double* tc, int32_t* fc);
/* Mesh.contains_naive(x)   Mesh.py:785-804 (point-in-solid by crossing parity; set-up at scale: rejection sampling of initial
 * positions, Population.py:209-246).  inside (n) uint8: 1 inside, 0 outside.  All faces count (no interface faces). */
int nk_contains(nk_ctx* ctx, int64_t n, const double* x, uint8_t* inside);
/* SubvolClassifier.predict(x) + Population.get_subvol_id counts  Geometry.py:1212, Population.py:671-683.
 * counts (n_subvols) int64 may be NULL. */
int nk_classify(nk_ctx* ctx, int64_t n, const double* x, int32_t* sv, int64_t* counts);
/* Phonon.calculate_occupation(T, omega)   Phonon.py:338-345 */
int nk_occupation(nk_ctx* ctx, int64_t n, const double* T, const double* omega, double* occ);
/* Phonon.lifetime_function([T, q, j])     Phonon.py:326-336 */
int nk_lifetime(nk_ctx* ctx, int64_t n, const double* T, const int32_t* mode, double* tau);
/* Phonon.temperature_function(E) / crystal_energy_function(T)   Phonon.py:387-390 */
int nk_temperature_of_energy(nk_ctx* ctx, int64_t n, const double* E, double* T);
int nk_energy_of_temperature(nk_ctx* ctx, int64_t n, const double* T, double* E);
/* temperature_interpolator(x) of Population.refresh_temperatures with the ctx's current T_sv
 * (Population.py:694-702) */
int nk_particle_temperature(nk_ctx* ctx, int64_t n, const double* x, double* T);

/* ---- the timestep ------------------------------------------------------------------------------ */

/* Population.timesteps_to_boundary for every live particle + get_collision_condition
 * (Population.py:308-316): fills tc, cfacet, cx,cy,cz from the bound SoA. */
int nk_init_collisions(nk_ctx* ctx);

/* n_steps x Population.run_timestep without the every-100-step output branch
 * (Population.py:1743-1769): drift, reservoir emission (fill_reservoirs + add_reservoir_particles),
 * boundary_scattering, refresh_temperatures, lifetime_scattering, and every n_dt_to_conv steps
 * calculate_heat_flux / calculate_kappa / adjust_reservoir_balance.  Enqueued; no host sync. */
int nk_step(nk_ctx* ctx, int n_steps);

/* Per-kernel device timing of the timestep (CUDA events recorded on the ctx stream around every
 * launch of nk_step / nk_step_local / nk_step_finalize between begin and end; used by bench.py for
 * the roofline of the streaming kernel).  ms[4] = total milliseconds in k_step, k_emit, k_boundary,
 * k_finalize; n_steps = timesteps covered.  nk_profile_end synchronises the stream. */
int nk_profile_begin(nk_ctx* ctx);
int nk_profile_end(nk_ctx* ctx, double* ms, int64_t* n_steps);

/* Make `occ` hold the post-lifetime_scattering occupation (the step kernel defers the relaxation of
 * step k to the head of step k+1; outputs that read occupation call this first). */
int nk_flush_relaxation(nk_ctx* ctx);

/* Per-subvolume results of the last completed step (host buffers, any may be NULL):
 *   T_sv, E_sv (S)  subvol_temperature / subvol_energy   Population.py:692, :704-728
 *   N_sv (S) int64  subvol_N_p                            Population.py:679
 *   flux (S,3)      subvol_heat_flux of the last convergence step [W/m2]  Population.py:730-747
 *   kappa_sv (S), kappa (1)                               Population.py:749-771 (slice)
 *   res_E_bal (R), res_flux (R,3): adjust_reservoir_balance values of the last convergence step
 *   N_leaving (R) int64: Population.N_leaving of the last step
 *   total_energy (1): sum of particle energies [eV]       Population.py:2034                      */
int nk_get_results(nk_ctx* ctx, double* T_sv, double* E_sv, int64_t* N_sv, double* flux,
                   double* kappa_sv, double* kappa, double* res_E_bal, double* res_flux,
                   int64_t* N_leaving, double* total_energy);

/* The same without stalling the device: nk_snapshot_results enqueues a copy of the results block as it will be after the steps
 * enqueued so far (pinned host ring of four; *ticket identifies it) and returns at once, so the caller can enqueue the next
 * batch of steps before it looks at the numbers; nk_get_snapshot waits for that copy only.  This is how the command line
 * writes its convergence.txt rows (every n_dt_to_conv steps, Population.py:1762-1767) while the GPU keeps stepping. */
int nk_snapshot_results(nk_ctx* ctx, int* ticket);
int nk_get_snapshot(nk_ctx* ctx, int ticket, double* T_sv, double* E_sv, int64_t* N_sv, double* flux,
                    double* kappa_sv, double* kappa, double* res_E_bal, double* res_flux,
                    int64_t* N_leaving, double* total_energy);

/* Population.contains_check (Population.py:1712-1722), detection half: the slots of the live particles that lie outside
 * the mesh bounding box by more than tol (upstream: 1e-10) are written to slots_dev (device, capacity `cap`; unordered),
 * their number to *n_found (may exceed cap: call again with a larger buffer).  The caller re-draws those particles
 * (Mesh.sample_volume) and gives them a first collision through nk_find_boundary. */
int nk_outside_slots(nk_ctx* ctx, double tol, int32_t* slots_dev, int64_t cap, int64_t* n_found);

/* ---- set-up helper (no ctx) ------------------------------------------------------------------------------------ */

/* Phonon.initialise_temperature_function's energy table (Phonon.py:352-384) on the device: out[i] =
 * sum over active modes of hbar*omega*n0(T[i], omega) / dens_norm + zero_point.  Host pointers; omega, active (M).
 * The sum order differs from NumPy's (1e-16 relative).  Error text: nk_last_error(NULL). */
int nk_energy_table(int device, int n_modes, const double* omega, const uint8_t* active, int nT, const double* T,
                    double hbar, double kb, double dens_norm, double zero_point, double* out);

/* ---- host-buffer end-to-end call ---------------------------------------------------------------- */

/* Upload n particles (host SoA, pinned or pageable), run n_steps, download the state and the per-SV
 * results.  This is the call bench.py times as `e2e`: host<->device copies are inside it.
 * The host arrays are the caller's copy of Population's particle arrays (Population.py:1724-1800 reads and writes
 * self.positions, self.n_timesteps, self.occupation, self.modes, ... in place); on return they hold the state after
 * the step(s).  For one step of >= 2^20 particles the call is pipelined: slots are cut into chunks, chunk c+1 is
 * uploaded while chunk c streams through the kernel and chunk c-1's positions/clocks are downloaded.  Only the fields the
 * streaming kernel reads travel densely; collision facet / position and ids go up as a sparse patch for the particles
 * that collide in this step (found by a host scan of tc < 1) and come back as a sparse patch of the slots the rare path
 * touched.  During the call the device arrays are a scratch image of the host arrays: do not mix this entry point with
 * device-resident stepping (nk_step) on the same ctx without re-uploading the state.  NK_HOST_SPARSE=0 uploads the cold
 * arrays densely, NK_HOST_PIPELINE=0 selects the plain upload-all / step / download-all sequence; all three leave
 * identical host arrays. */
int nk_advance_host(nk_ctx* ctx, int64_t n_in, int n_steps,
                    double* px, double* py, double* pz, double* tc, double* occ,
                    int32_t* mode, int32_t* omode, int32_t* cfacet,
                    double* cx, double* cy, double* cz, int64_t* pid,
                    int64_t* n_out, double* T_sv_out, double* E_sv_out, int64_t* N_sv_out);
/* Bytes the last nk_advance_host call moved over PCIe in each direction (counted from the copies it issued). */
int nk_last_transfer_bytes(nk_ctx* ctx, int64_t* h2d, int64_t* d2h);
/* Diagnostics (NK_TRACE=1 in the environment of nk_create): %globaltimer marks [ns] of the last step, then cleared:
 * 0 streaming kernel first block in, 1 last block out, 2 rare-path kernel first block in, 3 last item done,
 * 4 closing block enters the finalize, 5 leaves it; 6-7 unused. */
int nk_debug_trace(nk_ctx* ctx, uint64_t* out8);
/* Which kernels the last step used: 0 direct streaming kernel, 4 per-(mode, subvolume) table variant; + 8 when the rare path
 * ran with block-cooperative triangle tiles (meshes beyond 128 triangles). */
int nk_last_step_variant(nk_ctx* ctx);

/* ---- multi-GPU ------------------------------------------------------------------------------------ */

/* Particle shards: each rank owns a block of particles; every rank advances the whole reservoir table and the copies an
 * entry emits are dealt round-robin over the ranks (copy k of an entry that has emitted `fire` particles so far belongs to
 * rank (fire + k + mode) % world), so each rank injects 1/world of every mode and the union over ranks is exactly the
 * single-GPU emission; the only exchange is the per-step sum of the per-SV / per-reservoir accumulators.
 * nk_acc_buffer exposes that vector (device pointer + length in doubles) so the host layer can run
 * ncclAllReduce on it between the two halves of a step (nk_step_local / nk_step_finalize), or
 * register peer buffers for the fused one-shot exchange (nk_comm_*). */
int nk_set_rank(nk_ctx* ctx, int rank, int world);
int nk_acc_buffer(nk_ctx* ctx, double** dev_ptr, int64_t* n_doubles);
/* Run state that is neither a table nor a particle, for checkpoints that continue bit-exactly without rebuilding the
 * tables: res_counter (R, Q*J) f64; res_fire (R, Q*J) u8 emission deal counters; res_acc (4R) energy-balance / flux
 * accumulators of the current convergence window (Population.res_energy_balance / res_heat_flux between two
 * convergence rows); n_leaving (R) f64 Population.N_leaving of the previous step (feeds one_to_one); results
 * (nk_results_len doubles) the block behind nk_get_results.  Host pointers; NULL skips a field. */
int nk_results_len(nk_ctx* ctx);
int nk_get_run_state(nk_ctx* ctx, double* res_counter, uint8_t* res_fire, double* res_acc, double* n_leaving, double* results);
int nk_set_run_state(nk_ctx* ctx, const double* res_counter, const uint8_t* res_fire, const double* res_acc,
                     const double* n_leaving, const double* results);
int nk_step_local(nk_ctx* ctx);      /* kernels of one step up to the accumulators */
int nk_step_finalize(nk_ctx* ctx);   /* accumulators -> T_sv, results; closes the step */
/* Fused exchange over NVLink peer memory (replaces the all-reduce above): every rank exports its mailbox as a
 * 64-byte CUDA IPC handle, imports the handles of all ranks (own rank included, in rank order) and enables the
 * exchange; nk_step then needs no collective: the block that closes a step stores the rank's sums into every
 * mailbox, waits (bounded) for the peers' and adds them in rank order.  One box, <= 8 ranks. */
int nk_comm_export(nk_ctx* ctx, void* handle_out_64B);
int nk_comm_import(nk_ctx* ctx, int peer_rank, const void* handle_64B);
int nk_comm_enable(nk_ctx* ctx, int enable);   /* enable = 0 also unmaps the peers' mailboxes: do it on every rank
                                                  (then synchronise the ranks) before any rank calls nk_destroy */

#ifdef __cplusplus
}
#endif
#endif /* NK_B200_H */
