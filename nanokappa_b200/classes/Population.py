"""Particle population with the reference's ``Population`` surface, stepped on the GPU.

``Population(args, geometry, phonon)`` / ``run_timestep(geometry, phonon)`` / ``write_final_state`` /
``current_timestep`` / ``finish_sim`` / ``f`` / ``view`` behave as in ``classes/Population.py`` of the
reference; the per-particle arrays (``positions``, ``modes``, ``occupation``, ``omega``,
``group_vel``, ``n_timesteps``, ``collision_facets``, ``collision_positions``, ``subvol_id``,
``temperatures``) live on the device and materialise as NumPy arrays when read.

Set-up (particle count, reservoir entry probabilities, rough-wall tables, initial positions / modes /
occupation, output files) is host NumPy; everything inside ``run_timestep`` except the every-10 /
every-100-step text output runs in the CUDA kernels behind ``nanokappa_b200.engine.Engine``.
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

from .Constants import Constants
from .Visualisation import Visualisation
from ..engine import Engine
from ..routines import boundary_tables
from ..routines.rbf import interp_dims

BC_CODE = {'T': 0, 'P': 1, 'R': 2, 'F': 3}


def build_tables(args, geometry, phonon, pop):
    """The plain-array description of one run that ``Engine.set_tables`` uploads (see engine.py)."""
    mesh = geometry.mesh
    nf = mesh.n_of_facets
    tb = dict(face_normals=mesh.face_normals, face_k=mesh.face_k, face_lo=mesh.face_bounds[0], face_hi=mesh.face_bounds[1],
              face_origins=mesh.face_origins, face_basis=mesh.face_basis_matrix, face_facets=mesh.face_facets,
              face_vertices=mesh.vertices[mesh.faces], face_areas=mesh.face_areas,
              facet_bc=np.array([BC_CODE[c] for c in geometry.bound_cond]), facet_normal=geometry.facets_normal,
              facet_centroid=geometry.facet_centroid, facet_area=geometry.facets_area, bounds=geometry.bounds)
    partner = -np.ones(nf, dtype=int)
    for a, b in np.asarray(geometry.connected_facets, dtype=int).reshape(-1, 2):
        if partner[a] < 0:
            partner[a] = b
        if partner[b] < 0:
            partner[b] = a
    tb['facet_partner'] = partner
    fres = -np.ones(nf, dtype=int); fres[geometry.res_facets] = np.arange(len(geometry.res_facets))
    frough = -np.ones(nf, dtype=int); frough[geometry.rough_facets] = np.arange(len(geometry.rough_facets))
    tb['facet_res'], tb['facet_rough'] = fres, frough
    ptr, flat = [0], []
    for fct in mesh.facets:
        flat.extend(int(i) for i in fct); ptr.append(len(flat))
    tb['facet_faces_ptr'], tb['facet_faces'] = np.array(ptr), np.array(flat)
    tb.update(sv_centres=geometry.subvol_center, sv_volume=geometry.subvol_volume, sv_slice=geometry.subvol_type == 'slice',
              slice_axis=int(getattr(geometry, 'slice_axis', 0)), temp_interp=pop.temp_interp_type,
              interp_dims=interp_dims(geometry), res_gen=pop.res_gen)
    tb.update(omega=phonon.omega, group_vel=phonon.group_vel, tau=phonon.lifetime, T_grid=phonon.temperature_array,
              energy_array=phonon.energy_array, T_array=phonon.T_array, hbar=phonon.hbar, kb=phonon.kb,
              volume_unitcell=phonon.volume_unitcell, n_active=int(phonon.number_of_active_modes),
              eVpsa2_in_Wm2=pop.eVpsa2_in_Wm2, a_in_m=pop.a_in_m)
    Q, J = phonon.omega.shape
    R = pop.n_of_reservoirs
    tb.update(dt=pop.dt, norm_mean=pop.norm == 'mean', particle_density=pop.particle_density, n_dt_to_conv=pop.n_dt_to_conv,
              res_facet=np.asarray(pop.res_facet, dtype=int) if R else np.zeros(0, dtype=int),
              res_T=np.asarray(pop.res_facet_temperature, dtype=float) if R else np.zeros(0),
              enter_prob=pop.enter_prob if R else np.zeros((0, Q, J)),
              specularity=pop.specularity, true_specular=pop.true_specular, spec_out=pop.spec_out, roulette=pop.creation_roulette)
    return tb


class PopulationSetup(Constants):
    '''Host-only part of the set-up: particle count, reservoir tables, rough-wall LUTs.  Needs no GPU;
    bench.py's reference arm and the CPU tests use it to build the run tables for the oracle.'''

    def __init__(self, arguments, geometry, phonon, seed=None):
        super().__init__()
        self.setup_host(arguments, geometry, phonon, seed)

    def setup_host(self, arguments, geometry, phonon, seed=None):
        self.args = arguments
        self.results_folder_name = self.args.results_folder
        os.makedirs(self.results_folder_name, exist_ok=True)
        self.n_dt_to_conv = 10
        self.norm = self.args.energy_normal[0]
        self.n_of_subvols = geometry.n_of_subvols
        self.empty_subvols = list(self.args.empty_subvols)
        self.n_of_empty_subvols = len(self.empty_subvols)
        self.particle_type = self.args.particles[0]
        n_act = phonon.number_of_active_modes
        if self.particle_type == 'pmps':
            self.particles_pmps = float(self.args.particles[1])
            self.N_p = int(np.ceil(self.particles_pmps * n_act * self.n_of_subvols))
            self.particle_density = self.N_p / geometry.volume
        elif self.particle_type == 'total':
            self.N_p = int(np.ceil(float(self.args.particles[1])))
            self.particles_pmps = self.N_p / (n_act * self.n_of_subvols)
            self.particle_density = self.N_p / geometry.volume
        elif self.particle_type == 'pv':
            self.particle_density = float(self.args.particles[1])
            self.N_p = int(np.ceil(self.particle_density * geometry.volume))
            self.particles_pmps = self.N_p / (n_act * (self.n_of_subvols - self.n_of_empty_subvols))
        self.dt = float(self.args.timestep[0])
        self.t = 0.0
        if geometry.subvol_type == 'slice':
            self.slice_axis = geometry.slice_axis
            self.slice_length = geometry.slice_length
        self.subvol_volume = geometry.subvol_volume
        self.bound_cond = geometry.bound_cond
        self.res_gen = self.args.reservoir_gen[0]
        if self.args.reference_temp[0] != 'local':
            raise Exception('--reference_temp with a fixed value is a debug mode that is not on the GPU path (local only).')
        self.T_reference = 'local'
        self.rough_facets = geometry.rough_facets
        self.rough_facets_values = geometry.rough_facets_values
        self.connected_facets = geometry.connected_facets
        self.T_distribution = self.args.temp_dist[0]
        self.temp_interp_type = self.args.temp_interp[0]
        if self.temp_interp_type not in ('nearest', 'linear', 'radial'):
            raise Exception('Invalid T interpolator type.')
        if self.temp_interp_type == 'linear' and geometry.subvol_type != 'slice':
            print('Linear T interpolation is currently valid for slice subvolumes only. Defaulting to RBF interpolation to avoid extrapolation problems.')
            self.temp_interp_type = 'radial'
        if self.temp_interp_type == 'radial' and geometry.subvol_type == 'slice':
            raise Exception('Radial T interpolation needs grid or voronoi subvolumes: the centres of slices are collinear and the RBF system is singular.')
        self.colormap = self.args.colormap[0]
        self.fig_plot = self.args.fig_plot
        self.current_timestep = 0
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)

        print('Calculating diffuse scattering probabilities...')
        self.scat_model = self.args.bound_scat[0]
        luts = boundary_tables.build(geometry.facets_normal[self.rough_facets, :] if len(self.rough_facets) else np.zeros((0, 3)),
                                     self.rough_facets_values, phonon, self.scat_model)
        self.specularity, self.true_specular = luts['specularity'], luts['true_specular']
        self.spec_out, self.creation_roulette = luts['spec_out'], luts['roulette']
        self.correspondent_modes = luts['correspondent_modes']
        if len(self.rough_facets):
            np.savetxt(os.path.join(self.results_folder_name, 'specular_correspondences.txt'), self.correspondent_modes,
                       fmt='%.3f %.3f %.3f %d %d %d %d')

        print('Initialising reservoirs...')
        self.n_of_reservoirs = int((self.bound_cond == 'T').sum() + (self.bound_cond == 'F').sum())
        if self.n_of_reservoirs > 0:
            self.initialise_reservoirs(geometry, phonon)
        else:
            print('No reservoir to be initialised.')
            self.res_facet = np.zeros(0, dtype=int)
            self.res_facet_temperature = np.zeros(0)
            self.res_counter = np.zeros((0,) + phonon.omega.shape)

    def initialise_reservoirs(self, geometry, phonon):
        """Reservoir temperatures and per-mode entry probabilities (Population.py:146-161, :323-354)."""
        self.res_facet = geometry.res_facets
        self.res_bound_values = geometry.res_values
        self.res_bound_cond = geometry.res_bound_cond
        mask_T = geometry.res_bound_cond == 'T'
        mask_F = geometry.res_bound_cond == 'F'
        self.res_facet_temperature = np.full(self.n_of_reservoirs, np.nan)
        self.res_facet_temperature[mask_T] = geometry.res_values[mask_T]
        if mask_F.any():
            self.res_facet_temperature[mask_F] = geometry.res_values[mask_T].mean()
        self.enter_prob = self.enter_probability(geometry, phonon)
        self.res_counter = np.random.rand(*self.enter_prob.shape)
        self.N_leaving = np.sum(self.enter_prob, axis=(1, 2)).round().astype(int)
        self.res_energy_balance = np.zeros(self.n_of_reservoirs)
        self.res_heat_flux = np.zeros((self.n_of_reservoirs, 3))

    def enter_probability(self, geometry, phonon):
        thickness = phonon.number_of_active_modes / (self.particle_density * geometry.facets_area[self.res_facet])
        vel = np.transpose(phonon.group_vel, (0, 2, 1))
        normals = -geometry.facets_normal[self.res_facet, :]
        p = np.dot(normals, vel) * self.dt / thickness.reshape(-1, 1, 1)
        return np.where(p < 0, 0, p)

    def tables(self, geometry, phonon):
        return build_tables(self.args, geometry, phonon, self)


class Population(PopulationSetup):
    '''Class comprising the particles to be simulated.'''

    def __init__(self, arguments, geometry, phonon, device=None, seed=None, engine=None):
        Constants.__init__(self)
        self.setup_host(arguments, geometry, phonon, seed)

        # ---- device context; under torchrun (WORLD_SIZE > 1) every rank owns a shard of the particles:
        #      the host set-up above must be identical on all ranks (nanokappa.py seeds NumPy identically)
        self.world, self.rank, self.sharded = 1, 0, None
        if engine is None:
            self.world = int(os.environ.get('WORLD_SIZE', 1))
            self.rank = int(os.environ.get('RANK', 0))
            if device is None:
                device = int(os.environ.get('LOCAL_RANK', 0))
            if self.world > 1:
                import torch
                import torch.distributed as dist
                torch.cuda.set_device(device)
                if not dist.is_initialized():
                    dist.init_process_group('nccl', device_id=torch.device('cuda', device))
                common = [self.seed, self.res_counter]
                dist.broadcast_object_list(common, src=0)          # the Philox seed and the reservoir counters are global
                self.seed, self.res_counter = int(common[0]), np.asarray(common[1])
            engine = Engine(device, seed=self.seed)
        self.engine = engine
        self.tables = build_tables(self.args, geometry, phonon, self)
        hot = None
        if self.n_of_reservoirs == 0:
            hot = (295.0, 305.0)
        self.engine.set_tables(self.tables, res_counter=self.res_counter, hot_T=hot)
        geometry.attach_engine(self.engine)
        phonon.attach_engine(self.engine)
        if self.world > 1:
            from ..parallel import ShardedEngine
            self.sharded = ShardedEngine(self.engine, self.rank, self.world)
            self.sharded.enable_fused_exchange()        # falls back to the NCCL all-reduce between the step halves

        print('Initialising population...')
        self.initialise_all_particles(geometry, phonon)

        self.conv_crit = float(self.args.conv_crit[0])
        self.conv_count_min = int(self.args.conv_crit[1])
        self.initialise_residue(geometry)
        print('Creating convergence file...')
        self.open_convergence(geometry)
        self.write_convergence(geometry)
        self.view = Visualisation(self.args, geometry, phonon, self)
        print('Initialisation done!')

    # ---- set-up ------------------------------------------------------------------------------------
    def initialise_modes(self, phonon):
        """Tile the active modes when there is at least one particle per mode and subvolume, draw them
        at random otherwise (Population.py:127-144)."""
        print('Assigning modes...')
        self.unique_modes = np.vstack(np.where(~phonon.inactive_modes_mask)).T
        if self.particles_pmps >= 1:
            reps = int(np.ceil(self.particles_pmps * (self.n_of_subvols - self.n_of_empty_subvols)))
            modes = np.tile(self.unique_modes, (reps, 1))[:self.N_p, :]
        else:
            modes = self.unique_modes[np.random.randint(low=0, high=phonon.number_of_active_modes, size=self.N_p), :]
        return modes.astype(int)

    def generate_positions(self, number_of_particles, mesh, key):
        if key == 'random':
            return mesh.sample_volume(number_of_particles)
        return np.ones((number_of_particles, 3)) * mesh.center_mass

    def initialise_all_particles(self, geometry, phonon):
        """Positions, modes, initial temperatures and occupations on the host, then upload; the first
        boundary collisions of all particles are found on the GPU (Population.py:186-321)."""
        key = self.args.part_dist[0]
        S = self.n_of_subvols
        occupation = None
        if self._can_init_on_device(geometry, key):
            return self._initialise_on_device(geometry, phonon)
        if self._can_init_on_device_general(geometry, key):
            return self._initialise_on_device(geometry, phonon, general=True)
        if key.endswith('.npz'):
            # binary checkpoint written by write_final_state above NK_TEXT_DUMP_MAX particles (the text restart file of the
            # reference, particle_data.txt, is handled below): exact continuation, per rank in a sharded run
            self.engine.allocate(1024)
            self.load_checkpoint(key if self.world == 1 or self.rank == 0 else os.path.join(os.path.dirname(key), 'rank{}'.format(self.rank), os.path.basename(key)))
            r = self._pull_results()
            self.subvol_heat_flux = r['subvol_heat_flux']
            self.calculate_kappa(geometry)
            self.res_energy_balance = r['res_energy_balance']; self.res_heat_flux = r['res_heat_flux']
            return
        if key in ('random_domain', 'center_domain'):
            positions = self.generate_positions(self.N_p, geometry.mesh, key.split('_')[0])
        elif key == 'random_subvol':
            vol = geometry.subvol_volume
            n = np.ceil(self.N_p * vol / (vol.sum() - vol[self.empty_subvols].sum())).astype(int)
            n[self.empty_subvols] = 0
            parts = [np.zeros((0, 3)) for _ in range(S)]
            have = np.zeros(S, dtype=int)
            while np.any(have < n):
                x = self.generate_positions(int(min(max(n.sum() - have.sum(), 1) * 1.2 + 16, 4e6)), geometry.mesh, 'random')
                sv = geometry.subvol_classifier.predict(x)
                for i in np.nonzero(have < n)[0]:
                    take = np.nonzero(sv == i)[0][: n[i] - have[i]]
                    parts[i] = np.vstack((parts[i], x[take]))
                    have[i] = parts[i].shape[0]
            positions = np.vstack(parts)[:self.N_p, :]
        elif key == 'center_subvol':
            raise Exception('--part_dist center_subvol needs per-subvolume meshes, which this build does not generate.')
        else:
            try:
                data = np.loadtxt(key, delimiter=',', comments='#', dtype=float)
            except Exception:
                raise Exception('Wrong particle data file. Change the keyword or check whether the file exists.')
            positions = np.copy(data[:, [2, 3, 4]])
            modes = np.copy(data[:, [0, 1]]).astype(int)
            occupation = np.copy(data[:, 5])
        if occupation is None:
            modes = self.initialise_modes(phonon)
        ids = None
        if self.world > 1:
            if occupation is not None:
                raise Exception('Restarting from a particle file is a single-GPU feature; use the binary checkpoint per rank.')
            from ..parallel import shard_bounds
            lo_i, hi_i = shard_bounds(self.rank, self.world, positions.shape[0])       # every rank built the same arrays
            positions, modes = positions[lo_i:hi_i], modes[lo_i:hi_i]
            ids = np.arange(lo_i, hi_i, dtype=np.int64)
        self.N_p = positions.shape[0]
        J = phonon.number_of_branches
        flat = modes[:, 0] * J + modes[:, 1]
        sv = geometry.subvol_classifier.predict(positions)
        temperatures, self.subvol_temperature = self.assign_temperatures(sv, geometry)
        if occupation is None:
            occupation = phonon.calculate_occupation(temperatures, phonon.omega[modes[:, 0], modes[:, 1]])
        cap = int(self.N_p * float(os.environ.get('NK_CAPACITY_FACTOR', 1.25))) + 1024
        self.engine.allocate(cap)
        self.engine.set_sv_temperature(self.subvol_temperature)
        print('Getting first boundary collisions...')
        self.engine.load_particles(positions, flat, occupation, ids=ids)
        self.engine.set_timestep(0)
        if key not in ('random_domain', 'center_domain', 'random_subvol', 'center_subvol'):
            old = np.zeros(S)
            for _ in range(100):
                self.refresh_temperatures(geometry, phonon)
                if np.absolute((self.subvol_temperature - old) / self.subvol_temperature).max() <= 1e-6:
                    break
                old = np.copy(self.subvol_temperature)
        print('Initialising local quantities...')
        if self.world > 1:
            self._device_census(geometry, phonon)      # fresh populations start at equilibrium with their subvolume
        else:
            self._host_census(geometry, phonon)

    # ---- initialisation at scale: positions, modes and occupations created on the device ----
    def _can_init_on_device(self, geometry, key):
        thr = float(os.environ.get('NK_DEVICE_INIT_MIN', 2e5))
        return (key in ('random_subvol', 'random_domain') and geometry.shape in ('cuboid', 'box') and geometry.subvol_type == 'slice'
                and self.rotation_free(geometry) and self.n_of_empty_subvols == 0 and self.N_p >= thr and self.T_distribution != 'custom')

    @staticmethod
    def rotation_free(geometry):
        return geometry.rotation is None or not np.any(np.asarray(geometry.rotation, dtype=float) != 0)

    def _can_init_on_device_general(self, geometry, key):
        """Any mesh / any subvolume type: rejection sampling inside the mesh on the device (nk_contains + nk_classify)."""
        thr = float(os.environ.get('NK_DEVICE_INIT_MIN', 2e5))
        return key in ('random_subvol', 'random_domain') and self.N_p >= thr and self.T_distribution != 'custom'

    def _sample_positions_on_device(self, geometry, N, g):
        """Population.py:209-246 on the device: uniform candidates in the bounding box, kept when inside the mesh
        (Mesh.contains_naive -> nk_contains); for random_subvol every subvolume takes ceil(N V_s / V_filled) of them
        (SubvolClassifier.predict -> nk_classify) and the concatenation in subvolume order is cut to N, as upstream does.
        Returns (positions (N,3), subvolume index (N,)) as device tensors."""
        import torch
        from ..engine import _dp
        from .._lib import check
        eng, dev, S = self.engine, self.engine.device, self.n_of_subvols
        lo = torch.as_tensor(geometry.bounds[0], device=dev); ext = torch.as_tensor(np.ptp(geometry.bounds, axis=0), device=dev)
        by_sv = self.args.part_dist[0] == 'random_subvol'
        if by_sv:
            vol = np.asarray(geometry.subvol_volume, dtype=float)
            quota = np.ceil(N * vol / (vol.sum() - vol[self.empty_subvols].sum())).astype(np.int64)
            quota[self.empty_subvols] = 0
        else:
            quota = np.array([N], dtype=np.int64)
        need = torch.as_tensor(quota, device=dev)
        fill = float(geometry.volume / np.prod(np.ptp(geometry.bounds, axis=0)))
        kept_x, kept_sv = [], []
        while int(need.sum().item()) > 0:
            m = int(min(8e6, max(4096, 1.3 * int(need.sum().item()) / max(fill, 1e-3))))
            x = (lo + torch.rand((m, 3), generator=g, dtype=torch.float64, device=dev) * ext).contiguous()
            inside = torch.empty(m, dtype=torch.uint8, device=dev)
            check(eng.ctx, eng.L.nk_contains(eng.ctx, m, _dp(x), _dp(inside)), 'nk_contains')
            x = x[inside.bool()].contiguous()
            k = int(x.shape[0])
            if k == 0:
                continue
            if by_sv:
                sv = torch.empty(k, dtype=torch.int32, device=dev)
                check(eng.ctx, eng.L.nk_classify(eng.ctx, k, _dp(x), _dp(sv), None), 'nk_classify')
                sv = sv.long()
            else:
                sv = torch.zeros(k, dtype=torch.int64, device=dev)
            order = torch.argsort(sv, stable=True)
            svs = sv[order]
            first = torch.searchsorted(svs, torch.arange(need.numel(), device=dev))
            rank_in_sv = torch.arange(k, device=dev) - first[svs]
            keep = rank_in_sv < need[svs]
            take = order[keep]
            kept_x.append(x[take]); kept_sv.append(sv[take])
            need = need - torch.bincount(sv[take], minlength=need.numel())
        X = torch.cat(kept_x); SV = torch.cat(kept_sv)
        order = torch.argsort(SV, stable=True)[:N]                    # np.vstack(x)[:N]: subvolume order, the tail is cut
        X = X[order].contiguous()
        if not by_sv:
            sv = torch.empty(N, dtype=torch.int32, device=dev)
            check(eng.ctx, eng.L.nk_classify(eng.ctx, N, _dp(X), _dp(sv), None), 'nk_classify')
            return X, sv.long()
        return X, SV[order]

    def _initialise_on_device(self, geometry, phonon, general=False):
        """Box + slice subvolumes: the reference fills every slice with ceil(N V_s / V) uniform points and keeps
        the first N (Population.py:209-246); here each slice's quota is drawn directly inside the slice
        with the device generator (same distribution, no rejection), modes are tiled / drawn as in
        initialise_modes (:127-144), occupations are Bose-Einstein at the slice temperature."""
        import torch
        eng = self.engine
        dev = eng.device
        from ..parallel import shard_bounds
        lo_i, hi_i = shard_bounds(self.rank, self.world, int(self.N_p))
        N, S, ax = hi_i - lo_i, self.n_of_subvols, getattr(self, 'slice_axis', 0)
        g = torch.Generator(device=dev); g.manual_seed(self.seed + 7919 * self.rank)
        cap = int(N * float(os.environ.get('NK_CAPACITY_FACTOR', 1.25))) + 1024
        eng.allocate(cap)
        t = eng.t
        lo = geometry.bounds[0]; ext = np.ptp(geometry.bounds, axis=0)
        key = self.args.part_dist[0]
        if general:
            # arbitrary mesh / subvolumes: the tables must be on the device before nk_contains / nk_classify can run (they are:
            # set_tables ran in __init__); positions by rejection sampling
            X, _ = self._sample_positions_on_device(geometry, N, g)
            t['px'][:N] = X[:, 0]; t['py'][:N] = X[:, 1]; t['pz'][:N] = X[:, 2]
            del X
        else:
            for k, name in enumerate(('px', 'py', 'pz')):
                t[name][:N] = lo[k] + torch.rand(N, generator=g, dtype=torch.float64, device=dev) * ext[k]
        if key == 'random_subvol' and not general:
            quota = int(np.ceil(N * geometry.subvol_volume[0] / geometry.subvol_volume.sum()))
            sl = torch.clamp(torch.arange(N, device=dev, dtype=torch.int64) // quota, max=S - 1).to(torch.float64)
            u = torch.rand(N, generator=g, dtype=torch.float64, device=dev)
            t[('px', 'py', 'pz')[ax]][:N] = lo[ax] + (sl + u) * (ext[ax] / S)
            del sl, u
        print('Assigning modes...')
        act = torch.as_tensor(np.nonzero(~phonon.inactive_modes_mask.reshape(-1))[0].astype(np.int32), device=dev)
        if self.particles_pmps >= 1:
            idx = (torch.arange(N, device=dev, dtype=torch.int64) + lo_i) % act.numel()
        else:
            idx = torch.randint(0, act.numel(), (N,), generator=g, device=dev)
        t['mode'][:N] = act[idx]; t['omode'][:N] = act[idx]; t['mode'][N:] = -1
        t['pid'][:N] = torch.arange(N, device=dev, dtype=torch.int64) + lo_i
        del idx
        _, self.subvol_temperature = self.assign_temperatures(np.zeros(1, dtype=int), geometry)
        eng.set_sv_temperature(self.subvol_temperature)
        torch.cuda.synchronize(dev)
        from ..engine import _dp
        from .._lib import check
        check(eng.ctx, eng.L.nk_set_slot_count(eng.ctx, N), 'nk_set_slot_count')
        pos = torch.stack((t['px'][:N], t['py'][:N], t['pz'][:N]), dim=1).contiguous()
        Tp = torch.empty(N, dtype=torch.float64, device=dev)
        check(eng.ctx, eng.L.nk_particle_temperature(eng.ctx, N, _dp(pos), _dp(Tp)), 'nk_particle_temperature')
        if self.temp_interp_type != 'nearest' or general:      # initial occupation uses the subvolume temperature, not the interpolated one
            sv = torch.empty(N, dtype=torch.int32, device=dev)
            check(eng.ctx, eng.L.nk_classify(eng.ctx, N, _dp(pos), _dp(sv), None), 'nk_classify')
            Tp = torch.as_tensor(self.subvol_temperature, device=dev)[sv.long()]
        om = torch.as_tensor(phonon.omega.reshape(-1), device=dev)[t['mode'][:N].long()]
        check(eng.ctx, eng.L.nk_occupation(eng.ctx, N, _dp(Tp), _dp(om), _dp(t['occ'])), 'nk_occupation')
        del pos, Tp, om
        print('Getting first boundary collisions...')
        eng.set_timestep(0)
        eng.init_collisions()
        eng.sort_by_mode()
        print('Initialising local quantities...')
        self._device_census(geometry, phonon)

    def _device_census(self, geometry, phonon):
        """Initial per-subvolume counts / energies without pulling the particles to the host: every particle
        starts at equilibrium with its subvolume, so the deviational energy and heat flux are zero and the
        energy density is E(T_sv) (what Population.calculate_energy gives at start, Population.py:318-321)."""
        import torch
        eng = self.engine
        n, _ = eng.slot_count()
        t = eng.t
        live = t['mode'][:n] >= 0                   # the slot range holds free slots too (spare slots of the mode pools)
        pos = torch.stack((t['px'][:n][live], t['py'][:n][live], t['pz'][:n][live]), dim=1).contiguous()
        n = int(pos.shape[0])
        sv, counts = None, torch.zeros(self.n_of_subvols, dtype=torch.int64, device=eng.device)
        from ..engine import _dp
        from .._lib import check
        svt = torch.empty(max(n, 1), dtype=torch.int32, device=eng.device)
        check(eng.ctx, eng.L.nk_classify(eng.ctx, n, _dp(pos), _dp(svt), _dp(counts)), 'nk_classify')
        eng.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(counts)
        self.subvol_N_p = counts.cpu().numpy()
        self.N_p = int(self.subvol_N_p.sum())
        self.subvol_energy = np.interp(self.subvol_temperature, phonon.T_array, phonon.energy_array)
        self.subvol_heat_flux = np.zeros((self.n_of_subvols, 3))
        self.total_energy = 0.0
        self.calculate_kappa(geometry)
        self.res_energy_balance = np.zeros(self.n_of_reservoirs)
        self.res_heat_flux = np.zeros((self.n_of_reservoirs, 3))
        self.N_leaving = np.zeros(self.n_of_reservoirs, dtype=int)

    # ---- binary checkpoint: everything needed to continue bit-exactly ---------------------------
    def save_checkpoint(self, path):
        """Live particles (f64 positions, unlike the 1e-3 A text dump), collision clocks, reservoir counters and deal
        counters, the reservoir balances accumulated since the last convergence row, N_leaving (feeds one_to_one), the
        results block, subvolume temperatures, step counter and the Philox seed (``Engine.checkpoint``).  `np.savez`
        container; in a sharded run every rank writes its own shard."""
        z = self.engine.checkpoint()
        if self.current_timestep == 0:
            z['subvol_temperature'] = np.asarray(self.subvol_temperature, dtype=float)
        z['current_timestep'] = np.int64(self.current_timestep)
        z['world'] = np.int64(self.world); z['rank'] = np.int64(self.rank)
        np.savez(path, **z)

    def load_checkpoint(self, path):
        """Continue from ``save_checkpoint``.  The tables, the rank and the exchange buffers of the context are left
        untouched (only particles and run state are replaced), so it is valid inside a sharded run too: every rank loads
        the shard it wrote."""
        z = np.load(path)
        if int(z['seed']) != self.seed:
            raise Exception('checkpoint was written with seed {} but this run uses {}'.format(int(z['seed']), self.seed))
        if 'world' in z.files and (int(z['world']) != self.world or int(z['rank']) != self.rank):
            raise Exception('checkpoint belongs to rank {} of {} but this process is rank {} of {}'.format(
                int(z['rank']), int(z['world']), self.rank, self.world))
        self.engine.restore(z, capacity_factor=float(os.environ.get('NK_CAPACITY_FACTOR', 1.25)))
        self.subvol_temperature = np.array(z['subvol_temperature'])
        self.current_timestep = int(z['current_timestep'])
        self.t = self.current_timestep * self.dt
        self._cache_key = None
        if self.N_p >= float(os.environ.get('NK_RESORT_MIN', 1e6)):
            self.engine.sort_by_mode()

    def assign_temperatures(self, subvol_id, geometry):
        """Initial subvolume temperatures for --temp_dist (Population.py:565-655)."""
        print('Assigning temperatures...')
        key = self.T_distribution
        S = self.n_of_subvols
        if key == 'custom':
            sv_T = np.array(self.args.subvol_temp, dtype=float)
        else:
            bound_T = self.res_bound_values[self.res_bound_cond == 'T'] if self.n_of_reservoirs else np.array([300.0])
            if key == 'linear':
                pos = geometry.facet_centroid[self.res_facet[self.res_bound_cond == 'T'], :]
                if len(bound_T) > 2:
                    d = np.sum((geometry.subvol_center - pos[:, None, :]) ** 2, axis=2).T ** 0.5
                    w = 1 / d
                    w /= w.sum(axis=1, keepdims=True)
                    sv_T = np.sum(bound_T * w, axis=1)
                elif len(bound_T) == 2:
                    direction = pos[1] - pos[0]
                    alpha = ((geometry.subvol_center - pos[0]) * direction).sum(axis=1) / (direction ** 2).sum()
                    sv_T = bound_T[0] + alpha * (bound_T[1] - bound_T[0])
                else:
                    sv_T = np.ones(S) * bound_T
            elif key == 'random':
                sv_T = np.random.rand(S) * np.ptp(bound_T) + bound_T.min()
            elif key == 'hot':
                sv_T = np.ones(S) * bound_T.max()
            elif key == 'cold':
                sv_T = np.ones(S) * bound_T.min()
            elif key == 'mean':
                sv_T = np.ones(S) * bound_T.mean()
            else:
                raise Exception('Invalid --temp_dist.')
        return sv_T[subvol_id], sv_T

    def _host_census(self, geometry, phonon):
        """Per-subvolume energy / flux / kappa of the initial state (Population.py:318-321): computed
        once at set-up from the uploaded particles with the host formulas."""
        p = self.engine.particles(flush=False)
        J = phonon.number_of_branches
        om = phonon.omega.reshape(-1)[p['omega_modes']]
        v = phonon.group_vel[p['modes'][:, 0], p['modes'][:, 1], :]
        sv = geometry.subvol_classifier.predict(p['positions'])
        S = self.n_of_subvols
        self.subvol_N_p = np.bincount(sv, minlength=S)
        self.N_p = int(self.subvol_N_p.sum())
        dn = p['occupation'] - phonon._occupation_host(self.subvol_temperature[sv], om)
        e = self.hbar * om * dn
        with np.errstate(divide='ignore', invalid='ignore'):
            if self.norm == 'mean':
                norm = phonon.number_of_active_modes / self.subvol_N_p
                norm = np.where(np.isnan(norm), 0, norm)
            else:
                norm = phonon.number_of_active_modes / (self.particle_density * geometry.subvol_volume)
            self.subvol_energy = phonon.normalise_to_density(np.bincount(sv, weights=e, minlength=S) * norm) + \
                np.interp(self.subvol_temperature, phonon.T_array, phonon.energy_array)
            hf = np.stack([np.bincount(sv, weights=v[:, k] * e, minlength=S) for k in range(3)], axis=1) * norm.reshape(-1, 1)
        self.subvol_heat_flux = phonon.normalise_to_density(hf) * self.eVpsa2_in_Wm2
        self.total_energy = float(e.sum())
        self.calculate_kappa(geometry)
        self.res_energy_balance = np.zeros(self.n_of_reservoirs)
        self.res_heat_flux = np.zeros((self.n_of_reservoirs, 3))
        self.N_leaving = np.zeros(self.n_of_reservoirs, dtype=int)

    # ---- device state as reference attributes ------------------------------------------------------------
    def _particles(self):
        key = (self.current_timestep, 'p')
        if getattr(self, '_cache_key', None) != key:
            self._cache = self.engine.particles(flush=True)
            self._cache_key = key
        return self._cache

    positions = property(lambda self: self._particles()['positions'])
    modes = property(lambda self: self._particles()['modes'])
    occupation = property(lambda self: self._particles()['occupation'])
    n_timesteps = property(lambda self: self._particles()['n_timesteps'])
    collision_facets = property(lambda self: self._particles()['collision_facets'])
    collision_positions = property(lambda self: self._particles()['collision_positions'])

    @property
    def omega(self):
        return self.engine.tb['omega'].reshape(-1)[self._particles()['omega_modes']]

    @property
    def group_vel(self):
        m = self._particles()['modes']
        return self.engine.tb['group_vel'][m[:, 0], m[:, 1], :]

    @property
    def subvol_id(self):
        return self.engine.classify(self._particles()['positions'])

    @property
    def temperatures(self):
        return self.engine.particle_temperature(self._particles()['positions'])

    @property
    def collision_cond(self):
        return self.bound_cond[self._particles()['collision_facets']]

    def temperature_interpolator(self, x):
        x = np.asarray(x, dtype=float)
        if x.ndim == 1 or x.shape[-1] != 3:
            full = np.tile(np.mean(self.engine.tb['bounds'], axis=0), (x.reshape(-1).shape[0], 1))
            full[:, self.slice_axis] = x.reshape(-1)
            x = full
        return self.engine.particle_temperature(x)

    # ---- the reference's per-step methods -------------------------------------------------------------------
    def _pull_results(self):
        r = self.engine.results()
        self.subvol_temperature = r['subvol_temperature']
        self.subvol_energy = r['subvol_energy']
        self.subvol_N_p = r['subvol_N_p']
        self.N_p = r['N_p']
        self.N_leaving = r['N_leaving']
        self.total_energy = r['total_energy']
        return r

    def run_timestep(self, geometry, phonon):
        """drift -> emission -> boundary scattering -> temperatures -> lifetime scattering on the GPU
        (Population.py:1724-1769), with the reference's every-100 / every-10-step outputs."""
        if self.current_timestep == 0:
            print('Simulating...')
        if (self.current_timestep % 100) == 0:
            # particle dump: every 100 steps like the reference while it is the cheap text file.  For a large population the
            # dump is the binary checkpoint (1 GB per 1e7 particles) and 100 steps are milliseconds of GPU time, so it is
            # written by the wall clock instead -- when NK_DUMP_MIN_SECONDS (default 600) have passed since the last one --
            # or every NK_DUMP_EVERY steps when that is set; the end-of-run dump is always written
            small = self.N_p <= float(os.environ.get('NK_TEXT_DUMP_PERIODIC_MAX', 2e5))
            now = datetime.now()
            if not hasattr(self, '_last_dump'):
                self._last_dump = now
            if small:
                due = True
            elif 'NK_DUMP_EVERY' in os.environ:
                every = int(os.environ['NK_DUMP_EVERY'])
                due = every > 0 and (self.current_timestep % every) == 0
            else:
                due = (now - self._last_dump).total_seconds() >= float(os.environ.get('NK_DUMP_MIN_SECONDS', 600))
            if due:
                self._last_dump = now
                self.write_final_state(geometry, final=False)
            elif self.current_timestep > 0 and hasattr(self.view, 'mean_T'):
                self.write_subvolume_state(geometry)
            self.view.postprocess(verbose=False)
            self.update_residue(geometry)
            self.contains_check(geometry)
            info = 'Timestep {:>5d} - max residue: {:>9.3e} ({:<9s}) ['.format(int(self.current_timestep), self.max_residue, self.max_residue_qt)
            for sv in range(self.n_of_subvols):
                info += ' {:>7.3f}'.format(self.subvol_temperature[sv])
            print(info + ' ]')
        # maintenance: the per-mode slot pools keep emitted particles among their own mode, but a mode's pool runs dry or
        # fills up as its population fluctuates; the counting sort (a few ms at 1e8 particles) re-centres the regions every
        # NK_RESORT_EVERY steps (0 = never).  Results do not depend on it.
        every = int(os.environ.get('NK_RESORT_EVERY', 500))
        if every > 0 and self.current_timestep > 0 and (self.current_timestep % every) == 0 and \
                self.N_p >= float(os.environ.get('NK_RESORT_MIN', 1e6)):
            self.engine.sort_by_mode()
        if self.sharded is not None and self.current_timestep > 0 and (self.current_timestep % 100) == 0:
            self.sharded.rebalance()               # live counts drift with position-dependent absorption
        self._advance_device()
        self.current_timestep += 1
        self.t = self.current_timestep * self.dt
        if (self.current_timestep % self.n_dt_to_conv) == 0:
            if self.step_batching and self._tickets:
                # the next batch goes to the device BEFORE the host looks at this row, so the GPU never waits for the
                # formatting and the file write (not across a 100-step boundary: that branch inspects the particles)
                ticket = self._tickets.pop(0)
                if (self.current_timestep % 100) != 0:
                    self._enqueue_batch()
                r = self.engine.results(ticket)
                self.subvol_temperature, self.subvol_energy, self.subvol_N_p = r['subvol_temperature'], r['subvol_energy'], r['subvol_N_p']
                self.N_p, self.N_leaving, self.total_energy = r['N_p'], r['N_leaving'], r['total_energy']
            else:
                r = self._pull_results()
            self.subvol_heat_flux = r['subvol_heat_flux']
            if geometry.subvol_type == 'slice':
                self.subvol_kappa, self.kappa = r['subvol_kappa'], r['kappa']
            else:
                self.calculate_kappa(geometry)
            self.res_heat_flux, self.res_energy_balance = r['res_heat_flux'], r['res_energy_balance']
            self.write_convergence(geometry)
        elif (self.current_timestep % 100) == 99 and not self.step_batching:
            self._pull_results()

    # The command line sets `step_batching`: the timesteps between two convergence rows (n_dt_to_conv = 10) are enqueued as
    # ONE nk_step call followed by an asynchronous snapshot of the results block, and the calls of run_timestep in between
    # only advance the host's step counter.  The device then runs up to one batch ahead of `current_timestep`; it is level
    # with it at every multiple of n_dt_to_conv -- where rows are written -- and of 100 -- where particles are dumped,
    # checked and re-sorted.  Library users who inspect particles between arbitrary steps leave it off (the default).
    step_batching = False

    def _enqueue_batch(self):
        B = self.n_dt_to_conv
        nxt = (self._enqueued_until // B + 1) * B
        try:
            nxt = min(nxt, max(int(self.args.iterations[0]), self._enqueued_until + 1))
        except Exception:
            pass
        n = nxt - self._enqueued_until
        if self.sharded is not None:
            self.sharded.step(n)
        else:
            self.engine.step(n)
        self._enqueued_until = nxt
        if (nxt % B) == 0:
            self._tickets.append(self.engine.snapshot_results())

    def _advance_device(self):
        if not hasattr(self, '_enqueued_until'):
            self._enqueued_until, self._tickets = self.current_timestep, []
        if not self.step_batching or (self.sharded is not None and not self.sharded.fused and self.world > 1):
            if self._enqueued_until <= self.current_timestep:          # (a batch enqueued before batching was switched off)
                if self.sharded is not None:
                    self.sharded.step(1)
                else:
                    self.engine.step(1)
                self._enqueued_until = self.current_timestep + 1
            return
        if self._enqueued_until <= self.current_timestep:
            self._enqueue_batch()

    # ---- the reference's per-step methods as seams.  On the GPU drift, emission, boundary scattering and the
    # per-subvolume sums are ONE fused pass over the particles (k_step + k_rare) and the lifetime scattering is deferred to
    # the head of the next pass, so the sequence the reference's run_timestep spells out
    #     drift -> fill_reservoirs -> add_reservoir_particles -> boundary_scattering -> refresh_temperatures -> lifetime_scattering
    # maps to: drift() launches the fused step, the four methods in the middle find their work already done, and
    # lifetime_scattering() applies the deferred relaxation and closes the step.  Called in that order they advance the
    # population by exactly one timestep, like run_timestep (without its every-10 / every-100-step outputs).
    def drift(self):
        """Population.drift (Population.py:790-795) -- opens a fused timestep on the device."""
        if getattr(self, '_seam_open', False):
            raise Exception('drift() was already called for this timestep; finish it with lifetime_scattering().')
        if getattr(self, '_enqueued_until', self.current_timestep) > self.current_timestep:
            raise Exception('the device is ahead of current_timestep (step_batching): finish the batch with run_timestep().')
        batching, self.step_batching = self.step_batching, False
        try:
            self._advance_device()
        finally:
            self.step_batching = batching
        self._seam_open = True

    def _seam_done(self, name):
        if not getattr(self, '_seam_open', False):
            raise Exception(name + '() is part of the fused GPU timestep: call drift() first (or run_timestep()).')

    def fill_reservoirs(self, geometry, phonon):
        """Population.fill_reservoirs (Population.py:356-489): done by the emission scan of the fused step."""
        self._seam_done('fill_reservoirs')

    def add_reservoir_particles(self, geometry, phonon):
        """Population.add_reservoir_particles (Population.py:525-552): done by the rare-path kernel of the fused step."""
        self._seam_done('add_reservoir_particles')

    def boundary_scattering(self, geometry, phonon):
        """Population.boundary_scattering (Population.py:1546-1683): done by the rare-path kernel of the fused step."""
        self._seam_done('boundary_scattering')

    def lifetime_scattering(self, phonon):
        """Population.lifetime_scattering (Population.py:1701-1710): applies the deferred relaxation (nk_flush_relaxation)
        and closes the timestep opened by drift()."""
        self._seam_done('lifetime_scattering')
        self.engine.flush_relaxation()
        self._seam_open = False
        self.current_timestep += 1
        self.t = self.current_timestep * self.dt
        self._cache_key = None

    def calculate_energy(self, geometry, phonon):
        """Population.calculate_energy (Population.py:704-728): per-subvolume energy densities of the last step
        (block-private bins of k_step, closed by the finalize)."""
        self._pull_results()
        return self.subvol_energy

    def calculate_heat_flux(self, geometry, phonon):
        """Population.calculate_heat_flux (Population.py:730-747): the device evaluates it on convergence steps."""
        self.subvol_heat_flux = self.engine.results()['subvol_heat_flux']
        return self.subvol_heat_flux

    def refresh_temperatures(self, geometry, phonon):
        """Recompute subvolume energies/temperatures from the current particles without moving them
        (used by the --part_dist restart loop, Population.py:297-304).  Inside a drift() ... lifetime_scattering()
        sequence the fused step has already done it: only the results are pulled."""
        if getattr(self, '_seam_open', False):
            self._pull_results()
            return
        p = self.engine.particles(flush=False)
        om = phonon.omega.reshape(-1)[p['omega_modes']]
        sv = geometry.subvol_classifier.predict(p['positions'])
        S = self.n_of_subvols
        cnt = np.bincount(sv, minlength=S)
        dn = p['occupation'] - phonon._occupation_host(self.subvol_temperature[sv], om)
        e = np.bincount(sv, weights=self.hbar * om * dn, minlength=S)
        with np.errstate(divide='ignore', invalid='ignore'):
            norm = phonon.number_of_active_modes / cnt if self.norm == 'mean' else phonon.number_of_active_modes / (self.particle_density * geometry.subvol_volume)
            norm = np.where(np.isnan(norm), 0, norm)
        E = phonon.normalise_to_density(e * norm) + np.interp(self.subvol_temperature, phonon.T_array, phonon.energy_array)
        self.subvol_energy = E
        self.subvol_temperature = np.asarray(phonon.temperature_function(E))
        self.engine.set_sv_temperature(self.subvol_temperature)

    def calculate_kappa(self, geometry):
        """Population.py:749-788."""
        if geometry.subvol_type == 'slice':
            S = self.n_of_subvols
            T = np.zeros(S + 2)
            T[1:-1] = self.subvol_temperature
            if self.n_of_reservoirs == 2:
                T[[0, -1]] = self.res_facet_temperature
            phi = self.subvol_heat_flux[:, geometry.slice_axis]
            L = np.ptp(geometry.bounds[:, geometry.slice_axis])
            dx = 2 * L * self.a_in_m / S
            dT = T[2:] - T[:-2]
            DX = L * self.a_in_m * (1 + S) / S
            DT = T[-1] - T[0]
            with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                self.subvol_kappa = -phi * dx / dT
                self.kappa = -np.sum(phi * self.subvol_N_p) * (DX / DT) / self.N_p
            self.subvol_kappa[np.absolute(self.subvol_kappa) == np.inf] = 0
        else:
            i = geometry.subvol_connections[:, 0]
            j = geometry.subvol_connections[:, 1]
            dx = geometry.subvol_center[j, :] - geometry.subvol_center[i, :]
            n = dx / np.linalg.norm(dx, axis=1, keepdims=True)
            dT = self.subvol_temperature[j] - self.subvol_temperature[i]
            phi = (self.subvol_heat_flux[i, :] + self.subvol_heat_flux[j, :]) / 2
            with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                self.svcon_kappa = np.where(dT == 0, 0, -np.sum(phi * n, axis=1) * np.linalg.norm(dx, axis=1) * self.a_in_m / dT)

    def contains_check(self, geometry):
        """Population.py:1712-1722: particles that left the bounding box by more than 1e-10 are put back
        at a random interior point with a fresh first collision."""
        eng = self.engine
        n, _ = eng.slot_count()
        if n == 0:
            return
        import ctypes as C
        import torch
        from ..engine import _dp
        from .._lib import check
        t = eng.t
        buf = torch.empty(4096, dtype=torch.int32, device=eng.device)
        found = C.c_int64()
        check(eng.ctx, eng.L.nk_outside_slots(eng.ctx, 1e-10, _dp(buf), buf.numel(), C.byref(found)), 'nk_outside_slots')
        if found.value > buf.numel():
            buf = torch.empty(found.value, dtype=torch.int32, device=eng.device)
            check(eng.ctx, eng.L.nk_outside_slots(eng.ctx, 1e-10, _dp(buf), buf.numel(), C.byref(found)), 'nk_outside_slots')
        if found.value == 0:
            return
        idx = torch.sort(buf[:found.value].long()).values
        new = geometry.mesh.sample_volume(int(idx.numel()))
        md = t['mode'][idx].cpu().numpy().astype(int)
        v = self.engine.tb['group_vel'].reshape(-1, 3)[md]
        xc, tc, fc = eng.find_boundary(new, v)
        dev = eng.device
        newt = torch.as_tensor(new, device=dev)
        t['px'][idx] = newt[:, 0]; t['py'][idx] = newt[:, 1]; t['pz'][idx] = newt[:, 2]
        t['tc'][idx] = torch.as_tensor(tc / self.dt, device=dev)
        t['cfacet'][idx] = torch.as_tensor(fc.astype(np.int32), device=dev)
        xct = torch.as_tensor(xc, device=dev)
        t['cx'][idx] = xct[:, 0]; t['cy'][idx] = xct[:, 1]; t['cz'][idx] = xct[:, 2]

    # ---- residue / convergence / output files (formats are an interface: Visualisation parses them) ----------------
    def initialise_residue(self, geo):
        S, R = self.n_of_subvols, self.n_of_reservoirs
        n = 3 * S + R if geo.subvol_type == 'slice' else 4 * S + R + geo.n_of_subvol_con
        self.old_mean_large = np.ones(n)
        self.old_std_large = np.ones(n)
        self.conv_count = 0
        self.finish_sim = False
        self.max_residue = 1
        self.max_residue_qt = 'none'
        if geo.subvol_type == 'slice':
            ax = ['x', 'y', 'z'][self.slice_axis]
            self.residue_qts = ['T_{:d}'.format(i) for i in range(S)] + ['phi_{:s}_{:d}'.format(ax, j) for j in range(S)] + \
                               ['en_res_{:d}'.format(i) for i in range(R)] + ['k_{:d}'.format(i) for i in range(S)]
        else:
            self.residue_qts = ['T_{:d}'.format(i) for i in range(S)] + ['phi_{:s}_{:d}'.format(i, j) for j in range(S) for i in ['x', 'y', 'z']] + \
                               ['en_res_{:d}'.format(i) for i in range(R)] + ['k_{:d}'.format(i) for i in range(geo.n_of_subvol_con)]

    def update_residue(self, geo):
        """Relative change of the rolling means between two 100-step checks (Population.py:1797-1839)."""
        v = self.view
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            if geo.subvol_type == 'slice':
                sel = 3 * np.arange(self.n_of_subvols) + self.slice_axis
                new_mean = np.concatenate((v.mean_T, v.mean_sv_phi[sel], v.mean_en_res, v.mean_sv_k))
                new_std = np.concatenate((v.std_T, v.std_sv_phi[sel], v.std_en_res, v.std_sv_k))
            else:
                new_mean = np.concatenate((v.mean_T, v.mean_sv_phi, v.mean_en_res, v.mean_con_k))
                new_std = np.concatenate((v.std_T, v.std_sv_phi, v.std_en_res, v.std_con_k))
            residue = np.absolute((new_mean - self.old_mean_large) / self.old_mean_large)
        self.residue_all = np.where(new_std > np.absolute(new_mean), 0, residue)
        self.max_residue = np.nanmax(self.residue_all)
        self.max_residue_qt = self.residue_qts[int(np.nonzero(self.residue_all == self.max_residue)[0][0])]
        self.conv_count = self.conv_count + 1 if self.max_residue < self.conv_crit else 0
        if self.conv_count >= self.conv_count_min:
            self.finish_sim = True
        self.old_mean_large, self.old_std_large = new_mean, new_std
        with open(os.path.join(self.results_folder_name, 'residue.txt'), 'a+') as f:
            f.writelines(''.join('{:9.3e} '.format(i) for i in self.residue_all) + '\n')

    def open_convergence(self, geometry):
        S, R = self.n_of_subvols, self.n_of_reservoirs
        line = '# ' + 'Real Time                  ' + 'Timest. ' + 'Simul. Time ' + 'Total Energy '
        if R > 0:
            line += ''.join('En Bal Res {} '.format(i) for i in range(R))
            line += ''.join(' Hflux x Res {0}  Hflux y Res {0}  Hflux z Res {0} '.format(i) for i in range(R))
        line += ' No. Part. '
        line += ''.join(' T Sv {:>3d} '.format(i) for i in range(S))
        line += ''.join(' Energ Sv {:>2d} '.format(i) for i in range(S))
        line += ''.join(' Hflux x Sv {0:>2d}  Hflux y Sv {0:>2d}  Hflux z Sv {0:>2d} '.format(i) for i in range(S))
        line += ''.join(' Np Sv {:>3d} '.format(i) for i in range(S))
        if geometry.subvol_type == 'slice':
            line += ''.join(' Kappa Sv {:>2d} '.format(i) for i in range(S)) + ' Kappa total  '
        else:
            line += ''.join(' K Con {:>3d}-{:>3d} '.format(c[0], c[1]) for c in geometry.subvol_connections)
        self.f = open(os.path.join(self.results_folder_name, 'convergence.txt'), 'a+')
        self.f.write(line + '\n')
        self.f.close()

    def write_convergence(self, geometry):
        """One row of convergence.txt (Population.py:2027-2069).  The reference formats every vector with np.array2string and
        strips the brackets; the plain joins below give the same characters at a twentieth of the cost (a row every ten
        timesteps is the host's main job while the GPU steps)."""
        def a2s(a, f):
            return ' '.join(f.format(v) for v in np.asarray(a, dtype=float).reshape(-1)) + ' '
        line = datetime.now().strftime('%Y-%m-%dT%H:%M:%S.%f ')
        line += '{:>8d} '.format(int(self.current_timestep))
        line += '{:>12.5e} '.format(self.t)
        line += '{:>12.5e} '.format(getattr(self, 'total_energy', 0.0))
        if self.n_of_reservoirs > 0:
            line += a2s(self.res_energy_balance, '{:>12.5e}')
            for i in range(self.n_of_reservoirs):
                line += a2s(self.res_heat_flux[i, :], '{:>14.6e}')
        line += '{:>10d} '.format(int(self.N_p))
        line += a2s(self.subvol_temperature, '{:>9.3f}')
        line += a2s(self.subvol_energy, '{:>12.5e}')
        for i in range(self.n_of_subvols):
            line += a2s(self.subvol_heat_flux[i, :], '{:>14.6e}')
        line += ' '.join('{:>10d}'.format(int(v)) for v in np.asarray(self.subvol_N_p).reshape(-1)) + ' '
        if geometry.subvol_type == 'slice':
            line += a2s(self.subvol_kappa, '{:>12.5e}')
            line += '{:>13.6e} '.format(self.kappa)
        else:
            line += a2s(self.svcon_kappa, '{:>14.7e}')
        self.f = open(os.path.join(self.results_folder_name, 'convergence.txt'), 'a+')
        self.f.write(line.replace('\n', ' ') + '\n')
        self.f.close()

    def write_final_state(self, geometry, final=True):
        """particle_data.txt / subvolumes.txt / subvol_connections.txt (Population.py:2071-2151).  `final=False` marks the
        every-100-steps call of run_timestep."""
        time = datetime.now().strftime('%Y-%m-%dT%H:%M:%S.%f')
        if getattr(self, '_enqueued_until', self.current_timestep) > self.current_timestep:
            # a run stopped by --max_sim_time in the middle of a batch: the device state is a few steps ahead; what is saved
            # is that state, under its own step number
            self.current_timestep = self._enqueued_until
            self.t = self.current_timestep * self.dt
        # the reference dumps ~60 bytes of text per particle every 100 steps (0.5 s per 1e5 particles -- a thousand times the
        # cost of the 100 timesteps themselves here -- and 6 GB at 1e8): the exact binary checkpoint (also a valid restart
        # point) replaces the text file above NK_TEXT_DUMP_MAX particles, and above NK_TEXT_DUMP_PERIODIC_MAX for the
        # periodic dumps; the dump at the end of a run keeps the reference's format up to NK_TEXT_DUMP_MAX
        limit = float(os.environ.get('NK_TEXT_DUMP_MAX', 5e6))
        if not final:
            limit = min(limit, float(os.environ.get('NK_TEXT_DUMP_PERIODIC_MAX', 2e5)))
        if self.N_p > limit:
            self.save_checkpoint(os.path.join(self.results_folder_name, 'particle_data.npz'))
        else:
            p = self._particles()
            header = 'Particles final state data \n' + 'Date and time: {}\n'.format(time) + \
                     'hdf file = {}, POSCAR file = {}\n'.format(self.args.hdf_file, self.args.poscar_file) + \
                     'q-point, branch, pos x [angs], pos y [angs], pos z [angs], occupation'
            data = np.hstack((p['modes'], p['positions'], p['occupation'].reshape(-1, 1)))
            np.savetxt(os.path.join(self.results_folder_name, 'particle_data.txt'), data, '%d, %d, %.3f, %.3f, %.3f, %.6e', delimiter=',', header=header)
        if self.current_timestep > 0 and hasattr(self.view, 'mean_T'):
            self.write_subvolume_state(geometry)

    def write_subvolume_state(self, geometry):
        """subvolumes.txt / subvol_connections.txt (Population.py:2093-2151)."""
        time = datetime.now().strftime('%Y-%m-%dT%H:%M:%S.%f')
        v = self.view
        S = self.n_of_subvols
        head = 'subvols final state data \nDate and time: {}\nhdf file = {}, POSCAR file = {}\n'.format(time, self.args.hdf_file, self.args.poscar_file)
        cols = [np.arange(S).reshape(-1, 1), geometry.subvol_center, self.subvol_volume.reshape(-1, 1), v.mean_T.reshape(-1, 1),
                v.std_T.reshape(-1, 1), v.mean_sv_phi.reshape(-1, 3), v.std_sv_phi.reshape(-1, 3)]
        if geometry.subvol_type == 'slice':
            cols += [v.mean_sv_k.reshape(-1, 1), v.std_sv_k.reshape(-1, 1)]
            head += 'subvol id, subvol x, subvol y, subvol z, subvol volume, T [K], sigma T [K], HF x [W/m^2], HF y [W/m^2], HF z [W/m^2], sigma HF x [W/m^2], sigma HF y [W/m^2], sigma HF z [W/m^2], kappa [W/m K], sigma kappa [W/m K]'
            fmt = '%d, %.3e, %.3e, %.3e, %.3e, %.3f, %.3e, %.3e, %.3e, %.3e, %.3e, %.3e, %.3e, %.3e, %.3e'
        else:
            head += 'subvol id, subvol position, subvol volume, T [K], sigma T [K], HF x [W/m^2], HF y [W/m^2], HF z [W/m^2], sigma HF x [W/m^2], sigma HF y [W/m^2], sigma HF z [W/m^2]'
            fmt = '%d, %.3e, %.3e, %.3e, %.3e, %.3f, %.3e, %.3e, %.3e, %.3e, %.3e, %.3e, %.3e'
        np.savetxt(os.path.join(self.results_folder_name, 'subvolumes.txt'), np.hstack(cols), fmt, delimiter=',', header=head)
        if geometry.subvol_type != 'slice' and geometry.n_of_subvol_con > 0:
            head = 'connections final state data \nDate and time: {}\nhdf file = {}, POSCAR file = {}\n'.format(time, self.args.hdf_file, self.args.poscar_file) + \
                   'connection id, sv 1, sv 2, con dx, con dy, con dz, dT [K], sigma dT [K], HF [W/m^2], sigma HF [W/m^2], kappa [W/m K], sigma kappa [W/m K]'
            data = np.hstack((np.arange(geometry.n_of_subvol_con).reshape(-1, 1), geometry.subvol_connections, geometry.subvol_con_vectors,
                              v.mean_con_dT.reshape(-1, 1), v.std_con_dT.reshape(-1, 1), v.mean_con_phi.reshape(-1, 1),
                              v.std_con_phi.reshape(-1, 1), v.mean_con_k.reshape(-1, 1), v.std_con_k.reshape(-1, 1)))
            np.savetxt(os.path.join(self.results_folder_name, 'subvol_connections.txt'), data,
                       '%d, %d, %d, %.3e, %.3e, %.3e, %.3f, %.3e, %.3e, %.3e, %.3e, %.3e', delimiter=',', header=head)

    def save_plot_real_time(self):
        """Called by nanokappa.py:105 but missing upstream (AttributeError after the results are
        saved); provided as a no-op so the driver script runs to the end."""
        return None

    def plot_figures(self, *a, **k):
        return None
