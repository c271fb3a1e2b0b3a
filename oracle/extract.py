"""TEST INFRASTRUCTURE -- turn live reference objects (Geometry, Phonon, Population built through
``ref_harness``) into the plain-array tables / state ``nk_oracle`` and the CUDA path consume.

Everything here *reads attributes the reference computed*; nothing is recomputed, so the tables are
the reference's own set-up results (Mesh.py:205-324, Geometry.py:446-726, Phonon.py:326-401,
Population.py:146-161, :852-939, :1042-1461).
"""
from __future__ import annotations

import numpy as np

from . import nk_oracle as nko


def tables_from_reference(geo, ph, pop):
    mesh = geo.mesh
    tb = {}
    tb["face_normals"] = np.array(mesh.face_normals, dtype=float)
    tb["face_k"] = np.array(mesh.face_k, dtype=float)
    tb["face_lo"] = np.array(mesh.face_bounds[0], dtype=float)
    tb["face_hi"] = np.array(mesh.face_bounds[1], dtype=float)
    tb["face_origins"] = np.array(mesh.face_origins, dtype=float)
    tb["face_basis"] = np.array(mesh.face_basis_matrix, dtype=float)
    tb["face_facets"] = np.array(mesh.face_facets, dtype=np.int64)
    tb["face_vertices"] = np.array(mesh.vertices[mesh.faces, :], dtype=float)
    tb["face_areas"] = np.array(mesh.face_areas, dtype=float)

    nf = mesh.n_of_facets
    tb["facet_bc"] = np.array([nko.BC_CODE[c] for c in geo.bound_cond], dtype=np.int64)
    tb["facet_normal"] = np.array(geo.facets_normal, dtype=float)
    tb["facet_centroid"] = np.array(geo.facet_centroid, dtype=float)
    tb["facet_area"] = np.array(geo.facets_area, dtype=float)
    partner = -np.ones(nf, dtype=np.int64)
    for a, b in np.asarray(geo.connected_facets, dtype=int).reshape(-1, 2):
        # Population.py:1468-1470: first row of connected_facets containing the facet wins
        if partner[a] < 0:
            partner[a] = b
        if partner[b] < 0:
            partner[b] = a
    tb["facet_partner"] = partner
    fres = -np.ones(nf, dtype=np.int64)
    fres[np.asarray(geo.res_facets, dtype=int)] = np.arange(len(geo.res_facets))
    tb["facet_res"] = fres
    frough = -np.ones(nf, dtype=np.int64)
    frough[np.asarray(geo.rough_facets, dtype=int)] = np.arange(len(geo.rough_facets))
    tb["facet_rough"] = frough
    ptr = [0]
    flat = []
    for fct in mesh.facets:
        flat.extend(int(i) for i in fct)
        ptr.append(len(flat))
    tb["facet_faces_ptr"] = np.array(ptr, dtype=np.int64)
    tb["facet_faces"] = np.array(flat, dtype=np.int64)
    tb["bounds"] = np.array(geo.bounds, dtype=float)

    tb["sv_centres"] = np.array(geo.subvol_center, dtype=float)
    tb["sv_volume"] = np.array(geo.subvol_volume, dtype=float)
    tb["sv_slice"] = bool(geo.subvol_type == "slice")
    tb["slice_axis"] = int(getattr(geo, "slice_axis", 0))
    tb["temp_interp"] = str(pop.temp_interp_type)
    tb["res_gen"] = str(pop.res_gen)
    # coordinates the non-slice interpolator sees (Population.py:697-702): a grid with a collapsed direction drops it
    if geo.subvol_type == "grid" and np.any(np.asarray(geo.grid) == 1):
        tb["interp_dims"] = np.nonzero(np.asarray(geo.grid) != 1)[0].astype(np.int64)
    else:
        tb["interp_dims"] = np.arange(3, dtype=np.int64)

    tb["omega"] = np.array(ph.omega, dtype=float)
    tb["group_vel"] = np.array(ph.group_vel, dtype=float)
    tb["tau"] = np.array(ph.lifetime, dtype=float)
    tb["T_grid"] = np.array(ph.temperature_array, dtype=float)
    tb["energy_array"] = np.array(ph.energy_array, dtype=float)
    tb["T_array"] = np.array(ph.temperature_function.y, dtype=float)
    tb["hbar"] = float(ph.hbar)
    tb["kb"] = float(ph.kb)
    tb["volume_unitcell"] = float(ph.volume_unitcell)
    tb["n_active"] = int(ph.number_of_active_modes)
    tb["eVpsa2_in_Wm2"] = float(pop.eVpsa2_in_Wm2)
    tb["a_in_m"] = float(pop.a_in_m)

    tb["dt"] = float(pop.dt)
    tb["norm_mean"] = bool(pop.norm == "mean")
    tb["particle_density"] = float(pop.particle_density)
    tb["n_dt_to_conv"] = int(pop.n_dt_to_conv)
    R = int(pop.n_of_reservoirs)
    Q, J = ph.omega.shape
    if R > 0:
        tb["res_facet"] = np.array(pop.res_facet, dtype=np.int64)
        tb["res_T"] = np.array(pop.res_facet_temperature, dtype=float)
        tb["enter_prob"] = np.array(pop.enter_prob, dtype=float)
    else:
        tb["res_facet"] = np.zeros(0, dtype=np.int64)
        tb["res_T"] = np.zeros(0)
        tb["enter_prob"] = np.zeros((0, Q, J))

    Fr = len(geo.rough_facets)
    tb["specularity"] = np.array(pop.specularity, dtype=float).reshape(Fr, Q, J)
    tb["true_specular"] = np.array(pop.true_specular, dtype=bool).reshape(Fr, Q, J)
    spec_out = -np.ones((Fr, Q, J), dtype=np.int64)
    if Fr > 0 and pop.correspondent_modes.shape[0] > 0:
        for i_f, facet in enumerate(geo.rough_facets):
            qq, jj = np.nonzero(tb["true_specular"][i_f])
            if qq.shape[0] == 0:
                continue
            a = np.hstack((np.tile(-geo.facets_normal[facet, :], (qq.shape[0], 1)), qq.reshape(-1, 1), jj.reshape(-1, 1)))
            out = pop.specular_function(a).astype(int)               # Population.py:959-961
            spec_out[i_f, qq, jj] = out[:, 0] * J + out[:, 1]
    tb["spec_out"] = spec_out
    tb["roulette"] = np.array(getattr(pop, "creation_roulette", np.zeros((Fr, Q * J))), dtype=float).reshape(Fr, Q * J)
    return tb


def state_from_reference(ph, pop):
    J = ph.omega.shape[1]
    R = int(pop.n_of_reservoirs)
    n = pop.positions.shape[0]
    st = nko.State(
        positions=np.array(pop.positions, dtype=float),
        modes=np.array(pop.modes, dtype=int),
        omega=np.array(pop.omega, dtype=float),
        group_vel=np.array(pop.group_vel, dtype=float),
        occupation=np.array(pop.occupation, dtype=float),
        n_timesteps=np.array(pop.n_timesteps, dtype=float),
        collision_facets=np.array(pop.collision_facets, dtype=int),
        collision_positions=np.array(pop.collision_positions, dtype=float),
        collision_cond=np.array([nko.BC_CODE[c] for c in pop.collision_cond], dtype=np.int8),
        temperatures=np.array(pop.temperatures, dtype=float),
        ids=np.arange(n, dtype=np.int64),
        subvol_temperature=np.array(pop.subvol_temperature, dtype=float),
        res_counter=np.array(pop.res_counter, dtype=float) if R > 0 else np.zeros((0,) + ph.omega.shape),
    )
    st.omega_modes = st.modes[:, 0] * J + st.modes[:, 1]
    st.subvol_id = np.array(pop.subvol_id, dtype=int)
    st.energies = np.array(pop.energies, dtype=float)
    st.subvol_energy = np.array(pop.subvol_energy, dtype=float)
    st.subvol_N_p = np.array(pop.subvol_N_p, dtype=int)
    st.N_p = int(pop.N_p)
    st.subvol_heat_flux = np.array(pop.subvol_heat_flux, dtype=float)
    st.res_energy_balance = np.array(pop.res_energy_balance, dtype=float) if R > 0 else np.zeros(0)
    st.res_heat_flux = np.array(pop.res_heat_flux, dtype=float) if R > 0 else np.zeros((0, 3))
    st.N_leaving = np.array(pop.N_leaving, dtype=int) if R > 0 else np.zeros(0, dtype=int)     # Population.py:344
    st.current_timestep = int(pop.current_timestep)
    return st


def reference_step(pop, geo, ph):
    """``run_timestep`` minus the every-100-step output branch (Population.py:1743-1769)."""
    pop.drift()
    if pop.n_of_reservoirs > 0:
        if pop.res_gen == "one_to_one":                                  # Population.py:1746-1749
            pop.fill_reservoirs(geo, ph, n_leaving=pop.N_leaving)
        else:
            pop.fill_reservoirs(geo, ph)
        pop.add_reservoir_particles(geo)
    pop.boundary_scattering(geo, ph)
    pop.refresh_temperatures(geo, ph)
    pop.lifetime_scattering(ph)
    pop.current_timestep += 1
    pop.t = pop.current_timestep * pop.dt
    conv = None
    if (pop.current_timestep % pop.n_dt_to_conv) == 0:
        pop.subvol_heat_flux = pop.calculate_heat_flux(geo, ph)
        pop.calculate_kappa(geo)
        pop.adjust_reservoir_balance(geo, ph)
        conv = dict(subvol_heat_flux=np.array(pop.subvol_heat_flux), res_heat_flux=np.array(pop.res_heat_flux),
                    res_energy_balance=np.array(pop.res_energy_balance),
                    subvol_kappa=np.array(getattr(pop, "subvol_kappa", np.zeros(0))), kappa=float(getattr(pop, "kappa", 0.0)))
        pop.restart_reservoir_balance()
    return conv
