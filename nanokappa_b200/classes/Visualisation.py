"""Post-processing of ``convergence.txt`` (the part of the reference's ``Visualisation`` that feeds
back into the simulation: ``read_convergence`` -> rolling means / standard deviations used by
``Population.update_residue`` and ``write_final_state``; Visualisation.py:122-212).  The PNG plots of
the reference are produced only when matplotlib is importable and ``--fig_plot`` asks for them; they
are not part of the accelerated path."""
from __future__ import annotations

import os
import warnings

import numpy as np

from .Constants import Constants


class Visualisation(Constants):
    def __init__(self, args, geometry, phonon, population=None):
        super().__init__()
        print('Initialising visualisation class...')
        self.args = args
        self.phonon = phonon
        self.geometry = geometry
        self.population = population
        self.folder = self.args.results_folder
        self.convergence_file = os.path.join(self.folder, 'convergence.txt')
        self.particle_file = os.path.join(self.folder, 'particle_data.txt')
        self.dt = self.args.timestep[0]

    def update_population(self, population, verbose=False):
        self.population = population

    def postprocess(self, verbose=True):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if verbose:
                print('Reading convergence data')
            self.read_convergence()

    def read_convergence(self):
        """Positional parse of convergence.txt: column order is an interface (Population.open_convergence)."""
        # the file only grows while a run is alive: parse the rows that are new since the last call (the every-100-steps
        # postprocess of a 10 000-step run would otherwise re-parse up to a thousand rows a hundred times)
        cache = getattr(self, '_conv_cache', None)
        size = os.path.getsize(self.convergence_file)
        if cache is None or cache['file'] != self.convergence_file or size < cache['offset']:
            cache = self._conv_cache = dict(file=self.convergence_file, offset=0, stamps=[], nums=[])
        with open(self.convergence_file, 'r') as f:
            if cache['offset'] == 0:
                f.readline()                                        # header
            else:
                f.seek(cache['offset'])
            new = f.read()
            cache['offset'] = f.tell()
        for ln in new.splitlines():
            p = ln.split()
            if p:
                cache['stamps'].append(p[0])
                cache['nums'].append(np.array(p[1:], dtype=float))
        stamps = np.array(cache['stamps'])
        num = np.array(cache['nums'], dtype=float).reshape(len(cache['nums']), -1)

        class _Cols:                                                # data[:, a:b] of the text table, already numeric
            def __getitem__(_, key):
                rows, cols = key
                if isinstance(cols, int):
                    return stamps[rows] if cols == 0 else num[rows, cols - 1]
                return num[rows, slice(cols.start - 1, cols.stop - 1)]
        data = _Cols()
        S = self.n_of_subvols = self.geometry.n_of_subvols
        R = self.n_of_reservoirs = self.geometry.n_of_reservoirs
        C = self.n_of_subvol_con = self.geometry.n_of_subvol_con
        self.datetime = data[:, 0].astype('datetime64[us]')
        self.timestep = data[:, 1].astype(int)
        self.sim_time = data[:, 2].astype(float)
        self.total_en = data[:, 3].astype(float)
        c = 4
        self.en_res = data[:, c:c + R].astype(float); c += R
        self.phi_res = data[:, c:c + 3 * R].astype(float); c += 3 * R
        self.N_p = data[:, c].astype(int); c += 1
        self.T = data[:, c:c + S].astype(float); c += S
        self.sv_en = data[:, c:c + S].astype(float); c += S
        self.sv_phi = data[:, c:c + 3 * S].astype(float); c += 3 * S
        self.sv_Np = data[:, c:c + S].astype(float); c += S
        slice_type = self.geometry.subvol_type == 'slice'
        if slice_type:
            self.sv_k = data[:, c:c + S].astype(float); c += S
            self.k = data[:, c].astype(float)
        else:
            self.con_k = data[:, c:c + C].astype(float)
        N = self.n_mean = int(self.args.n_mean[0])
        for name, arr in (('total_en', self.total_en), ('en_res', self.en_res), ('phi_res', self.phi_res), ('Np', self.N_p),
                          ('T', self.T), ('sv_en', self.sv_en), ('sv_phi', self.sv_phi), ('sv_Np', self.sv_Np)):
            setattr(self, 'mean_' + name, arr[-N:].mean(axis=0))
            setattr(self, 'std_' + name, arr[-N:].std(axis=0))
        if slice_type:
            self.mean_sv_k = np.nanmean(self.sv_k[-N:, :], axis=0)
            self.std_sv_k = np.nanstd(self.sv_k[-N:, :], axis=0)
        else:
            con = self.geometry.subvol_connections
            self.mean_con_k = np.nanmean(self.con_k[-N:, :], axis=0)
            self.std_con_k = np.nanstd(self.con_k[-N:, :], axis=0)
            dT = self.T[-N:, con[:, 1]] - self.T[-N:, con[:, 0]]
            self.mean_con_dT, self.std_con_dT = np.nanmean(dT, axis=0), np.nanstd(dT, axis=0)
            self.mean_con_phi, self.std_con_phi = np.zeros(C), np.zeros(C)
            for i, cn in enumerate(con):
                phi = (self.sv_phi[-N:, 3 * cn[0]:3 * (cn[0] + 1)] + self.sv_phi[-N:, 3 * cn[1]:3 * (cn[1] + 1)]) / 2
                dx = np.copy(self.geometry.subvol_con_vectors[i, :])
                dx /= np.linalg.norm(dx)
                proj = np.sum(phi * dx, axis=1)
                self.mean_con_phi[i], self.std_con_phi[i] = np.nanmean(proj), np.nanstd(proj)
            weak = np.absolute(self.mean_con_k) < self.std_con_k
            self.mean_con_k[weak] = np.nan
            self.std_con_k[weak] = np.nan
