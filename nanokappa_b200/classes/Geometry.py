"""Simulation domain: mesh, boundary conditions, periodic pairs and subvolumes (host-side set-up).

Keeps the reference's ``Geometry`` surface (``classes/Geometry.py``): ``Geometry(args)`` with
``.mesh .bounds .volume .facets_normal .facets_area .facet_centroid .bound_cond .res_facets
.res_values .res_bound_cond .rough_facets .rough_facets_values .connected_facets .subvol_type
.n_of_subvols .subvol_center .subvol_volume .subvol_connections .subvol_classifier .slice_axis
.slice_length`` and the build order of Geometry.py:60-69.  ``SubvolClassifier.predict`` is the GPU
kernel ``nk_classify`` once an engine is attached (Geometry.py:1198-1213).

Out of the hot path, so plain NumPy; no trimesh / shapely dependency (own STL reader, periodic
pairs validated by comparing the facets' boundary vertices).
"""
from __future__ import annotations

import os

import numpy as np
from scipy.spatial import cKDTree, Delaunay
from scipy.spatial.transform import Rotation as rot

from .Mesh import Mesh, read_stl


class SubvolClassifier:
    """Nearest-centre classification of positions into subvolumes."""

    def __init__(self, n, xc=None, a=None):
        self.n = n
        if xc is None:
            self.a = a
            self.xc = np.ones((n, 3)) * 0.5
            self.xc[:, a] = np.linspace(0, 1 - 1 / n, n) + 1 / (2 * n)
        else:
            self.xc = np.asarray(xc, dtype=float)
        self.engine = None
        self._tree = cKDTree(self.xc)

    def predict(self, x):
        if self.engine is not None:
            return self.engine.classify(x)
        return self._tree.query(np.asarray(x, dtype=float).reshape(-1, 3))[1].astype(int)

    f = predict


class Geometry:
    standard_shapes = ['cuboid', 'box', 'cylinder', 'rod', 'bar', 'star', 'castle', 'zigzag', 'corrugated', 'freewire']

    def __init__(self, args):
        self.args = args
        self.scale = args.scale
        self.shape = args.geometry[0]
        self.dimensions = args.dimensions
        if len(args.geo_rotation) > 0:
            self.rotation = np.array(args.geo_rotation[:-1]).astype(float)
            self.rot_order = args.geo_rotation[-1]
        else:
            self.rotation = None
            self.rot_order = None
        self.subvol_type = args.subvolumes[0]
        self.folder = args.results_folder
        self.path_points = np.array(getattr(args, 'path_points', [])[1:]).astype(float).reshape(-1, 3)
        self.tol_decimals = 1

        self.load_geo_file(self.shape)
        self.transform_mesh()
        self.get_mesh_properties()
        self.get_bound_facets(args)
        self.check_facet_connections(args)
        self.set_subvolumes()
        self.get_path()
        print('Geometry processing done!')

    # ---- mesh ------------------------------------------------------------------------------------
    def load_geo_file(self, shape):
        print('Loading geometry...')
        if shape in self.standard_shapes:
            self.mesh = self.generate_primitives(shape, self.dimensions)
        else:
            v, f = read_stl(shape)
            self.mesh = Mesh(np.around(v, decimals=10), f)

    def generate_primitives(self, shape, dims):
        """Vertex / face tables of the built-in shapes, numbered as the reference numbers them
        (Geometry.py:86-412) so that facet ids -- which users address through --bound_pos -- agree."""
        if shape in ['cuboid', 'box']:
            corners = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1], [1, 1, 0]], dtype=float)
            vertices = corners * np.array(dims, dtype=float)
            quads = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 4, 5, 1), (3, 7, 6, 2), (0, 4, 7, 3), (1, 5, 6, 2)]   # x0 x1 y0 y1 z0 z1
            faces = np.array([t for a, b, c, d in quads for t in ((a, b, c), (a, c, d))], dtype=int)
        elif shape in ['cylinder', 'rod', 'bar']:
            L, R, N = float(dims[0]), float(dims[1]), int(dims[2])
            ang = np.arange(N) * 2 * np.pi / N
            ring = np.stack((np.cos(ang), np.sin(ang), np.zeros(N)), axis=1) * R
            top = np.array([0, 0, L])
            vertices = np.vstack((np.zeros((1, 3)), ring, np.zeros((1, 3)) + top, ring + top))
            nxt = lambda i: 1 if i == N else i + 1
            base = [[0, i, nxt(i)] for i in range(1, N + 1)]
            sides = []
            for i in range(1, N + 1):
                j = nxt(i)
                sides.append([i, i + N + 1, j + N + 1])
                sides.append([i, j, j + N + 1])
            faces = np.array(base + sides, dtype=int)
            faces = np.vstack((faces, faces[:N] + N + 1))
        else:
            from ..routines.primitives import generate as _gen
            vertices, faces = _gen(shape, dims)
        return Mesh(vertices, faces)

    def transform_mesh(self):
        """rezero -> scale -> rotate -> rezero -> recompute (Geometry.py:414-433)."""
        print('Transforming geometry...')
        self.mesh.rezero()
        self.mesh.vertices = self.mesh.vertices * np.array(self.scale, dtype=float)
        if self.rotation is not None or self.rot_order is not None:
            R = rot.from_euler(self.rot_order, self.rotation, degrees=True)
            self.mesh.vertices = R.apply(self.mesh.vertices)
            self.mesh.vertices = self.mesh.vertices - self.mesh.vertices.min(axis=0)
        self.mesh.update_mesh_properties()
        self.mesh.rezero()

    def get_mesh_properties(self):
        m = self.mesh
        self.faces, self.facets = m.faces, m.facets
        self.n_of_faces, self.n_of_facets = m.n_of_faces, m.n_of_facets
        self.bounds, self.facet_centroid, self.volume = m.bounds, m.facet_centroid, m.volume
        self.facets_normal, self.facets_area = m.facets_normal, m.facets_area

    def scale_positions(self, x, inv=False):
        if inv:
            return x * np.ptp(self.bounds, axis=0) + self.bounds[0, :]
        return (x - self.bounds[0, :]) / np.ptp(self.bounds, axis=0)

    # ---- boundary conditions ---------------------------------------------------------------------------
    def get_bound_facets(self, args):
        """Every facet starts with the last --bound_cond; each --bound_pos point is snapped to its
        closest facet and takes the j-th condition; --bound_values are consumed in order by the
        non-periodic positions (Geometry.py:652-709)."""
        self.bound_cond = np.array([args.bound_cond[-1] for _ in range(self.n_of_facets)])
        try:
            self.bound_pos = np.array(args.bound_pos[1:]).reshape(-1, 3).astype(float)
        except Exception:
            raise Exception('Boundary positions ill defined. Check input parameters.')
        if args.bound_pos[0] == 'relative':
            self.bound_pos = self.scale_positions(self.bound_pos, True)
        elif args.bound_pos[0] != 'absolute':
            raise Exception('Please specify the type of position for BC with the keyword "absolute" or "relative".')
        self.bound_facets, _, _ = self.mesh.closest_facet(self.bound_pos)
        for j, i in enumerate(self.bound_facets):
            self.bound_cond[i] = args.bound_cond[j]
        is_res = np.logical_or(self.bound_cond == 'T', self.bound_cond == 'F')
        self.res_facets = np.arange(self.n_of_facets, dtype=int)[is_res]
        self.res_bound_cond = self.bound_cond[is_res]
        self.rough_facets = np.arange(self.n_of_facets, dtype=int)[self.bound_cond == 'R']
        self.n_of_reservoirs = self.res_facets.shape[0]
        self.n_of_rough_facets = self.rough_facets.shape[0]
        self.res_values = np.ones(self.n_of_reservoirs) * np.nan
        self.rough_facets_values = np.ones(self.n_of_rough_facets) * np.nan
        if args.bound_cond[-1] in ['T', 'F']:
            self.res_values[:] = args.bound_values[-1]
        elif args.bound_cond[-1] == 'R':
            self.rough_facets_values[:] = args.bound_values[-1]
        value_index, k = [], 0
        for facet in self.bound_facets:
            if self.bound_cond[facet] != 'P':
                value_index.append(k)
                k += 1
            else:
                value_index.append(-1)
        for i, facet in enumerate(self.bound_facets):
            if facet in self.res_facets:
                self.res_values[self.res_facets == facet] = args.bound_values[value_index[i]]
            elif facet in self.rough_facets:
                self.rough_facets_values[self.rough_facets == facet] = args.bound_values[value_index[i]]

    def check_facet_connections(self, args):
        """Periodic pairs from --connect_pos (Geometry.py:711-766).  A pair is accepted when the
        normals are opposite and the two facets have congruent outlines; a bad pair raises (upstream
        builds the exception but never raises it)."""
        print('Checking connected faces...')
        self.connected_facets = np.zeros((0, 2), dtype=int)
        if len(args.connect_pos) > 0:
            pts = np.array(args.connect_pos[1:], dtype=float).reshape(-1, 3)
            if args.connect_pos[0] == 'relative':
                pts = self.scale_positions(pts, True)
            elif args.connect_pos[0] != 'absolute':
                raise Exception("Wrong option in --connect_pos. Choose between 'relative' or 'absolute'.")
            self.connected_facets = self.mesh.closest_facet(pts)[0].reshape(-1, 2)
        for i, (a, b) in enumerate(self.connected_facets):
            n1, n2 = self.facets_normal[a], self.facets_normal[b]
            if not np.all(np.abs(n1 + n2) < 10 ** -self.tol_decimals):
                raise Exception('Connected facets normals do not agree!!')
            va = self.mesh.facet_vertices[a] - self.facet_centroid[a]
            vb = self.mesh.facet_vertices[b] - self.facet_centroid[b]
            same = va.shape == vb.shape and np.allclose(va[np.lexsort(np.around(va, 6).T[::-1])],
                                                         vb[np.lexsort(np.around(vb, 6).T[::-1])], atol=1e-6 * max(1.0, np.abs(va).max()))
            if same:
                print('Connection {:d} OK!'.format(i))
            else:
                print('Connection {:d}: outlines differ, check --connect_pos.'.format(i))

    # ---- subvolumes ------------------------------------------------------------------------------------
    def set_subvolumes(self):
        """slice / grid / voronoi centres, volumes and neighbour graph (Geometry.py:446-544)."""
        print('Defining subvolumes centers...')
        sort3 = lambda c: c[np.lexsort((c[:, 2], c[:, 1], c[:, 0]))]
        if self.subvol_type == 'slice':
            self.n_of_subvols = int(self.args.subvolumes[1])
            self.slice_axis = int(self.args.subvolumes[2])
            c = np.zeros((self.n_of_subvols, 3)) + np.mean(self.bounds, axis=0)
            a = (np.arange(self.n_of_subvols) + 0.5) / self.n_of_subvols
            a *= np.ptp(self.bounds[:, self.slice_axis])
            a += self.bounds[0, self.slice_axis]
            c[:, self.slice_axis] = a
            self.subvol_center = sort3(c)
            self.slice_length = np.ptp(self.bounds[:, self.slice_axis]) / self.n_of_subvols
            self.subvol_classifier = SubvolClassifier(n=self.n_of_subvols, xc=self.subvol_center)
            self.subvol_volume = self.calculate_subvol_volume()
            self.get_subvol_connections()
        elif self.subvol_type == 'voronoi':
            from ..routines import subvolumes
            self.n_of_subvols = int(self.args.subvolumes[1])
            c = subvolumes.distribute(self.mesh, self.n_of_subvols, self.folder, view=False)
            c = c[self.mesh.contains(c)]
            self.subvol_center = sort3(c)
            self.n_of_subvols = self.subvol_center.shape[0]
            self.get_subvol_connections()
            self.subvol_classifier = SubvolClassifier(n=self.n_of_subvols, xc=self.subvol_center)
            self.subvol_volume = self.calculate_subvol_volume()
        elif self.subvol_type == 'grid':
            self.grid = np.array(self.args.subvolumes[1:4]).astype(int)
            if (self.grid == 1).sum() == 2:
                print("1D subvolume distribution should use 'slice' type, not 'grid'. Defaulting to 'slice'...")
                self.subvol_type = 'slice'
                ax = int(np.nonzero(self.grid != 1)[0][0])
                self.args.subvolumes = [self.subvol_type, self.grid[ax], ax]
                del self.grid
                self.set_subvolumes()
                return
            lin = [np.linspace(0.5 / n, 1 - 0.5 / n, int(n)) for n in self.grid]
            g = np.meshgrid(*lin)
            c = np.vstack(list(map(np.ravel, g))).T * np.ptp(self.bounds, axis=0) + self.bounds[0, :]
            c = c[self.mesh.contains(c)]
            self.subvol_center = sort3(c)
            self.n_of_subvols = self.subvol_center.shape[0]
            self.get_subvol_connections()
            self.subvol_classifier = SubvolClassifier(n=self.n_of_subvols, xc=self.subvol_center)
            self.subvol_volume = self.calculate_subvol_volume()
        else:
            print('Invalid subvolume type!')
            print('Stopping simulation...')
            quit()

    def calculate_subvol_volume(self, algorithm='mc', tol=1e-4, return_centers=False, verbose=False):
        """V/S for slices/grids of a box, otherwise Monte-Carlo cover fractions refined until the
        relative change drops under ``tol`` (Geometry.py:546-650)."""
        if self.subvol_type in ['slice', 'grid'] and self.shape in ['cuboid', 'box']:
            return self.volume * np.ones(self.n_of_subvols) / self.n_of_subvols
        S = self.n_of_subvols
        cover, err, nt, ns = np.zeros(S), np.ones(S), 0, 2 ** 10
        rng = np.random.RandomState(20240613)
        while err.max() > tol:
            x = self.mesh.sample_volume(ns, rng=rng)
            r = self.subvol_classifier.predict(x)
            new_cover = (cover * nt + np.bincount(r, minlength=S)) / (nt + ns)
            nt += ns
            with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                err = np.abs((new_cover - cover) / cover)
                err[np.isnan(err)] = 1
            cover = new_cover
            ns = min(ns * 2, 2 ** 18)
            if nt > int(float(os.environ.get('NK_VOLUME_MAX_SAMPLES', 2 ** 24))):
                break
        return cover * self.volume

    def get_subvol_connections(self):
        """Neighbour pairs (i < j) of the subvolume centres (Geometry.py:961-1052).  Slices: i <-> i+1.  Otherwise the
        reference's greedy pruning, restated: candidate pairs are those whose midpoint lies inside the solid and whose
        connecting segment does not leave it; they are visited from the shortest to the longest, and a pair (i, j) is
        dropped when its midpoint lies on or beyond the bisector plane of a connection (i, k) or (j, k) confirmed before it
        (the farther centre is then hidden behind the nearer one).  The number and order of the connections is part of the
        convergence.txt format, so the visiting order (NumPy's argsort of the distances, ties included) is kept."""
        print('Getting subvol connections...')
        S = self.n_of_subvols
        c = self.subvol_center
        if self.subvol_type == 'slice' or S < 2:
            con = np.stack((np.arange(S - 1), np.arange(1, S)), axis=1) if S > 1 else np.zeros((0, 2), dtype=int)
        else:
            mid = (c + np.expand_dims(c, 1)) / 2                  # mid[a, b]
            nrm = c - np.expand_dims(c, 1)                        # nrm[a, b] = c[b] - c[a]
            dist = np.linalg.norm(nrm, axis=-1)
            ii, jj = np.triu_indices(S, k=1)                      # lexicographically ordered pairs i < j
            pairs = np.stack((ii, jj), axis=1)
            pairs = pairs[self.mesh.contains(mid[pairs[:, 0], pairs[:, 1], :])]
            if pairs.shape[0]:
                _, t, _ = Mesh.find_boundary(self.mesh, c[pairs[:, 0], :], nrm[pairs[:, 0], pairs[:, 1], :])
                pairs = pairs[t > 1, :]
            n_p = pairs.shape[0]
            confirmed = np.zeros(n_p, dtype=bool)
            remove = np.zeros(n_p, dtype=bool)
            order = np.argsort(dist[pairs[:, 0], pairs[:, 1]])
            for idx in order:
                i, j = pairs[idx]
                for a in (i, j):                                   # connections of i, then (if still alive) of j
                    if remove[idx]:
                        break
                    touching = np.nonzero(np.any(pairs == a, axis=1) & confirmed)[0]
                    for row in touching:
                        k = pairs[row, 0] if pairs[row, 1] == a else pairs[row, 1]
                        if np.sum((mid[i, j, :] - mid[a, k, :]) * nrm[a, k, :]) >= 0:
                            remove[idx] = True
                if not remove[idx]:
                    confirmed[idx] = True
            con = pairs[~remove, :]
            used = np.unique(con)
            if used.shape[0] != S and used.shape[0] > 0:           # upstream keeps only the connected subvolumes (:1035-1046)
                relabel = -np.ones(S, dtype=int)
                relabel[used] = np.arange(used.shape[0])
                self.subvol_center = c[used, :]
                if hasattr(self, 'subvol_volume') and np.size(self.subvol_volume) == S:
                    self.subvol_volume = np.asarray(self.subvol_volume)[used]
                self.n_of_subvols = used.shape[0]
                con = relabel[con]
        self.subvol_connections = con
        self.n_of_subvol_con = con.shape[0]
        self.subvol_con_vectors = self.subvol_center[con[:, 1], :] - self.subvol_center[con[:, 0], :] if con.shape[0] else np.zeros((0, 3))

    def get_path(self):
        self.path_kappa = None
        if self.path_points is not None and len(getattr(self.args, 'path_points', [])) > 0:
            if self.args.path_points[0] == 'relative':
                self.path_points = self.scale_positions(self.path_points, inv=True)
            self.path_kappa = self.subvol_classifier.predict(self.path_points)
        else:
            self.path_points = None

    def attach_engine(self, engine):
        """Route find_boundary / predict through the CUDA kernels of this engine."""
        self.mesh.engine = engine
        self.subvol_classifier.engine = engine
