"""Triangle mesh of the simulation domain (host-side set-up).

Same public surface as the reference's ``classes/Mesh.py`` for what the particle loop and the
geometry set-up consume: the per-triangle plane tables of ``find_boundary`` (Mesh.py:205-243,
:314-324), coplanar facets (:244-308), outward winding (:114-160), closest face/facet queries
(:686-738), ``contains``, ``sample_surface`` (:923-951) and ``sample_volume`` (:890-904).

Written from scratch: facets are found with a union-find over coplanar edge-adjacent triangles,
winding by counting forward crossings of the face normal, the volume by the divergence theorem and
``sample_volume`` by rejection in the bounding box -- no Delaunay tetrahedralisation is needed.
``find_boundary`` is the GPU kernel (``nk_find_boundary``) when an engine is attached; the NumPy
version below is only used during set-up, before any device context exists.
"""
from __future__ import annotations

import os

import numpy as np


class Mesh:
    def __init__(self, vertices, faces, remove_unref=True, triangulate_volume=True):
        vertices = np.asarray(vertices, dtype=float)
        if vertices.shape[1] == 2:
            vertices = np.hstack((vertices, np.zeros((vertices.shape[0], 1))))
        self.vertices = vertices.astype(float)
        self.faces = np.asarray(faces).astype(int)
        self.tol = 1e-10
        self.engine = None
        self.update_mesh_properties(remove_unref)

    # ------------------------------------------------------------------------------------------
    def update_mesh_properties(self, remove_unref=True, triangulate_volume=True):
        if remove_unref:
            self.remove_unref_vertices()
        self.n_of_vertices = self.vertices.shape[0]
        self.bounds = np.vstack((self.vertices.min(axis=0), self.vertices.max(axis=0)))
        self.extents = np.ptp(self.bounds, axis=0)
        self._face_tables()
        self._edges()
        self._facets()
        self._interfaces()
        self.check_winding()
        self._volume()

    def remove_unref_vertices(self):
        used = np.unique(self.faces)
        if used.shape[0] == self.vertices.shape[0]:
            return
        remap = -np.ones(self.vertices.shape[0], dtype=int)
        remap[used] = np.arange(used.shape[0])
        self.vertices = self.vertices[used]
        self.faces = remap[self.faces]

    def _face_tables(self):
        v = self.vertices
        f = self.faces
        self.n_of_faces = f.shape[0]
        b1 = v[f[:, 1]] - v[f[:, 0]]
        b2 = v[f[:, 2]] - v[f[:, 0]]
        cr = np.cross(b1, b2)
        nrm = np.linalg.norm(cr, axis=1)
        # the reference takes the 1-D norm face by face (Mesh.py:220: sqrt(dot(x, x)), BLAS), which differs from the
        # axis-wise norm by one ulp for some triangles; areas feed the emission entry probabilities, so match it exactly
        self.face_areas = np.array([np.linalg.norm(c) for c in cr]) / 2 if cr.shape[0] else nrm / 2
        self.area = float(self.face_areas.sum())
        self.face_centroid = v[f].mean(axis=1)
        self.face_normals = cr / nrm[:, None]
        self.face_basis = np.stack((b1, b2), axis=0)
        # columns b1, b2, n  (reference face_basis_matrix, Mesh.py:231-232)
        self.face_basis_matrix = np.stack((b1, b2, self.face_normals), axis=2)
        self.face_origins = v[f[:, 0]].copy()
        self.face_k = -np.sum(self.face_normals * self.face_origins, axis=1)
        self.face_bounds = np.stack((v[f].min(axis=1), v[f].max(axis=1)), axis=0)   # (2,F,3)

    def get_face_k(self):
        self.face_k = -np.sum(self.face_normals * self.face_origins, axis=1)

    def get_facets_k(self):
        self.facets_k = -np.sum(self.facets_normal * self.facets_origin, axis=1)

    def _edges(self):
        f = self.faces
        e = np.sort(np.vstack((f[:, [0, 1]], f[:, [0, 2]], f[:, [1, 2]])), axis=1)
        self.edges, inv = np.unique(e, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        F = f.shape[0]
        self.n_of_edges = self.edges.shape[0]
        self.face_edges = np.stack((inv[:F], inv[F:2 * F], inv[2 * F:]), axis=1)
        self.edges_faces = [[] for _ in range(self.n_of_edges)]
        for fi in range(F):
            for ei in self.face_edges[fi]:
                self.edges_faces[ei].append(fi)
        self.edges_faces = [np.array(sorted(set(x)), dtype=int) for x in self.edges_faces]
        adj = []
        for fs in self.edges_faces:
            for a in range(len(fs)):
                for b in range(a + 1, len(fs)):
                    adj.append((fs[a], fs[b]))
        self.face_adjacency = np.array(sorted(set(adj)), dtype=int).reshape(-1, 2)

    def _facets(self):
        """Maximal sets of edge-adjacent coplanar triangles (reference criterion Mesh.py:260-267:
        |n1.n2| > 1 - tol and |k1| - |k2| < tol)."""
        F = self.n_of_faces
        parent = np.arange(F)

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        n = self.face_normals
        k = -np.sum(n * self.vertices[self.faces[:, 0]], axis=1)
        for a, b in self.face_adjacency:
            if abs(float(np.dot(n[a], n[b]))) > 1 - self.tol and (abs(k[a]) - abs(k[b])) < self.tol:
                ra, rb = find(a), find(b)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
        roots = np.array([find(i) for i in range(F)])
        order = sorted(set(roots.tolist()))          # facets numbered by their lowest face index
        self.facets = [np.nonzero(roots == r)[0].astype(int) for r in order]
        self.n_of_facets = len(self.facets)
        self.face_facets = np.zeros(F, dtype=int)
        for i, fct in enumerate(self.facets):
            self.face_facets[fct] = i
        self._facet_tables()

    def _facet_tables(self):
        self.facets_normal = np.array([self.face_normals[fct[0]] for fct in self.facets])
        self.facets_area = np.array([self.face_areas[fct].sum() for fct in self.facets])
        self.facet_centroid = np.array([np.sum(self.face_centroid[fct] * self.face_areas[fct].reshape(-1, 1), axis=0) / self.facets_area[i]
                                        for i, fct in enumerate(self.facets)])
        self.facets_origin = np.array([self.vertices[self.faces[fct[0], 0]] for fct in self.facets])
        self.facet_vertices = [self.vertices[np.unique(self.faces[fct])] for fct in self.facets]
        self.facets_edges, self.facets_boundary = [], []
        for fct in self.facets:
            e, c = np.unique(self.face_edges[fct].ravel(), return_counts=True)
            self.facets_edges.append(e.astype(int))
            self.facets_boundary.append(e[c == 1].astype(int))
        self.get_facets_k()

    def _interfaces(self):
        """Internal facets (all boundary edges shared by > 2 faces), reference Mesh.py:329-352."""
        multi = {e for e, fs in enumerate(self.edges_faces) if len(fs) > 2}
        self.interfacets = np.array([i for i, b in enumerate(self.facets_boundary) if len(b) > 0 and all(int(e) in multi for e in b)], dtype=int)
        if self.interfacets.shape[0] > 0:
            self.interfaces = np.unique(np.concatenate([self.facets[i] for i in self.interfacets])).astype(int)
        else:
            self.interfaces = np.array([], dtype=int)

    def faces_to_facets(self, index_faces):
        return self.face_facets[index_faces]

    def face_to_facet(self, f):
        return int(self.face_facets[f]) if 0 <= f < self.n_of_faces else -1

    # ------------------------------------------------------------------------------------------
    def check_winding(self):
        """Make every face normal point out of the solid: a normal that crosses the rest of the
        surface an odd number of times points inward (reference Mesh.py:114-160)."""
        o = self.face_centroid
        n = self.face_normals
        flip = np.zeros(self.n_of_faces, dtype=bool)
        for f in range(self.n_of_faces):
            _, t, faces_hit, pts = self._ray_all(o[f], n[f], skip=f)
            if len(t) == 0:
                continue
            pts = np.unique(np.around(pts, decimals=8), axis=0)
            if pts.shape[0] % 2 == 1:
                flip[f] = True
        if flip.any():
            self.faces[flip] = self.faces[flip][:, [1, 0, 2]]
            self._face_tables()
            self._edges()
            self._facets()
            self._interfaces()

    def _ray_all(self, x, v, skip=None):
        """All forward crossings of one ray with the surface (set-up helper)."""
        n, k = self.face_normals, self.face_k
        with np.errstate(divide='ignore', invalid='ignore'):
            t = -(n @ x + k) / (n @ v)
        ok = np.isfinite(t) & (t > self.tol)
        if skip is not None:
            ok[skip] = False
        if len(self.interfaces):
            ok[self.interfaces] = False
        idx = np.nonzero(ok)[0]
        if idx.shape[0] == 0:
            return x, np.zeros(0), idx, np.zeros((0, 3))
        c = x + t[idx, None] * v
        bar = np.linalg.solve(self.face_basis_matrix[idx], (c - self.face_origins[idx])[..., None])[..., 0]
        a, b = bar[:, 0], bar[:, 1]
        w = 1 - a - b
        inside = (np.around(a, 10) >= 0) & (np.around(b, 10) >= 0) & (np.around(w, 10) >= 0) & \
                 (np.around(a, 10) <= 1) & (np.around(b, 10) <= 1) & (np.around(w, 10) <= 1)
        return x, t[idx][inside], idx[inside], c[inside]

    def _volume(self):
        v = self.vertices
        f = self.faces
        ref = self.bounds.mean(axis=0)
        a, b, c = v[f[:, 0]] - ref, v[f[:, 1]] - ref, v[f[:, 2]] - ref
        vol6 = np.einsum('ij,ij->i', a, np.cross(b, c))
        keep = np.ones(f.shape[0], dtype=bool)
        if len(self.interfaces):
            keep[self.interfaces] = False
        self.volume = float(abs(vol6[keep].sum()) / 6.0)
        cent = (a + b + c) / 4.0
        if self.volume > 0:
            self.center_mass = ref + (cent[keep] * vol6[keep, None]).sum(axis=0) / vol6[keep].sum()
        else:
            self.center_mass = self.facet_centroid[0].copy()
        self.n_of_simplices = 0

    def rezero(self):
        dx = self.vertices.min(axis=0)
        self.vertices = self.vertices - dx
        for name in ("face_centroid", "facet_centroid", "face_origins", "facets_origin", "center_mass", "bounds"):
            setattr(self, name, getattr(self, name) - dx)
        self.face_bounds = self.face_bounds - dx
        self.get_face_k()
        self.get_facets_k()

    def rotate(self, rot_order, rotation, degrees=True):
        from scipy.spatial.transform import Rotation as rot
        self.vertices = rot.from_euler(rot_order, rotation, degrees=degrees).apply(self.vertices)

    # ------------------------------------------------------------------------------------------
    def closest_face(self, x):
        """Closest face onto which each point projects inside the triangle (lowest index on ties);
        -1 when none.  Reference Mesh.py:686-720."""
        x = np.asarray(x, dtype=float).reshape(-1, 3)
        P = x.shape[0]
        dist = np.sum(self.face_normals[None] * (x[:, None, :] - self.face_origins[None]), axis=2)      # (P,F)
        pj = x[:, None, :] - self.face_normals[None] * dist[..., None]
        valid = np.all(pj >= self.face_bounds[0] - self.tol, axis=2) & np.all(pj <= self.face_bounds[1] + self.tol, axis=2)
        ip, jf = valid.nonzero()
        if ip.shape[0]:
            bar = np.linalg.solve(self.face_basis_matrix[jf], (pj[ip, jf] - self.face_origins[jf])[..., None])[..., 0][:, :2]
            bar = np.concatenate((bar, 1 - bar.sum(axis=1, keepdims=True)), axis=1)
            valid[ip, jf] = np.all((bar >= -self.tol) & (bar <= 1 + self.tol), axis=1)
        d = np.where(valid, np.abs(dist), np.inf)
        f = np.argmin(d, axis=1)
        dmin = d[np.arange(P), f]
        f = np.where(np.isinf(dmin), -1, f)
        return f.astype(int), dmin, pj[np.arange(P), np.maximum(f, 0)]

    def closest_facet(self, x):
        f, d, xc = self.closest_face(x)
        ok = f >= 0
        f[ok] = self.face_facets[f[ok]]
        return f, d, xc

    def contains(self, x):
        """Point-in-solid by crossing parity.  A ray that grazes an edge shared by two triangles is counted twice (or not at
        all), and "generic" directions with simple ratios do meet edges of meshes with simple aspect ratios (a (3, 2, 1) ray from
        the centre line of a 2:1 face runs along its diagonal), so three unrelated directions vote."""
        x = np.asarray(x, dtype=float).reshape(-1, 3)
        dirs = np.array([[0.6827316519, 0.5341127043, 0.4985163227],
                         [-0.3711942835, 0.8260437511, 0.4241559872],
                         [0.2903178467, -0.4478291235, 0.8456771093]])
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        out = np.zeros(x.shape[0], dtype=bool)
        inb = np.nonzero(np.all(x >= self.bounds[0] - self.tol, axis=1) & np.all(x <= self.bounds[1] + self.tol, axis=1))[0]
        keep = np.ones(self.n_of_faces, dtype=bool)
        if len(self.interfaces):
            keep[self.interfaces] = False
        n, k = self.face_normals[keep], self.face_k[keep]
        inv = np.linalg.inv(self.face_basis_matrix[keep])                      # (F,3,3)
        org = self.face_origins[keep]
        chunk = max(1, int(2e6 // max(1, n.shape[0])))
        for s in range(0, inb.shape[0], chunk):
            idx = inb[s:s + chunk]
            p = x[idx]
            def parity(q, d):
                den = n @ d
                with np.errstate(divide='ignore', invalid='ignore'):
                    t = -(q @ n.T + k) / den                                      # (P,F)
                c = q[:, None, :] + t[..., None] * d - org[None]
                bar = np.einsum('fij,pfj->pfi', inv, c)
                a, b = bar[..., 0], bar[..., 1]
                hit = np.isfinite(t) & (t > self.tol) & (a >= 0) & (b >= 0) & (a + b <= 1)
                return (hit.sum(axis=1) % 2) == 1
            first, second = parity(p, dirs[0]), parity(p, dirs[1])
            res = first & second
            tie = np.nonzero(first != second)[0]                                   # a grazed edge: the third ray decides
            if tie.shape[0]:
                res[tie] = parity(p[tie], dirs[2])
            out[idx] = res
        return out

    contains_naive = contains

    def find_boundary(self, x, v):
        """(xc, tc, fc) of the nearest forward hit (reference Mesh.py:806-856).  GPU kernel when an
        engine is attached; otherwise a NumPy evaluation for set-up-time use."""
        if self.engine is not None:
            return self.engine.find_boundary(x, v)
        x = np.asarray(x, dtype=float).reshape(-1, 3)
        v = np.asarray(v, dtype=float).reshape(-1, 3)
        n, k = self.face_normals, self.face_k
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            t = -(np.sum(x[:, None, :] * n, axis=2) + k) / np.sum(v[:, None, :] * n, axis=2)
        ok = (t >= self.tol) & np.isfinite(t)
        ip, jf = ok.nonzero()
        c = x[ip] + t[ip, jf, None] * v[ip]
        inb = np.all(c >= self.face_bounds[0, jf] - self.tol, axis=1) & np.all(c <= self.face_bounds[1, jf] + self.tol, axis=1)
        ok[ip, jf] = inb
        ip, jf, c = ip[inb], jf[inb], c[inb]
        if ip.shape[0]:
            bar = np.linalg.solve(self.face_basis_matrix[jf], (c - self.face_origins[jf])[..., None])[..., 0][:, :2]
            bar = np.concatenate((bar, 1 - bar.sum(axis=1, keepdims=True)), axis=1)
            ok[ip, jf] = np.all((bar >= -self.tol) & (bar <= 1 + self.tol), axis=1)
        t = np.where(ok, t, np.inf)
        tc = t.min(axis=1)
        fc = self.face_facets[np.argmax(t == tc[:, None], axis=1)].astype(int)
        fc[np.isinf(tc)] = -1
        with np.errstate(invalid='ignore'):
            xc = x + tc[:, None] * v
        return xc, tc, fc

    def sample_surface(self, n, faces=None, facets=None, rng=None):
        """Uniform points on the given facets (reference Mesh.py:923-951)."""
        rng = np.random if rng is None else rng
        if facets is None and faces is None:
            faces = np.arange(self.n_of_faces, dtype=int)
        elif facets is not None:
            facets = np.atleast_1d(np.asarray(facets, dtype=int))
            faces = np.concatenate([self.facets[f] for f in facets])
        p = self.face_areas[faces] / self.face_areas[faces].sum()
        f = rng.choice(faces, size=n, p=p)
        v = self.vertices[self.faces[f]]
        s = np.sqrt(rng.random((n, 1)))
        r = rng.random((n, 1))
        return (1 - s) * v[:, 0] + (1 - r) * s * v[:, 1] + r * s * v[:, 2]

    def sample_volume(self, n, rng=None):
        """Uniform points inside the solid by rejection in the bounding box."""
        rng = np.random if rng is None else rng
        if self.volume <= 0:
            raise Exception('The mesh has no volume to sample from.')
        convex_box = self.n_of_facets == 6 and abs(self.volume - float(np.prod(self.extents))) <= 1e-9 * self.volume
        out = np.zeros((0, 3))
        while out.shape[0] < n:
            m = int((n - out.shape[0]) * (1.0 if convex_box else float(np.prod(self.extents)) / self.volume * 1.1)) + 8
            x = rng.random((m, 3)) * self.extents + self.bounds[0]
            if not convex_box:
                x = x[self.contains(x)]
            out = np.vstack((out, x))
        return out[:n]

    def export_stl(self, name, path=None):
        path = os.getcwd() if path is None else path
        name = name.replace('.stl', '')
        lines = ['solid {:s}'.format(name)]
        for f in range(self.n_of_faces):
            lines.append('facet normal {:.6e} {:.6e} {:.6e}'.format(*self.face_normals[f]))
            lines.append('    outer loop')
            for k in range(3):
                lines.append('        vertex {:.6e} {:.6e} {:.6e}'.format(*self.vertices[self.faces[f, k]]))
            lines.append('    endloop')
            lines.append('endfacet')
        lines.append('endsolid {:s}'.format(name))
        with open(os.path.join(path, name + '.stl'), 'w') as fh:
            fh.write('\n'.join(lines))


def read_stl(path):
    """ASCII or binary STL -> (vertices, faces) with coincident vertices merged (the reference uses
    trimesh.load for this, Geometry.py:82-84)."""
    with open(path, 'rb') as fh:
        data = fh.read()
    tri = None
    head = data[:512].lstrip().lower()
    if head.startswith(b'solid') and b'facet' in data[:4096].lower():
        vals = []
        for line in data.decode('ascii', errors='ignore').splitlines():
            p = line.split()
            if len(p) == 4 and p[0].lower() == 'vertex':
                vals.append([float(p[1]), float(p[2]), float(p[3])])
        tri = np.array(vals, dtype=float).reshape(-1, 3, 3)
    else:
        n = int(np.frombuffer(data[80:84], dtype='<u4')[0])
        rec = np.frombuffer(data[84:84 + 50 * n], dtype=np.dtype([('n', '<f4', 3), ('v', '<f4', (3, 3)), ('a', '<u2')]))
        tri = rec['v'].astype(float)
    pts = np.around(tri.reshape(-1, 3), decimals=10)
    verts, inv = np.unique(pts, axis=0, return_inverse=True)
    return verts, np.asarray(inv).reshape(-1, 3)
