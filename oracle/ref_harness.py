"""TEST INFRASTRUCTURE -- loads the UNMODIFIED Nano-kappa reference from ``/root/reference``.

The reference is NumPy-1-era, plot-heavy Python with six third-party imports that are not
installed here.  Nothing under ``/root/reference`` is touched or copied: the sources are read
where they lie, patched *in memory* at import time, and executed.  Three shims:

1. ``sys.modules`` stubs for matplotlib / mpl_toolkits / trimesh / shapely / h5py / phonopy /
   imageio (MagicMock based; ``plt.subplots`` returns a 2-tuple so ``fig, ax = ...`` unpacks).
2. AST rewrite ``expr.ptp(args)`` -> ``np.ptp(expr, args)`` (``ndarray.ptp`` left NumPy 2).
3. AST rewrite ``np.linalg.solve(A, b)`` -> a wrapper restoring the NumPy-1 rule that a ``b`` with
   ``b.ndim == A.ndim - 1`` is a stack of vectors (``Mesh.py:706, :840``).

This module only works where ``/root/reference`` exists (the build container).  The GPU box never
imports it: there the checker is ``oracle/nk_oracle.py`` plus the fixtures in ``tests/golden``.
"""
from __future__ import annotations

import ast
import importlib.abc
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("NK_REFERENCE_ROOT", "/root/reference")
_REF_TOPLEVEL = ("classes", "routines", "argument_parser")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "classes", "Population.py"))


# ----------------------------------------------------------------------------------------------
# shim 3: NumPy-1 broadcasting of np.linalg.solve
# ----------------------------------------------------------------------------------------------
def _solve_np1(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.ndim >= 3 and b.ndim == a.ndim - 1:
        if a.shape[0] == 0:
            return np.zeros(b.shape, dtype=float)
        return np.linalg.solve(a, b[..., None])[..., 0]
    return np.linalg.solve(a, b)


# ----------------------------------------------------------------------------------------------
# shims 2+3 as an AST pass
# ----------------------------------------------------------------------------------------------
class _Np2Rewriter(ast.NodeTransformer):
    def visit_Call(self, node: ast.Call):
        self.generic_visit(node)
        f = node.func
        if isinstance(f, ast.Attribute) and f.attr == "ptp":
            # np.ptp(...) itself stays; <expr>.ptp(...) becomes np.ptp(<expr>, ...)
            if not (isinstance(f.value, ast.Name) and f.value.id in ("np", "numpy")):
                new = ast.Call(
                    func=ast.Attribute(value=ast.Name(id="np", ctx=ast.Load()), attr="ptp", ctx=ast.Load()),
                    args=[f.value] + list(node.args),
                    keywords=list(node.keywords),
                )
                return ast.copy_location(new, node)
        if (
            isinstance(f, ast.Attribute)
            and f.attr == "solve"
            and isinstance(f.value, ast.Attribute)
            and f.value.attr == "linalg"
            and isinstance(f.value.value, ast.Name)
            and f.value.value.id in ("np", "numpy")
        ):
            new = ast.Call(func=ast.Name(id="__nk_solve_np1__", ctx=ast.Load()), args=list(node.args), keywords=list(node.keywords))
            return ast.copy_location(new, node)
        return node


class _RefLoader(importlib.abc.Loader):
    def __init__(self, path: str, is_pkg: bool):
        self.path = path
        self.is_pkg = is_pkg

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        module.__dict__["__nk_solve_np1__"] = _solve_np1
        if self.is_pkg and not os.path.isfile(self.path):
            return  # namespace-like package (reference has no __init__.py)
        with open(self.path, "r", encoding="utf-8") as fh:
            src = fh.read()
        tree = ast.parse(src, filename=self.path)
        tree = _Np2Rewriter().visit(tree)
        ast.fix_missing_locations(tree)
        code = compile(tree, self.path, "exec")
        exec(code, module.__dict__)


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top not in _REF_TOPLEVEL:
            return None
        rel = fullname.replace(".", os.sep)
        base = os.path.join(REFERENCE_ROOT, rel)
        if os.path.isdir(base):
            init = os.path.join(base, "__init__.py")
            spec = importlib.util.spec_from_loader(fullname, _RefLoader(init, True), is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        if os.path.isfile(base + ".py"):
            return importlib.util.spec_from_loader(fullname, _RefLoader(base + ".py", False))
        return None


# ----------------------------------------------------------------------------------------------
# shim 1: stubs for absent third-party packages
# ----------------------------------------------------------------------------------------------
class _StubModule(types.ModuleType):
    """Module whose every attribute is a MagicMock (cached), importable as a package."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # looks like a package so 'import a.b' works
        self._mocks = {}

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        m = self._mocks.get(item)
        if m is None:
            m = mock.MagicMock(name=f"{self.__name__}.{item}")
            self._mocks[item] = m
        return m


def _install_stubs():
    names = [
        "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.cm",
        "matplotlib.ticker", "matplotlib.animation", "matplotlib.gridspec", "matplotlib.lines",
        "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
        "trimesh", "shapely", "shapely.geometry", "h5py", "phonopy", "phonopy.interface",
        "phonopy.interface.calculator", "imageio",
    ]
    for n in names:
        if n in sys.modules and not isinstance(sys.modules[n], _StubModule):
            continue  # a real install wins
        try:
            if n not in sys.modules:
                importlib.import_module(n)
                continue
        except Exception:
            pass
        sys.modules[n] = _StubModule(n)
    for n in names:  # wire parents -> children
        if "." in n:
            parent, child = n.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _StubModule):
                sys.modules[parent].__dict__[child] = sys.modules[n]
    plt = sys.modules["matplotlib.pyplot"]
    if isinstance(plt, _StubModule):
        def _subplots(*a, **k):
            return mock.MagicMock(name="fig"), mock.MagicMock(name="ax")
        plt.__dict__["subplots"] = _subplots


_LOADED = None


def load_reference():
    """Import the reference classes.  Returns a namespace with Geometry, Mesh, Phonon, Population,
    Visualisation, Constants, SubvolClassifier and the argument parser module."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    _install_stubs()
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    ns = types.SimpleNamespace()
    ns.Constants = importlib.import_module("classes.Constants").Constants
    ns.Mesh = importlib.import_module("classes.Mesh").Mesh
    geo = importlib.import_module("classes.Geometry")
    ns.Geometry = geo.Geometry
    ns.SubvolClassifier = geo.SubvolClassifier
    ns.Phonon = importlib.import_module("classes.Phonon").Phonon
    ns.Visualisation = importlib.import_module("classes.Visualisation").Visualisation
    ns.Population = importlib.import_module("classes.Population").Population
    ns.argument_parser = importlib.import_module("argument_parser")
    # pure-plot methods are no-oped (they only draw); the numeric ones stay the reference's
    for name in ("flux_contribution", "plot_convergence_general", "convergence_energy_balance", "plot_kappa_path"):
        if hasattr(ns.Visualisation, name):
            setattr(ns.Visualisation, name, lambda self, *a, **k: None)
    ns.Population.plot_figures = lambda self, *a, **k: None   # would also consume np.random.rand(N)
    ns.Geometry.plot_mesh_bc = lambda self, *a, **k: None
    _LOADED = ns
    return ns


# ----------------------------------------------------------------------------------------------
# argument namespaces: parse a parameters file exactly as the reference CLI does
# ----------------------------------------------------------------------------------------------
def parse_args(tokens, results_folder):
    """``argument_parser.initialise_parser(False).parse_args(tokens)`` with the results folder fixed
    (reference: argument_parser.py:110-140; generate_results_folder is bypassed on purpose so that no
    ``_N`` directories are scattered around)."""
    ref = load_reference()
    parser = ref.argument_parser.initialise_parser(False)
    args = parser.parse_args(list(tokens))
    os.makedirs(results_folder, exist_ok=True)
    args.results_folder = results_folder
    return args


def parse_parameters_text(text, results_folder, overrides=None):
    tokens = text.split()
    args = parse_args(tokens, results_folder)
    for k, v in (overrides or {}).items():
        setattr(args, k, v)
    return args


# ----------------------------------------------------------------------------------------------
# Phonon without __init__ (h5py / phonopy / the hdf5 blobs are absent): fill the raw tables from a
# synthetic full-BZ table and let the reference's own methods derive everything else.
# ----------------------------------------------------------------------------------------------
def make_phonon(args, table):
    """table: dict with omega (Q,J) [rad THz], group_vel (Q,J,3) [A THz], gamma (NT,Q,J) [THz],
    temperature_array (NT,), q_points (Q,3) reduced, lattice (3,3) rows = cell vectors [A],
    data_mesh (3,).  Follows Phonon.load_base_properties (Phonon.py:66-149) from line 102 on."""
    ref = load_reference()
    ph = object.__new__(ref.Phonon)
    ref.Constants.__init__(ph)
    ph.args = args
    ph.mat_index = 0
    ph.mat_folder = ""
    lattice = np.asarray(table["lattice"], dtype=float)
    reciprocal_lattice = np.linalg.inv(lattice) * 2 * np.pi            # Phonon.py:72
    ph.volume_unitcell = float(abs(np.linalg.det(lattice)))            # Phonon.py:83
    ph.data_mesh = np.asarray(table["data_mesh"])
    ph.frequency = np.asarray(table["omega"], dtype=float) / (2 * ph.pi)
    ph.omega = np.asarray(table["omega"], dtype=float).copy()
    ph.group_vel = np.around(np.asarray(table["group_vel"], dtype=float), decimals=10)  # :102
    ph.temperature_array = np.asarray(table["temperature_array"], dtype=float)
    gamma = np.asarray(table["gamma"], dtype=float)
    ph.gamma = np.where(gamma > 0, gamma, -1)                           # :324
    ph.q_points = np.asarray(table["q_points"], dtype=float).copy()
    ph.weights = np.ones(ph.q_points.shape[0])
    ph.number_of_qpoints = ph.q_points.shape[0]
    ph.number_of_branches = ph.omega.shape[1]
    ph.number_of_modes = ph.number_of_qpoints * ph.number_of_branches
    ph.inactive_modes_mask = np.all(ph.group_vel == 0, axis=2)
    ph.number_of_inactive_modes = ph.inactive_modes_mask.sum()
    ph.number_of_active_modes = ph.number_of_modes - ph.number_of_inactive_modes
    ph.reciprocal_lattice = np.around(reciprocal_lattice, decimals=6)  # :129
    ph.unique_modes = np.stack(np.meshgrid(np.arange(ph.number_of_qpoints), np.arange(ph.number_of_branches)), axis=-1).reshape(-1, 2).astype(int)
    ph.get_wavevectors()
    ph.get_norms()
    ph.calculate_lifetime()
    ph.zero_point = ph.calculate_zeropoint()
    ph.initialise_temperature_function()
    ph.initialise_density_of_states()
    return ph


def make_geometry(args):
    ref = load_reference()
    return ref.Geometry(args)


def make_population(args, geo, ph, seed=None):
    ref = load_reference()
    if seed is not None:
        np.random.seed(seed)
    return ref.Population(args, geo, ph)
