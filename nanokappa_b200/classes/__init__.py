"""Reference-shaped host classes (Geometry, Mesh, Phonon, Population, Visualisation, Constants).

Same class / method / attribute names as ``/root/reference/classes`` so that ``nanokappa.py`` and
user scripts written against Nano-kappa keep working; set-up runs on the host in NumPy, the
per-timestep path runs on the GPU through ``nanokappa_b200.engine``."""
