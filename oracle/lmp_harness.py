"""Compatibility alias: the LAMMPS-side harness (atoms, ghosts, full neighbour lists, brick
decomposition) lives in lmpshim/harness.py -- it emulates the CALLER of the pair style, not the
path's arithmetic, and is shared by the tests, the reference shim driver and bench.py."""
from lmpshim.harness import *  # noqa: F401,F403
from lmpshim.harness import _image_shifts  # noqa: F401
