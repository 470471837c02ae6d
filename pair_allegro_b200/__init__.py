"""pair_allegro_b200: B200-native Allegro force evaluation behind the LAMMPS `pair_style allegro`
interface.  The product is csrc/ (CUDA kernels + C-ABI, include/allegro_b200.h) and the host-side
mirror of the reference pair style; `export` is the offline weight exporter."""
__version__ = "0.1.0"
