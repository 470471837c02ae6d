#!/bin/bash
# install_into_lammps.sh [-l] /path/to/lammps
#
# Puts the B200 pair style and compute of this repo into a LAMMPS source tree (the role the reference's
# patch_lammps.sh plays for its libtorch build): pair_allegro_b200.{h,cpp}, compute_allegro_b200.{h,cpp} and
# the C-ABI header go to <lammps>/src, and <lammps>/cmake/CMakeLists.txt is told to link
# pair_allegro_b200/liballegro_b200.so (built beforehand with `python -c "import __graft_entry__ as g; g.build()"`).
# No libtorch, no C++ standard bump.  -l symlinks instead of copying.
#
# NOT exercised in this repository's CI (no LAMMPS tree in the image): the sources themselves are compiled
# and tested against lmpshim/ (make -C src); this script only automates the copy + two CMake lines that
# INTEGRATION.md section 1 describes.
set -euo pipefail
link=false
while getopts "hl" opt; do
  case $opt in
    l) link=true ;;
    h) sed -n 2,13p "$0"; exit 0 ;;
    *) exit 2 ;;
  esac
done
shift $((OPTIND - 1))
lammps_dir=${1:-}
here=$(cd "$(dirname "$0")/.." && pwd)
[ -n "$lammps_dir" ] || { echo "usage: $0 [-l] /path/to/lammps" >&2; exit 1; }
[ -d "$lammps_dir/cmake" ] && [ -d "$lammps_dir/src" ] || { echo "$lammps_dir does not look like a LAMMPS source directory" >&2; exit 1; }
lib="$here/pair_allegro_b200/liballegro_b200.so"
[ -f "$lib" ] || { echo "build $lib first (python -c 'import __graft_entry__ as g; g.build()')" >&2; exit 1; }
if grep -q "liballegro_b200" "$lammps_dir/cmake/CMakeLists.txt"; then
  echo "this LAMMPS tree already references liballegro_b200 -- not patching CMakeLists.txt again" >&2
  patched=true
else
  patched=false
fi
for f in "$here"/src/pair_allegro_b200.h "$here"/src/pair_allegro_b200.cpp "$here"/src/compute_allegro_b200.h \
         "$here"/src/compute_allegro_b200.cpp "$here"/include/allegro_b200.h; do
  if $link; then ln -sf "$f" "$lammps_dir/src/$(basename "$f")"; else cp "$f" "$lammps_dir/src/$(basename "$f")"; fi
done
# pair_style allegro/kk goes where the reference's patch_lammps.sh puts its Kokkos twin: <lammps>/src/KOKKOS (patch_lammps.sh:61-66)
if [ -d "$lammps_dir/src/KOKKOS" ]; then
  for f in "$here"/src/pair_allegro_b200_kokkos.h "$here"/src/pair_allegro_b200_kokkos.cpp; do
    if $link; then ln -sf "$f" "$lammps_dir/src/KOKKOS/$(basename "$f")"; else cp "$f" "$lammps_dir/src/KOKKOS/$(basename "$f")"; fi
  done
fi
if ! $patched; then
  cat >> "$lammps_dir/cmake/CMakeLists.txt" <<EOF2

# --- pair_style allegro (B200 C-ABI build) ---
message(STATUS "<< allegro_b200: linking $lib >>")
target_link_libraries(lammps PUBLIC "$lib")
set_property(TARGET lammps APPEND PROPERTY BUILD_RPATH "$(dirname "$lib")")
EOF2
fi
echo "done: configure LAMMPS as usual (newton on; the pair style needs no package flags)"
