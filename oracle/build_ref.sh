#!/bin/bash
# Builds oracle/_ref/libdfki_ref_h<N>.so: the REFERENCE'S OWN sources on the MPC path, compiled where they lie under
# /root/reference (nothing is copied into the repo), behind the C entry points of oracle/ref_harness.cpp.
# TEST INFRASTRUCTURE ONLY - the product never loads these libraries.
#
# The reference's horizon is a compile-time constant (MPC_PREDICTION_HORIZON = 20, mit_controller_params.hpp:6; the
# reference user edits that header).  One library is built per horizon in $HORIZONS: a scratch include tree of symlinks
# to the reference's headers is made under $TMP, in which mit_controller_params.hpp alone is a generated copy with that
# one constant changed.  Everything else is compiled from the original files.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${DFKI_REFERENCE:-/root/reference}"
SRC="$REF/ws/src"
OUT="$HERE/_ref"
HORIZONS="${HORIZONS:-10 16 20}"
CXX="${REF_CXX:-$(test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)}"
CC="${REF_CC:-$(test -x /usr/bin/gcc && echo /usr/bin/gcc || echo gcc)}"
if [ ! -d "$SRC/controllers/src/mit_controller" ]; then
  echo "build_ref.sh: $REF not present - keeping whatever is already in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/dfkiref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
# strict IEEE evaluation, like the reference's aarch64 build (-O3, controllers/CMakeLists.txt:17); its x86 build uses
# -Ofast (:19), which is not a defined arithmetic.  x86-64-v3 so the library also runs on the GPU box's host.
# Hidden visibility + -Bsymbolic + no STB_GNU_UNIQUE: the per-horizon libraries define the same C++ symbols with different
# layouts (arrays sized by the horizon) and must not interpose each other when loaded into one process.
CXXFLAGS="-std=c++17 -O3 -ffp-contract=off -fno-fast-math -march=x86-64-v3 -fPIC -w -fvisibility=hidden -fvisibility-inlines-hidden -fno-gnu-unique"
$CC -O3 -march=x86-64-v3 -fopenmp -fPIC -std=gnu11 -fvisibility=hidden -c "$HERE/mpc_ref.c" -o "$TMP/mpc_ref.o"
for H in $HORIZONS; do
  T="$TMP/h$H"
  mkdir -p "$T/include/mit_controller" "$T/include/potato_sim" "$T/include/common"
  for f in "$SRC"/controllers/include/mit_controller/*.hpp; do ln -s "$f" "$T/include/mit_controller/$(basename "$f")"; done
  for f in "$SRC"/controllers/include/potato_sim/*.hpp; do ln -s "$f" "$T/include/potato_sim/$(basename "$f")"; done
  for f in "$SRC"/common/include/common/*.hpp; do ln -s "$f" "$T/include/common/$(basename "$f")"; done
  rm "$T/include/mit_controller/mit_controller_params.hpp"
  sed -E "s/(MPC_PREDICTION_HORIZON[[:space:]]*=[[:space:]]*)[0-9]+;/\1$H;/" \
      "$SRC/controllers/include/mit_controller/mit_controller_params.hpp" > "$T/include/mit_controller/mit_controller_params.hpp"
  grep -q "MPC_PREDICTION_HORIZON = $H;" "$T/include/mit_controller/mit_controller_params.hpp"
  INC="-I$HERE/ref_shim -I$T/include -I$T/include/mit_controller"
  OBJS=""
  for f in gait simple_gait_sequencer adaptive_gait_sequencer mpc_trajectory_planner raibert_foot_step_planner mpc; do
    $CXX $CXXFLAGS $INC -c "$SRC/controllers/src/mit_controller/$f.cpp" -o "$T/$f.o"
    OBJS="$OBJS $T/$f.o"
  done
  $CXX $CXXFLAGS $INC -c "$HERE/ref_harness.cpp" -o "$T/ref_harness.o"
  $CXX -shared -fopenmp -Wl,-Bsymbolic -o "$OUT/libdfki_ref_h$H.so" $OBJS "$T/ref_harness.o" "$TMP/mpc_ref.o" -lm
  echo "built $OUT/libdfki_ref_h$H.so"
done
