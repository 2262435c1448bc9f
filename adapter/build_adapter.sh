#!/usr/bin/env bash
# Builds ls1-MarDyn WITH the B200 drop-in: the unmodified reference objects of oracle/_ref/<prec>/obj (oracle/build_ref.sh)
# + the two adapter classes of this directory + a build-time patched copy of Simulation.cpp that knows two more XML values:
#     <datastructure type="LinkedCellsB200">   (next to Simulation.cpp:476-500)
#     <integrator type="LeapfrogB200">         (next to Simulation.cpp:176-189)
# Output: adapter/_build/<prec>/MarDyn_b200 (git-ignored), linked against ls1_mardyn_b200/libls1b200.so.
# Nothing of the reference is copied into the repository: the patched translation unit lives only under adapter/_build.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$HERE")"
REF="${LS1_REF:-/root/reference}"
PRECS="${1:-mixed double}"
CXX="${REF_CXX:-/usr/bin/g++}"
[ -d "$REF/src" ] || { echo "build_adapter: $REF/src not found -- nothing to do"; exit 0; }
for prec in $PRECS; do
  case "$prec" in mixed) PDEF=-DMARDYN_SPDP;; double) PDEF=-DMARDYN_DPDP;; esac
  OBJ="$ROOT/oracle/_ref/$prec/obj"; GEN="$ROOT/oracle/_ref/$prec/gen"
  [ -d "$OBJ" ] || "$ROOT/oracle/build_ref.sh" "$prec"
  OUT="$HERE/_build/$prec"; mkdir -p "$OUT/inc/particleContainer" "$OUT/inc/integrators"
  # the adapter headers are included as ls1 would include them once the files are dropped into src/
  cp "$HERE/LinkedCellsB200.h" "$OUT/inc/particleContainer/"; cp "$HERE/LeapfrogB200.h" "$OUT/inc/integrators/"
  python3 - "$REF/src/Simulation.cpp" "$OUT/Simulation_b200.cpp" <<'PY'
import sys
s = open(sys.argv[1]).read()
inc = '#include "particleContainer/LinkedCells.h"'
assert inc in s
s = s.replace(inc, inc + '\n#include "particleContainer/LinkedCellsB200.h"\n#include "integrators/LeapfrogB200.h"', 1)
a = 'if(datastructuretype == "LinkedCells") {'
assert s.count(a) == 1
s = s.replace(a, 'if(datastructuretype == "LinkedCellsB200") {\n\t\t\t\t_moleculeContainer = new LinkedCellsB200();\n\t\t\t\t_moleculeContainer->setCutoff(_cutoffRadius);\n\t\t\t} else ' + a)
b = 'if(integratorType == "Leapfrog") {'
assert s.count(b) == 1
s = s.replace(b, 'if(integratorType == "LeapfrogB200") {\n\t\t\t_integrator = new LeapfrogB200();\n\t\t} else ' + b)
open(sys.argv[2], "w").write(s)
PY
  FLAGS="-std=c++17 -O3 -DNDEBUG -fopenmp -DENABLE_OPENMP -fopenmp-simd -mavx2 -mfma -DENABLE_VECTORIZED_CODE=1 $PDEF -DTIMERS -w \
-I$OUT/inc -I$GEN -I$REF/src -I$REF/libs/rapidxml -I$REF/libs -I$ROOT/include"
  pids=()
  $CXX $FLAGS -c "$OUT/Simulation_b200.cpp" -o "$OUT/Simulation_b200.o" & pids+=($!)
  $CXX $FLAGS -c "$HERE/LinkedCellsB200.cpp" -o "$OUT/LinkedCellsB200.o" & pids+=($!)
  $CXX $FLAGS -c "$HERE/LeapfrogB200.cpp" -o "$OUT/LeapfrogB200.o" & pids+=($!)
  for p in "${pids[@]}"; do wait "$p"; done      # a failed compile must fail the build (plain `wait` returns 0)
  OBJS=$(ls "$OBJ"/*.o | grep -v '/Simulation.o$')
  $CXX -fopenmp $OBJS "$OUT/Simulation_b200.o" "$OUT/LinkedCellsB200.o" "$OUT/LeapfrogB200.o" \
      -L"$ROOT/ls1_mardyn_b200" -lls1b200 -Wl,-rpath,'$ORIGIN/../../../ls1_mardyn_b200' -Wl,-rpath,/usr/local/cuda/lib64 -o "$OUT/MarDyn_b200"
  echo "build_adapter: built $OUT/MarDyn_b200"
done
