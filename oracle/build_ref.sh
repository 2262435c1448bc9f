#!/usr/bin/env bash
# Builds the UNMODIFIED ls1-MarDyn reference (sources where they lie under $LS1_REF, default
# /root/reference) into oracle/_ref/<precision>/ with plain g++ -- no cmake, no copies of sources.
# Test infrastructure only: the product (ls1_mardyn_b200) never links or runs anything built here.
#
#   oracle/_ref/mixed/MarDyn    PRECISION=MIXED  (-DMARDYN_SPDP), AVX2+FMA, OpenMP
#   oracle/_ref/double/MarDyn   PRECISION=DOUBLE (-DMARDYN_DPDP), AVX2+FMA, OpenMP
#   oracle/_ref/mixed512/MarDyn PRECISION=MIXED with AVX-512 (only on request: build_ref.sh mixed512; bench.py's CPU legs prefer it on AVX-512 hosts)
#   oracle/_ref/rmm/MarDyn      PRECISION=SINGLE + REDUCED_MEMORY_MODE (only on request: build_ref.sh rmm). STATUS: the unmodified
#                               reference does NOT compile in this configuration with g++ 13.3 (-std=c++17 or c++23):
#                               particleContainer/ParticleIterator.h:78 and RegionParticleIterator.h:75 default a constexpr copy
#                               constructor whose implicit version is not constexpr once Molecule = MoleculeRMM. The rmm build
#                               therefore compiles from a shadow tree of symbolic links under _ref/rmm/shadow with that keyword
#                               removed from those two declarations (generated at build time, never copied into the repo).
#   oracle/_ref/<prec>/refdump  our dump driver (oracle/refdump.cpp) linked against the same objects
#
# The source list mirrors what the reference's CMakeLists select for ENABLE_MPI=OFF, ENABLE_ADIOS2=OFF,
# ENABLE_VTK=OFF, ENABLE_AUTOPAS=OFF, ENABLE_UNIT_TESTS=OFF, OPENMP=ON, VECTOR_INSTRUCTIONS=AVX2
# (src/CMakeLists.txt and src/*/CMakeLists.txt; flags from cmake/modules/{options,vectorization,OpenMP}.cmake).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LS1_REF:-/root/reference}"
PRECS="${1:-mixed double}"
JOBS="${JOBS:-$(nproc)}"
CXX="${REF_CXX:-/usr/bin/g++}"   # NB: /opt/gcc/bin/g++ (env CXX in this image) lacks libgomp.spec
[ -d "$REF/src" ] || { echo "build_ref: $REF/src not found (reference absent) -- nothing to do"; exit 0; }

MPI_ONLY="CommunicationBuffer CommunicationPartner DomainDecompMPIBase DomainDecomposition GeneralDomainDecomposition \
KDDecomposition KDDStaticValues KDNode MPIKDNode NeighborAcquirer NeighbourCommunicationScheme NonBlockingMPIHandlerBase \
NonBlockingMPIMultiStepHandler ParticleDataFull ParticleDataRMM ParticleForceData ResilienceComm \
StaticIrregDomainDecomposition ALLLoadBalancer"

list_sources() {
  ( cd "$REF/src" && find . -name '*.cpp' \
      ! -path '*/tests/*' ! -path './io/vtk/*' ! -path './bhfmm/fft/*' \
      ! -path '*/AutoPasTemplateInstantiations/*' ! -name 'AutoPas*.cpp' ! -name 'Adios2*.cpp' \
      | sed 's|^\./||' | sort ) | while read -r f; do
        b="$(basename "$f" .cpp)"; skip=0
        case "$f" in parallel/*) for m in $MPI_ONLY; do [ "$b" = "$m" ] && skip=1; done;; esac
        [ $skip = 0 ] && echo "$f"
      done
}

for prec in $PRECS; do
  VEC="-mavx2 -mfma"
  case "$prec" in
    mixed)  PDEF=-DMARDYN_SPDP;;
    double) PDEF=-DMARDYN_DPDP;;
    mixed512) PDEF=-DMARDYN_SPDP; VEC="-mavx512f -mavx512cd -mavx512vl -mavx512bw -mavx512dq -mavx2 -mfma";;      # VCP_VEC_AVX512F (SIMD_TYPES.h:70-75): the CPU baseline on AVX-512 hosts
    rmm)    PDEF="-DMARDYN_SPSP -DENABLE_REDUCED_MEMORY_MODE=1";;      # the reference's reduced-memory mode: all FP32, 1CLJ only (VCP1CLJRMM)
    *) echo "unknown precision $prec"; exit 2;;
  esac
  OUT="$HERE/_ref/$prec"; OBJ="$OUT/obj"; mkdir -p "$OBJ" "$OUT/gen"
  # generated header (cmake's configure_file of src/MarDyn_version.h.in): written here, not copied
  cat > "$OUT/gen/MarDyn_version.h" <<H
#pragma once
#include <string>
const std::string MARDYN_VERSION = "1.2.0_oracle_ref_$prec";
H
  SRC="$REF/src"
  if [ "$prec" = rmm ]; then
    # g++ 13.3 rejects two `constexpr ... = default` copy constructors once Molecule = MoleculeRMM (see the header of this file).
    # The sources are compiled from a shadow tree under _ref/ (git-ignored, generated here): symbolic links to every reference
    # file, except the two headers, which are written by sed with that one keyword removed. Nothing is copied into the repo.
    SRC="$OUT/shadow"; rm -rf "$SRC"; mkdir -p "$SRC"
    ( cd "$REF/src" && find . -type d ) | while read -r d; do mkdir -p "$SRC/$d"; done
    ( cd "$REF/src" && find . -type f ) | while read -r f; do ln -s "$REF/src/$f" "$SRC/$f"; done
    for h in particleContainer/ParticleIterator.h particleContainer/RegionParticleIterator.h; do
      rm "$SRC/$h"; sed 's/^\(\s*\)constexpr \(\(Region\)\?ParticleIterator\s*(const\)/\1\2/' "$REF/src/$h" > "$SRC/$h"
    done
  fi
  FLAGS="-std=c++17 -O3 -DNDEBUG -fopenmp -DENABLE_OPENMP -fopenmp-simd $VEC -DENABLE_VECTORIZED_CODE=1 \
$PDEF -DTIMERS -w -I$OUT/gen -I$SRC -I$REF/libs/rapidxml -I$REF/libs"
  MK="$OUT/Makefile.gen"
  { echo "CXX=$CXX"; echo "FLAGS=$FLAGS"; echo "OBJS="; } > "$MK"
  for f in $(list_sources); do
    o="$OBJ/$(echo "$f" | tr '/' '_' | sed 's/\.cpp$/.o/')"
    echo "OBJS+=$o" >> "$MK"
    printf '%s: %s\n\t@$(CXX) $(FLAGS) -c $< -o $@\n' "$o" "$SRC/$f" >> "$MK"
  done
  cat >> "$MK" <<M
LIBOBJS=\$(filter-out %/MarDyn.o,\$(OBJS))
all: $OUT/MarDyn $OUT/refdump
$OUT/MarDyn: \$(OBJS)
	@\$(CXX) -fopenmp \$(OBJS) -o \$@
$OUT/refdump.o: $HERE/refdump.cpp
	@\$(CXX) \$(FLAGS) -c \$< -o \$@
$OUT/refdump: \$(LIBOBJS) $OUT/refdump.o
	@\$(CXX) -fopenmp \$(LIBOBJS) $OUT/refdump.o -o \$@
M
  if [ -f "$HERE/refdump.cpp" ] && [ "$prec" != rmm ] && [ "$prec" != mixed512 ]; then make -f "$MK" -j"$JOBS" all; else make -f "$MK" -j"$JOBS" "$OUT/MarDyn"; fi      # refdump reads F / M / Vi, which MoleculeRMM does not store
  echo "build_ref: built $OUT/MarDyn"
done
