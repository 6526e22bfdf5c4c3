#!/bin/bash
# Builds the reference's own example programs against the B200 estimator: an overlay of symlinks to the read-only
# reference tree is created under build/overlay, with include/cauchy_estimator.hpp replaced by this repository's
# drop-in header; the reference's src/*.cpp are then compiled UNCHANGED from the overlay and linked with
# libmce_b200.so.  Needs /root/reference (build time only); the binaries land in build/dropin/.
set -e
REF=${REF:-/root/reference}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OV=$ROOT/build/overlay
rm -rf "$OV" && mkdir -p "$OV/include" "$OV/src" "$OV/scripts/swig/cauchy" "$OV/tests" "$ROOT/build/dropin"
for f in "$REF"/include/*; do      # sub-directories (models/) are re-created so that their "../x.hpp" includes stay inside the overlay
  if [ -d "$f" ]; then mkdir -p "$OV/include/$(basename "$f")"; for g in "$f"/*; do ln -s "$g" "$OV/include/$(basename "$f")/"; done
  else ln -s "$f" "$OV/include/"; fi
done
for f in "$REF"/src/*; do ln -s "$f" "$OV/src/"; done
ln -s "$REF/scripts/swig/cauchy/pycauchy.hpp" "$OV/scripts/swig/cauchy/pycauchy.hpp"
for f in "$ROOT"/tests/dropin/*.cpp; do ln -s "$f" "$OV/tests/"; done
rm "$OV/include/cauchy_estimator.hpp"
ln -s "$ROOT/include/cauchy_estimator.hpp" "$OV/include/cauchy_estimator.hpp"
for t in cauchy_estimator leo_satellite_7state_gps window_manager homing_missile leo_satellite_5state; do
  g++ -O3 -w -ffp-contract=off -I"$ROOT/include" "$OV/src/$t.cpp" -o "$ROOT/build/dropin/$t" \
      -L"$ROOT/cauchyfriendly_b200" -lmce_b200 -Wl,-rpath,'$ORIGIN/../../cauchyfriendly_b200' -lm -lpthread
  echo "built build/dropin/$t"
done
# the device cpdf dispatcher checked against the reference's own CPU cpdf code over the host mirror (tests/dropin/cpdf1d_dropin.cpp)
# ... the Swig shim driven from C++ (tests/dropin/pycauchy_dropin.cpp), the window bank with logging (winbank_dropin.cpp) and the
# relative-system readers of cauchy_prediction.hpp over two estimators (rsys_dropin.cpp)
for t in cpdf1d_dropin pycauchy_dropin winbank_dropin; do
  g++ -O3 -w -ffp-contract=off -I"$OV/include" -I"$ROOT/include" "$OV/tests/$t.cpp" -o "$ROOT/build/dropin/$t" \
      -L"$ROOT/cauchyfriendly_b200" -lmce_b200 -Wl,-rpath,'$ORIGIN/../../cauchyfriendly_b200' -lm -lpthread
  echo "built build/dropin/$t"
done
# Two estimators in one program (rsys_dropin.cpp): the constructor draws root_point / b_pert with libc rand() once per helper thread (est:125-150), so the
# second estimator's draws depend on NUM_CPUS.  The golden comes from the NUM_CPUS = 1 reference; this program is therefore built from a second overlay whose
# cauchy_constants.hpp says NUM_CPUS = 1 (generated exactly like oracle/Makefile does; everything else is the overlay above).
OV1=$ROOT/build/overlay_cpu1
rm -rf "$OV1" && cp -a "$OV" "$OV1"
rm "$OV1/include/cauchy_constants.hpp"
sed 's/^const int NUM_CPUS = [0-9]*;/const int NUM_CPUS = 1;/' "$REF/include/cauchy_constants.hpp" > "$OV1/include/cauchy_constants.hpp"
grep -q 'const int NUM_CPUS = 1;' "$OV1/include/cauchy_constants.hpp"
for t in rsys_dropin; do
  g++ -O3 -w -ffp-contract=off -I"$OV1/include" -I"$ROOT/include" "$OV1/tests/$t.cpp" -o "$ROOT/build/dropin/$t" \
      -L"$ROOT/cauchyfriendly_b200" -lmce_b200 -Wl,-rpath,'$ORIGIN/../../cauchyfriendly_b200' -lm -lpthread
  echo "built build/dropin/$t"
done
gcc -O2 -shared -fPIC -o "$ROOT/build/dropin/fixed_time.so" "$ROOT/tests/dropin/fixed_time.c"
