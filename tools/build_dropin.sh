#!/bin/bash
# Builds the reference's own example programs against the B200 estimator: an overlay of symlinks to the read-only
# reference tree is created under build/overlay, with include/cauchy_estimator.hpp replaced by this repository's
# drop-in header; the reference's src/*.cpp are then compiled UNCHANGED from the overlay and linked with
# libmce_b200.so.  Needs /root/reference (build time only); the binaries land in build/dropin/.
set -e
REF=${REF:-/root/reference}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OV=$ROOT/build/overlay
rm -rf "$OV" && mkdir -p "$OV/include" "$OV/src" "$ROOT/build/dropin"
for f in "$REF"/include/*; do ln -s "$f" "$OV/include/"; done
for f in "$REF"/src/*; do ln -s "$f" "$OV/src/"; done
rm "$OV/include/cauchy_estimator.hpp"
ln -s "$ROOT/include/cauchy_estimator.hpp" "$OV/include/cauchy_estimator.hpp"
for t in cauchy_estimator leo_satellite_7state_gps window_manager; do
  g++ -O3 -w -ffp-contract=off -I"$ROOT/include" "$OV/src/$t.cpp" -o "$ROOT/build/dropin/$t" \
      -L"$ROOT/cauchyfriendly_b200" -lmce_b200 -Wl,-rpath,'$ORIGIN/../../cauchyfriendly_b200' -lm -lpthread
  echo "built build/dropin/$t"
done
# the device cpdf dispatcher checked against the reference's own CPU cpdf code over the host mirror (tests/dropin/cpdf1d_dropin.cpp)
mkdir -p "$OV/tests"
ln -sf "$ROOT/tests/dropin/cpdf1d_dropin.cpp" "$OV/tests/cpdf1d_dropin.cpp"
g++ -O3 -w -ffp-contract=off -I"$OV/include" -I"$ROOT/include" "$OV/tests/cpdf1d_dropin.cpp" -o "$ROOT/build/dropin/cpdf1d_dropin" \
    -L"$ROOT/cauchyfriendly_b200" -lmce_b200 -Wl,-rpath,'$ORIGIN/../../cauchyfriendly_b200' -lm -lpthread
echo "built build/dropin/cpdf1d_dropin"
