#!/bin/sh
# Builds tests/emu/_build/libmce_emu.so (CPU emulation of the kernel bodies, test-only). No FMA contraction.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/_build"
g++ -O2 -g -std=c++17 -ffp-contract=off -fPIC -shared -Wall -Wno-unused-function -Wno-unknown-pragmas \
    -I"$HERE/../../cauchyfriendly_b200/csrc" "$HERE/emu_capi.cpp" -o "$HERE/_build/libmce_emu.so"
