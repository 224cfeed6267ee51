#!/usr/bin/env bash
# TEST INFRASTRUCTURE: host build of the engine's phase functions (tests/emu/rach_emu.cpp).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$here/../.."
mkdir -p "$here/_build"
g++ -O2 -Wall -Wno-unused-function -fPIC -shared -std=c++17 -I"$root/include" -I"$root/oracle" \
    -I"$root/5g-nr-randomaccess_b200/csrc" "$here/rach_emu.cpp" \
    "$root/5g-nr-randomaccess_b200/csrc/rach_host.cpp" -o "$here/_build/librach_emu.so"
echo "$here/_build/librach_emu.so"
