#!/bin/bash
# Stage 0 of the repo's own build (reference: scripts/build/xmake.py compiles build/xmake.cc, lines 53-61):
# compiles tools/xmake/xmake.cc -> tools/xmake/.out/xmake. Stage 1 is `xmake <target>`.
set -e
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$here/.out"
${CXX:-g++} -O3 -std=c++20 -Wall "$here/xmake.cc" -o "$here/.out/xmake"
echo "$here/.out/xmake"
