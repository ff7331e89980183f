#!/bin/sh
# Builds the CPU oracle (test infrastructure) -> oracle/_build/libsynseg_oracle.so
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_build"
gcc -O2 -fPIC -shared -fvisibility=hidden -o "$here/_build/libsynseg_oracle.so" "$here/synseg_oracle.c" -lm
echo "built $here/_build/libsynseg_oracle.so"
