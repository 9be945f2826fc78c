#!/bin/bash
# tools/mk.sh <source-stem> <output-name> [extra nvcc flags...]   -- build a developer probe into tools/_build/
set -e -o pipefail
cd "$(dirname "$0")/.."
mkdir -p tools/_build
src=$1; out=$2; shift 2
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Imobilesuperresolution_b200/csrc "$@" tools/$src.cu -o tools/_build/$out -lcuda > tools/_build/$out.log 2>&1 || { cat tools/_build/$out.log; exit 1; }
grep -v "warning\|^$\|\^\|declared but never\|Remark\|was set but" tools/_build/$out.log || true
