#!/bin/sh
# Builds tools/dropin_bench against the drop-in headers and the in-tree C-ABI library (plain g++: a user of the
# reference's headers does not need nvcc).  The binary travels to the GPU box with the snapshot.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
PKG="$HERE/../alllsatisfiabilitysolver_b200"
CXX=g++; [ -x /usr/bin/g++ ] && CXX=/usr/bin/g++
$CXX -std=c++20 -O2 -w -I"$PKG/include" "$HERE/dropin_bench.cpp" -o "$HERE/dropin_bench" -L"$PKG" -lalll_b200 -Wl,-rpath,"$PKG" -pthread
echo built "$HERE/dropin_bench"
