#!/bin/bash
# builds tests/cpp/test_histogram against the drop-in class, the CUDA library and the C oracle (test infrastructure)
set -e
here=$(cd "$(dirname "$0")" && pwd); root=$(cd "$here/../.." && pwd)
make -s -C "$root/oracle" libshf_oracle.so
g++ -std=c++17 -O1 -Wall -I "$root/include" -I "$root/include/compat" \
    "$here/test_histogram.cpp" "$root/superterrainplus_b200/host/STPSingleHistogramFilter.cpp" \
    -L "$root/superterrainplus_b200" -lshf_b200 -L "$root/oracle" -lshf_oracle \
    -Wl,-rpath,"$root/superterrainplus_b200" -Wl,-rpath,"$root/oracle" -o "$here/test_histogram"
echo "$here/test_histogram"
