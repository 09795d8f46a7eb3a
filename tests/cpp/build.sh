#!/bin/bash
# builds tests/cpp/test_surface, test_histogram, test_batcher and test_biome against the drop-in classes, the CUDA library and the C oracle
# (test infrastructure)
set -e
here=$(cd "$(dirname "$0")" && pwd); root=$(cd "$here/../.." && pwd)
make -s -C "$root/oracle" libshf_oracle.so libbiome_oracle.so
g++ -std=c++17 -O1 -Wall -pthread -I "$root/include" -I "$root/include/compat" \
    "$here/test_surface.cpp" "$root/superterrainplus_b200/host/STPSingleHistogramFilter.cpp" \
    -L "$root/superterrainplus_b200" -lshf_b200 -Wl,-rpath,"$root/superterrainplus_b200" -o "$here/test_surface"
for t in test_histogram test_batcher; do
g++ -std=c++17 -O1 -Wall -pthread -I "$root/include" -I "$root/include/compat" \
    "$here/$t.cpp" "$root/superterrainplus_b200/host/STPSingleHistogramFilter.cpp" \
    "$root/superterrainplus_b200/host/STPSingleHistogramBatcher.cpp" \
    -L "$root/superterrainplus_b200" -lshf_b200 -L "$root/oracle" -lshf_oracle \
    -Wl,-rpath,"$root/superterrainplus_b200" -Wl,-rpath,"$root/oracle" -o "$here/$t"
done
g++ -std=c++17 -O1 -Wall -pthread -I "$root/include" -I "$root/include/compat" -I /usr/local/cuda/include \
    "$here/test_biome.cpp" "$root/superterrainplus_b200/host/STPSingleHistogramFilter.cpp" \
    "$root/superterrainplus_b200/host/STPBiomeFactoryDevice.cpp" \
    -L "$root/superterrainplus_b200" -lshf_b200 -L "$root/oracle" -lshf_oracle -lbiome_oracle -L /usr/local/cuda/lib64 -lcudart \
    -Wl,-rpath,"$root/superterrainplus_b200" -Wl,-rpath,"$root/oracle" -o "$here/test_biome"
echo "$here/test_surface $here/test_histogram $here/test_batcher $here/test_biome"
