#!/bin/bash
# scan-kernel variant experiments: tools/build_scan_variant.sh NAME "<extra nvcc flags>"  ->  build/libgpud_NAME.so  (use with GPUD_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
NAME=$1; shift
S=gpud_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $@ -c $S/kmsg_scan.cu -o build/scan_$NAME.o 2> build/scan_$NAME.log
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libgpud_$NAME.so build/scan_$NAME.o $S/api.o $S/ring.o $S/select.o $S/ib_scan.o $S/fabric.o $S/catalog.o $S/host_component.o $S/component_abi.o $S/kmsg_stateful.o $S/poller.o $S/store_sqlite.o -lcudart -ldl
grep -A1 "k_scan_match" build/scan_$NAME.log | grep -E "Used|spill" | head -2
