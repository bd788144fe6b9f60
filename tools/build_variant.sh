#!/bin/bash
# Builds a variant of libb200det.so with extra -D flags for conv_tc.cu (experiments): tools/build_variant.sh NAME -DFOO=1 ...
set -e
cd "$(dirname "$0")/../object_detection_tracking_b200"
name=$1; shift
python build.py >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -O2 \
  -Xcompiler -Wno-attributes -Xptxas -v "$@" -c csrc/conv_tc.cu -o build/conv_tc_$name.o 2>&1 | grep -E "error|registers|spill" | grep -B1 -A0 "" | head -8
objs=$(ls build/*.o | grep -v "conv_tc" )
nvcc -shared -o libb200det_$name.so $objs build/conv_tc_$name.o -gencode arch=compute_100a,code=sm_100a -cudart static
ls -la libb200det_$name.so
