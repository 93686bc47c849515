#!/bin/bash
# build a kernel variant of the library out of tree: build_variant.sh NAME "-DFLAG=.. -DFLAG2=.."  -> build_variants/lib_NAME.so
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=/tmp/tb2_variant_$NAME
rm -rf $T; mkdir -p $T/traffic_b200 $T/include
cp -r $ROOT/traffic_b200/csrc $T/traffic_b200/; cp $ROOT/include/*.h $T/include/
make -C $T/traffic_b200/csrc clean > /dev/null
make -C $T/traffic_b200/csrc VARIANT="$FLAGS" > $T/make.log 2>&1
mkdir -p $ROOT/build_variants; cp $T/traffic_b200/csrc/libtraffic_b200.so $ROOT/build_variants/lib_$NAME.so
python $ROOT/scripts/ptxas_info.py $T/make.log 'pipe_kernel<double, false, false' | cut -c1-230
