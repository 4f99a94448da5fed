#!/bin/bash
# Build a variant of the library with extra -D flags for srb_api.cu (the SRBTrack step kernels) into ipcdnnwalk_b200/_ab/lib<NAME>.so:
#   tools/build_variant.sh NAME "-DSRB_OPT_PIPE -DSRB_OPT_CLAMP"
# The other translation units are taken from the current build (ipcdnnwalk_b200/csrc/_obj).  A/B with tools/ab_multi.sh.
set -e
NAME=$1; FLAGS=$2
D=ipcdnnwalk_b200
mkdir -p $D/_ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --use_fast_math $FLAGS -c -o $D/_ab/srb_api_$NAME.o $D/csrc/srb_api.cu
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $D/_ab/lib$NAME.so $D/_ab/srb_api_$NAME.o $D/csrc/_obj/fk_api.o $D/csrc/_obj/cdm_api.o $D/csrc/_obj/mmsto_api.o
rm -f $D/_ab/srb_api_$NAME.o
echo built $D/_ab/lib$NAME.so
