#!/bin/bash
# Variant of the library with extra -D flags for cdm_api.cu (the AdaptiveSRB step kernels) into ipcdnnwalk_b200/_ab/lib<NAME>.so:
#   tools/build_variant_cdm.sh NAME "-DCDM_WPB=4"
set -e
NAME=$1; FLAGS=$2
D=ipcdnnwalk_b200
mkdir -p $D/_ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --use_fast_math $FLAGS -c -o $D/_ab/cdm_api_$NAME.o $D/csrc/cdm_api.cu
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $D/_ab/lib$NAME.so $D/_ab/cdm_api_$NAME.o $D/csrc/_obj/fk_api.o $D/csrc/_obj/srb_api.o $D/csrc/_obj/mmsto_api.o
rm -f $D/_ab/cdm_api_$NAME.o
echo built $D/_ab/lib$NAME.so
