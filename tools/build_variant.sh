#!/bin/bash
# tools/build_variant.sh NAME [nvcc flags]: rebuild csrc/ccb_rollout_rows.cu with extra flags into build_var/libccb_NAME.so (A/B runs through CCB_LIB)
set -e
ROOT=/root/repo; CS=$ROOT/constrainedcontrol.jl_b200/csrc; OUT=$ROOT/build_var
name=$1; shift
mkdir -p $OUT
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -Xptxas -v -c $CS/ccb_rollout_rows.cu -o $OUT/rows_$name.o 2> $OUT/rows_$name.log
objs=""; for f in ccb_mech ccb_api ccb_rollout ccb_rollout_chain ccb_rollout_mw ccb_linearize ccb_riccati ccb_riccati_project ccb_riccati_dmma ccb_riccati_big ccb_are ccb_synth; do objs="$objs $CS/$f.o"; done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libccb_$name.so $objs $OUT/rows_$name.o
