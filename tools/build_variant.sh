#!/bin/bash
# Builds a variant of the product library for an A/B run on the GPU box (tools/ab_cfg2.py): the working tree's scht_device.cu
# compiled with extra flags, linked with the objects of the regular build.
# usage: tools/build_variant.sh NAME [nvcc flags...]   ->   build_ab/libscht_NAME.so   (build_ab/ holds only *.so / *.o: git-ignored)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=semi-convex_hull_tree_b200/csrc
make -s -C $C >/dev/null
mkdir -p build_ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o build_ab/scht_device_$name.o $C/scht_device.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_ab/libscht_$name.so build_ab/scht_device_$name.o \
  $C/build/scht_create.o $C/build/scht_multi.o $C/build/scht_comm.o $C/build/scht_build.o $C/build/tree_builder.o $C/build/scht_error.o $C/build/scht_io.o \
  -cudart shared -lpthread -ldl
ls -la build_ab/libscht_$name.so
