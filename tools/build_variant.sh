#!/bin/bash
# Builds qchem-rs_b200/libqcfock_<name>.so: the block-kernel classes recompiled with extra nvcc flags (kernel A/B runs),
# everything else taken from the default build directory.   usage: tools/build_variant.sh <name> <flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/../qchem-rs_b200/csrc"
mkdir -p build_$name
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC"
BLOCK="0000 1000 1010 1100 1110 1111 2000 2010 2011 2020"
SLAB="2100 2110 2111 2120 2121 2200 2210 2211 2220 2221 2222"
for c in $BLOCK; do
  nvcc $NVFLAGS "$@" -DQCF_LA=${c:0:1} -DQCF_LB=${c:1:1} -DQCF_LC=${c:2:1} -DQCF_LD=${c:3:1} -c eri_class.cu -o build_$name/class_$c.o &
done
wait
objs=""
for c in $BLOCK; do objs="$objs build_$name/class_$c.o"; done
for c in $SLAB; do objs="$objs build/class_$c.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libqcfock_$name.so $objs build/engine.o build/scf_device.o build/loaders.o \
  -L/usr/local/cuda/lib64 -lcusolver -lcublas -Xlinker -rpath=/usr/local/cuda/lib64
echo built ../libqcfock_$name.so
