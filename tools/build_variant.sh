#!/bin/bash
# Builds a tuning variant of liborphx.so: tools/build_variant.sh NAME -DOX_KB_ROWS64=2 ...
# -> orphics_b200/_lib/liborphx_NAME.so (select with ORPHX_LIB=...); only ox_fused.cu is recompiled.
set -e
cd "$(dirname "$0")/../orphics_b200/csrc"
name=$1; shift
make -s >/dev/null
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c ox_fused.cu -o build/ox_fused_$name.o
qe=build/ox_qe.o
if [ -n "$ALSO_QE" ]; then   # the row / column kernels are also instantiated in ox_qe.cu
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c ox_qe.cu -o build/ox_qe_$name.o
  qe=build/ox_qe_$name.o
fi
objs="build/ox_core.o build/ox_binner.o build/ox_sim.o build/ox_power.o build/ox_pipeline.o $qe build/ox_splits.o"
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_lib/liborphx_$name.so $objs build/ox_fused_$name.o -L/usr/local/cuda/lib64 -lcufft -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo built liborphx_$name.so
