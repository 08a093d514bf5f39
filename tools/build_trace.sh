# Diagnostic build of the native library with the level-1 kernel's hand-off trace (tools/coarse_trace.py reads it).
set -e
cd "$(dirname "$0")/.."
OUT=libcluster_b200/_lib/trace
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fopenmp,-Wall,-Wno-unused-function -DLCB_COARSE_TRACE"
for f in kernels.cu tc_kernels.cu engine.cu host_model.cpp c_api.cpp; do
  nvcc $FLAGS -x cu -c libcluster_b200/csrc/$f -o $OUT/${f%.*}.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/liblcb200_trace.so $OUT/*.o -Xcompiler -fopenmp -lcudart_static -ldl -lpthread -lrt
echo built $OUT/liblcb200_trace.so
