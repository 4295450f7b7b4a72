#!/bin/bash
# Build an A/B variant of the library: recompile ONE source with extra flags, link with the in-tree objects of the others.
#   bash tools/build_variant.sh NAME knn.cu "-DDFB_KNN_MINB=6"   ->  ab_libs/lib_NAME.so  (select with DEEPFIT_B200_LIB)
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; EXTRA=$3
mkdir -p ab_libs/obj_$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
nvcc $FLAGS $EXTRA -Xptxas -v -c deepfit_b200/csrc/$SRC -o ab_libs/obj_$NAME/${SRC%.cu}.o 2> ab_libs/obj_$NAME/ptxas.log
OBJS=""
for f in api knn pca wls mlp pipeline textio; do
  if [ "$f.cu" == "$SRC" ]; then OBJS="$OBJS ab_libs/obj_$NAME/$f.o"; else OBJS="$OBJS deepfit_b200/csrc/build/$f.o"; fi
done
nvcc -shared -o ab_libs/lib_$NAME.so $OBJS -gencode arch=compute_100a,code=sm_100a -cudart static
echo ab_libs/lib_$NAME.so
