#!/bin/bash
# build_variant.sh NAME "EXTRA_NVCC_FLAGS": builds librfwb200.so with extra compile flags (e.g. -DSHADE_MINB=6) out of tree and
# leaves it as rendering-fw_b200/_exp/librfwb200_NAME.so for A/B runs (RFWB200_LIB=... python tools/variant_sweep.py)
set -e
NAME=$1; EXTRA=$2
REPO=$(cd "$(dirname "$0")/.." && pwd)
B=/tmp/rfwb200_build_$NAME
rm -rf $B && mkdir -p $B/rendering-fw_b200 $B/include
cp -r $REPO/rendering-fw_b200/csrc $REPO/rendering-fw_b200/Makefile $REPO/rendering-fw_b200/data $B/rendering-fw_b200/ 2>/dev/null || true
rm -rf $B/rendering-fw_b200/data/_baked
cp $REPO/include/rfwb200.h $B/include/
rm -f $B/rendering-fw_b200/csrc/*.o
make -C $B/rendering-fw_b200 -j8 NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas -DBLUENOISE_PATH='\"$B/rendering-fw_b200/data/bluenoise_256spp.bin\"' -Xptxas -v $EXTRA" > $B/build.log 2>&1 || (tail -20 $B/build.log; exit 1)
mkdir -p $REPO/rendering-fw_b200/_exp
cp $B/rendering-fw_b200/librfwb200.so $REPO/rendering-fw_b200/_exp/librfwb200_$NAME.so
grep -h "k_shade\b\|Used" $B/rendering-fw_b200/csrc/kernels_shade.ptxas.log | grep -A1 "k_shadeE" | tail -1
echo built $REPO/rendering-fw_b200/_exp/librfwb200_$NAME.so
