#!/bin/bash
# Compiles the reference's OWN CUDA sources, unmodified, where they lie under /root/reference, together with the two ROS
# stub headers, the cuTT shim and the ROS-free driver of oracle/ref_harness/.  Outputs only into oracle/_ref/.
#   ref_driver_parity : IEEE float semantics (-fmad=false, no fast-math)  -> golden fixtures
#   ref_driver_fast   : the reference's Release flags (CMakeLists.txt:32)  -> timing arm of bench.py
# Test infrastructure only; never part of the product.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
[ -d "$REF/src" ] || { echo "no reference at $REF"; exit 0; }
mkdir -p "$OUT/obj_parity" "$OUT/obj_fast"
INC="-I$HERE/ref_harness/stubs -I$REF/include -I$REF/include/par_wave -I$REF/src"
COMMON="-std=c++17 -gencode arch=compute_100a,code=sm_100a -DNDEBUG -w $INC"
SRCS="$REF/src/kernel/edt/local_edt.cu $REF/src/kernel/edt/warmup.cu $REF/src/kernel/par_wave/glb_hash_map.cu \
$REF/src/kernel/point_cloud/pntcld_raycast.cu $REF/src/kernel/hokuyo/hokuyo_fast.cu $REF/src/kernel/vlp16/vlp16_fast.cu \
$REF/src/kernel/realsense/realsense_fast.cu $REF/src/kernel/pre_map/pre_map.cu \
$HERE/ref_harness/cutt_shim.cu $HERE/ref_harness/ref_driver.cu"
build() {  # $1 = flavour, $2 = flags
    local objs=""
    for s in $SRCS; do
        o="$OUT/obj_$1/$(basename "$s" .cu).o"
        objs="$objs $o"
        if [ ! -f "$o" ] || [ "$s" -nt "$o" ]; then
            ( $NVCC $COMMON $2 -c "$s" -o "$o" ) &
        fi
    done
    wait
    $NVCC -gencode arch=compute_100a,code=sm_100a -o "$OUT/ref_driver_$1" $objs
}
build parity "-O3 -fmad=false"
build fast "-O3 -use_fast_math -ftz=true -prec-div=false -prec-sqrt=false"
ls -la "$OUT"/ref_driver_*
