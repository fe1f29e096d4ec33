#!/bin/bash
# Build experimental variants of the small fused-pass kernel as separate shared objects (same C ABI) so that one GPU
# call can time them side by side:  tools/k1_variants.sh build   (here, no GPU)
#                                   tools/k1_variants.sh run "c2_ncu,c2" (on the GPU box; one jsonl per variant)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/dlsa_b200/csrc
VD=$CS/build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
build_one() {  # name, source, extra flags
    name=$1; src=$2; shift 2
    nvcc $FLAGS "$@" -c "$src" -o $VD/$name.o
    nvcc $ARCH -shared -o $VD/libdlsa_$name.so $VD/$name.o $CS/build/abi.o $CS/build/irls_tiled.o $CS/build/irls_cat.o \
        $CS/build/linalg.o $CS/build/lars.o $CS/build/misc.o -lcudart
}
case "$1" in
build)
    make -C $CS -j8 >/dev/null
    mkdir -p $VD
    if [ -n "$K1_BASE_REV" ]; then
        git -C $ROOT show $K1_BASE_REV:dlsa_b200/csrc/irls_small.cu > $VD/base_irls_small.cu
        sed -i 's#"common.cuh"#"../../common.cuh"#' $VD/base_irls_small.cu
        build_one base $VD/base_irls_small.cu
    fi
    R72="-DDLSA_K1_REGS_ROW=72 -DDLSA_K1_REGS_MMA=184"
    build_one nofb $CS/irls_small.cu $R72 -DDLSA_K1_FRAGBUF=0 -DDLSA_K1_XCHUNK=0 &
    build_one nofb_noahead $CS/irls_small.cu $R72 -DDLSA_K1_FRAGBUF=0 -DDLSA_K1_XCHUNK=0 -DDLSA_K1_DMUL_AHEAD=0 &
    build_one nofb_noahead_nog $CS/irls_small.cu $R72 -DDLSA_K1_FRAGBUF=0 -DDLSA_K1_XCHUNK=0 -DDLSA_K1_DMUL_AHEAD=0 -DDLSA_K1_GSAC=0 &
    build_one noahead $CS/irls_small.cu $R72 -DDLSA_K1_DMUL_AHEAD=0 &
    wait
    ls -la $VD/*.so
    ;;
run)
    cases=${2:-c2_ncu,c2}
    for so in $VD/libdlsa_*.so; do
        v=$(basename $so .so); v=${v#libdlsa_}
        echo "== variant $v"
        DLSA_B200_LIB=$so python $ROOT/tools/k1_sweep.py --cases "$cases" --reps 5 --out $ROOT/gpurun_out/k1_variant_$v.jsonl
    done
    ;;
esac
