#!/bin/bash
# build_variant.sh NAME SRC.cu [nvcc flags...]  -> variants/lib_NAME.so : the library with ONE translation unit rebuilt
# from SRC (a file in hmvec_b200/csrc or a path) with extra flags -- A/B measurement builds for tools/kbench.py.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; SRC=$2; shift 2
CS=$ROOT/hmvec_b200/csrc
make -s -C $CS -j8
mkdir -p $ROOT/variants/obj
UNIT=$(basename $SRC .cu)
case $UNIT in _old_*) REPL=${UNIT#_old_};; *) REPL=$UNIT;; esac
[ -f "$SRC" ] || SRC=$CS/$SRC
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$CS "$@" -c $SRC -o $ROOT/variants/obj/${NAME}_$REPL.o
OBJS=$(ls $CS/build/*.o | grep -v "/$REPL.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/lib_$NAME.so $OBJS $ROOT/variants/obj/${NAME}_$REPL.o
echo built variants/lib_$NAME.so
