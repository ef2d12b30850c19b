#!/bin/bash
# alternative builds of the library for the tuning sweeps (profiles/tune_*.py): build/variants/<name>.so = the current objects with
# one source (SRC, default decode.cu) recompiled under extra -D flags.
# usage: [SRC=nms.cu] build_variants.sh name "-DYB_DC_THREADS=64 -DYB_DC_MINBLOCKS=8" [name flags ...]
set -e
cd "$(dirname "$0")/.."
python -m pytorch_yolo_b200.build > /dev/null
SRC=${SRC:-decode.cu}
mkdir -p build/variants build/vobj
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -c -o build/vobj/v_$name.o pytorch_yolo_b200/csrc/$SRC
  objs=$(ls build/obj/*.cu.o | grep -v /$SRC.o)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/$name.so build/vobj/v_$name.o $objs
  echo built build/variants/$name.so
done
