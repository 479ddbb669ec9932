#!/bin/bash
# Builds a tuning variant of the library: build_variant.sh <name> <-D flags...>  ->  oriented_object_detection_b200/lib/variants/<name>.so
# (run with GM_LIB_PATH=<that file>; the default build is __graft_entry__.build()).
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p oriented_object_detection_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 550 "$@" \
  -shared -o oriented_object_detection_b200/lib/variants/$NAME.so oriented_object_detection_b200/csrc/*.cu
