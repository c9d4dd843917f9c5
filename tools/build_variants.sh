#!/bin/bash
# Build A/B variants of the library into lighter_b200/variants/ (git-ignored .so files travel to the GPU box).
#   usage: tools/build_variants.sh name1="-DFOO=1 -DBAR=2" name2="..."
set -e
cd "$(dirname "$0")/../lighter_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
    name="${spec%%=*}"; flags="${spec#*=}"
    make -j"$(nproc)" OUT=../variants/lib_$name.so BUILD=../../build/obj_$name EXTRA_NVFLAGS="$flags" >/dev/null
    echo "built variants/lib_$name.so  [$flags]"
done
