#!/bin/bash
# Builds kernel variants into offline_raytracer_b200/variants/<name>.so (development: A/B timing on the GPU box
# with tools/variant_bench.py).  Usage: tools/build_variants.sh name1="-DFLAG=1 ..." name2="..." ...
set -e
cd "$(dirname "$0")/../offline_raytracer_b200"
mkdir -p variants
for spec in "$@"; do
    name="${spec%%=*}"; flags="${spec#*=}"
    rm -rf "vb_$name"; mkdir "vb_$name"
    cp csrc/*.cu csrc/*.cuh csrc/*.h csrc/*.cpp csrc/Makefile "vb_$name"/
    ( cd "vb_$name" && make -s OUT="../variants/$name.so" EXTRA_NVFLAGS="$flags" EXTRA_CXXFLAGS="$(echo "$flags" | tr " " "\n" | grep "^-D" | tr "\n" " ")" >/dev/null 2>"../variants/$name.log"; cp ptxas.log "../variants/$name.ptxas" 2>/dev/null; cd .. && rm -rf "vb_$name" ) &
done
wait
ls -la variants/*.so
