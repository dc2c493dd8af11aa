#!/bin/bash
# Build A/B copies of the library with -D switches:  scripts/build_variants.sh name1 "-DX=1" name2 "-DY=2" ...
# -> camliflow_b200/_build/variants/libcamli_<name>.so ; select one with CAMLI_LIB_PATH=<path>.
mkdir -p camliflow_b200/_build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( CAMLI_LIB_PATH=$PWD/camliflow_b200/_build/variants/libcamli_$name.so CAMLI_NVCC_EXTRA="$flags" \
      python -c "from camliflow_b200.build import build_library; print(build_library(force=True))" ) &
done
wait
