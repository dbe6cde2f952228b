#!/bin/bash
# builds the A/B variants of libhbtunbind.so that tools/ab_walk.py compares (tools/ab/ is git-ignored but travels to the GPU box)
set -e
cd "$(dirname "$0")/../hbtplus_b200/csrc"
OUT=$(cd ../../tools && pwd)/ab
mkdir -p "$OUT"
rm -f "$OUT"/*.so
build() { # name, flags
  make -s -j8 BUILD=build_$1 LIB=$OUT/lib_$1.so EXTRA="$2" > /dev/null && echo "built $1"
}
# the variants of the last sweep of the round (profiles/r02_walk_notes.md section 1); edit to taste, then
#   python tools/ab_walk.py lib=default lib=tools/ab/lib_pre.so lib=tools/ab/lib_u2.so,walk_masked_blocks=8 ...    on the GPU box
build l196 "-DHBT_MASKED_SMEM_TOTAL=200704 -DHBT_MASKED_CARVEOUT=86" &
build l164 "-DHBT_MASKED_SMEM_TOTAL=167936 -DHBT_MASKED_CARVEOUT=72" &
wait
ls -la $OUT
