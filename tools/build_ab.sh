#!/bin/bash
# builds the A/B variants of libhbtunbind.so that tools/ab_walk.py compares (tools/ab/ is git-ignored but travels to the GPU box)
set -e
cd "$(dirname "$0")/../hbtplus_b200/csrc"
OUT=$(cd ../../tools && pwd)/ab
mkdir -p "$OUT"
build() { # name, flags
  make -s -j8 BUILD=build_$1 LIB=$OUT/lib_$1.so EXTRA="$2" > /dev/null && echo "built $1"
}
build u2 "-DHBT_M_UNROLL=2" &
build big "-DHBT_A_PEND=32 -DHBT_M_PEND=16" &
wait
build d4 "-DHBT_M_UNROLL_D=4" &
build u2big "-DHBT_M_UNROLL=2 -DHBT_A_PEND=32 -DHBT_M_PEND=16" &
wait
ls -la $OUT
