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
build pre "-DHBT_M_PREFETCH=1" &
build exold "-DHBT_M_UNROLL_D=4" &
wait
build big "-DHBT_A_PEND=24 -DHBT_M_PEND=12" &
build u8 "-DHBT_M_UNROLL=8" &
wait
ls -la $OUT
