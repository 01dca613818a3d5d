#!/bin/bash
# everything that can be checked without a GPU, in the order the round-end driver would hit it:
#   build (nvcc cross-compile for sm_100a + the C oracle), header as plain C, CPU test suite, the same suite against the ASan/UBSan host build
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
gcc -x c -std=c99 -Wall -Wextra -pedantic -Werror -fsyntax-only include/gpud_b200.h
python -m pytest tests -x -q -m "not gpu"
tools/build_asan.sh > /dev/null
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 GPUD_B200_LIB=$PWD/build/libgpud_asan.so python -m pytest tests -x -q -m "not gpu"
rm -rf build
echo "check_all: ok"
