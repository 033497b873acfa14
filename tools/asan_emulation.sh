#!/bin/bash
# compute-sanitizer is closed on this GPU pool; instead the kernel's device code
# (bayesair_b200/csrc/ba_sim.cuh, ba_scan2.cuh), compiled for the host by tests/hostemu, is run under
# AddressSanitizer + UBSan over every golden case, both list variants, spill and overflow.
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/_build
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -std=c++17 \
    -fPIC -shared -Wno-unknown-pragmas -o tests/_build/libba_hostemu_asan.so \
    tests/hostemu/ba_hostemu.cpp bayesair_b200/csrc/ba_plan_build.cpp
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
ASAN_OPTIONS=detect_leaks=0 python tools/asan_emulation.py
