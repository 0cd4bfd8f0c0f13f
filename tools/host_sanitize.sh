#!/bin/bash
# AddressSanitizer + ThreadSanitizer runs of the host expansion driver (no GPU: the refine() stand-in of the test hooks runs in
# the "GPU" thread, so the two-rounds-in-flight loop, the inline cell-id container and the two-generation scratch are exercised).
# usage: tools/host_sanitize.sh        (from the repo root; prints one line per mode and any sanitizer report)
set -e
R=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
SRC="$R/tools/host_sanitize_main.cpp $R/pais-mvs_b200/host/tmvs_hooks.cpp $R/pais-mvs_b200/host/tmvs_lib.cpp"
LNK="-L$R/pais-mvs_b200/lib -lpmvs_b200 -Wl,-rpath,$R/pais-mvs_b200/lib -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64"
g++ -O1 -g -std=c++17 -pthread -fopenmp -ffp-contract=off -fsanitize=address -fno-omit-frame-pointer -o $T/asan $SRC $LNK
g++ -O1 -g -std=c++17 -pthread -fopenmp -ffp-contract=off -fsanitize=thread -o $T/tsan $SRC $LNK
for m in 1 2 0; do ASAN_OPTIONS=detect_leaks=1:protect_shadow_gap=0 $T/asan $m 2>&1 | grep -E "ERROR|SUMMARY|^mode"; done
# libgomp is not instrumented: one OpenMP thread, so that only the driver's own std::thread is under test
for m in 1 2; do OMP_NUM_THREADS=1 $T/tsan $m 2>&1 | grep -E "WARNING|SUMMARY|^mode"; done
rm -rf $T
