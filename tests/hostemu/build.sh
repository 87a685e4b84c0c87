#!/bin/sh
# Host-emulation build of the kernel sources (TEST INFRASTRUCTURE ONLY, never loaded by khepri_b200):
# the same .cu/.cuh files compiled as plain C++ with one virtual thread per CTA, so that the kernel
# logic can be exercised through the C ABI on a machine without a GPU.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
g++ -O2 -std=c++17 -fPIC -shared -DKH_HOST_EMU -x c++ "$ROOT/khepri_b200/csrc/kh_api.cu" -o "$HERE/libkh_hostemu.so"
