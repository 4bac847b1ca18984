#!/bin/bash
# Tuning build: the engine with stage stamps in the single-system kernels (-DRB_STAMPS), as
# variants/lib_stamps.so (RGPOT_B200_ENGINE selects it; tools/small_calls.py --stamps prints the stages).
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off"
nvcc $FLAGS -DRB_STAMPS -c -o variants/engine_stamps.o rgpot_b200/csrc/engine.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_stamps.so variants/engine_stamps.o rgpot_b200/csrc/tile.o -lcudart
