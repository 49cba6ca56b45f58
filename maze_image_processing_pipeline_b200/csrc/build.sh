#!/bin/sh
# Build libmaze_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -Xcompiler -fPIC -shared -I"$HERE/../../include" -I"$HERE" \
    -o "$HERE/libmaze_b200.so" \
    "$HERE/maze_morph.cu" "$HERE/maze_label.cu" "$HERE/maze_props.cu" "$HERE/maze_merge.cu" "$HERE/maze_merge_win.cu" "$HERE/maze_synth.cu" "$HERE/maze_prof.cu" "$HERE/maze_fused.cu" "$HERE/maze_host.cpp" "$HERE/maze_step.cu" "$HERE/maze_shape.cu" "$HERE/maze_bands.cu" "$HERE/maze_wide.cu" \
    "$@"
