#!/bin/bash
# builds the measurement tools under tools/probe into build/ (in-tree: they travel to the GPU box)
set -e
cd "$(dirname "$0")/../.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -I include -I uda_poseestimation_b200/csrc \
  -shared -o build/peer_probe.so tools/probe/peer_probe.cu -cudart static
echo build/peer_probe.so
