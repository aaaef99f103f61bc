#!/bin/sh
# development aid: build libspsg_raycast_<tag>.so variants with extra -D flags:  tools/variants.sh tag1 "-DX=1" tag2 "-DX=2" ...
cd "$(dirname "$0")/.."
P=self-supervised-scene-generation-with-semantic-segmentation_b200
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC,-fvisibility=hidden $2 -I include $P/csrc/*.cu -o $P/lib/libspsg_raycast_$1.so || exit 1
  shift 2
done
