#!/bin/sh
# development build with march event counters (tools/stats_run.py); never shipped as the product library
cd "$(dirname "$0")/.."
P=self-supervised-scene-generation-with-semantic-segmentation_b200
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC,-fvisibility=hidden,-fopenmp -lgomp -DSPSG_STATS=${SPSG_STATS:-1} -I include $P/csrc/*.cu -o $P/lib/libspsg_raycast_stats.so
