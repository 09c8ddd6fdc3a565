#!/bin/bash
# builds tools/umma_probe/libprobe.so (development tool; sm_100a)
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -shared -Xcompiler -fPIC --cudart shared -o libprobe.so probe.cu
