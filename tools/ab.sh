#!/bin/bash
# tools/ab.sh WORKLOADS LIB...   -- on the GPU box: run tools/ab_probe.py for every workload with every library variant
# (variants are built beforehand in the build container:  python tools/ab_build.py NAME -DX=1 ...)
WLS=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  for wl in $WLS; do
    OMNI_B200_LIB=$PWD/omnirevolve-image-processor_b200/lib/$lib timeout 300 python tools/ab_probe.py $wl 20 2>&1 | tail -1 | tee -a gpurun_out/ab.jsonl
  done
done
