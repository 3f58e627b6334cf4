#!/bin/bash
# tools/gpu_generic.sh -- parity of the generic kernels at any even M, then their throughput (new library, and the one
# before the mixed-radix transform on the geometries where a direct DFT finishes in reasonable time)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "generic or any_even or any_M or geometry or autotest or random_prototype or roundtrip or firpfbch" > gpurun_out/gen_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gen_pytest.log
tail -n 6 gpurun_out/gen_pytest.log
( echo "== tiled generic kernels"; timeout 600 python tools/bench_generic.py 2>&1 | grep path
  echo "== one-frame-per-block kernels (YG_GENERIC_TILED=0)"; YG_GENERIC_TILED=0 timeout 600 python tools/bench_generic.py 2>&1 | grep path
  if [ -f yagi_b200/lib/libyagi_b200_old.so ]; then
    echo "== before (tools/build_variant.sh old ...), N = 2^22"; YG_LIB=old YG_LOG2N=22 YG_GEN_GEOM=10:5,24:5,48:5,100:5,240:4,1000:4 timeout 600 python tools/bench_generic.py 2>&1 | grep path
  fi ) | tee gpurun_out/gen_bench.log
