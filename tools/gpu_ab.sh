#!/bin/bash
# tools/gpu_ab.sh "BENCH ARGS" NAME... -- same-box A/B of library variants (tools/build_variant.sh) on tools/bench_kernels.py;
# "-" is the stock library.  Two interleaved rounds per size so box noise shows.
mkdir -p gpurun_out
args=$1; shift
for round in 1 2; do for n in ${YG_AB_SIZES:-24 26}; do for v in "$@"; do
  echo "== round $round YG_LOG2N=$n lib=$v"
  if [ "$v" = "-" ]; then YG_LOG2N=$n timeout 300 python tools/bench_kernels.py $args 2>&1 | grep '^{'
  else YG_LIB=$v YG_LOG2N=$n timeout 300 python tools/bench_kernels.py $args 2>&1 | grep '^{'; fi
done; done; done | tee gpurun_out/ab_$(echo "$@" | tr ' -' '__').log
