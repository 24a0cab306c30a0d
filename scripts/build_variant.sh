#!/bin/bash
# build_variant.sh NAME "-DFLAG ..." : builds trace.jl_b200/csrc/libtrace_cuda_NAME.so with extra nvcc flags (A/B experiments;
# select it with TRACE_CUDA_LIB=...), then rebuilds the default library.
set -e
cd "$(dirname "$0")/../trace.jl_b200/csrc"
make clean > /dev/null
make EXTRA="$2" > /dev/null
cp libtrace_cuda.so /tmp/libtrace_cuda_$1.so
make clean > /dev/null
make > /dev/null
cp /tmp/libtrace_cuda_$1.so libtrace_cuda_$1.so
ls -la libtrace_cuda*.so
