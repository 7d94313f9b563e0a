#!/bin/bash
# round 2, session 3, call C: setmaxnreg in all 16-epilogue-warp kernels, A/B against the previous library (lib_old), full GPU tests
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/r3c_tc_check.log
: > $L
run() { echo "\$ $*" >> $L; timeout 300 "$@" 2>&1 | grep -v "sum\[\|value-only\|clock start" >> $L; }
for lib in lib_old lib lib_old lib; do
export LD_LIBRARY_PATH=$PWD/scaled-mmd-gan_b200/$lib
echo "=== $lib" >> $L
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 16384 16384 512 10 0
run $B mmd mix_rq 32768 32768 1024 4 0
SMMD_SYM=0 run $B mmd mix_rq 16384 16384 256 10 0
SMMD_SYM=0 run $B mmd mix_rq 16384 16384 1024 5 0
done
cat $L
export LD_LIBRARY_PATH=
python -X faulthandler -m pytest tests -m gpu -q -x 2>&1 | grep -v "^  File \"/opt" | tail -5 | tee gpurun_out/r3c_pytest.log
