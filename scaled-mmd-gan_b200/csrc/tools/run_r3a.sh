#!/bin/bash
# round 2, session 3, call A: setmaxnreg in tc_symf_kernel, A/B against the previous library (lib_old)
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/r3a_tc_check.log
: > $L
run() { echo "\$ $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for rep in 1 2; do
for lib in lib_old lib; do
export LD_LIBRARY_PATH=$PWD/scaled-mmd-gan_b200/$lib
echo "=== $lib" >> $L
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 65536 65536 256 4 0
run $B mmd rbf 32768 32768 256 5 0
done
done
export LD_LIBRARY_PATH=$PWD/scaled-mmd-gan_b200/lib
run $B suite
grep -vE "^   sum\[" $L
python -X faulthandler -m pytest tests/test_gpu_sym.py -m gpu -q -x 2>&1 | tail -3
