#!/bin/bash
# pass 1 of the two-pass path with and without CTA-pair multicast (same process tree, for one ncu session)
B=scaled-mmd-gan_b200/build/tc_check
SMMD_WZ_PAIR=0 $B mmd mix_rq 32768 32768 1024 1 0
SMMD_WZ_PAIR=1 $B mmd mix_rq 32768 32768 1024 1 0
