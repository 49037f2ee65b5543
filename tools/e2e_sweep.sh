#!/bin/bash
# e2e knob sweep of the host-pointer call (development aid): timeline of the default, then chromosome-group cuts and
# sweep packing plans.  Output: gpurun_out/e2e_sweep.log
out=gpurun_out/e2e_sweep.log
mkdir -p gpurun_out; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 40 python tools/e2e_probe.py --reps 8 >> $out 2>&1; }
run EDB200_TIMELINE=1
run EDB200_PARTS=4
run EDB200_PARTS=5
run EDB200_CUTS=0.12,0.35,0.6,0.8,0.93,1.0
run EDB200_CUTS=0.2,0.45,0.7,0.88,0.97,1.0
run EDB200_CUTS=0.3,0.55,0.75,0.9,0.97,1.0
run EDB200_CUTS=0.25,0.5,0.7,0.85,0.98,1.0
run EDB200_PACKPLAN=122222
run EDB200_PACKPLAN=111222
run EDB200_PACKPLAN=121111
run EDB200_PACKPLAN=222222
grep -v timeline $out | grep "==\|calls+stats " 
