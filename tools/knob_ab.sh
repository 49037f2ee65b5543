#!/bin/bash
# A/B of the experiment knobs that are built but unmeasured: run on a B200 box, e.g.
#   gpurun --timeout 240 -- 'bash tools/knob_ab.sh'
# 1. bit-identity of every knob against the default (opt-in test), 2. device-resident bench per knob (value, per-kernel
# times), 3. the host-pointer call per knob.  Output: gpurun_out/knob_ab.log
out=gpurun_out/knob_ab.log
mkdir -p gpurun_out; : > $out
echo "== bit-identity" >> $out
EDB200_TEST_EXPERIMENTS=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -k experiment_knobs >> $out 2>&1
for knob in "" EDB200_EMISSION_WARPROWS=1 EDB200_CRIT_WARPS=2 EDB200_CRIT_WARPS=1 "EDB200_EMISSION_WARPROWS=1 EDB200_CRIT_WARPS=2"; do
    echo "== bench ${knob:-default}" >> $out
    env $knob timeout 60 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-aux 2>> $out | python -c "
import json, sys
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g  ms/step %.4f  frac %s/%s  kernels %s' % (j['value'], j['ms_per_step'], round(j['roofline']['frac'], 4), round(j['roofline_other']['frac'], 4), j['kernel_ms_per_step']))
" >> $out 2>&1
    echo "== e2e ${knob:-default}" >> $out
    env $knob timeout 40 python tools/e2e_probe.py --reps 8 2>&1 | grep "calls+stats " >> $out
done
cat $out
