#!/bin/bash
mkdir -p gpurun_out
for ns in "$@"; do
  AFB200_MFCC_SKEW_NS=$ns timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/skew_$ns.json 2> gpurun_out/skew_$ns.err
  python -c "
import json
d=json.load(open('gpurun_out/skew_$ns.json')); print('skew_ns=$ns', round(d['value']/1e6,1),'Mframes/s', round(d['ms_per_step'],3),'ms')"
done
