#!/bin/bash
mkdir -p gpurun_out
for kb in "$@"; do
  AFB200_CWT_LEG_KB=$kb timeout 200 python tools/bench_cqt_cwt.py --cqt-batch 8 --cwt-batch 8 > gpurun_out/cwt_$kb.json 2> gpurun_out/cwt_$kb.err
  python -c "
import json
d=json.load(open('gpurun_out/cwt_$kb.json'))['cwt']; print('leg_kb=$kb', round(d['ms']/d['batch'],3),'ms/clip', round(d['compulsory_GBs']),'GB/s')"
done
