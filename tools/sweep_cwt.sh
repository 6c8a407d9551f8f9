#!/bin/bash
# usage: tools/sweep_cwt.sh <group sizes...>   (items per cols/rows launch pair in the CWT fast path)
mkdir -p gpurun_out
for g in "$@"; do
  AFB200_CWT_GROUP=$g timeout 200 python tools/bench_cqt_cwt.py --cqt-batch 8 --cwt-batch 8 > gpurun_out/cwt_g$g.json 2> gpurun_out/cwt_g$g.err
  python -c "
import json
d=json.load(open('gpurun_out/cwt_g$g.json'))['cwt']; print('group=$g', round(d['ms']/d['batch'],3),'ms/clip', round(d['compulsory_GBs']),'GB/s')"
done
