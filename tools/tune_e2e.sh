#!/bin/bash
# on the GPU box: host-pointer throughput (pinned / pageable) against workers per device and chunk rows
for w in 3 4 6 8; do for c in 262144 524288 1048576; do
  IGLB200_HOST_WORKERS=$w IGLB200_STREAM_CHUNK=$c python bench.py --steps 3 --warmup 1 --no-cpu --no-c5 > gpurun_out/e2e_${w}_${c}.json 2>gpurun_out/e2e_${w}_${c}.err
  python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('workers',sys.argv[2],'chunk',sys.argv[3],'dev %.1fM pinned %.1fM pageable %.1fM'%(d['value']/1e6,e['value']/1e6,e['pageable']['value']/1e6))
" gpurun_out/e2e_${w}_${c}.json $w $c
done; done
