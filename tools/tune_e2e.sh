#!/bin/bash
# on the GPU box: host-pointer throughput (pinned / pageable) against workers per device, chunk rows and edge chunks
run() { env "$@" python bench.py --steps 3 --warmup 1 --no-cpu --no-c5 > gpurun_out/e2e_tmp.json 2>gpurun_out/e2e_tmp.err
  python -c "
import json,sys
for l in open('gpurun_out/e2e_tmp.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print(' '.join(sys.argv[1:]),'dev %.1fM pinned %.1fM pageable %.1fM'%(d['value']/1e6,e['value']/1e6,e['pageable']['value']/1e6))
" "$@"; }
run X=default
for w in 3 6; do for e in 1 2 6; do run IGLB200_HOST_WORKERS=$w IGLB200_STREAM_EDGE=$e; done; run IGLB200_HOST_WORKERS=$w IGLB200_STREAM_EDGE_OFF=1; done
run IGLB200_HOST_WORKERS=6 IGLB200_STREAM_EDGE_OFF=1 IGLB200_STREAM_CHUNK=2097152
run IGLB200_HOST_WORKERS=4 IGLB200_STREAM_EDGE_OFF=1
