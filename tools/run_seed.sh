#!/bin/bash
for so in build/variants/libiglb200_c32g4.so build/variants/libiglb200_c32g2.so; do for c in c2 c3 c5; do
  export IGLB200_LIBRARY=$PWD/$so
  python bench.py --config $c --steps 3 --warmup 2 --no-cpu --no-c5 --e2e-steps 1 > gpurun_out/seed_$c.json 2>gpurun_out/seed_$c.err; python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[3], sys.argv[2], 'value %.1fM ms %.2f'%(d['value']/1e6,d['ms_per_step']), d['kernel_ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6))
" gpurun_out/seed_$c.json $c $so; done; done
