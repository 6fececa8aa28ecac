#!/bin/bash
tools/runvar.sh
for c in c5 c3; do python bench.py --config $c --steps 3 --warmup 2 --no-cpu --e2e-steps 1 > gpurun_out/seed_$c.json 2>gpurun_out/seed_$c.err; python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], 'value %.1fM ms %.2f'%(d['value']/1e6,d['ms_per_step']), d['kernel_ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6))
" gpurun_out/seed_$c.json $c; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
