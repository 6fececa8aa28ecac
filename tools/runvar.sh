#!/bin/bash
# on the GPU box: bench.py (short) against the default build and every variant in build/variants/
ARGS="${@:---steps 4 --warmup 2 --no-cpu --no-c5 --e2e-steps 1}"
summ() { python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], 'value %.1fM ms %.2f'%(d['value']/1e6,d['ms_per_step']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'e2e %.1fM'%(d['e2e']['value']/1e6 if d['e2e'].get('value') else 0))
" $1 $2; }
python bench.py $ARGS > gpurun_out/var_default.json 2> gpurun_out/var_default.err || tail -5 gpurun_out/var_default.err
summ gpurun_out/var_default.json default
for so in build/variants/*.so; do
  n=$(basename $so .so); n=${n#libiglb200_}
  IGLB200_LIBRARY=$PWD/$so python bench.py $ARGS > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err || tail -5 gpurun_out/var_$n.err
  summ gpurun_out/var_$n.json $n
done
