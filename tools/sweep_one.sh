#!/bin/bash
# on the GPU box: the default bench at N = $1 GPUs
n=$1
if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps ${STEPS:-10} --warmup 3 > gpurun_out/sweep_n$n.json 2> gpurun_out/sweep_n$n.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 > gpurun_out/sweep_n$n.json 2> gpurun_out/sweep_n$n.err; fi
grep -h "Error\|error" gpurun_out/sweep_n$n.err | tail -3
python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); c5=d.get('c5',{}); e=d['e2e']
        print('N=%s c2 %.1fM q/s %.2f ms | e2e %.1fM pageable %s single_process %s | c5 %s'%(sys.argv[2], d['value']/1e6, d['ms_per_step'], (e.get('value') or 0)/1e6, (e.get('pageable') or {}).get('value'), (e.get('single_process') or {}).get('value'), ('%.1fM q/s %.1f ms'%(c5['value']/1e6,c5['ms_per_step'])) if 'value' in c5 else c5), d['kernel_ms_per_step'])
" gpurun_out/sweep_n$n.json $n
