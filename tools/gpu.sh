#!/bin/bash
# build everything that travels to the GPU box, then run a command there:  tools/gpu.sh [--gpus N] TIMEOUT 'command'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" >/dev/null
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
exec /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@"
