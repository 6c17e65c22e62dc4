#!/bin/bash
# scratch/gpu.sh <gpurun --timeout seconds> [--gpus N] -- '<command>': rebuild both libraries, then run on a B200 box
set -e
cd "$(dirname "$0")/.."
python -c 'import __graft_entry__ as g; g.build()' > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
T=$1; shift
exec timeout $((T + 1900)) /usr/local/graft/bin/gpurun --timeout $T "$@"
