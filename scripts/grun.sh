#!/bin/bash
# Local helper: rebuild the native library, then run a command on the GPU box.
set -e
cd /root/repo
python -c "from cytospace_b200 import _native; _native.build()"
make -s -C oracle liboracle.so
T=${GRUN_TIMEOUT:-1200}
exec /usr/local/graft/bin/gpurun --timeout $T -- "$@"
