#!/bin/sh
# MNIST example of the LogisticRegression application.
#   ./run.sh <mnist dir> [ranks]
# 1 rank: local model; N ranks: parameter-server model on the host runtime (tools/mvrun.py forks
# the ranks; there is no mpirun dependency).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
python "$HERE/convert.py" "${1:-.}"
N=${2:-1}
if [ "$N" -gt 1 ]; then
  sed 's/^use_ps            = false/use_ps            = true/' "$HERE/mnist_softmax.config" > mnist_ps.config
  python "$ROOT/tools/mvrun.py" -n "$N" -- "$ROOT/build/bin/logreg" mnist_ps.config
else
  "$ROOT/build/bin/logreg" "$HERE/mnist_softmax.config"
fi
