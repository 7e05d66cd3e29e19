#!/bin/sh
# Train word embeddings on a text corpus (counterpart of Applications/WordEmbedding/example/
# run.bat: dim 300, CBOW, 5 negatives, window 5, lr 0.01-ish, 16 threads).
#   ./run.sh corpus.txt [ranks]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
CORPUS=${1:?usage: run.sh corpus.txt [ranks]}
N=${2:-1}
"$ROOT/build/bin/word_count" -train_file "$CORPUS" -save_vocab vocab.txt -min_count 5
ARGS="-train_file $CORPUS -read_vocab vocab.txt -output vectors.bin -binary 1 -size 300 -cbow 1 -negative 5 \
 -window 5 -alpha 0.025 -sample 1e-4 -min_count 5 -epoch 5 -threads 16 -data_block_size 10000000 -is_pipeline 1"
if [ "$N" -gt 1 ]; then
  python "$ROOT/tools/mvrun.py" -n "$N" --timeout 86400 -- "$ROOT/build/bin/wordembedding" $ARGS
else
  "$ROOT/build/bin/wordembedding" $ARGS
fi
# GPU (device backend), same flags:
#   torchrun --nproc-per-node 8 -m multiverso_b200.apps.wordembedding $ARGS
