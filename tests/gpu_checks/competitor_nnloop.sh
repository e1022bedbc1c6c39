#!/bin/bash
# Same-box NN-boundary comparison: NeuralNet::getOutput in a loop, reference CUDA/cuDNN backend vs libkgb200,
# identical caller (oracle/ref_nnloop_driver.cpp).  On the GPU box:
#   bash tests/gpu_checks/competitor_nnloop.sh [model name for modelgen | path to a .bin/.bin.gz] [batch] [iters] [dump prefix]
# With a dump prefix the first 8 rows' inputs and outputs of every run are written to <prefix>_<backend>_fp16<0|1>.bin
# (compared with each other and with the numpy oracle by tests/gpu_checks/competitor_parity.py).
MODEL=${1:-b18c384nbt}; BATCH=${2:-256}; ITERS=${3:-30}; DUMP=${4:-}
if [ -f "$MODEL" ]; then
  FILE=$MODEL
else
  D=$(mktemp -d)
  python - <<PY
import sys; sys.path.insert(0, '.')
from katago_b200 import modelgen
modelgen.write_model('$D/$MODEL.bin.gz', '$MODEL', seed=0)
PY
  FILE=$D/$MODEL.bin.gz
fi
for b in cuda b200; do
  for fp16 in 1 0; do
    echo "== kgref_nnloop_$b fp16=$fp16 $FILE"
    timeout 300 oracle/_ref/kgref_nnloop_$b $FILE $BATCH $ITERS $fp16 1 1 ${DUMP:+${DUMP}_${b}_fp16${fp16}.bin} 2>&1 | tail -3
  done
done
