#!/bin/bash
# A/B run of library variants on the GPU box: tools/ab.sh name1 name2 ...  ("" = the product library)
# each variant: GPU parity tests, then the HMult+Relin / rotate loop and the NTT loop
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "base" ]; then unset PFHE_B200_LIB; else export PFHE_B200_LIB=$PWD/phantom-fhe_b200/libpfhe_b200_$v.so; fi
  echo "== variant $v"
  python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  python tools/hmult_loop.py 2>&1 | tail -2
  python tools/ntt_loop.py 2>&1 | tail -2
done
