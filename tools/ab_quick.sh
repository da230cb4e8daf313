#!/bin/bash
# timing-only A/B of library variants on one GPU box, interleaved twice: tools/ab_quick.sh name1 name2 ...
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "base" ]; then unset PFHE_B200_LIB; else export PFHE_B200_LIB=$PWD/phantom-fhe_b200/libpfhe_b200_$v.so; fi
  echo "== variant $v (pass $rep)"
  python tools/hmult_loop.py 2>&1 | tail -2
  python tools/ntt_loop.py 2>&1 | tail -2
  python tools/two_lane.py 200 2>&1 | sed -n 2p
done
done
