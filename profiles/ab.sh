#!/bin/bash
# Same-box A/B of library builds (the gpurun boxes differ by up to 20 % on identical builds, so versions are only
# compared inside one call).  Put the builds as gpurun_ab/lib_<name>.so (git-ignored, travels with the snapshot) and run
#     gpurun -- 'VARIANTS="old new" bash profiles/ab.sh'
# Every variant is timed twice, interleaved, on one bench slab (train + adjust, CUDA events).
last=""
for rep in 1 2; do
  for v in $VARIANTS; do
    cp gpurun_ab/lib_$v.so xsdba_b200/libxsdba_b200.so
    echo "== $v"; python profiles/train_only.py 48 3 adjust 2>&1 | tail -1
    last=$v
  done
done
cp gpurun_ab/lib_$last.so xsdba_b200/libxsdba_b200.so
