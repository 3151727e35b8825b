#!/bin/bash
# same-box A/B of two library builds on the training pass: _lib/libpixie_b200_head.so vs _lib/libpixie_b200_new.so
SHAPES=${SHAPES:-"5241600 32 10 10;3355392 40 20 20"}
IFS=';' read -ra SH <<< "$SHAPES"
for rep in 1 2; do
for shape in "${SH[@]}"; do
  echo "head: $(PIXIE_LIB_PATH=$PWD/ark_analysis_b200/_lib/libpixie_b200_head.so python scripts/prof_train_pass.py $shape 5 2>&1 | tail -2 | tr '\n' ' ')"
  echo "new : $(PIXIE_LIB_PATH=$PWD/ark_analysis_b200/_lib/libpixie_b200_new.so python scripts/prof_train_pass.py $shape 5 2>&1 | tail -2 | tr '\n' ' ')"
done; done
