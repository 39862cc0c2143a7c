#!/bin/bash
# kNN f16 sweep with / without threshold seeding: kernel durations from ncu, then plain timings.
for s in 0 1; do
  echo "== BGNN_F16_NOSEED=$s"
  BGNN_F16_NOSEED=$s ncu --metrics gpu__time_duration.sum -k regex:"knn_cosine_f16|knn_merge|knn_simt|seed|gather_sample" --clock-control none -c 12 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^\s+(void )?(bgnn::)?[a-z_0-9]+(<[0-9, ]+>)?\(|gpu__time|algo=" | sed -E 's/\(.*//' | paste - - | tail -8
  BGNN_F16_NOSEED=$s python tools/profile_knn.py f16 262144 786432 128 20 3
done
