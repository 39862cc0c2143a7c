#!/bin/bash
# Who waits for whom in knn_cosine_f16_kernel: BGNN_F16_DBG bit3 prints the wait cycles of the TMA producer and
# the MMA issuer of CTAs 0/1 (first call only is interesting).  $1 = pair (0/1), $2 = extra dbg bits.
export BGNN_F16_PAIR=${1:-1}
export BGNN_F16_DBG=$((8 + ${2:-0}))
python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta|algo=" | grep -E "3072 slot|3072 tiles" | sort | uniq | head -14
