O=gpurun_out
export BGNN_F16_EW=2
for pr in 0 1; do
echo "== PAIR=$pr waits dbg=8" | tee -a $O/r02v.log
BGNN_F16_PAIR=$pr BGNN_F16_DBG=8 python tools/profile_knn.py f16 37888 786432 128 20 1 2>&1 | grep -E "^cta" | grep -E "slot loads|3072 tiles" | sort | uniq | grep -E "warp 2:|warp 5:|warp 9:|issuer|producer" | head -12 | tee -a $O/r02v.log
done
