O=gpurun_out
export BGNN_F16_EW=2 BGNN_F16_PAIR=0
python tools/profile_knn.py f16 37888 786432 128 20 1 > $O/r02w_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:knn_cosine_f16_kernel -s 1 -c 1 -o $O/r02w_prof_knn \
    python tools/profile_knn.py f16 37888 786432 128 20 1 > $O/r02w_ncu.log 2>&1
tail -2 $O/r02w_ncu.log
