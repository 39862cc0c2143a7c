#!/bin/bash
# Round evidence on one B200: bench lines, launch list, full ncu captures (each only after the same command
# exited 0 without ncu).  Usage: tools/evidence.sh <tag> [parts]   (writes gpurun_out/<tag>_*)
#   parts: any of "bench launches gat dense knn" (default: all).  gpurun brings back at most 64 MiB per call, and
#   the three --set full reports together exceed that: capture "gat knn" and "dense" in separate calls.
set -u
T=${1:-r01m}
P=${2:-"bench launches gat dense knn"}
O=gpurun_out
OURS='regex:knn_|normalize_|gather_sample|gatv2_|reduce_partials|reduce_columns|adapted_|domain_colsum|make_keys|flag_heads|compact_kernel|rowptr_kernel|degree_keys|spmm_csr|set_int|copy_int|zero_rowptr|rowpanel_gemm|wgrad_|bn_reduce|bn_apply|bn_fwd|bn_bwd|select_|edge_validity|tf32_planes'
has() { [[ " $P " == *" $1 "* ]]; }
if has bench; then
  python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err || exit 1
  python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
fi
if has launches || has gat || has dense; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > $O/${T}_plain.log 2>&1 || exit 1
fi
if has launches; then
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > $O/${T}_ncu_launches.log 2>&1
fi
if has gat; then
  ncu --set full --import-source on --clock-control none -k 'regex:gatv2_(heads_)?(fwd|bwd_dst|bwd_dst_hub|bwd_src)_kernel' -c 6 -o $O/${T}_prof_gat \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > $O/${T}_ncu_gat.log 2>&1
fi
if has dense; then
  ncu --set full --import-source on --clock-control none -k 'regex:rowpanel_gemm_kernel|wgrad_gemm_kernel|bn_reduce_kernel|bn_apply_kernel|adapted_skinny_(fwd|bwd)_kernel' -s 19 -c 19 -o $O/${T}_prof_dense \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-sync16m > $O/${T}_ncu_dense.log 2>&1
fi
if has knn; then
  python tools/profile_knn.py f16 262144 786432 128 20 2 > $O/${T}_knn_plain.log 2>&1 || exit 1
  ncu --set full --import-source on --clock-control none -k regex:knn_cosine_f16_kernel -s 3 -c 1 -o $O/${T}_prof_knn \
    python tools/profile_knn.py f16 262144 786432 128 20 1 > $O/${T}_ncu_knn.log 2>&1
fi
has bench && tail -1 $O/${T}_bench.json | cut -c1-400
