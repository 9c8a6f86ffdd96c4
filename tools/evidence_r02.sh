#!/bin/bash
# Round-2 evidence on one B200: GPU test suite, headline bench (+ reference arm), ncu launch list of the bench,
# one `ncu --set full` capture per hot kernel, model-level benches.  Raw outputs land in gpurun_out/r02_*;
# tools/ncu_summary.py turns the .ncu-rep files into the JSON summaries committed under profiles/.
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2>> $O/r02_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --sustained-seconds 0 --cpu-seconds 0.1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn_tc4_kernel -s 3 -c 1 -o $O/r02_knn_tc4 python bench.py --steps 2 --warmup 3 --sustained-seconds 0 --cpu-seconds 0.1 > /dev/null 2>&1
timeout 200 python tools/ab_knn_paths.py > $O/r02_ab_knn_paths.json 2>/dev/null   # one-tile-per-CTA kernel vs four-tile kernel, same layer
timeout 400 ncu --set full --clock-control none -k regex:genconv_aggregate_kernel -s 2 -c 1 -o $O/r02_aggregate_products python bench_sparse.py --products --steps 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:rowlinear_tc_kernel -s 2 -c 1 -o $O/r02_rowlinear python tools/time_sparse_block.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:dist_rows_kernel -s 1 -c 1 -o $O/r02_slab_dist_rows python tools/profile_bigk.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:select_rows_fast_kernel -s 1 -c 1 -o $O/r02_slab_select_fast python tools/profile_bigk.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches_slab_path.csv python tools/profile_bigk.py > /dev/null 2>&1
timeout 300 python bench_sparse.py --products > $O/r02_bench_sparse_products.json 2>/dev/null
timeout 300 python bench_sparse.py > $O/r02_bench_sparse_arxiv.json 2>/dev/null
timeout 300 python tools/time_sparse_block.py > $O/r02_sparse_block_pieces.json 2>/dev/null
timeout 500 python bench_models.py --which c2,c3 > $O/r02_bench_models_c2c3.json 2>> $O/r02_bench.err
timeout 300 python bench_models.py --which c4 > $O/r02_bench_models_c4.json 2>> $O/r02_bench.err
tail -2 $O/r02_tests.log; tail -c 300 $O/r02_bench.json; ls -la $O/r02_*
