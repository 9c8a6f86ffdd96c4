#!/bin/bash
# Round evidence on one B200: full GPU test suite, headline bench, ncu launch list + full capture of the
# dominant kernel, model-level benches.  Outputs land in gpurun_out/ (copied to profiles/ afterwards).
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev_bench_ref.json 2>> gpurun_out/ev_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/ev_launches.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 3 -c 1 -o gpurun_out/ev_knn_tc python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
timeout 400 python bench_models.py --which c2,c3 > gpurun_out/ev_models_c2c3.json 2>> gpurun_out/ev_bench.err
timeout 300 python bench_models.py --which c4 > gpurun_out/ev_models_c4.json 2>> gpurun_out/ev_bench.err
tail -c 400 gpurun_out/ev_bench.json; tail -3 gpurun_out/ev_bench.err
