timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tests/multigpu_sparse_check.py > gpurun_out/r2m_check8.log 2>&1; tail -3 gpurun_out/r2m_check8.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_bench_8gpu.err
tail -c 1500 gpurun_out/r2m_bench_8gpu.json
