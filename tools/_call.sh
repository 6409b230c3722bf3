mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "block_product or lanczos or affinity_vs_oracle") > gpurun_out/r02_gputest7.log 2>&1; echo TESTRC=$? >> gpurun_out/r02_gputest7.log
timeout 300 python tests/latency_small.py > gpurun_out/r02_latency_small_n.txt 2>&1
timeout 200 python bench.py --size 16384 --speakers 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n16384.json 2> gpurun_out/r02_bench_n16384.err
timeout 200 python bench.py --size 32768 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n32768.json 2> gpurun_out/r02_bench_n32768.err
tail -3 gpurun_out/r02_gputest7.log; cat gpurun_out/r02_latency_small_n.txt; head -c 300 gpurun_out/r02_bench_n16384.json
