mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "affinity or diffuse or fused_chain") > gpurun_out/r02_gemm2cta_tests.log 2>&1; echo TESTRC=$? >> gpurun_out/r02_gemm2cta_tests.log
for p in single split2 split3; do for c in 1 0; do SCB_GEMM_2CTA=$c timeout 120 python tools/time_diffuse.py --n 65536 --precision $p; done; done > gpurun_out/r02_gemm2cta_timing.txt 2>&1
for c in 1 0; do SCB_GEMM_2CTA=$c timeout 120 python tools/time_diffuse.py --n 16384 --precision single --iters 10; done >> gpurun_out/r02_gemm2cta_timing.txt 2>&1
(time timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_fullsize.py tests/test_gpu_predict.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py -m gpu -q --durations=10 -k "not 16384-graphcut and not 8192") > gpurun_out/r02_gputest3.log 2>&1; echo TESTRC=$? >> gpurun_out/r02_gputest3.log
SCB_LANCZOS_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
bash tools/sanitize.sh r02
tail -3 gpurun_out/r02_gemm2cta_tests.log; cat gpurun_out/r02_gemm2cta_timing.txt; tail -5 gpurun_out/r02_gputest3.log
