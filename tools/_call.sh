mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_predict.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py tests/test_gpu_callers.py tests/test_gpu_constraints.py -m gpu -q --durations=5) > gpurun_out/r02_gputest6.log 2>&1; echo TESTRC=$? >> gpurun_out/r02_gputest6.log
o=gpurun_out/r02_ab_stages4.txt; : > $o
for rep in 1 2; do
  timeout 200 python tools/time_stages.py --tag "default(rep$rep)" >> $o 2>&1
  SCB_AFFINITY_CHUNK_KB=1 timeout 200 python tools/time_stages.py --tag "aff-chain1(rep$rep)" >> $o 2>&1
  SCB_SYMM_STAGES=2 timeout 200 python tools/time_stages.py --tag "symm-2stage-2cta(rep$rep)" >> $o 2>&1
done
tail -4 gpurun_out/r02_gputest6.log; cat $o
