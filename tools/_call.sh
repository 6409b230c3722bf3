mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_predict.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py tests/test_gpu_parity_fullsize.py tests/test_gpu_nonsymmetric.py -m gpu -q --durations=5 -k "not 16384-graphcut and not 8192") > gpurun_out/r02_gputest5.log 2>&1; echo TESTRC=$? >> gpurun_out/r02_gputest5.log
o=gpurun_out/r02_ab_stages3.txt; : > $o
for rep in 1 2; do
  timeout 200 python tools/time_stages.py --tag "symm-dmma(rep$rep)" >> $o 2>&1
  SCB_SYMM_V2=0 timeout 200 python tools/time_stages.py --tag "symm-dfma(rep$rep)" >> $o 2>&1
done
SCB_BLUR_TILES_PER_CTA=1 timeout 200 python tools/time_stages.py --tag "blur-tpc1" >> $o 2>&1
timeout 200 python tools/time_stages.py --n 16384 --tag "n16384 symm-dmma" >> $o 2>&1
SCB_SYMM_V2=0 timeout 200 python tools/time_stages.py --n 16384 --tag "n16384 symm-dfma" >> $o 2>&1
tail -4 gpurun_out/r02_gputest5.log; cat $o
