mkdir -p gpurun_out
(time timeout 1100 python -m pytest tests -m gpu -x -q --durations=10) > gpurun_out/r02_gputest_final.log 2>&1; echo TESTRC=$? >> gpurun_out/r02_gputest_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
bash tools/gpu_profile_pack.sh r02f 65536
tail -4 gpurun_out/r02_gputest_final.log; cat gpurun_out/r02_smoke.txt | tail -2; head -c 600 gpurun_out/r02_bench_final.json
