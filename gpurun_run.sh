mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short > gpurun_out/t_full.log 2>&1
timeout 300 python tools/latency_small.py > gpurun_out/latency_small.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gemm_tcgen05 -o gpurun_out/prof_gemm_65k python tools/profile_step.py --n 65536 --stop-after diffuse > gpurun_out/ncu_gemm65k.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_65k_final.csv python tools/profile_step.py --n 65536 > gpurun_out/ncu_launch.log 2>&1
tail -n 6 gpurun_out/t_full.log; cat gpurun_out/latency_small.txt | tail -4; tail -c 300 gpurun_out/bench_default.err; head -c 400 gpurun_out/bench_default.json
