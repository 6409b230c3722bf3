mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --tb=short > gpurun_out/t_full.log 2>&1
timeout 600 python bench.py --n 16384 --steps 2 --warmup 3 > gpurun_out/b16k.json 2> gpurun_out/b16k.err
timeout 900 python bench.py --n 65536 --steps 1 --warmup 3 > gpurun_out/b65k.json 2> gpurun_out/b65k.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_16k.csv python tools/profile_step.py --n 16384 > gpurun_out/ncu16k.log 2>&1
tail -n 5 gpurun_out/t_full.log
tail -c 600 gpurun_out/b16k.err gpurun_out/b65k.err
