mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 --tb=short > gpurun_out/t_sharded1.log 2>&1
timeout 600 python bench.py --workload sharded-predict --n 65536 --steps 2 --warmup 2 > gpurun_out/sp1.json 2> gpurun_out/sp1.err
tail -n 12 gpurun_out/t_sharded1.log; tail -c 600 gpurun_out/sp1.err; head -c 700 gpurun_out/sp1.json
