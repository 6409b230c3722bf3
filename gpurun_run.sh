mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_ops.py -m gpu -q --timeout 600 --tb=short > gpurun_out/t_sharded.log 2>&1
tail -n 30 gpurun_out/t_sharded.log
