mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 500 --tb=short -k "two_ranks" > gpurun_out/t_sharded2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 2 --steps 3 --warmup 2 --workload sharded-predict --size 65536 > gpurun_out/sp2_65k_v2.json 2> gpurun_out/sp2_65k_v2.err
tail -n 3 gpurun_out/t_sharded2.log; tail -c 300 gpurun_out/sp2_65k_v2.err | grep -v "^\*\|OMP\|^$"; head -c 330 gpurun_out/sp2_65k_v2.json
