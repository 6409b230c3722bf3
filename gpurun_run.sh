mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --size 16384 > gpurun_out/rep2_16k.json 2> gpurun_out/rep2_16k.err; echo "rc=$?" >> gpurun_out/rep2_16k.err
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --workload sharded-refine --size 65536 > gpurun_out/sh2.json 2> gpurun_out/sh2.err
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --workload sharded-refine --size 131072 > gpurun_out/sh2_131k.json 2> gpurun_out/sh2_131k.err
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/rep2.json 2> gpurun_out/rep2.err; echo "rc=$?" >> gpurun_out/rep2.err
tail -c 1500 gpurun_out/rep2_16k.err gpurun_out/sh2.err gpurun_out/sh2_131k.err gpurun_out/rep2.err
