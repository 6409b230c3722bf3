mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L | wc -l > gpurun_out/ngpu.txt
run() { # nproc port out args...
  np=$1; port=$2; out=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
}
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 500 --tb=short > gpurun_out/t_sharded8.log 2>&1
run 8 29601 sp8_65k --steps 3 --warmup 2 --workload sharded-predict --size 65536
run 4 29602 sp4_65k --steps 3 --warmup 2 --workload sharded-predict --size 65536
run 2 29603 sp2_65k --steps 3 --warmup 2 --workload sharded-predict --size 65536
run 8 29604 sp8_131k --steps 3 --warmup 2 --workload sharded-predict --size 131072
run 4 29605 sp4_131k --steps 2 --warmup 2 --workload sharded-predict --size 131072
run 8 29606 sr8_131k --steps 2 --warmup 2 --workload sharded-refine --size 131072
run 8 29607 rep8 --steps 2 --warmup 3 --no-cpu-baseline
tail -n 4 gpurun_out/t_sharded8.log
for f in sp8_65k sp4_65k sp2_65k sp8_131k sp4_131k sr8_131k rep8; do echo "== $f"; tail -c 300 gpurun_out/$f.err | grep -v "^\*\|OMP_NUM\|^$"; head -c 260 gpurun_out/$f.json; echo; done
