mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for kb in 8 16 32 128 256; do SCB_GEMM_PACE_KB=$kb timeout 300 python tools/time_diffuse.py --n 65536 --iters 2 2>&1 | sed "s/^/pace_kb=$kb /" >> gpurun_out/pace_sweep.txt; done
cat gpurun_out/pace_sweep.txt
