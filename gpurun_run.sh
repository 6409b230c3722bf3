mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for t in 1 2 4 8 16 64 512; do SCB_BLUR_TILES_PER_CTA=$t timeout 300 python tools/time_blur.py --n 65536 >> gpurun_out/blur_sweep.txt 2>&1; done
for t in 1 4 16; do SCB_BLUR_TILES_PER_CTA=$t timeout 300 python tools/time_blur.py --n 16384 >> gpurun_out/blur_sweep.txt 2>&1; done
cat gpurun_out/blur_sweep.txt
