mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --tb=short > gpurun_out/t_full.log 2>&1
timeout 900 python bench.py --n 65536 --steps 2 --warmup 3 > gpurun_out/b65k.json 2> gpurun_out/b65k.err
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_blur -o gpurun_out/prof_blur_16k python tools/profile_step.py --n 16384 --stop-after diffuse > gpurun_out/ncu_blur.log 2>&1
tail -n 8 gpurun_out/t_full.log
tail -c 600 gpurun_out/b65k.err
