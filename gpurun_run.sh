mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gemm_tcgen05 -o gpurun_out/prof_gemm_65k_paced python tools/profile_step.py --n 65536 --stop-after diffuse > gpurun_out/ncu_gemm65k.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_65k_final.csv python tools/profile_step.py --n 65536 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_gemm65k.log
