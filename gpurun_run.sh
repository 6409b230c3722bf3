mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --tb=short > gpurun_out/t_full.log 2>&1
tail -n 4 gpurun_out/t_full.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 180 -x -k "fused_chain_equals_operator_chain or (diffuse_vs_oracle and 513) or (affinity_vs_oracle and 200) or (gaussian_blur_vs_scipy and 515) or (kmeans_vs_oracle and 450) or (lanczos_matches_dense and rownorm)" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -n 8 gpurun_out/sanitizer_memcheck.log
