mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_scale_sp2_65k_final.json 2> gpurun_out/r02_scale_sp2_65k_final.err
echo "stdout lines: $(wc -l < gpurun_out/r02_scale_sp2_65k_final.json)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-stagewise-n 0 > gpurun_out/r02_ref_arm_2gpu.json 2> gpurun_out/r02_ref_arm_2gpu.err
echo "ref stdout lines: $(wc -l < gpurun_out/r02_ref_arm_2gpu.json)"
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:'k_symm_dmma' -c 1 -o gpurun_out/r02f_symm_dmma -f python tools/profile_step.py --n 65536 > gpurun_out/r02f_symm_dmma.log 2>&1
python tools/ncu_summary.py gpurun_out/r02f_symm_dmma.ncu-rep > gpurun_out/r02f_ncu_symm_dmma.txt 2>&1
grep -h "NCCL INFO" gpurun_out/r02_scale_sp2_65k_final.err | grep -i "nranks" | head -4
head -c 300 gpurun_out/r02_scale_sp2_65k_final.json
