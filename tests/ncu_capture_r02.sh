set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 400 $NCU -k regex:"igemm_halo|d0_" -c 10 -o gpurun_out/r02_engine_kernels python tests/profile_step.py gates > gpurun_out/ncu1.log 2>&1
timeout 400 $NCU -k regex:"slab_gates_fwd" -c 5 -o gpurun_out/r02_slab_gates_fwd python tests/profile_step.py step > gpurun_out/ncu2.log 2>&1
timeout 400 $NCU -k regex:"slab_gates_bwd" -c 5 -o gpurun_out/r02_slab_gates_bwd python tests/profile_step.py step > gpurun_out/ncu3.log 2>&1
timeout 400 $NCU -k regex:"slab_inorm" --launch-skip 6 -c 6 -o gpurun_out/r02_slab_inorm python tests/profile_step.py step > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log gpurun_out/ncu4.log
ls -la gpurun_out/*.ncu-rep
