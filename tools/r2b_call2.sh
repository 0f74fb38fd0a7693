#!/bin/bash
# call 2: what binds the halo kernel?  flag sweep (timing experiments) + ncu --set full of the four B=64 shapes
mkdir -p gpurun_out
timeout 300 python tools/halo_bench.py 10 0,4,1,2,8,11,15 > gpurun_out/b2_halo_sweep.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/b2_halo_sweep.log | cut -c1-250
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -c 8 -o gpurun_out/b2_halo python tools/halo_bench.py 1 0,4 ncu > gpurun_out/b2_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 5 gpurun_out/b2_ncu.log | cut -c1-200
ls -la gpurun_out/
