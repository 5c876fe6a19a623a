#!/bin/bash
# round-2 evidence, part 2: ncu --set full of every kernel of one step of configs[3] (TT estimator 4096^2, 8 realisations) and of
# configs[2] (IQU 2048^2, 16 realisations); summarised on the box (the reports exceed what gpurun copies back)
mkdir -p gpurun_out/profiles_r02b
B="python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 --no-extras"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"fused_row|fused_col_kernel|qe_finish" --launch-skip 300 --launch-count 7 -o gpurun_out/prof_r02_qechain -f $B --configs 3 > gpurun_out/ncu_r02_qechain.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_qechain.ncu-rep > gpurun_out/profiles_r02b/r02_ncu_qe_chain_summary.body 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"fused_sim_col_kernel<double, \(int\)2048, \(int\)3|fused_row_tma_kernel|fused_col_bin_kernel<double, \(int\)2048, \(int\)3" --launch-skip 20 --launch-count 3 -o gpurun_out/prof_r02_iqu3 -f $B --configs 2 > gpurun_out/ncu_r02_iqu3.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_iqu3.ncu-rep > gpurun_out/profiles_r02b/r02_ncu_iqu_kernels_summary.body 2>&1
rm -f gpurun_out/prof_r02_qechain.ncu-rep gpurun_out/prof_r02_iqu3.ncu-rep
tail -3 gpurun_out/ncu_r02_qechain.log gpurun_out/ncu_r02_iqu3.log
grep -c "====" gpurun_out/profiles_r02b/*.body
