#!/bin/bash
# round-2 evidence: launch list of the bench command, ncu --set full of the three pipeline kernels (batch 64, as timed) and of the
# dominant kernel of each other configuration at the bench's own batch sizes.  Read with tools/make_profiles_r02.py.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $B --configs none > gpurun_out/ncu_r02_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fused_(sim_col|row_tma|col_bin)_kernel" --launch-skip 9 --launch-count 3 -o gpurun_out/prof_r02_T -f $B --configs none > gpurun_out/ncu_r02_T.log 2>&1
# configs[2] IQU: K_A with three components dominates
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"fused_sim_col_kernel<double, \(int\)2048, \(int\)3" --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02_IQU -f $B --configs 2 > gpurun_out/ncu_r02_IQU.log 2>&1
# configs[3] TT 4096^2: Q3a rows c2r (persistent TMA kernel, 2-row tiles)
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"fused_row_tma_kernel<double, \(int\)2048, \(int\)2, \(int\)3>" --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02_TT -f $B --configs 3 > gpurun_out/ncu_r02_TT.log 2>&1
# configs[4] EB 8192^2: Q3a rows c2r (fp64), Q2b legs columns (fp32)
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"fused_row_kernel<double, \(int\)4096, \(int\)[0-9]+, \(int\)3>" --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02_EB64 -f $B --configs 4 > gpurun_out/ncu_r02_EB64.log 2>&1
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"fused_col_kernel<float, \(int\)8192, \(int\)1," --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02_EB32 -f $B --configs 4 > gpurun_out/ncu_r02_EB32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 120 --csv --log-file gpurun_out/launches_r02_qe.csv $B --configs 3 > gpurun_out/ncu_r02_launches_qe.log 2>&1
# the reports together exceed what gpurun copies back (64 MiB): summarise them here, keep only the pipeline capture
python tools/make_profiles_r02.py gpurun_out/profiles_r02 > gpurun_out/make_profiles_r02.log 2>&1
rm -f gpurun_out/prof_r02_IQU.ncu-rep gpurun_out/prof_r02_TT.ncu-rep gpurun_out/prof_r02_EB64.ncu-rep gpurun_out/prof_r02_EB32.ncu-rep
ls -la gpurun_out gpurun_out/profiles_r02 | tail -24
