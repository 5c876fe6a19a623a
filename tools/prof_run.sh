# round-1 evidence: launch list of the bench command, ncu --set full of the three pipeline kernels

ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 --batch 16 --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ --launch-skip 9 --launch-count 3 -o gpurun_out/prof_r01c -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 16 --no-extras > gpurun_out/ncu_full_c.log 2>&1
ls -la gpurun_out
