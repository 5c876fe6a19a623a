ncu --set full --clock-control none --import-source on -k regex:"fused_col_kernel|fused_row_kernel|qe_finish" --launch-skip 24 --launch-count 8 -o gpurun_out/prof_qe -f python tools/bench_qe.py 4096 8 f64 TT > gpurun_out/ncu_qe.log 2>&1
tail -2 gpurun_out/ncu_qe.log | cut -c1-200; ls -la gpurun_out
