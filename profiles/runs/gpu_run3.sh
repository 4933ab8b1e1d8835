timeout 900 python -m pytest tests/test_check_fp32_gpu.py -x -q -s 2>&1 | tail -40 > gpurun_out/r2g_check_fp32.txt
VP_AB_SWEEP=1 VP_ATTN_KERNEL=2 timeout 300 python profiles/attn_bench.py 32 2>&1 | grep sweep > gpurun_out/r2g_sweep_new.txt
VP_AB_SWEEP=1 VP_ATTN_KERNEL=0 timeout 300 python profiles/attn_bench.py 32 2>&1 | grep sweep > gpurun_out/r2g_sweep_old.txt
cat gpurun_out/r2g_check_fp32.txt gpurun_out/r2g_sweep_new.txt gpurun_out/r2g_sweep_old.txt
