timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2h_pytest_gpu.txt
timeout 600 python bench.py --steps 10 > gpurun_out/r2h_bench_base.json 2> gpurun_out/r2h_bench_base.err
timeout 600 python bench.py --workload retrieval --global-batch 64 --global-queries 256 --steps 5 > gpurun_out/r2h_bench_retr.json 2> gpurun_out/r2h_bench_retr.err
timeout 300 python profiles/timeline.py 32 base > gpurun_out/r2h_timeline_base.txt 2>&1
tail -5 gpurun_out/r2h_pytest_gpu.txt; cat gpurun_out/r2h_bench_base.json | cut -c1-1500; tail -3 gpurun_out/r2h_bench_base.err; cat gpurun_out/r2h_bench_retr.json | cut -c1-1500; tail -3 gpurun_out/r2h_bench_retr.err; cat gpurun_out/r2h_timeline_base.txt
