timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -5 > gpurun_out/r2f_pytest_attn.txt
VP_ATTN_TRACE=1 timeout 200 python profiles/attn_trace.py 2> gpurun_out/r2f_trace.txt
for v in "2 0" "2 2" "2 4" "2 6"; do
  set -- $v
  VP_ATTN_KERNEL=$1 VP_ATTN_POLY=$2 timeout 180 python profiles/attn_bench.py 32 > gpurun_out/r2f_ab_$1_$2.txt 2>&1
done
cat gpurun_out/r2f_pytest_attn.txt; grep -h "spatial\|auxil" gpurun_out/r2f_ab_*.txt
