set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -15 > gpurun_out/r2a_pytest_attn.txt
for v in "0 4" "2 0" "2 2" "2 3" "2 4" "2 5" "2 6"; do
  set -- $v
  VP_ATTN_KERNEL=$1 VP_ATTN_POLY=$2 timeout 180 python profiles/attn_bench.py 32 > gpurun_out/r2a_ab_$1_$2.txt 2>&1
done
VP_ATTN_KERNEL=2 VP_ATTN_POLY=4 timeout 180 python profiles/attn_bench.py 32 1.5 > gpurun_out/r2a_ab_sust_2_4.txt 2>&1
VP_ATTN_KERNEL=0 timeout 180 python profiles/attn_bench.py 32 1.5 > gpurun_out/r2a_ab_sust_0.txt 2>&1
cat gpurun_out/r2a_*.txt
