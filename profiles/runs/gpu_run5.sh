timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2i_pytest_gpu.txt
(timeout 300 python profiles/probes/latency_small_batch.py base; VP_GRAPHS=0 timeout 300 python profiles/probes/latency_small_batch.py base) > gpurun_out/r2i_latency.txt 2>&1
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2i_bench_base.json 2> gpurun_out/r2i_bench_base.err
VP_GRAPHS=0 timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2i_bench_base_nograph.json 2> gpurun_out/r2i_bench_base_nograph.err
timeout 600 python bench.py --steps 10 --no-cpu-baseline --global-batch 4 > gpurun_out/r2i_bench_b4.json 2> gpurun_out/r2i_bench_b4.err
VP_GRAPHS=0 timeout 600 python bench.py --steps 10 --no-cpu-baseline --global-batch 4 > gpurun_out/r2i_bench_b4_nograph.json 2> gpurun_out/r2i_bench_b4_nograph.err
cat gpurun_out/r2i_pytest_gpu.txt gpurun_out/r2i_latency.txt; tail -2 gpurun_out/r2i_bench_*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, 'value',round(d['value'],1),'e2e',round(e['value'],1),'blocking',round(e['blocking_call']['value'],1),'u8',round(e['uint8_frames']['value'],1),'u8bf16',round(e['uint8_frames_bf16_features']['value'],1),'launches',d['gpu_launches'])
    except Exception as ex: print(f, 'ERR', ex)
PY
