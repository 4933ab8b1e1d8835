# NOTE: this is the call that spent the rest of the round's GPU budget: the retrieval bench of that commit ran two extra steps WITH the
# all-gather on rank 0 alone (trace leg), NCCL waited 10 minutes per run, and 8 GPUs are charged 8x.  Fixed in bench.py (the trace
# leg runs the forward without the collective; 4-minute NCCL timeout); the first three bench lines below are valid measurements.
export MASTER_ADDR=127.0.0.1
timeout 600 python -m pytest tests/test_retrieval_nccl_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "nccl or two_devices" -s 2>&1 | tail -12 > gpurun_out/r2j_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2j_bench_base_8gpu.json 2> gpurun_out/r2j_bench_base_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --model large > gpurun_out/r2j_bench_large_8gpu.json 2> gpurun_out/r2j_bench_large_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload retrieval --global-batch 256 --global-queries 1024 --steps 5 > gpurun_out/r2j_bench_lvt_base_8gpu.json 2> gpurun_out/r2j_bench_lvt_base_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload retrieval --model large --global-batch 256 --global-queries 1024 --steps 5 > gpurun_out/r2j_bench_lvt_large_8gpu.json 2> gpurun_out/r2j_bench_lvt_large_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2j_bench_base_4gpu.json 2> gpurun_out/r2j_bench_base_4gpu.err
cat gpurun_out/r2j_pytest_multi.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(e['value'],1), {k:round(v['value'],1) for k,v in e.items() if isinstance(v,dict)}, 'clocks',d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as ex: print(f, 'ERR', ex)
PY
for f in gpurun_out/r2j_bench_*.err; do echo $f; tail -n 3 $f; done
