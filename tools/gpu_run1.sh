set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_sparse.json 2> gpurun_out/bench_sparse.err; tail -c 1800 gpurun_out/bench_sparse.json
OMNI_B200_EDGE_DENSE=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_dense.json 2>> gpurun_out/bench_sparse.err
timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_sparse_c3.json 2>> gpurun_out/bench_sparse.err
OMNI_B200_EDGE_DENSE=1 timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_dense_c3.json 2>> gpurun_out/bench_sparse.err
python - <<'PY'
import json
for f in ("bench_sparse","bench_dense","bench_sparse_c3","bench_dense_c3"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
