set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r8_c2.json 2> gpurun_out/bench_r8.err
timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_r8_c3.json 2>> gpurun_out/bench_r8.err
timeout 300 python bench.py --no-cpu-baseline --workload config4 --steps 20 --warmup 5 > gpurun_out/bench_r8_c4.json 2>> gpurun_out/bench_r8.err
python - <<'PY'
import json
for f in ("bench_r8_c2","bench_r8_c3","bench_r8_c4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["step_ms"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"], d.get("thinning_kernel"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_r8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r8.log 2>&1
tail -2 gpurun_out/ncu_r8.log | cut -c1-300
