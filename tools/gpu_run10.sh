set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r10_c2.json 2> gpurun_out/bench_r10.err
timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_r10_c3.json 2>> gpurun_out/bench_r10.err
python - <<'PY'
import json
for f in ("bench_r10_c2","bench_r10_c3"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        t=d.get("thinning_kernel") or {}
        print(f, d["ms_per_step"], d["step_ms"]["median"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"], t.get("ms"), t.get("iterations_max"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_r10.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fk_thin" --launch-skip 1 --launch-count 1 -o gpurun_out/prof_r1q_thin -f python tools/profile_thin.py > gpurun_out/ncu_r1q.log 2>&1
tail -4 gpurun_out/ncu_r1q.log
