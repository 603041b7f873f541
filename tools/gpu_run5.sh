set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -7 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r5_c2.json 2> gpurun_out/bench_r5.err
timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_r5_c3.json 2>> gpurun_out/bench_r5.err
timeout 300 python bench.py --no-cpu-baseline --workload config4 --steps 20 --warmup 5 > gpurun_out/bench_r5_c4.json 2>> gpurun_out/bench_r5.err
python - <<'PY'
import json
for f in ("bench_r5_c2","bench_r5_c3","bench_r5_c4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"], {k:(v["ms"],round(v["frac_of_peak"],3)) for k,v in (d.get("resize_kernel") or {}).items()})
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fk_assign|fk_hyst" --launch-skip 4 --launch-count 2 -o gpurun_out/prof_r1p -f python tools/profile_once.py > gpurun_out/ncu_r1p.log 2>&1
tail -3 gpurun_out/ncu_r1p.log
