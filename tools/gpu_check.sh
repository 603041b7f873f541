# Quick GPU check: all GPU tests + one bench line per workload.   gpurun -- 'bash tools/gpu_check.sh'
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in config2 config3 config4 config5; do
timeout 300 python bench.py --no-cpu-baseline --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_check_$wl.json 2> gpurun_out/bench_check_$wl.err
done
python - <<'PY'
import json
for wl in ("config2","config3","config4","config5"):
    try:
        d=json.loads(open(f"gpurun_out/bench_check_{wl}.json").read().strip().splitlines()[-1])
        print(wl, round(d["value"]), d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], round(d["roofline"]["step"]["frac"],4), d["e2e"]["ms_per_step"], (d.get("thinning_kernel") or {}).get("ms"))
    except Exception as e: print(wl, "ERR", e); print(open(f"gpurun_out/bench_check_{wl}.err").read()[-800:])
PY
