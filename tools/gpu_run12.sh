set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -x -q -k "fused or golden or stage or full_size" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -3 gpurun_out/pytest_gpu_sub.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r12_c2.json 2> gpurun_out/bench_r12.err
OMNI_B200_LIB=$PWD/omnirevolve-image-processor_b200/lib/libomni_minb8.so timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r12_c2_minb8.json 2>> gpurun_out/bench_r12.err
OMNI_B200_LIB=$PWD/omnirevolve-image-processor_b200/lib/libomni_minb8.so timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_r12_c3_minb8.json 2>> gpurun_out/bench_r12.err
python - <<'PY'
import json
for f in ("bench_r12_c2","bench_r12_c2_minb8","bench_r12_c3_minb8"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        t=d.get("thinning_kernel") or {}
        print(f, d["ms_per_step"], d["step_ms"]["median"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"], t.get("ms"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_r12.err
