set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r9_c2.json 2> gpurun_out/bench_r9.err
timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_r9_c3.json 2>> gpurun_out/bench_r9.err
OMNI_B200_LIB=$PWD/omnirevolve-image-processor_b200/lib/libomni_tr64.so timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r9_c2_tr64.json 2>> gpurun_out/bench_r9.err
OMNI_B200_LIB=$PWD/omnirevolve-image-processor_b200/lib/libomni_tr64.so timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_r9_c3_tr64.json 2>> gpurun_out/bench_r9.err
OMNI_B200_LIB=$PWD/omnirevolve-image-processor_b200/lib/libomni_tr64.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or golden or sparse" > gpurun_out/pytest_tr64.log 2>&1; tail -2 gpurun_out/pytest_tr64.log
python - <<'PY'
import json
for f in ("bench_r9_c2","bench_r9_c2_tr64","bench_r9_c3","bench_r9_c3_tr64"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        t=d.get("thinning_kernel") or {}
        print(f, d["ms_per_step"], d["step_ms"]["median"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"], t.get("ms"), t.get("iterations_max"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_r9.err
