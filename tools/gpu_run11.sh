set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r11_c2.json 2> gpurun_out/bench_r11.err
python - <<'PY'
import json
for f in ("bench_r11_c2",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        t=d.get("thinning_kernel") or {}
        print(f, d["ms_per_step"], d["step_ms"]["median"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"], t.get("ms"), t.get("iterations_max"), d.get("resize_kernel"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_r11.err
# memory checker over the small-shape tests of every kernel family (slow: sanitizer serialises)
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse_tile_runs or thinning_shapes or thinning_in_place or resize_area or resize_strided or edges_binary_masks or edges_strided or layer_masks or assign_lab and not all_colours" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -6 gpurun_out/sanitizer_memcheck.log
