# Final evidence run of a round: GPU tests, smoke, default bench (+ cpu_baseline), reference arm, the other configs,
# the ncu launch list of the bench command and one `ncu --set full` capture of the step's kernels.  Outputs: gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 400 gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; tail -c 300 gpurun_out/bench_reference_arm.json
timeout 300 python bench.py --no-cpu-baseline --workload config4 --steps 20 --warmup 5 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fk_assign_slices|fk_morph_lab|fk_edges3|fk_hyst|fk_build_tables3|fk_label_open" --launch-skip 12 --launch-count 6 -o gpurun_out/prof_final_step -f python tools/profile_once.py config5 > gpurun_out/ncu_full_step.log 2>&1
tail -2 gpurun_out/ncu_full_step.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fk_thin|fk_bytes_to_bits|fk_expand_bits" --launch-skip 3 --launch-count 3 -o gpurun_out/prof_final_thin -f python tools/profile_thin.py > gpurun_out/ncu_full_thin.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fk_resize" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_final_resize -f python tools/profile_resize.py > gpurun_out/ncu_full_resize.log 2>&1
python - <<'PY'
import json
for f in ("bench_default","bench_reference_arm","bench_config4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("roofline") or {}).get("kernels_ms_per_step"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
