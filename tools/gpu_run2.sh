set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in new; do
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_${v}_c2.json 2> gpurun_out/bench_${v}.err
timeout 300 python bench.py --no-cpu-baseline --workload config3 --steps 10 --warmup 3 > gpurun_out/bench_${v}_c3.json 2>> gpurun_out/bench_${v}.err
timeout 300 python bench.py --no-cpu-baseline --workload config4 --steps 20 --warmup 5 > gpurun_out/bench_${v}_c4.json 2>> gpurun_out/bench_${v}.err
done
OMNI_B200_ASSIGN_LABCELL=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_oldassign_c2.json 2>> gpurun_out/bench_new.err
python - <<'PY'
import json
for f in ("bench_new_c2","bench_oldassign_c2","bench_new_c3","bench_new_c4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fk_ --launch-skip 12 --launch-count 5 -o gpurun_out/prof_r1m -f python tools/profile_once.py > gpurun_out/ncu_r1m.log 2>&1
tail -3 gpurun_out/ncu_r1m.log
