#!/bin/bash
# Final evidence run of a round on one B200: bench lines of every BASELINE config, the reference arm, the ncu launch list of
# one steady-state env step and a full-set capture of one substep's kernels.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_final.sh TAG'
TAG=${1:-r2}
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
python tools/exp_reset_trace.py > gpurun_out/reset_trace_$TAG.log 2>&1; grep -a 'full reset\|masked' gpurun_out/reset_trace_$TAG.log
python tools/exp_usage.py > gpurun_out/usage_$TAG.log 2>&1; tail -12 gpurun_out/usage_$TAG.log
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ref_$TAG.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --jump-frac 0.05 2>&1 | tail -1 > gpurun_out/bench_jump_$TAG.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --preroll 0 2>&1 | tail -1 > gpurun_out/bench_nopreroll_$TAG.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --envs-per-gpu 8192 2>&1 | tail -1 > gpurun_out/bench_8192_$TAG.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --env UR5Reach-v0 --envs-per-gpu 4096 2>&1 | tail -1 > gpurun_out/bench_reach4096_$TAG.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --env pandaPick-v0 --envs-per-gpu 16384 2>&1 | tail -1 > gpurun_out/bench_pick16384_$TAG.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*$TAG.json")):
    try:
        d = json.load(open(f)); r = d.get("roofline", {})
        print(f, "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "setup %.2f pgs %.2f" % (r.get("setup_kernels_ms", 0), r.get("pgs_kernels_ms", 0)),
              "reset", d.get("reset", {}).get("full_batch_ms"), d.get("reset", {}).get("rounds"), "ovf", d.get("capacity_overflow_env_steps"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"prb_" -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/b_launch_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | tee gpurun_out/launch_summary_$TAG.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"prb_setup|prb_pgs" -s 16 -c 8 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/b_ncu_$TAG.log 2>&1
tail -c 200 gpurun_out/b_ncu_$TAG.log
