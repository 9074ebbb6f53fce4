#!/bin/bash
# Evidence run on ONE B200 (driven through gpurun): GPU parity tests, smoke, both bench arms, the ncu launch
# list of the bench command, one `--set full` capture of a C3 step, and the other single-GPU configs.
# Usage (from the repo root):  tools/record_run.sh <tag>      outputs land in gpurun_out/*_<tag>.*
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_${tag}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -f --set full --clock-control none --import-source on --profile-from-start off -c 3 \
    -o gpurun_out/prof_${tag}_step python tools/profile_c3.py > gpurun_out/ncu_full_${tag}.log 2>&1
# gpurun copies back at most 64 MiB: keep the raw metric page, drop the report
ncu -i gpurun_out/prof_${tag}_step.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_step_raw.csv 2>/dev/null
rm -f gpurun_out/prof_${tag}_step.ncu-rep
python tools/bench_configs.py c1 c2 c4 c5 > gpurun_out/configs_${tag}.jsonl 2>> gpurun_out/bench_${tag}.err
# one SETDRK4 step of C5 (512^3, one GPU): 20 launches of the 3-D passes
FSM_NCU_WINDOW=1 timeout 600 ncu -f --set full --clock-control none --import-source on --profile-from-start off -c 20 \
    -o gpurun_out/prof_${tag}_c5 python tools/bench_configs.py c5 > gpurun_out/ncu_c5_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}_c5.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_c5_raw.csv 2>/dev/null
rm -f gpurun_out/prof_${tag}_c5.ncu-rep
tail -2 gpurun_out/pytest_gpu_${tag}.log; cat gpurun_out/smoke_${tag}.log | tail -1; cut -c1-300 gpurun_out/bench_${tag}.json
