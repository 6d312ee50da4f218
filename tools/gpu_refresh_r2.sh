#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/gpu_refresh_r2.sh'): every measurement artefact of round 2 into gpurun_out/.
# Plain runs first; the ncu captures only after the same command has exited 0 without the profiler.
set -uo pipefail
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_pytest_gpu.log
python bench.py --steps 200 --warmup 20 > $O/r2_bench.json 2> $O/r2_bench.err
python bench.py --workload c2 --steps 200 --warmup 20 > $O/r2_bench_c2.json 2>> $O/r2_bench.err
python bench.py --workload c4 --steps 96 --warmup 24 > $O/r2_bench_c4.json 2>> $O/r2_bench.err
python bench.py --workload c5 --steps 40 --warmup 10 > $O/r2_bench_c5.json 2>> $O/r2_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_ref.json 2>> $O/r2_bench.err
python tools/env_throughput.py > $O/r2_env_throughput.txt 2>&1
python tools/e2e_sweep.py 1 2 4 6 > $O/r2_e2e_sweep.txt 2>&1
python tools/same_actions.py > $O/r2_same_actions.log 2>&1
python tools/block_tail.py > $O/r2_block_tail.txt 2>&1
python tools/binning_study.py > $O/r2_binning_study.txt 2>&1
python tools/contact_profile.py > $O/r2_contact_profile.txt 2>&1
T=40 python tools/validate_c2.py > $O/r2_validation.txt 2>&1
python bench.py --profile --steps 20 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 100 --csv --log-file $O/r2_launches.csv python bench.py --profile --steps 20 --warmup 3 > $O/ncu_r2_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qg_step_kernel -s 110 -c 1 -f -o $O/r2_step_kernel python bench.py --profile --steps 20 --warmup 3 > $O/ncu_r2_step.log 2>&1
python tools/profile_env_kernels.py > /dev/null 2>&1 && \
ncu --set full --clock-control none -k "regex:qg_(walk|po|reset|bin)" -s 100 -c 5 -f -o $O/r2_env_kernels python tools/profile_env_kernels.py > $O/ncu_r2_env.log 2>&1
ls -la $O | tail -30
