#!/bin/bash
# One gpurun call that produces every artefact scripts/summarize_profiles.py reads (tag r02): bench lines, ncu launch list,
# per-phase cycles, per-config numbers, sanitizer logs and the full-set capture of one bench-batch launch.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -c 600 gpurun_out/bench_r02_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
(python scripts/profile_phases.py 8; python scripts/profile_phases.py 4096) > gpurun_out/phases_r02.txt 2>&1
python scripts/config_sweep.py > gpurun_out/config_sweep.log 2>&1; mv gpurun_out/configs_r01.json gpurun_out/configs_r02.json
for cfg in cfg2 cfg3; do
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/prof_one.py 8 $cfg > gpurun_out/r02_racecheck_$cfg.txt 2>&1
  timeout 600 compute-sanitizer --tool memcheck python scripts/prof_one.py 8 $cfg > gpurun_out/r02_memcheck_$cfg.txt 2>&1
done
tail -3 gpurun_out/r02_racecheck_cfg2.txt gpurun_out/r02_racecheck_cfg3.txt gpurun_out/r02_memcheck_cfg2.txt gpurun_out/r02_memcheck_cfg3.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sqp_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_full python scripts/prof_one.py 4096 cfg2 distinct 2>&1 | tail -3
ls -la gpurun_out
