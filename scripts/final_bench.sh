set -x
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
(python scripts/profile_phases.py 8; python scripts/profile_phases.py 4096) > gpurun_out/phases_r02.txt 2>&1
python scripts/config_sweep.py > gpurun_out/config_sweep.log 2>&1; mv gpurun_out/configs_r01.json gpurun_out/configs_r02.json
python -c "import __graft_entry__ as g; g.smoke()"
tail -c 300 gpurun_out/bench_r02_n1.json
