set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; tail -2 gpurun_out/r4_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; tail -3 gpurun_out/r4_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r4_bench_n1.json 2> gpurun_out/r4_bench_n1.err
python bench.py --steps 20 --warmup 5 --partition-sms 0 --no-cpu-baseline > gpurun_out/r4_bench_n1_seq.json 2> gpurun_out/r4_bench_n1_seq.err
python bench.py --steps 10 --warmup 3 --workload beam5 --no-cpu-baseline > gpurun_out/r4_beam5_n1.json 2> gpurun_out/r4_beam5_n1.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_step.csv python scripts/profile_step.py > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_beam.csv python scripts/dev/beam_profile.py > /dev/null 2>&1
