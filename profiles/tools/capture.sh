cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r62_pytest.log
python bench.py > gpurun_out/r62_C2.log 2>&1
python bench.py --workload C3 --no-cpu-baseline > gpurun_out/r62_C3.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r62_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v34.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r62_ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_propagate -s 1 -c 1 -o gpurun_out/prof_r1_v34 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r62_ncu2.log 2>&1
cat gpurun_out/r62_pytest.log; tail -n 1 gpurun_out/r62_C2.log | cut -c1-200
