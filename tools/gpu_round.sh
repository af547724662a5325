#!/bin/bash
# one GPU-box visit: parity tests, bench line, ncu launch list, one full capture of the sweeps
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
CMD="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_filter2|k_pass._v3" -s 15 -c 5 -o gpurun_out/${TAG}_sweeps $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
