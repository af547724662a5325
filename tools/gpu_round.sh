#!/bin/bash
# one GPU-box visit: parity tests, bench line, (optionally) ncu launch list + one full capture of the sweeps
#   tools/gpu_round.sh TAG [ncu]
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
NCU=${2:-}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_r02.json gpurun_out/${TAG}_parity.json 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
if [ -n "$NCU" ]; then
CMD="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --state lattice"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_filter2|k_pass._v3|k_brick" -s 15 -c 6 -o gpurun_out/${TAG}_sweeps $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
fi
tail -15 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
