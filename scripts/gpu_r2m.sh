#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
timeout 300 python scripts/tune/m2_profile.py 127/255 1 > gpurun_out/r2m_m2_b1.log 2>&1
timeout 300 python scripts/tune/m2_profile.py 256/512 64 > gpurun_out/r2m_m2_b64.log 2>&1
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 --no-cpu > gpurun_out/r2m_stream.log 2>&1
tail -3 gpurun_out/r2m_pytest.log; head -40 gpurun_out/r2m_m2_b1.log | cut -c1-200; tail -2 gpurun_out/r2m_stream.log | cut -c1-600
