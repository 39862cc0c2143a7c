python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -5 gpurun_out/r02n_pytest.log | cut -c1-600
bash tools/evidence.sh r02n "launches gat"
ls -la gpurun_out/r02n_*
