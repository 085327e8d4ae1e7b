# round 2 (third session), call 9: the whole GPU suite on the final library
mkdir -p gpurun_out
timeout 75 python -m pytest tests -m gpu -x -q > gpurun_out/r3c9_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3c9_pytest.log
tail -5 gpurun_out/r3c9_pytest.log
