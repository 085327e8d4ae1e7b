# round 2 (third session), call 8: single-device self-test of k_peer_allreduce (ranks played by streams)
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_zz_peer_allreduce.py -x -q > gpurun_out/r3c8_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3c8_pytest.log
tail -15 gpurun_out/r3c8_pytest.log
