cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_spartacus.py -x -q 2>&1 | tail -40 > gpurun_out/sp_test1.log
cat gpurun_out/sp_test1.log
