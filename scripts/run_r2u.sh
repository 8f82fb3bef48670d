# compact sharded upload (int32 gid, row-pointer source row): exactness inside the sharded worker (gloo ranks on cuda:0)
timeout 120 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "mlp_scorer" > gpurun_out/r2u_tests.log 2>&1; tail -15 gpurun_out/r2u_tests.log
