#!/bin/bash
mkdir -p gpurun_out/last
timeout 45 python -m pytest tests/test_e2e_gpu.py::test_bootea_lifecycle tests/test_gnn_gpu.py::test_alinet_lifecycle tests/test_gnn_gpu.py::test_rdgcn_lifecycle tests/test_zz_triple_ext_gpu.py::test_adadelta_steps_equal_dense_tf_steps -q -p no:cacheprovider > gpurun_out/last/tests.txt 2>&1
tail -15 gpurun_out/last/tests.txt
