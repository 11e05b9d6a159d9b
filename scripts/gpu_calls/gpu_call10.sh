mkdir -p gpurun_out/r2
EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py cfg4 > gpurun_out/r2/c10_trace_cfg4_a.txt 2>&1
grep -E "mean cycles|mma_ready ->|full_ready" gpurun_out/r2/c10_trace_cfg4_a.txt
EVAVOS_SAMPLE_STRIDE=2 TRACE_I0=150 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py cfg4 > gpurun_out/r2/c10_trace_cfg4_b.txt 2>&1
grep -E "mean cycles|mma_ready ->|full_ready" gpurun_out/r2/c10_trace_cfg4_b.txt
timeout 600 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2/c10_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/c10_pytest.txt
tail -30 gpurun_out/r2/c10_pytest.txt
