mkdir -p gpurun_out/r2
EVAVOS_SAMPLE_STRIDE=2 TRACE_I0=150 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py cfg4 > gpurun_out/r2/c7_trace_cfg4_b.txt 2>&1
EVAVOS_SAMPLE_STRIDE=2 TRACE_I0=20 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py cfg4 > gpurun_out/r2/c7_trace_cfg4_a.txt 2>&1
tail -5 gpurun_out/r2/c7_trace_cfg4_b.txt
