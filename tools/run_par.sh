python tools/parity_report.py > gpurun_out/parity_r01.md 2> gpurun_out/parity_err.log; tail -3 gpurun_out/parity_err.log; cat gpurun_out/parity_r01.md
python -m pytest tests/test_model_gpu.py -q -k "16khz or long_form or priormix_with_network or cli" 2>&1 | tail -4
