python -m pytest tests/test_model_gpu.py -q -k "other_plugins or ald_rejects or probability_flow or registry" 2>&1 | tail -5
