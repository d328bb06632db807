python -m pytest tests/test_gpu_parity.py -x -q -k "stream_state or many_threads or error_behaviour" 2>&1 | tail -4
