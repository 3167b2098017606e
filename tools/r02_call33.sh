#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "concurrent_host_threads or side_stream" 2>&1 | tail -5
