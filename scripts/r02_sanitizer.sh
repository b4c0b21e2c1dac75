#!/bin/bash
# compute-sanitizer over the round-2 kernels: memcheck on the trace / builder / service / render tests, racecheck on the wavefront
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -q -x -k "not full_size and not large_and_degenerate and not 200000 and not config" 2>&1 | grep -v "^$" | tail -8 ) > gpurun_out/r02_sanitizer.txt 2>&1
( timeout 600 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_gpu_render.py -q -x -k "same_samples_as_oracle and cornell or coherent_camera" 2>&1 | grep -v "^$" | tail -6 ) >> gpurun_out/r02_sanitizer.txt 2>&1
( timeout 600 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_gpu_trace.py -q -x -k "known_answers or golden_vectors or edge_cases or axis_aligned or compact_wire" 2>&1 | grep -v "^$" | tail -6 ) >> gpurun_out/r02_sanitizer.txt 2>&1
# the streaming host-buffer launch (StreamGate) with 64 Ki-ray chunks: memcheck + racecheck (its per-warp bookkeeping lives in shared memory)
( LMB200_E2E_CHUNK_LOG2=16 CHECK_TRIS=20000 CHECK_N=300001,262144 timeout 500 compute-sanitizer --tool memcheck python scripts/r02_stream_check.py 2>&1 | grep -v "^$" | tail -3 ) >> gpurun_out/r02_sanitizer.txt 2>&1
( LMB200_E2E_CHUNK_LOG2=16 CHECK_TRIS=20000 CHECK_N=300001,262144 timeout 500 compute-sanitizer --tool racecheck --racecheck-report all python scripts/r02_stream_check.py 2>&1 | grep -v "^$" | tail -3 ) >> gpurun_out/r02_sanitizer.txt 2>&1
cat gpurun_out/r02_sanitizer.txt
