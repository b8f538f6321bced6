mkdir -p gpurun_out/t1g
timeout 300 python tools/e2e_chunks.py > gpurun_out/t1g/chunks.log 2>&1; cat gpurun_out/t1g/chunks.log
