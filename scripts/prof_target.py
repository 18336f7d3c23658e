#!/usr/bin/env python3
"""Profiling target: 1M-bead bench workload; the region between cudaProfilerStart/Stop holds two ungated
rebuilds and a few MD steps.  Run under `ncu --profile-from-start off ...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
s, e = bench.prepared_engine(n, n // 100, 12345, 0, 300)
e.run(105)                 # past one unload + load event
torch.cuda.synchronize()
torch.cuda.profiler.start()
e.force_rebuild()
e.run(steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", e.stats()["last_run_gpu_ms"] / steps, "ms/step")
e.close()
