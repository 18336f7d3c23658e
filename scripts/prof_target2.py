#!/usr/bin/env python3
"""Profiling target without torch: 1M-bead bench workload; the region between cudaProfilerStart/Stop holds one ungated
rebuild and STEPS MD steps starting at timestep 105 (STEPS >= 98 includes one ex_unload and one ex_load event).
Run under `ncu --profile-from-start off ...` with LE_B200_DIRECT=1.  python scripts/prof_target2.py [BEADS] [STEPS]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rt = None
for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
s, e = bench.prepared_engine(n, n // 100, 12345, 0, 300)
e.run(105)                 # past one unload + load event
if rt: rt.cudaProfilerStart()
e.force_rebuild()
e.run(steps)
if rt: rt.cudaProfilerStop()
print("ok", e.stats()["last_run_gpu_ms"] / steps, "ms/step", "cudart" if rt else "no cudart", flush=True)
e.close()
