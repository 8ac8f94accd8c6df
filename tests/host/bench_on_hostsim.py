"""TEST INFRASTRUCTURE ONLY.  Runs bench.py's PRODUCT arm against the host-sim build of the library (D3Q19_LIB) so that
the arm's own logic -- argument handling, warm-up / timed region, counters, roofline / e2e / cpu_baseline objects, the one
JSON line -- is exercised on the GPU-less box.  torch's CUDA queries are stubbed for this process only; every number
in the line is meaningless (a CPU is stepping the lattice) and the test that calls this looks at structure alone."""
import os
import runpy
import sys

import torch

os.environ["D3Q19_HOSTSIM_BENCH_STRUCTURE_TEST"] = "1"      # bench.py refuses the host-sim build otherwise
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
