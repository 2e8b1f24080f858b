"""Where does the end-to-end leg's time go?  Runs bench.run_e2e on one config with the diagnostic switches
(TSGU_E2E_TRACE / _WHATIF_WARM / _WHATIF_NO_D2H / _SLOTS) and prints ms per step for each combination.

    python scripts/e2e_probe.py [config] [steps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torchsparsegradutils_b200 import _pattern, sparse_mm  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = bench.CONFIGS[cfg_name]
A, B, G = bench.build_inputs(cfg, dev)
os.environ["TSGU_E2E_TRACE"] = "1"
combos = [
    ("cold, 4 slots", {}),
    ("cold, 2 slots (one item ahead)", {"TSGU_E2E_SLOTS": "2"}),
    ("cold, 6 slots", {"TSGU_E2E_SLOTS": "6"}),
    ("cold, no D2H", {"TSGU_E2E_WHATIF_NO_D2H": "1"}),
    ("warm (stale patterns)", {"TSGU_E2E_WHATIF_WARM": "1"}),
]
for name, env in combos:
    for k in ("TSGU_E2E_SLOTS", "TSGU_E2E_WHATIF_NO_D2H", "TSGU_E2E_WHATIF_WARM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    _pattern._VERIFY = "0" if "TSGU_E2E_WHATIF_WARM" in env else os.environ.get("TSGU_B200_VERIFY_PATTERN", "sampled")  # the warm what-if reuses stale patterns on purpose
    r = bench.run_e2e(A.detach(), B.detach(), G, steps, dev, None, sparse_mm)
    print(f"[{name}] ms/step {r['ms_per_step']:.3f} verified {r['results_verified']}", file=sys.stderr, flush=True)
