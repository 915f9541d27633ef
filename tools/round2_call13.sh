#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== eager host cost per launch"; timeout -s KILL 300 python - <<'PY'
import time, torch
from paroquant_b200.checkpoint import make_synthetic_activations, make_synthetic_layer
from paroquant_b200.linear import ParoLinearKernel
k = ParoLinearKernel.from_buffers(make_synthetic_layer(4096, [4096], seed=1, device="cuda"), torch.bfloat16, check_pairs=False)
x = make_synthetic_activations(1, 4096, seed=2, device="cuda")
y = torch.empty(1, 4096, dtype=torch.bfloat16, device="cuda")
for name, fn in (("ParoLinearKernel.__call__ (torch.ops.paro.linear)", lambda: k(x)), ("forward_into (ctypes -> C-ABI)", lambda: k.forward_into(x, y))):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name}: {(t1 - t0) / 2000 * 1e6:.1f} us host time per eager launch (GPU kernel ~7 us)")
PY
echo "== done"
