#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_extra_kernels_gpu.py -q -m gpu -k "nvfp4" > gpurun_out/nvfp4_test2.log 2>&1; echo "nvfp4 tests rc=$?"; grep -E "passed|failed|Error|FAILED|assert " gpurun_out/nvfp4_test2.log | tail -12 | cut -c1-300
python - <<'PY' 2>&1 | tail -3
import torch, sys
sys.path.insert(0, ".")
from megatron_b200 import ops
x = torch.randn(8192, 8192, device="cuda").bfloat16()
for _ in range(3): ops.nvfp4_quantize(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = (x.abs().amax().float() / (6 * 448)).reshape(1)
e0.record()
for _ in range(20): ops.ext().nvfp4_quant(x, t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"nvfp4 quantise 8192x8192 bf16: {ms*1e3:.0f} us ({x.numel() * (2 + 0.5 + 1/16) / ms / 1e9:.2f} TB/s)")
PY
