#!/usr/bin/env python
"""FP64 peaks of the device: DFMA chains, DMMA m8n8k4 chains, and both side by side."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covest_b200.models import BasicModel  # noqa: E402

ctx = BasicModel(21, 100, {j: 1 for j in range(1, 33)}, 0, max_error=8).device_context
print(json.dumps({'dfma_tflops': ctx.fp64_peak(0), 'dmma_tflops': ctx.fp64_peak(1),
                  'mixed_tflops': ctx.fp64_peak(2)}))
