#!/usr/bin/env python
"""Latency of the drop-in call every reference caller makes: VoxelGenerator.generate(points) for ONE
nuScenes 10-sweep frame, numpy in -> numpy out (H2D, kernels, D2H of the padded voxels tensor and a
synchronisation inside the timed region), next to the CPU oracle port of the reference's numba path."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (CPU comparison only)
from partner_b200 import VoxelGenerator, synth  # noqa: E402

g = synth.GRIDS["NUSC-PILLAR"]
polar = oracle.transform_points(synth.nusc_frame(2000))
vg = VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
for _ in range(5):
    out = vg.generate(polar)
reps = 30
t0 = time.perf_counter()
for _ in range(reps):
    out = vg.generate(polar)
ours = (time.perf_counter() - t0) / reps * 1e3
ref = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
ref.generate(polar)
t0 = time.perf_counter()
for _ in range(3):
    ref.generate(polar)
cpu = (time.perf_counter() - t0) / 3 * 1e3
print(json.dumps({"workload": "VoxelGenerator.generate, one nuScenes 10-sweep frame, numpy -> numpy",
                  "points": int(polar.shape[0]), "voxels": int(out[0].shape[0]), "ms_gpu_drop_in": ours,
                  "ms_cpu_oracle_port": cpu, "voxels_tensor_MB": out[0].nbytes / 1e6}))
