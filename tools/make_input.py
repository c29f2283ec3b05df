#!/usr/bin/env python3
"""Write a synthetic config as binary PLY / bnpts: make_input.py <config> <out> [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonrecon_gpu_b200 import plyio, synth
cfg, out = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else None
p, nr, d = synth.make(cfg, n)
(plyio.write_bnpts if out.endswith('.bnpts') else plyio.write_points_ply)(out, p, nr)
print(cfg, p.shape[0], 'points depth', d, '->', out)
