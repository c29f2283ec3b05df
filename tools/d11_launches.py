import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonrecon_gpu_b200 import PoissonRecon, synth
p, n, D = synth.make(sys.argv[1] if len(sys.argv) > 1 else "multi20m_d11")
pr = PoissonRecon(D)
for _ in range(2):
    pr.set_points(p, n); pr.run()
print(pr.stats())
