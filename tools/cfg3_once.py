import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rogtk_b200
from rogtk_b200 import synth
n3 = 4000
rp = synth.skewed_reads_per_umi(synth.config_seed(3), n3)
data, offs, goff = synth.synth_reads(synth.config_seed(3), 0, n3, rp, 150, 500)
eng = rogtk_b200.Engine()
for _ in range(2):
    g = eng.count_prune(data, offs, goff, 31, 2)
print(g.ms_total, g.n_spilled_groups)
