import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from ldpc_toolbox_b200 import codes
from ldpc_toolbox_b200.ber import BerEngine
path = codes.cached_alist_path("dvbs2:R1_2")
eng = BerEngine(path, "Minstarapproxi8")
def fr(): return round(torch.cuda.mem_get_info()[0] / 1e9, 1)
a = eng.run(1.2, 25, 0, 151552); print("lane0 151552 @0     :", a[:6].tolist(), "free GB", fr())
b = eng.run(1.2, 25, 151552, 151552); print("lane1 151552 @151552:", b[:6].tolist(), "free GB", fr())
c = eng.run(1.2, 25, 151552, 75776); eng.run(1.2, 25, 151552 + 75776, 75776, counters=c); print("split  @151552      :", c[:6].tolist(), "free GB", fr())
d = eng.run(1.2, 25, 151552, 151552); print("again 151552 @151552:", d[:6].tolist(), "free GB", fr())
