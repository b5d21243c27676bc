"""Development aid: which forward kernel wins at which batch (small / ring / split), N = 1000 and 2000."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200"), os.path.join(ROOT, "scripts")]
from split_time import run
for N in (1000, 2000):
    for B in (148, 296, 512, 768, 888, 1024, 1536, 2048, 3552, 4096):
        for kern in ("small", "ring", "split"):
            try:
                run(B, N, kern, iters=20)
            except Exception as ex:
                print(kern, B, N, "failed:", ex)
