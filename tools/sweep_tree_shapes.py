"""One-off extended fuzz on the CPU: random insert / insert_at_lod / update / clear sequences (tests/test_host_octree_shape.py)
with many more seeds than the suite runs; the product host octree must build the oracle's tree every time. ~7 minutes."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
import test_host_octree_shape as T
t0 = time.time(); n = 0
for seed in range(100, 100000):
    size, dim = [(8, 1), (8, 2), (16, 2), (16, 4), (32, 4), (32, 8), (64, 8), (16, 1), (32, 2), (64, 16), (4, 1), (8, 4)][seed % 12]
    T.test_random_edit_sequences_have_identical_shape(size, dim, seed)
    n += 1
    if time.time() - t0 > 420: break
print("sequences", n, "all identical", round(time.time() - t0), "s")
