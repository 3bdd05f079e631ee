"""One-off extended fuzz on the CPU: MIP bricks of the product host octree against the oracle under random edits, random
resampling strategies and MIPs enabled before / midway (tests/test_mipmap.py with many more seeds). ~5 minutes."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
import test_mipmap as T
from test_mipmap import OracleOctree, ProductOctree, BOX, POINT, POINT_BD, POSTERIZE, POSTERIZE_BD
t0 = time.time(); n = 0
ALL = [BOX, POINT, POINT_BD, POSTERIZE, POSTERIZE_BD]
for seed in range(1000, 100000):
    rng = np.random.default_rng(seed)
    size, dim = [(8, 1), (16, 2), (32, 4), (64, 8), (16, 1), (32, 2), (8, 2), (16, 4)][seed % 8]
    methods = {int(l): (ALL[int(rng.integers(0, 5))], float(rng.choice([0.0, 0.05, 0.2, 0.5]))) for l in range(1, int(rng.integers(1, 5)))}
    simplify = bool(seed & 8)
    a, b = OracleOctree(size, dim), ProductOctree(size, dim)
    for t in (a, b):
        t.set_auto_simplify(simplify)
        if seed % 3: t.switch_albedo_mip_maps(True)
        for lvl, (m, thr) in methods.items():
            t.set_method_at(lvl, m, thr)
    for round_ in range(3):
        ops = T._random_ops(rng, size, 150)
        T._apply(a, ops); T._apply(b, ops)
        if seed % 3 == 0 and round_ == 1:
            for t in (a, b): t.switch_albedo_mip_maps(True)
        assert a.structure_hash() == b.structure_hash(), seed
        assert a.mip_hash() == b.mip_hash(), (seed, size, dim, methods, simplify, round_)
    a.recalculate_mips(); b.recalculate_mips()
    assert a.mip_hash() == b.mip_hash(), seed
    n += 1
    if time.time() - t0 > 300: break
print("mip sequences", n, "all identical", round(time.time() - t0), "s")
