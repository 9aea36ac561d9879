"""Randomised parity sweep: CUDA path vs the CPU oracle on generated scenes (tools, not a test: minutes).

    python tools/fuzz_parity.py [first_seed] [n_seeds]

Two generators: scenes.build_fuzz (every blend mode, overhanging layers, anisotropic scale) and an
occlusion / whole-pixel mix (opaque NORMAL layers at opacity 1, integer and fractional shifts, layers larger
than the frame, translucent and opaque backgrounds) that aims at the one-tap and hidden-layer paths."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv  # noqa: E402
from movis_b200 import scenes  # noqa: E402
from oracle import scene_eval  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from scene_zoo import occlusion_fuzz  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40


bad = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed + 7)
    cases = [("fuzz", scenes.build_fuzz(mv, seed, size=(int(rng.integers(20, 300)), int(rng.integers(20, 200))),
                                        n_layers=int(rng.integers(1, 36)))),
             ("occl", occlusion_fuzz(mv, seed))]
    for tag, comp in cases:
        for t in (0.0, float(rng.uniform(0, 2.0)), 1.0):
            for bg in ((0, 0, 0, 0), (int(rng.integers(0, 256)), 9, 200, 255)):
                for level in (1, 2):
                    if min(comp.size) < 2 * level:
                        continue
                    with comp.preview(level=level):
                        got = comp(t, bg_color=bg)
                    want = scene_eval.render(comp, t, bg, preview_level=level)
                    if got.shape != want.shape or not np.array_equal(got, want):
                        bad += 1
                        d = np.abs(got.astype(int) - want.astype(int)) if got.shape == want.shape else None
                        print(f"MISMATCH {tag} seed {seed} t {t} bg {bg} L {level}: "
                              f"{'shape' if d is None else f'max {d.max()}, {int((d > 0).sum())} values'}")
print(f"seeds {first}..{first + count - 1}: {bad} mismatching frames")
