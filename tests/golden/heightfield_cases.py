"""Seeded cases of the heightfield consumer shared by the golden generator (run on a GPU box, where the reference's own
device code can execute) and the tests."""
# w, h: chunk map; r: filter radius; b: biomes; kind: map structure; table_seed: generator tables; offset: noise offset
HEIGHT_CASES = [
    dict(w=24, h=16, r=4, b=6, kind="iid", seed=21, table_seed=1, offset=(0.0, 0.0), max_octave=8),
    dict(w=32, h=40, r=8, b=12, kind="blocky", seed=22, table_seed=2, offset=(1536.0, -512.0), max_octave=6),
    dict(w=48, h=48, r=16, b=5, kind="blocky", seed=23, table_seed=3, offset=(-20480.5, 777.25), max_octave=8),
    dict(w=64, h=20, r=10, b=30, kind="rare", seed=24, table_seed=4, offset=(100000.0, 250000.0), max_octave=4),
]


def make_height_case(case):
    import numpy as np

    from golden.cases import make_case
    from helpers import make_generator_tables

    m = make_case(dict(w=case["w"], h=case["h"], nn=(3, 3), b=case["b"], kind=case["kind"], seed=case["seed"]))
    table, perm, grad = make_generator_tables(case["table_seed"], case["b"], grad_size=8 + case["table_seed"],
                                              max_octave=case["max_octave"])
    return m, table, perm, grad
