"""Synthetic stream cases shared by the CPU and GPU parity tests."""
from espflix_b200 import synth

# (name, kwargs) — the coverage set adds what the reference's own fixtures lack
COVERAGE = [
    ("bench12", dict(n_pictures=12, slices=12)),
    ("slices5", dict(n_pictures=12, slices=5)),
    ("slices1", dict(n_pictures=12, slices=1)),
    ("mbquant", dict(n_pictures=12, flags=synth.MBQUANT)),
    ("fullpel", dict(n_pictures=12, flags=synth.FULLPEL)),
    ("fcode3", dict(n_pictures=12, flags=synth.FCODE3)),
    ("biglevels", dict(n_pictures=6, gop=6, flags=synth.BIGLEVELS)),
    ("matrices", dict(n_pictures=12, flags=synth.MATRICES)),
    ("intra_in_p", dict(n_pictures=12, slices=5, flags=synth.INTRA_IN_P)),
    ("static", dict(n_pictures=12, flags=synth.STATIC, noise=0)),
    ("mixed_two_gops", dict(n_pictures=24, qscale=4, flags=synth.MBQUANT | synth.MATRICES | synth.INTRA_IN_P)),
    ("smooth_q6", dict(n_pictures=12, noise=0)),
]


def make(idx, kw):
    return synth.generate(synth.SEED0 + 1000 + idx, **kw)

# Outside the reference's defined domain (its clamp table is indexed out of bounds when a sample
# leaves [-256,511], which any coefficient saturated at +-2048 causes): the CUDA path is compared
# with the oracle restatement only, which clamps every integer to [0,248].
UNPINNED = [
    ("overdrive", dict(n_pictures=6, gop=6, flags=synth.BIGLEVELS | synth.OVERDRIVE)),
]
