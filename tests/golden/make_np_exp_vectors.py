#!/usr/bin/env python3
"""Golden vectors for the exp the reference's logistic uses (CROPSR.py:313): x and np.exp(x) as
evaluated by numpy in the build container (numpy 2.3.5, AVX512_SKX dispatch -> SVML __svml_exp8_ha).
Written once, committed as tests/golden/np_exp_vectors.npz; oracle/np_exp.c and the device
replica (csrc/npexp.cuh) are checked against them bit for bit wherever the tests run."""
import numpy as np

feats = np._core._multiarray_umath.__cpu_features__
assert feats.get("AVX512_SKX"), "generate on a host where numpy takes the AVX-512 SVML path"
rng = np.random.default_rng(20261018)
x = np.concatenate([rng.uniform(-18.0, 9.0, 6000),            # the range of CROPSR's pre-activation
                    rng.uniform(-700.0, 700.0, 1500), rng.normal(0.0, 1.0, 500), rng.normal(0.0, 1e-6, 100),
                    np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 6.54, -3.71, 1e-300, 5e-324, 700.0, -700.0,
                              np.log(2.0), -np.log(2.0), 0.6931471805599453 / 16])])
np.savez_compressed(__file__.replace("make_np_exp_vectors.py", "np_exp_vectors.npz"), x=x, exp=np.exp(x),
                    score=1.0 / (1.0 + np.exp(x)), numpy=np.__version__)
print(len(x), "vectors")
