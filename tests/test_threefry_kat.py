"""Pins the PRNG restatement (oracle/threefry.py) against published known-answer vectors:
Random123's Threefry-2x32 (20 rounds) test vectors and the values printed in the JAX documentation
for ``jax.random.split`` / ``jax.random.uniform`` with the default (non-partitionable) Threefry.
Derived vectors further down come from the restatement itself and guard
against regressions of the layout (odd-length pad, split rows, randint multiplier)."""
import numpy as np

from oracle import rollout, threefry as tf


def test_random123_threefry2x32_kats():
    vec = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, out in vec:
        o0, o1 = tf.threefry2x32(key[0], key[1], [ctr[0]], [ctr[1]])
        assert (int(o0[0]), int(o1[0])) == out


def test_jax_docs_split_and_uniform():
    assert tf.split(tf.prng_key(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert tf.split(tf.prng_key(42)).tolist() == [[2465931498, 3679230171], [255383827, 267815257]]
    assert abs(float(tf.uniform01(tf.prng_key(0))) - 0.41845703) < 1e-8


def test_derived_vectors_regression():
    assert tf.random_bits32(tf.prng_key(0), 5).tolist() == [2467461003, 428148500, 1688610540, 3840466878, 2562233961]
    assert tf.randint(tf.prng_key(0), 20, 0, 2).tolist() == [0, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0]
    assert tf.randint(tf.prng_key(42), 20, 0, 2).tolist() == [1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 1]
    assert tf.randint(tf.prng_key(7), 7, 0, 2).tolist() == [0, 1, 1, 0, 1, 0, 1]           # odd-length pad path
    assert tf.randint(tf.prng_key(0), 8, 0, 1000).tolist() == [748, 46, 18, 904, 866, 200, 503, 278]


def test_batched_reset_draw_equals_generic_randint():
    keys = np.stack([tf.prng_key(s) for s in (0, 7, 42, 123456789)])
    for n in (1, 2, 7, 20, 35, 100, 251):
        got = tf.randint01_many(keys, n)
        for i, k in enumerate(keys):
            assert np.array_equal(got[i], tf.randint(k, n, 0, 2))


def test_rollout_chain_vector():
    k = rollout.rollout_keys(tf.prng_key(42), 4, 10)
    assert k["rng"].tolist() == [4213474292, 2545835346]
    assert k["act_key"].tolist() == [255383827, 267815257]
    assert k["step_key"].tolist() == [3923418436, 1366451097]
    assert k["prob_key"].tolist() == [3141285288, 505661898]
    assert k["reset_key"].tolist() == [3484599284, 2782933294]
    assert k["new_problem_indices"].tolist() == [6, 1, 3, 4]
    assert k["reset_keys"].tolist() == [[3906952687, 1665899524], [1667047876, 3343579328],
                                        [2484920080, 3058425514], [1690051172, 139544230]]
