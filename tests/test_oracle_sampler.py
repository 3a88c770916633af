"""CPU: the Philox restatement against the Random123 known-answer vectors, and the on-device batch builder's oracle
(`oracle/sasrec_np.py::seq_batch_build`) against the reference's per-sample semantics (trainset.py:40-75)."""
import random

import numpy as np

from oracle import philox_np as PH
from oracle import sasrec_np as O


def _kat(counter, key):
    ctr = np.array([counter[0] | (counter[1] << 32)], dtype=np.uint64)
    return [int(x) for x in PH.philox4x32_10(ctr, counter[2], key[0] | (key[1] << 32), c3=counter[3])[0]]


def test_philox4x32_10_random123_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert _kat((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_dropout_mask_rate_and_scale():
    for p in (0.1, 0.5):
        m = PH.rowwise_keep_scale(400, 512, p, seed=11, stream=3)
        kept = m != 0
        assert abs(kept.mean() - (1 - p)) < 0.01
        assert np.allclose(m[kept], 1.0 / (1.0 - np.float32(p)), rtol=1e-6)
    assert (PH.rowwise_keep_scale(5, 64, 0.0, 1, 1) == 1).all()


def _windows(g, n, W, item_num):
    padded = np.zeros((n, W), dtype=np.int64)
    for r in range(n):
        ln = int(g.integers(2, W + 1))
        padded[r, W - ln:] = g.choice(np.arange(1, item_num), size=ln, replace=False)
    return padded


def test_batch_builder_oracle_has_the_reference_sample_layout():
    g = np.random.default_rng(0)
    item_num, W = 60, 11                       # small catalog: rejections really happen
    padded = _windows(g, 40, W, item_num)
    sel = g.permutation(40)[:32]
    items, mask = O.seq_batch_build(padded, sel, item_num, seed=77)
    assert items.shape == (32, 2, W) and mask.shape == (32, W - 1) and items.dtype == np.int64
    rnd = random.Random(0)
    for b, r in enumerate(sel):
        seq = padded[r][padded[r] != 0]
        ref_items, ref_mask = O.seq_train_sample(seq, item_num, W - 1, rnd)      # trainset.py:65-75, its own RNG
        assert np.array_equal(items[b, 0], ref_items[0])                          # positives, left-padded
        assert np.array_equal(mask[b], ref_mask)                                  # one 1 per transition
        neg = items[b, 1]
        assert np.array_equal(neg != 0, ref_items[1] != 0)                        # neg[t] exists exactly where pos[t] has a predecessor
        live = neg[neg != 0]
        assert ((live >= 1) & (live < item_num)).all() and not set(live.tolist()) & set(seq.tolist())


def test_batch_builder_oracle_is_a_pure_function_of_seed_and_position():
    g = np.random.default_rng(1)
    padded = _windows(g, 10, 8, 1000)
    sel = np.arange(10)
    a = O.seq_batch_build(padded, sel, 1000, seed=5)
    b = O.seq_batch_build(padded, sel, 1000, seed=5)
    c = O.seq_batch_build(padded, sel, 1000, seed=6)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0][:, 1], c[0][:, 1]) and np.array_equal(a[0][:, 0], c[0][:, 0])
    neg = np.concatenate([O.seq_batch_build(padded, sel, 1000, seed=s)[0][:, 1].ravel() for s in range(40)])
    neg = neg[neg != 0]
    assert abs(neg.mean() - 500) < 25          # uniform on [1, item_num - 1]


def test_seq_train_sample_equals_the_reference_dataset_under_a_seeded_random():
    """A1 pin: tests/golden/seqtrain_ref.npz holds what the UNMODIFIED SEQTrainDataset.__getitem__
    (trainset.py:65-75) returned under random.seed(s) (oracle/make_golden.py::make_seqtrain_case); the oracle must
    reproduce items AND negatives bit for bit from the same stream."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "seqtrain_ref.npz"))
    n_cases = len([k for k in z.files if k.endswith("_meta")])
    assert n_cases == 3
    long_seen = False
    for ci in range(n_cases):
        item_num, L, n_seq, seed = (int(x) for x in z[f"c{ci}_meta"])
        flat, offs = z[f"c{ci}_flat"], z[f"c{ci}_offs"]
        rnd = random.Random(seed)                       # same Mersenne stream as random.seed(seed) + module functions
        for i in range(n_seq):
            seq = flat[offs[i]:offs[i + 1]]
            long_seen |= len(seq) > L + 1
            items, mask = O.seq_train_sample(seq, item_num, L, rnd)
            assert np.array_equal(items, z[f"c{ci}_items"][i]), (ci, i)
            assert np.array_equal(mask, z[f"c{ci}_mask"][i]), (ci, i)
    assert long_seen                                    # windows longer than L+1 keep their tail (trainset.py:46-50)
