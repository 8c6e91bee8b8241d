"""Index plumbing of the label-sorted key bank on CPU tensors: the torch specification the CUDA layout kernel is
tested against (tests/bank_layout_spec.py) and the sidecar round trip."""
import numpy as np
import pytest
import torch

from bank_layout_spec import hard_bank_layout_spec


@pytest.mark.parametrize("n_keys,n_classes,seed", [(1, 1, 0), (17, 3, 1), (1000, 37, 2), (5000, 1000, 3), (300, 5, 4)])
def test_layout_invariants(n_keys, n_classes, seed):
    from summer_clip_b200 import ops
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(-1, n_classes + 1, (n_keys,), generator=g).int()      # -1 and n_classes are invalid
    bank = hard_bank_layout_spec(labels, n_classes)
    perm, gcls = bank.perm.numpy(), bank.gcls.numpy()
    bits = bank.kbits.numpy().astype(np.int64) & 0xFFFFFFFF
    valid = (labels >= 0) & (labels < n_classes)
    assert perm.size % 256 == 0 and gcls.size == perm.size // 16 and bits.size == perm.size // 32
    assert bank.n_sorted % 16 == 0 and bank.n_sorted <= perm.size and bank.n_keys == n_keys
    real = perm[perm >= 0]
    assert sorted(real.tolist()) == torch.nonzero(valid).flatten().tolist()       # every valid key exactly once
    lab = labels.numpy()
    for gi in range(gcls.size):                                                    # groups are single-class
        members = perm[16 * gi: 16 * gi + 16]
        members = members[members >= 0]
        if gcls[gi] < 0:
            assert members.size == 0
        else:
            assert members.size > 0 and np.all(lab[members] == gcls[gi])
    seen = gcls[gcls >= 0]
    assert np.all(np.diff(seen) >= 0)                                              # classes ascending: one run each
    for c in np.unique(seen):                                                      # stable within a class
        mem = perm[np.repeat(gcls == c, 16)]
        mem = mem[mem >= 0]
        assert np.all(np.diff(mem) > 0)
    unpacked = ((bits[:, None] >> np.arange(32)[None, :]) & 1).reshape(-1).astype(bool)
    assert np.array_equal(unpacked, perm >= 0)                                     # bit j of word w = key 32 w + j
    assert np.all(perm[bank.n_sorted:] == -1)


def test_sidecar_round_trip(tmp_path):
    """bank_io: a gathered bank survives save -> load bit for bit; a sidecar built for another key is refused."""
    from summer_clip_b200 import bank_io, ops
    g = torch.Generator().manual_seed(9)
    labels = torch.randint(0, 7, (300,), generator=g).int()
    bank = hard_bank_layout_spec(labels, 7)
    rows = torch.randn(300, 64, generator=g).half()
    src = bank.perm.clamp_min(0)
    bank.rows = rows[src]
    bank.rows[bank.perm < 0] = 0
    feats = tmp_path / "features.pt"
    torch.save(rows, feats)
    key = bank_io.bank_key([feats], 7, torch.float16, idx=torch.arange(300))
    assert key != bank_io.bank_key([feats], 8, torch.float16, idx=torch.arange(300))
    assert key != bank_io.bank_key([feats], 7, torch.float16, idx=torch.arange(299))
    d = bank_io.save_hard_bank(bank, tmp_path / "bank", key)
    back = bank_io.load_hard_bank(d, "cpu", key)
    assert back is not None and back.n_sorted == bank.n_sorted and back.n_keys == 300 and back.n_classes == 7
    for a, b in ((back.rows, bank.rows), (back.perm, bank.perm), (back.gcls, bank.gcls), (back.kbits, bank.kbits)):
        assert a.dtype == b.dtype and torch.equal(a, b)
    assert bank_io.load_hard_bank(d, "cpu", "another-key") is None
    assert bank_io.load_hard_bank(tmp_path / "missing", "cpu") is None
    bank_io.save_hard_bank(bank, tmp_path / "bank", key)                      # overwrite in place
    assert bank_io.load_hard_bank(d, "cpu", key) is not None
    # the opt-in 8-bit operand type travels the same way (rows are stored as raw 16-bit words)
    bank.rows = (bank.rows.float() * 256).to(torch.float8_e4m3fn)
    key8 = bank_io.bank_key([feats], 7, torch.float8_e4m3fn, idx=torch.arange(300))
    assert key8 != key
    d8 = bank_io.save_hard_bank(bank, tmp_path / "bank8", key8)
    back8 = bank_io.load_hard_bank(d8, "cpu", key8)
    assert back8 is not None and back8.rows.dtype == torch.float8_e4m3fn and back8.rows.shape == bank.rows.shape
    assert torch.equal(back8.rows.view(torch.uint8), bank.rows.view(torch.uint8))


def test_dense_and_query_sidecars_round_trip(tmp_path):
    """bank_io: dense-value caches (keys + transposed values) and query banks (rows + zero-shot logits) survive
    save -> load bit for bit; kinds and keys are not confused."""
    from summer_clip_b200 import bank_io
    g = torch.Generator().manual_seed(10)
    k, vt = torch.randn(300, 64, generator=g).half(), torch.rand(48, 304, generator=g).half()
    d = bank_io.save_dense_bank(k, vt, 300, 40, tmp_path / "dense", "k1")
    back = bank_io.load_dense_bank(d, "cpu", "k1")
    assert back is not None and back[2:] == (300, 40)
    assert back[0].dtype == torch.float16 and torch.equal(back[0], k) and torch.equal(back[1], vt)
    assert bank_io.load_dense_bank(d, "cpu", "k2") is None and bank_io.load_hard_bank(d, "cpu", "k1") is None
    q, z = torch.randn(77, 64, generator=g).bfloat16(), torch.randn(77, 40, generator=g)
    dq = bank_io.save_query_bank(q, tmp_path / "queries", "q1", clip_logits=z)
    qb, zb = bank_io.load_query_bank(dq, "cpu", "q1")
    assert qb.dtype == torch.bfloat16 and torch.equal(qb, q) and torch.equal(zb, z)
    assert bank_io.load_query_bank(tmp_path / "dense", "cpu") is None
