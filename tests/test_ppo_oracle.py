"""CPU suite: pins oracle/ppo_oracle.py (first-party scalar code of the reference: ComputeGAE, trajectory concatenation,
ExperienceBuffer FIFO) with hand-derived known-answer cases — the reference ships no vectors for them."""
import numpy as np

from oracle import ppo_oracle as po


def test_gae_closed_form_no_terminal():
    # values 0, rewards 1, gamma 0.5, lambda 1, no done/trunc, returnStd 1, no clip:
    # delta_t = 1 ; A_t = sum_{k>=0} 0.5^k  over the remaining steps ; ret same
    n = 5
    adv, tgt, ret = po.compute_gae(np.ones(n), np.zeros(n), np.zeros(n), np.zeros(n + 1), 0.5, 1.0, 1.0, 0.0)
    exp = np.array([1.9375, 1.875, 1.75, 1.5, 1.0], dtype=np.float32)
    assert np.array_equal(adv, exp) and np.array_equal(ret, exp) and np.array_equal(tgt, exp)


def test_gae_done_and_truncation_cut_the_carry_but_only_done_cuts_bootstrap():
    # step 1 is terminal, step 3 truncated. gamma = lambda = 1, values = 10 everywhere (11 after the end)
    r = np.array([1, 2, 3, 4], dtype=np.float32)
    d = np.array([0, 1, 0, 0], dtype=np.float32)
    tr = np.array([0, 0, 0, 1], dtype=np.float32)
    v = np.array([10, 10, 10, 10, 11], dtype=np.float32)
    adv, tgt, ret = po.compute_gae(r, d, tr, v, 1.0, 1.0, 1.0, 0.0)
    # t=3 (truncated): delta = 4 + 11 - 10 = 5 (bootstrap kept: TorchFuncs.cpp:36 multiplies by `done` only), carry cut
    # t=2: delta = 3 + 10 - 10 = 3, A = 3 + 5 = 8 ; ret = 3 + 4 = 7
    # t=1 (done): delta = 2 + 0 - 10 = -8, A = -8 (carry cut) ; ret = 2
    # t=0: delta = 1, A = 1 - 8 = -7 ; ret = 1 + 2 = 3
    assert np.array_equal(adv, np.array([-7, -8, 8, 5], dtype=np.float32))
    assert np.array_equal(ret, np.array([3, 2, 7, 4], dtype=np.float32))
    assert np.array_equal(tgt, np.array([3, 2, 18, 15], dtype=np.float32))


def test_gae_return_scale_and_clip():
    r = np.array([100.0, -100.0], dtype=np.float32)
    adv, _, ret = po.compute_gae(r, [0, 0], [0, 1], [0, 0, 0], 0.0, 0.0, 5.0, 10.0)
    assert np.array_equal(adv, np.array([10.0, -10.0], dtype=np.float32))  # 100/5 = 20 -> clipped to 10
    assert np.array_equal(ret, r)  # returns are NOT normalised (TorchFuncs.cpp:38)
    adv, _, _ = po.compute_gae(r, [0, 0], [0, 1], [0, 0, 0], 0.0, 0.0, 0.0, 10.0)
    assert np.array_equal(adv, r)  # returnStd == 0 -> raw rewards (TorchFuncs.cpp:31-33)


def test_concat_order_and_truncation_marks():
    T, A, P = 3, 2, 2
    rew = np.arange(T * A * P, dtype=np.float32).reshape(T, A * P)
    done = np.zeros((T, A), dtype=np.uint8)
    done[2, 1] = 1  # arena 1 terminal on its last step; arena 0 merely cut short
    done[0, 0] = 1
    cat = po.concat_reference_order({"rewards": rew}, done, P)
    assert np.array_equal(cat["rewards"], rew.T.reshape(-1))  # row n's T steps back to back
    assert np.array_equal(cat["dones"].reshape(A * P, T), [[1, 0, 0], [1, 0, 0], [0, 0, 1], [0, 0, 1]])
    assert np.array_equal(cat["truncateds"].reshape(A * P, T), [[0, 0, 1], [0, 0, 1], [0, 0, 0], [0, 0, 0]])


def test_gae_seam_uses_next_rows_first_value():
    # two rows of T=1, no done: row 0's bootstrap is row 1's value (the reference's quirk), row 1's is the appended value
    rew = np.array([[1.0, 1.0]], dtype=np.float32)
    done = np.zeros((1, 2), dtype=np.uint8)
    val = np.array([[5.0, 7.0], [100.0, 9.0]], dtype=np.float32)  # slot T: only the LAST row's entry (9) is used
    adv, tgt, ret = po.gae_reference_order(rew, done, val, 1, 1.0, 1.0, 1.0, 0.0)
    assert np.array_equal(adv, [[1 + 7 - 5, 1 + 9 - 7]])
    assert np.array_equal(ret, [[1, 1]])


def test_experience_buffer_fifo():
    b = po.ExperienceBufferOracle(5)
    b.submit({"x": np.array([1, 2, 3], dtype=np.float32)})
    assert b.cur == 3 and np.array_equal(b.data["x"][:3], [1, 2, 3]) and np.isnan(b.data["x"][3:]).all()
    b.submit({"x": np.array([4, 5, 6], dtype=np.float32)})
    assert b.cur == 5 and np.array_equal(b.data["x"], [2, 3, 4, 5, 6])
    b.submit({"x": np.arange(10, 17, dtype=np.float32)})  # larger than the buffer: keeps the tail
    assert np.array_equal(b.data["x"], [12, 13, 14, 15, 16])


def test_policy_probs_clamp_and_temperature():
    logits = np.array([[0.0, 0.0, -1000.0]], dtype=np.float32)
    p = po.policy_probs(logits)
    assert p[0, 2] == np.float32(1e-11) and abs(p[0, 0] - 0.5) < 1e-7
    p2 = po.policy_probs(np.array([[2.0, 0.0]], dtype=np.float32), temperature=2.0)
    assert abs(p2[0, 0] - 1 / (1 + np.exp(-1.0))) < 1e-6
