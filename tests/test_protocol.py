"""The mbarrier protocols of the persistent attention kernels, checked on a discrete-event model (tests/protocol_sim.py):
no deadlock, no phase aliasing, no buffer hazard — for item mixes and interleavings a GPU run cannot enumerate."""
import random

import pytest

from protocol_sim import Deadlock, simulate_bwd, simulate_fwd


def _fwd_mixes():
    rng = random.Random(7)
    mixes = [
        [(4, True)] * 6,                                   # the bench shape: 6 items of 4 key blocks, two query tiles
        [(1, True)] * 7,                                   # single-block items back to back (S = 128..256)
        [(1, False)] * 5,                                  # one query tile, one key block (S = 128)
        [(3, True), (3, False)] * 3,                       # S = 300: every other item has no second tile
        [(2, False), (5, True), (1, True), (1, False), (4, True)],
        [(16, True)] * 2,                                  # S = 2048
    ]
    for _ in range(6):                                     # right-padded batches: the block count changes from item to item
        mixes.append([(rng.randint(1, 5), rng.random() < 0.8) for _ in range(rng.randint(1, 8))])
    return mixes


@pytest.mark.parametrize("mix", _fwd_mixes())
def test_forward_protocol(mix):
    for seed in range(12):
        simulate_fwd(mix, seed=seed)
    for seed in range(4):                      # attn_fwd3_kernel<.., ELECT = true>: one arrival per softmax warp, barrier counts 4
        simulate_fwd(mix, seed=seed, elect=True)


def _bwd_mixes():
    rng = random.Random(11)
    mixes = [([True] * 11, 4), ([True] * 5, 1), ([True, False, True, True, False, False, True], 4), ([True] * 3, 16),
             ([False, False, True], 2), ([True] * 4, 2), ([True] * 4, 3)]
    for _ in range(5):
        mixes.append(([rng.random() < 0.75 for _ in range(rng.randint(1, 9))] + [True], rng.randint(1, 6)))
    return mixes


@pytest.mark.parametrize("items,nq", _bwd_mixes())
def test_backward_protocol(items, nq):
    for seed in range(12):
        simulate_bwd(items, nq, seed=seed)
    for seed in range(4):                      # attn_bwd3_kernel<.., ELECT = true>
        simulate_bwd(items, nq, seed=seed, elect=True)


@pytest.mark.parametrize("mutate", ["no_k_empty", "no_s_free"])
def test_checker_catches_planted_forward_bugs(mutate):
    caught = 0
    for seed in range(20):
        try:
            simulate_fwd([(4, True)] * 4, seed=seed, mutate=mutate)
        except (AssertionError, Deadlock):
            caught += 1
    assert caught > 0, f"the model did not notice the planted bug {mutate!r}"


@pytest.mark.parametrize("mutate", ["no_kv_empty", "two_stages"])
def test_checker_catches_planted_backward_bugs(mutate):
    caught = 0
    for seed in range(20):
        try:
            simulate_bwd([True] * 4, 4, seed=seed, mutate=mutate)
        except (AssertionError, Deadlock):
            caught += 1
    assert caught > 0, f"the model did not notice the planted bug {mutate!r}"
