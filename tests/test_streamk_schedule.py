"""CPU check of the stream-K unit schedule of gemm2_f16_kernel<EPI_RESADD> (spokennlp_b200/csrc/gemm2.cuh: `sk`, `decode`,
`next_unit`).  The three warp roles of the kernel (TMA producer, MMA issuer, epilogue) walk the same arithmetic; this file
transcribes it and checks what the kernel relies on: every K block of every tile is processed exactly once, exactly one
segment per tile carries the bias (kb0 == 0), no segment is empty, and the shares differ by at most one K block."""
import os
import re

import pytest

from conftest import ROOT


def segments(m_tiles, n_tiles, k_blocks, n_pairs):
    total = m_tiles * n_tiles * k_blocks
    for pair in range(n_pairs):
        u, hi = total * pair // n_pairs, total * (pair + 1) // n_pairs          # u_begin, sk_hi
        while u < hi:                                                            # for (u = u_begin; u < u_end; u = next_unit(...))
            t = u // k_blocks
            nt, mt = t % n_tiles, t // n_tiles
            kb0 = u - t * k_blocks
            kb1 = min(k_blocks, kb0 + (hi - u))
            yield pair, mt, nt, kb0, kb1
            u += kb1 - kb0


@pytest.mark.parametrize("M,N,K,pairs", [(16384, 768, 3072, 74), (16384, 768, 768, 74), (16384, 768, 3072, 72), (1024, 768, 768, 74),
                                         (300, 768, 1536, 74), (256, 256, 64, 74), (16384, 3072, 768, 74), (4096, 768, 3072, 1)])
def test_streamk_covers_every_k_block_once_and_biases_each_tile_once(M, N, K, pairs):
    m_tiles, n_tiles, k_blocks = (M + 255) // 256, (N + 255) // 256, (K + 63) // 64
    n_pairs = min(pairs, m_tiles * n_tiles)          # launch_gemm2: grid = 2 * min(units, pairs)
    seen, bias, per_pair, segs_per_pair = {}, {}, [0] * n_pairs, [0] * n_pairs
    for pair, mt, nt, kb0, kb1 in segments(m_tiles, n_tiles, k_blocks, n_pairs):
        assert 0 <= mt < m_tiles and 0 <= nt < n_tiles and 0 <= kb0 < kb1 <= k_blocks
        for kb in range(kb0, kb1):
            assert (mt, nt, kb) not in seen
            seen[(mt, nt, kb)] = pair
        if kb0 == 0:
            bias[(mt, nt)] = bias.get((mt, nt), 0) + 1
        per_pair[pair] += kb1 - kb0
        segs_per_pair[pair] += 1
    assert len(seen) == m_tiles * n_tiles * k_blocks
    assert all(bias.get((mt, nt), 0) == 1 for mt in range(m_tiles) for nt in range(n_tiles))
    assert max(per_pair) - min(per_pair) <= 1
    # a pair touches at most (whole tiles in its share) + 2 partial tiles
    assert max(segs_per_pair) <= max(per_pair) // k_blocks + 2


def test_transcription_matches_the_kernel_source():
    src = open(os.path.join(ROOT, "spokennlp_b200", "csrc", "gemm2.cuh")).read()
    flat = re.sub(r"\s+", " ", src)
    for needle in ("const bool sk = RESADD && g.k_splits == -1;",
                   "sk_total * (pair + 1) / n_pairs", "sk_total * pair / n_pairs",
                   "const int t = u / k_blocks;", "kb0 = u - t * k_blocks;", "kb1 = min(k_blocks, kb0 + (sk_hi - u));",
                   "return sk ? u + (kb1 - kb0) : u + n_pairs;",
                   "const bool add_bias = !RESADD || kb0 == 0;"):
        assert needle in flat, needle
