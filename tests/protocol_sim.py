"""A small discrete-event model of the mbarrier protocols of the persistent attention kernels (attn_fwd3.cuh, attn_bwd3.cuh).

Every warp role of a CTA is transcribed as a Python generator that performs the SAME sequence of barrier operations as the
CUDA code (same counters, same parities, same order); TMA loads and tcgen05.commit arrivals complete after random delays
(in issue order per issuing thread).  A random scheduler interleaves the roles.  The model checks what a GPU test cannot
enumerate:
  * no deadlock for any item mix (single-block items, items without a second query tile, dead key blocks, one CTA walking
    many items, more CTAs than items);
  * no phase aliasing: a parity wait always observes exactly the phase it means (a waiter two phases behind would hang on
    hardware, one that runs ahead would read stale data);
  * no buffer hazard: a shared-memory / TMEM buffer is never overwritten while an operation that reads it is outstanding,
    and never read before the write it expects has landed.
TEST INFRASTRUCTURE ONLY (CPU); used by tests/test_protocol.py.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional


class Deadlock(AssertionError):
    pass


class Bar:
    """mbarrier: `count` arrivals complete a phase; wait(k) == try_wait.parity(k & 1) meaning "phase k has completed"."""

    def __init__(self, name: str, count: int):
        self.name, self.count, self.phase, self.pending = name, count, 0, 0

    def arrive(self, weight: int = 1) -> None:
        self.pending += weight
        assert self.pending <= self.count, f"{self.name}: more arrivals than the barrier expects in phase {self.phase}"
        if self.pending == self.count:
            self.phase += 1
            self.pending = 0

    def ready(self, k: int) -> bool:
        # hardware: succeeds iff the current (incomplete) phase has the other parity
        ok = (self.phase & 1) != (k & 1)
        if ok:
            assert self.phase == k + 1, f"{self.name}: waiter for phase {k} passed at phase {self.phase} (aliasing)"
        else:
            assert self.phase <= k, f"{self.name}: waiter for phase {k} is blocked although phase {self.phase} is current (would hang)"
        return ok


class Buf:
    """A buffer with outstanding asynchronous readers and a content tag."""

    def __init__(self, name: str):
        self.name, self.tag, self.readers = name, None, 0

    def write(self, tag) -> None:
        assert self.readers == 0, f"{self.name}: overwritten with {tag} while {self.readers} reader(s) of {self.tag} are outstanding"
        self.tag = tag

    def begin_read(self, tag) -> None:
        assert self.tag == tag, f"{self.name}: expected {tag}, holds {self.tag}"
        self.readers += 1

    def end_read(self) -> None:
        self.readers -= 1

    def read_now(self, tag) -> None:
        assert self.tag == tag, f"{self.name}: expected {tag}, holds {self.tag}"


class Sim:
    def __init__(self, seed: int):
        self.rng = random.Random(seed)
        self.roles: List = []
        self.pending = []            # (due_time, seq, fn) asynchronous completions
        self.now, self.seq = 0, 0
        self.fifo_tail: Dict[str, int] = {}

    def later(self, fn, queue: Optional[str] = None, lo: int = 1, hi: int = 40) -> None:
        """Run fn after a random delay; completions that share `queue` keep their issue order."""
        due = self.now + self.rng.randint(lo, hi)
        if queue is not None:
            due = max(due, self.fifo_tail.get(queue, 0) + 1)
            self.fifo_tail[queue] = due
        self.seq += 1
        self.pending.append((due, self.seq, fn))

    def spawn(self, name: str, gen) -> None:
        self.roles.append([name, gen, None])       # [name, generator, blocked-on predicate]

    def run(self, max_steps: int = 2_000_000) -> None:
        for _ in range(max_steps):
            self.now += 1
            self.pending.sort()
            while self.pending and self.pending[0][0] <= self.now:
                self.pending.pop(0)[2]()
            runnable = [r for r in self.roles if r[2] is None or r[2]()]
            if not self.roles:
                if not self.pending:
                    return
                continue
            if not runnable:
                if self.pending:
                    self.now = self.pending[0][0] - 1
                    continue
                raise Deadlock("deadlock: " + ", ".join(f"{r[0]}" for r in self.roles))
            r = self.rng.choice(runnable)
            r[2] = None
            try:
                r[2] = next(r[1])
            except StopIteration:
                self.roles.remove(r)
        raise AssertionError("simulation did not finish")


def wait(bar: Bar, k: int):
    return lambda: bar.ready(k)


# ======================================================================================================= forward (attn_fwd3.cuh)
def simulate_fwd(items, seed: int = 0, mutate: str = "", elect: bool = False):
    """items: list of (n_blocks, tileB) processed by ONE CTA in this order.  `mutate` plants a protocol bug (self-test of
    the checker): "no_k_empty" (the producer does not wait for the K stage to be released), "no_s_free" (the next score
    tile is issued without waiting for the softmax warps to have read the current one)."""
    sim = Sim(seed)
    q_full = [Bar(f"q_full{i}", 1) for i in range(2)]
    q_empty = [Bar(f"q_empty{i}", 2) for i in range(2)]
    k_full = [Bar(f"k_full{i}", 1) for i in range(2)]
    k_empty = [Bar(f"k_empty{i}", 2) for i in range(2)]
    v_full = [Bar(f"v_full{i}", 1) for i in range(2)]
    v_empty = [Bar(f"v_empty{i}", 2) for i in range(2)]
    s_full = [Bar(f"s_full{x}", 1) for x in range(2)]
    W = 1 if elect else 32            # arrivals a softmax warp contributes: one elected lane (ELECT kernels) or every lane
    s_free = [Bar(f"s_free{x}", 4 * W) for x in range(2)]
    p_full = [Bar(f"p_full{x}", 4 * W) for x in range(2)]
    o_full = [Bar(f"o_full{x}", 1) for x in range(2)]
    b_go = Bar("b_go", 1)
    Q = [Buf(f"Q{i}") for i in range(2)]
    K = [Buf(f"K{i}") for i in range(2)]
    V = [Buf(f"V{i}") for i in range(2)]
    S = [Buf(f"S{x}") for x in range(2)]           # TMEM score tile
    P = [Buf(f"P{x}") for x in range(2)]           # smem probabilities tile
    O = [Buf(f"O{x}") for x in range(2)]           # TMEM accumulator (tag = (item, blocks accumulated))

    def tma():
        n = kb = 0
        for it, (nb, tileB) in enumerate(items):
            buf = n & 1
            yield wait(q_empty[buf], (n >> 1) - 1) if (n >> 1) >= 1 else None
            Q[buf].write(("loading", it))
            sim.later(lambda buf=buf, it=it: (Q[buf].write(("Q", it)), q_full[buf].arrive()))
            for j in range(nb):
                s, use = kb & 1, kb >> 1
                if use >= 1 and mutate != "no_k_empty":
                    yield wait(k_empty[s], use - 1)
                K[s].write(("loading", it, j))
                sim.later(lambda s=s, it=it, j=j: (K[s].write(("K", it, j)), k_full[s].arrive()))
                if use >= 1:
                    yield wait(v_empty[s], use - 1)
                V[s].write(("loading", it, j))
                sim.later(lambda s=s, it=it, j=j: (V[s].write(("V", it, j)), v_full[s].arrive()))
                kb += 1
            n += 1

    def mma(x):
        n = kb = t = 0
        qname = f"mma{x}"

        def issue_s(it, j, s, buf, last):
            Q[buf].begin_read(("Q", it)); K[s].begin_read(("K", it, j))
            S[x].write(("computing", it, j))

            def done():
                Q[buf].end_read(); K[s].end_read()
                S[x].write(("S", it, j))
            sim.later(done, queue=qname)
            sim.later(s_full[x].arrive, queue=qname)
            sim.later(k_empty[s].arrive, queue=qname)
            if last:
                sim.later(q_empty[buf].arrive, queue=qname)

        for it, (nb, tileB) in enumerate(items):
            buf = n & 1
            if x == 1 and not tileB:
                yield wait(q_full[buf], n >> 1)
                q_empty[buf].arrive()
                for j in range(nb):
                    s, use = kb & 1, kb >> 1
                    yield wait(k_full[s], use)
                    k_empty[s].arrive()
                    yield wait(v_full[s], use)
                    v_empty[s].arrive()
                    kb += 1
                n += 1
                continue
            yield wait(q_full[buf], n >> 1)
            yield wait(k_full[kb & 1], kb >> 1)
            if t > 0:
                yield wait(s_free[x], t - 1)
            elif x == 1:
                yield wait(b_go, 0)
            issue_s(it, 0, kb & 1, buf, nb == 1)
            for j in range(nb):
                s, use = kb & 1, kb >> 1
                if j + 1 < nb:
                    s1, use1 = (kb + 1) & 1, (kb + 1) >> 1
                    yield wait(k_full[s1], use1)
                    if mutate != "no_s_free":
                        yield wait(s_free[x], t)
                    issue_s(it, j + 1, s1, buf, j + 2 == nb)
                yield wait(p_full[x], t)
                yield wait(v_full[s], use)
                P[x].begin_read(("P", it, j)); V[s].begin_read(("V", it, j))
                prev = O[x].tag
                assert j == 0 or prev == ("O", it, j), f"O{x}: accumulating block {j} of item {it} onto {prev}"
                O[x].write(("accumulating", it, j))

                def done(s=s, it=it, j=j):
                    P[x].end_read(); V[s].end_read()
                    O[x].write(("O", it, j + 1))
                sim.later(done, queue=qname)
                sim.later(o_full[x].arrive, queue=qname)
                sim.later(v_empty[s].arrive, queue=qname)
                t += 1
                kb += 1
            n += 1

    def softmax(x, w):
        t = 0
        for it, (nb, tileB) in enumerate(items):
            if x == 1 and not tileB:
                continue
            for j in range(nb):
                yield wait(s_full[x], t)
                S[x].read_now(("S", it, j))                     # tcgen05.ld + wait::ld
                yield None
                s_free[x].arrive(W)
                if x == 0 and t == 0 and w == 0:
                    b_go.arrive()
                if j > 0:
                    yield wait(o_full[x], t - 1)
                    O[x].read_now(("O", it, j))                 # lazy rescale may read / write O here
                yield None
                if w == 0:
                    P[x].write(("P", it, j))                    # (all four warps write their own rows; one tag is enough)
                else:
                    assert P[x].readers == 0, f"P{x} written while the previous P.V is still reading it"
                p_full[x].arrive(W)
                t += 1
            yield wait(o_full[x], t - 1)
            O[x].read_now(("O", it, nb))
            yield None
            assert P[x].readers == 0                            # context staging reuses this warp's rows of the P tile

    sim.spawn("tma", tma())
    sim.spawn("mmaA", mma(0))
    sim.spawn("mmaB", mma(1))
    for x in range(2):
        for w in range(4):
            sim.spawn(f"softmax{x}.{w}", softmax(x, w))
    sim.run()


# ======================================================================================================= backward (attn_bwd3.cuh)
def simulate_bwd(items, nq: int, seed: int = 0, mutate: str = "", elect: bool = False):
    """items: list of booleans (True = live key block, False = dead block) processed by ONE CTA; nq query blocks each.
    `mutate`: "no_kv_empty" (K / V of the next item loaded without waiting for the last gradient MMAs), "two_stages"
    (the producer believes the Q / dO ring has its 3 stages while the consumer side releases only what it used — modelled
    by dropping the wait on grad_done before P / dS are rewritten)."""
    NST = 3
    sim = Sim(seed)
    kv_full, kv_empty = Bar("kv_full", 1), Bar("kv_empty", 1)
    qdo_full = [Bar(f"qdo_full{i}", 1) for i in range(NST)]
    qdo_empty = [Bar(f"qdo_empty{i}", 1) for i in range(NST)]
    s_full = [Bar(f"s_full{g}", 1) for g in range(2)]
    W = 1 if elect else 32
    s_free = [Bar(f"s_free{g}", 4 * W) for g in range(2)]
    ds_full, grad_done = Bar("ds_full", 8 * W), Bar("grad_done", 1)
    KV = Buf("KV")
    QDO = [Buf(f"QdO{i}") for i in range(NST)]
    SD = [Buf(f"S/dP{g}") for g in range(2)]
    PDS = Buf("P/dS")
    DQ = [Buf("dQ0"), Buf("dQ1")]
    live = [i for i, ok in enumerate(items) if ok]

    def tma():
        n = qs = 0
        for it in live:
            def load_qdo(i, it=it):
                nonlocal qs
                st, use = qs % NST, qs // NST
                if use >= 1:
                    yield wait(qdo_empty[st], use - 1)
                QDO[st].write(("loading", it, i))
                sim.later(lambda: (QDO[st].write(("QdO", it, i)), qdo_full[st].arrive()))
                qs += 1
            yield from load_qdo(0)
            if n >= 1 and mutate != "no_kv_empty":
                yield wait(kv_empty, n - 1)
            KV.write(("loading", it))
            sim.later(lambda it=it: (KV.write(("KV", it)), kv_full.arrive()))
            for i in range(1, nq):
                yield from load_qdo(i)
            n += 1

    def mma():
        n = qs = ir = 0

        def scores(it, i, st, g):
            QDO[st].begin_read(("QdO", it, i)); KV.begin_read(("KV", it))
            SD[g].write(("computing", it, i))

            def done():
                QDO[st].end_read(); KV.end_read()
                SD[g].write(("S", it, i))
            sim.later(done, queue="mma")
            sim.later(s_full[g].arrive, queue="mma")

        for it in live:
            yield wait(kv_full, n)
            yield wait(qdo_full[qs % NST], qs // NST)
            if ir > 0:
                yield wait(s_free[0], ir - 1)
            scores(it, 0, qs % NST, 0)
            for i in range(nq):
                st = qs % NST
                if i == 0:
                    if ir > 0:
                        yield wait(s_free[1], ir - 1)
                    else:
                        yield wait(s_free[0], 0)
                    scores(it, 0, st, 1)
                if i + 1 < nq:
                    nst = (qs + 1) % NST
                    yield wait(qdo_full[nst], (qs + 1) // NST)
                    for g in range(2):
                        yield wait(s_free[g], ir)
                        scores(it, i + 1, nst, g)
                yield wait(ds_full, ir)
                PDS.begin_read(("PdS", it, i)); QDO[st].begin_read(("QdO", it, i)); KV.begin_read(("KV", it))
                DQ[ir & 1].write(("computing", it, i))

                def done(st=st, it=it, i=i, b=ir & 1):
                    PDS.end_read(); QDO[st].end_read(); KV.end_read()
                    DQ[b].write(("dQ", it, i))
                sim.later(done, queue="mma")
                sim.later(qdo_empty[st].arrive, queue="mma")
                sim.later(grad_done.arrive, queue="mma")
                if i + 1 == nq:
                    sim.later(kv_empty.arrive, queue="mma")
                ir += 1
                qs += 1
            n += 1

    def softmax(g, w):
        ir = 0
        for it in live:
            for i in range(nq):
                yield wait(s_full[g], ir)
                SD[g].read_now(("S", it, i))
                yield None
                s_free[g].arrive(W)
                yield None                                       # exp / dS math
                if i > 0 and mutate != "two_stages":
                    yield wait(grad_done, ir - 1)
                if g == 0 and w == 0:
                    PDS.write(("PdS", it, i))
                else:
                    assert PDS.readers == 0, "P / dS written while gradient MMAs still read them"
                ds_full.arrive(W)
                if i > 0:
                    DQ[(ir - 1) & 1].read_now(("dQ", it, i - 1))
                    yield None
                ir += 1
            yield wait(grad_done, ir - 1)
            DQ[(ir - 1) & 1].read_now(("dQ", it, nq - 1))
            yield None                                           # dK / dV drain (staging patches, named barriers: CTA-local)

    sim.spawn("tma", tma())
    sim.spawn("mma", mma())
    for g in range(2):
        for w in range(4):
            sim.spawn(f"softmax{g}.{w}", softmax(g, w))
    sim.run()
