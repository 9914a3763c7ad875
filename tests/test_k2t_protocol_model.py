"""CPU model of the mbarrier protocol of k2t (csrc/cross_attention.cu: short_kv_attn_tc_kernel): the TMA producer, the S issuer, the
P V issuer and the three softmax / epilogue groups of ONE CTA as cooperating state machines over the same barriers, phases and parities
as the kernel, run under a randomised scheduler with randomly delayed asynchronous completions (TMA transactions, tcgen05.commit
arrivals — delivered in issue order per issuing warp, as the hardware does).

Checked for many unit counts, (b, h) run lengths and seeds: no deadlock, every unit is computed exactly once, and no buffer is
overwritten before its consumer is done with it (Q ring stages, K/V ring slots, the S/P and O columns of a TMEM slot, the staging
tile) — i.e. the waits in the kernel are sufficient, and the parities never alias."""
import random
import re
import os

import pytest

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tweediemix_b200", "csrc", "cross_attention.cu")


def consts():
    m = re.search(r"constexpr int kStages = (\d+), kSlots = (\d+), kKvSlots = (\d+);", open(SRC).read())
    assert m, "ring constants not found in cross_attention.cu"
    return tuple(int(v) for v in m.groups())


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.completed = count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending, self.completed = 0, self.completed + 1

    def passed(self, parity):                      # mbarrier.try_wait.parity: the phase of that parity has completed
        return self.completed % 2 != parity


def simulate(n_units, QT, u_begin, seed, fault=None):
    NST, NSLOT, NKV = consts()
    rng = random.Random(seed)
    bars = {}
    for i in range(NST):
        bars["full", i], bars["empty", i] = Bar(1), Bar(1)
    for i in range(NKV):
        bars["kv_full", i], bars["kv_empty", i] = Bar(1), Bar(1)
    for i in range(NSLOT):
        bars["s_full", i], bars["p_full", i], bars["o_full", i], bars["slot_free", i] = Bar(1), Bar(4), Bar(1), Bar(4)
    # resources: what each buffer currently holds
    qstage = [None] * NST            # unit whose Q tile is (being) loaded / resident, None = free
    kvslot = [None] * NKV            # run index resident
    sp = [("free", None)] * NSLOT    # ("S", u) after the S MMA, ("P", u) after the softmax, ("free", u) after P V consumed it
    oc = [("free", None)] * NSLOT    # ("O", u) after P V, ("free", u) after the epilogue loaded it
    staging = [None] * NSLOT         # unit whose store may still be reading the staging tile
    done = []
    async_q = {"tma": [], "s": [], "pv": []}      # in-order completion queues: (callable)

    def bh(i):
        return (u_begin + i) // QT

    def run_index(i):                              # (b, h) run index of local unit i
        return len({bh(k) for k in range(i + 1)}) - 1

    def tma():
        prev, kvi = None, -1
        for i in range(n_units):
            st = i % NST
            if bh(i) != prev:
                prev, kvi = bh(i), kvi + 1
                g = kvi % NKV
                yield ("wait", ("kv_empty", g), ((kvi // NKV) & 1) ^ 1)
                assert kvslot[g] is None, f"K/V slot {g} overwritten while run {kvslot[g]} is in use"
                kvslot[g] = kvi
                async_q["tma"].append(lambda g=g: bars["kv_full", g].arrive())
            if fault != "tma_skips_empty":
                yield ("wait", ("empty", st), ((i // NST) & 1) ^ 1)
            assert qstage[st] is None, f"Q stage {st} overwritten while unit {qstage[st]} is in use"
            qstage[st] = i
            async_q["tma"].append(lambda st=st: bars["full", st].arrive())

    def s_issuer():
        for i in range(n_units):
            st, j = i % NST, i % NSLOT
            kvi = run_index(i)
            g = kvi % NKV
            yield ("wait", ("full", st), (i // NST) & 1)
            yield ("wait", ("kv_full", g), (kvi // NKV) & 1)
            if i >= NSLOT and fault != "s_skips_o_full":
                yield ("wait", ("o_full", j), ((i // NSLOT) - 1) & 1)
            assert qstage[st] == i and kvslot[g] == kvi
            assert sp[j][0] == "free", f"S({i}) issued into slot {j} holding {sp[j]}"
            sp[j] = ("busy", i)

            def complete(i=i, st=st, j=j):
                sp[j] = ("S", i)
                bars["s_full", j].arrive()
                qstage[st] = None
                bars["empty", st].arrive()
            async_q["s"].append(complete)

    def pv_issuer():
        for k in range(n_units):
            j = k % NSLOT
            kvi = run_index(k)
            g = kvi % NKV
            last = k + 1 >= n_units or bh(k + 1) != bh(k)
            yield ("wait", ("p_full", j), ((k // NSLOT) & 1) ^ (1 if fault == "pv_wrong_parity" and k >= NSLOT else 0))
            if k >= NSLOT:
                yield ("wait", ("slot_free", j), ((k // NSLOT) - 1) & 1)
            assert sp[j] == ("P", k) and kvslot[g] == kvi
            assert oc[j][0] == "free", f"P V({k}) issued into O columns holding {oc[j]}"
            oc[j] = ("busy", k)

            def complete(k=k, j=j, g=g, last=last):
                oc[j] = ("O", k)
                sp[j] = ("free", k)
                bars["o_full", j].arrive()
                if last:
                    kvslot[g] = None
                    bars["kv_empty", g].arrive()
            async_q["pv"].append(complete)

    def softmax_warp(slot, q):
        for i in range(slot, n_units, NSLOT):
            par = (i // NSLOT) & 1
            yield ("wait", ("s_full", slot), par)
            assert sp[slot][1] == i and sp[slot][0] in ("S", "P")      # (another warp of the group may already have stored its P rows)
            if q == 0:
                sp[slot] = ("P", i)                                       # group-level state: set once all four warps arrive below
            bars["p_full", slot].arrive()
            yield ("wait", ("o_full", slot), par)
            assert oc[slot][1] == i
            if q == 0:                                                    # the issuing thread: previous store has read the staging tile
                staging[slot] = None
            yield ("barrier", ("A", slot, i))
            bars["slot_free", slot].arrive()
            yield ("barrier", ("B", slot, i))
            if q == 0:
                oc[slot] = ("free", i)
                assert staging[slot] is None
                staging[slot] = i
                done.append(i)

    roles = {"tma": tma(), "s": s_issuer(), "pv": pv_issuer()}
    for slot in range(NSLOT):
        for q in range(4):
            roles[f"smx{slot}{q}"] = softmax_warp(slot, q)
    blocked = {}                                                          # role -> pending wait
    named = {}                                                            # named barrier -> arrivals
    for name, gen in list(roles.items()):
        try:
            blocked[name] = next(gen)
        except StopIteration:
            del roles[name]
    steps = 0
    while roles:
        steps += 1
        assert steps < 200000, "runaway simulation"
        moves = []
        for name, w in blocked.items():
            if w[0] == "wait" and bars[w[1]].passed(w[2]):
                moves.append(("role", name))
            elif w[0] == "barrier" and named.get(w[1], 0) == 4:
                moves.append(("role", name))
            elif w[0] == "barrier" and name not in named.setdefault(("in", w[1]), set()):
                named[("in", w[1])].add(name)
                named[w[1]] = named.get(w[1], 0) + 1
                if named[w[1]] == 4:
                    moves.append(("role", name))
        moves += [("async", k) for k, v in async_q.items() if v]
        assert moves, f"deadlock with {len(roles)} roles blocked: {blocked}"
        kind, name = rng.choice(moves)
        if kind == "async":
            async_q[name].pop(0)()
            continue
        try:
            blocked[name] = next(roles[name])
        except StopIteration:
            del roles[name], blocked[name]
    while any(async_q.values()):
        for v in async_q.values():
            if v:
                v.pop(0)()
    assert sorted(done) == list(range(n_units))
    assert all(s is None for s in qstage) and all(s is None for s in kvslot)
    return True


@pytest.mark.parametrize("n_units", [1, 2, 3, 4, 5, 8, 9, 17, 40])
@pytest.mark.parametrize("QT,u_begin", [(8, 0), (8, 5), (32, 31), (1, 0), (2, 1), (3, 7)])
def test_k2t_protocol_never_deadlocks_or_overwrites(n_units, QT, u_begin):
    for seed in range(6):
        assert simulate(n_units, QT, u_begin, seed)


@pytest.mark.parametrize("fault", ["s_skips_o_full", "tma_skips_empty", "pv_wrong_parity"])
def test_the_model_catches_a_broken_protocol(fault):
    """Teeth: dropping one of the kernel's waits (or flipping a parity) must trip an overwrite / ordering assertion or deadlock."""
    failures = 0
    for seed in range(8):
        try:
            simulate(17, 8, 5, seed, fault=fault)
        except AssertionError:
            failures += 1
    assert failures > 0


def test_constants_match_the_documented_pipeline():
    assert consts() == (8, 3, 2)
