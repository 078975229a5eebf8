"""Model of the NVLS reduce protocol of b200q_reduce.cu / the fused TP mat-vec (k_mmvq_ring, tp.in / tp.out), executed under random
interleavings of the ranks: two parity buffers per rank, a flag word per rank that every rank's completion increments (multicast),
a rank-local sequence counter, and the per-parity DIRTY EXTENT that tells the next user how much of a buffer must be zeroed.
The model checks the invariants the kernels rely on: a consumer that passed its flag wait sees the complete sum, and no add can land in a
buffer before its owner has zeroed the part that is still dirty — for reduces of MIXED lengths (tg: n_embd floats, pp: 512 x n_embd), which is
exactly the case that the first implementation (zeroing only the current length) got wrong on the GPU."""
import random

import pytest


class Rank:
    def __init__(self, r, world, stride):
        self.r, self.world = r, world
        self.buf = [[0.0] * stride, [0.0] * stride]
        self.flag = 0
        self.seq = 0
        self.dirty = [0, 0]
        self.pc = 0            # index of the next op of this rank's program
        self.stage = 0         # sub-step inside the op


def run(world, lengths, seed, zero_dirty=True):
    """lengths[i] = length of reduce i; every rank runs: producer(i) ; consumer(i) for i in order.  Returns the consumer results."""
    rnd = random.Random(seed)
    stride = max(lengths)
    ranks = [Rank(r, world, stride) for r in range(world)]
    partial = lambda r, i, j: float((r + 1) * 1000 + i * 7 + (j % 5))       # integer-valued: sums are exact in any order
    expect = lambda i, j: sum(partial(r, i, j) for r in range(world))
    seen = [[None] * len(lengths) for _ in range(world)]
    # every rank's program: for each reduce i: [zero, add..., signal] then [wait, read]
    done = 0
    pending_adds = {r: [] for r in range(world)}       # adds issued by rank r but not yet delivered (delivered in random order, before r's signal)
    while done < world:
        R = rnd.choice([x for x in ranks if x.pc < 2 * len(lengths)])
        i, kind = divmod(R.pc, 2)
        n = lengths[i]
        if kind == 0:                                   # ---- producer of reduce i (the row-parallel mat-vec / k_allreduce_nvls) ----
            if R.stage == 0:
                s = R.seq; assert s == i
                p = s & 1
                ext = R.dirty[p ^ 1] if zero_dirty else n          # (old protocol: zero only the current length)
                for j in range(ext):
                    R.buf[p ^ 1][j] = 0.0
                pending_adds[R.r] = [(q, j) for q in range(world) for j in range(n)]
                rnd.shuffle(pending_adds[R.r])
                R.stage = 1
            elif R.stage == 1:                          # deliver some of the multimem.red adds (they reach the ranks in any order)
                for _ in range(rnd.randint(1, max(1, len(pending_adds[R.r])))):
                    if not pending_adds[R.r]:
                        break
                    q, j = pending_adds[R.r].pop()
                    ranks[q].buf[i & 1][j] += partial(R.r, i, j)
                if not pending_adds[R.r]:
                    R.stage = 2
            else:                                       # last CTA: bookkeeping, then the release-flag on every rank (after all adds: fence + release)
                p = i & 1
                R.dirty[p ^ 1] = 0; R.dirty[p] = n; R.seq = i + 1
                for q in ranks:
                    q.flag += 1
                R.stage = 0; R.pc += 1
        else:                                           # ---- consumer of reduce i (prologue of the next mat-vec / copy-out) ----
            if R.flag >= world * R.seq and R.seq == i + 1:
                seen[R.r][i] = [R.buf[i & 1][j] for j in range(n)]
                R.pc += 1
                if R.pc == 2 * len(lengths):
                    done += 1
            # else: still spinning on the flag; another rank gets scheduled
    return seen, expect


@pytest.mark.parametrize("world", [2, 4, 8])
def test_protocol_holds_for_mixed_lengths_under_random_interleavings(world):
    lengths = [8, 8, 64, 64, 8, 8, 8, 64, 16, 8, 64, 8]       # tg-sized and pp-sized reduces sharing the parity buffers
    for seed in range(40):
        seen, expect = run(world, lengths, seed)
        for r in range(world):
            for i, n in enumerate(lengths):
                assert seen[r][i] == [expect(i, j) for j in range(n)], (world, seed, r, i)


def test_zeroing_only_the_current_length_is_wrong():
    """The first implementation zeroed only [0, n) of the other parity buffer: a long reduce after short ones then adds into stale tails."""
    lengths = [64, 8, 16]          # long on parity 0, short on parity 1 (zeroes only 8 floats of parity 0), medium on parity 0: [8, 16) is stale
    bad = 0
    for seed in range(20):
        seen, expect = run(2, lengths, seed, zero_dirty=False)
        bad += any(seen[r][i] != [expect(i, j) for j in range(n)] for r in range(2) for i, n in enumerate(lengths))
    assert bad == 20


# ------------------------------------------------------------------------------------------------------------------------------------
# Tagged-slot exchange of the fused decode reduce (round 2: k_mmvq_ring<..., TP>, include/b200q.h b200q_nvls_comm): every entry is
# {value, number of the reduce}, written with one store; deliveries to the peers are NOT ordered and take arbitrarily long.
# ------------------------------------------------------------------------------------------------------------------------------------
def run_tagged(world, n_reduces, length, seed, parities=2):
    """Every rank runs produce(1); consume(1); produce(2); consume(2) ...; a produce only ISSUES the stores (one per element and destination),
    the network delivers them later in random order.  Returns (results, deadlock)."""
    rnd = random.Random(seed)
    slots = [[[[(0.0, 0)] * length for _ in range(world)] for _ in range(parities)] for _ in range(world)]     # [dst][parity][src][e]
    in_flight = []                                   # (dst, parity, src, e, value, tag)
    val = lambda r, i, e: float((r + 1) * 100 + i * 3 + e % 7)
    pc = [0] * world                                 # 2 * (i - 1) = about to produce i, 2 * (i - 1) + 1 = consuming i
    got = [[None] * (n_reduces + 1) for _ in range(world)]
    while any(p < 2 * n_reduces for p in pc):
        choices = [("rank", r) for r in range(world) if pc[r] < 2 * n_reduces] + ([("net", None)] if in_flight else [])
        kind, r = choices[rnd.randrange(len(choices))]
        if kind == "net":
            dst, par, src, e, v, tag = in_flight.pop(rnd.randrange(len(in_flight)))
            slots[dst][par][src][e] = (v, tag)
            continue
        i = pc[r] // 2 + 1
        if pc[r] % 2 == 0:                           # reduce_out launch: issue, do not wait
            for e in range(length):
                for dst in range(world):
                    in_flight.append((dst, (i - 1) % parities, r, e, val(r, i, e), i))
            pc[r] += 1
        else:                                        # reduce_in launch: completes once every slot carries tag i
            def complete(q):
                j = pc[q] // 2 + 1
                return all(slots[q][(j - 1) % parities][src][e][1] == j for src in range(world) for e in range(length))
            if complete(r):
                mine = slots[r][(i - 1) % parities]
                got[r][i] = [sum(mine[src][e][0] for src in range(world)) for e in range(length)]
                pc[r] += 1
            elif not in_flight and all(pc[q] >= 2 * n_reduces or (pc[q] % 2 == 1 and not complete(q)) for q in range(world)):
                return got, True                     # nothing in flight, every rank is waiting for a tag that will never come
    return got, False


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tagged_slot_exchange_two_parities(world):
    for seed in range(40):
        got, dead = run_tagged(world, n_reduces=7, length=5, seed=seed)
        assert not dead
        for r in range(world):
            for i in range(1, 8):
                assert got[r][i] == [sum(float((q + 1) * 100 + i * 3 + e % 7) for q in range(world)) for e in range(5)], (seed, r, i)


def test_tagged_slot_exchange_needs_two_parities():
    """With a single buffer a fast rank's reduce i+1 overwrites entries a slow rank has not consumed yet: that rank then waits for tag i forever."""
    assert any(run_tagged(2, n_reduces=6, length=4, seed=s, parities=1)[1] for s in range(40))


# ------------------------------------------------------------------------------------------------------------------------------------
# q8 hand-off (k_mmvq_ring<..., Q8 = 2>): the fused up/gate launch quantises its own result in the CTA tails.  A CTA owns a static row range;
# 32-row blocks inside the range are quantised directly, blocks shared with a neighbour through an arrival counter (rows, not CTAs, are
# counted): the CTA whose rows complete the block quantises it and re-arms the counter for the next launch.
# ------------------------------------------------------------------------------------------------------------------------------------
def run_q8_tail(M, grid, rpu, seed, launches=3):
    rnd = random.Random(seed)
    n_units = (M + rpu - 1) // rpu
    nblk = (M + 31) // 32
    cnt = [0] * nblk
    for launch in range(launches):
        written = [False] * M
        quantised = [0] * nblk
        ctas = list(range(grid))
        rnd.shuffle(ctas)
        # every CTA: (a) write its rows, (b) tail.  Tails of different CTAs interleave arbitrarily; the atomicAdd is the only shared step.
        events = []
        for b in ctas:
            c0, c1 = n_units * b // grid, n_units * (b + 1) // grid
            events.append(("rows", b, min(rpu * c0, M), min(rpu * c1, M)))
        rnd.shuffle(events)
        pending_tail = []
        for ev in events:
            _, b, r0, r1 = ev
            for r in range(r0, r1):
                written[r] = True
            pending_tail.append((b, r0, r1))
            # randomly run some of the pending tails now (a tail only starts after ITS OWN rows are written)
            rnd.shuffle(pending_tail)
            while pending_tail and rnd.random() < 0.5:
                tb, t0, t1 = pending_tail.pop()
                if t1 <= t0:
                    continue
                for blk in range(t0 >> 5, ((t1 - 1) >> 5) + 1):
                    lo, hi = max(t0, 32 * blk), min(t1, 32 * blk + 32)
                    need, own = min(32, M - 32 * blk), hi - lo
                    mine = own == need
                    if not mine:
                        old = cnt[blk]; cnt[blk] += own              # atomicAdd
                        mine = old + own == need
                        if mine:
                            cnt[blk] = 0
                    if mine:
                        assert all(written[32 * blk: 32 * blk + need]), ("block quantised before all its rows were written", blk)
                        quantised[blk] += 1
        for tb, t0, t1 in pending_tail:                               # the remaining tails
            if t1 <= t0:
                continue
            for blk in range(t0 >> 5, ((t1 - 1) >> 5) + 1):
                lo, hi = max(t0, 32 * blk), min(t1, 32 * blk + 32)
                need, own = min(32, M - 32 * blk), hi - lo
                mine = own == need
                if not mine:
                    old = cnt[blk]; cnt[blk] += own
                    mine = old + own == need
                    if mine:
                        cnt[blk] = 0
                if mine:
                    assert all(written[32 * blk: 32 * blk + need]), blk
                    quantised[blk] += 1
        assert quantised == [1] * nblk, (launch, quantised)
        assert cnt == [0] * nblk, "arrival counters must be re-armed for the next launch"


@pytest.mark.parametrize("M,grid,rpu", [(14336, 296, 2), (7168, 296, 2), (1792, 82, 2), (4096, 296, 1), (200, 7, 2), (64, 3, 2)])
def test_q8_handoff_tail_protocol(M, grid, rpu):
    for seed in range(10):
        run_q8_tail(M, grid, rpu, seed)
