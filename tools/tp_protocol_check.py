"""Randomised interleaving check of the cross-GPU exchange PROTOCOLS of the experimental tensor-parallel paths — a
model of csrc/mega.cu (tp_reduce_phase = "mode1", tp_reduce_cta = "mode2", the LL push exchange tp_reduce_ll = "ll") and csrc/tp_twoshot.cu ("twoshot"), not the
CUDA code itself: flags, epochs, parity double-buffering of the partials, the single-buffered gather of the two-shot.
Each CTA is a generator; `yield` = a point where any other CTA may run.  Reads assert they see exactly the value of the
allreduce they belong to (not a stale one, not one overwritten by a later allreduce)."""
import random, sys

def run(protocol, R, C, K, seed):
    rng = random.Random(seed)
    partial = [[[None] * C for _ in range(2)] for _ in range(R)]          # [rank][parity][cta range]
    flag = [[0] * R for _ in range(R)]                                     # mode 1: flag[recv][src]
    cflag = [[[0] * C for _ in range(R)] for _ in range(R)]                # mode 2: [recv][src][cta]
    bar = [[0] for _ in range(R)]                                          # grid barrier arrival counter per rank
    gather = [[None] * C for _ in range(R)]                                # two-shot: single-buffered gather per rank (slice = rank's)
    flag2 = [[0] * R for _ in range(R)]
    done = [0] * R
    ll = [[[[(None, 0)] * C for _ in range(R)] for _ in range(2)] for _ in range(R)]   # "ll": [recv][parity][src][cta] = (value, epoch)
    def grid_barrier(r, n):            # n-th barrier of this rank (monotonic counter like the kernel's)
        bar[r][0] += 1
        while bar[r][0] < n * C: yield
    def cta(r, c):
        nb = 0
        for k in range(1, K + 1):
            par = k & 1
            yield
            partial[r][par][c] = (r, k)                                    # projection epilogue
            yield
            if protocol == "ll":
                # mode 3 of mega.cu: push {value, epoch} into every rank's slot (own included), no flag; the receiver polls
                # the slot for EQUALITY with the epoch (a slot overwritten by allreduce k+2 before it was read = deadlock here)
                for p in range(R):
                    yield
                    ll[p][par][r][c] = ((r, k), k)
                for p in range(R):
                    while ll[r][par][p][c][1] != k:
                        assert ll[r][par][p][c][1] < k, ("ll: slot overwritten before it was read", r, c, k, p, ll[r][par][p][c])
                        yield
                    assert ll[r][par][p][c][0] == (p, k)
                nb += 1; yield from grid_barrier(r, nb)
            elif protocol == "mode2":
                for p in range(R):
                    if p != r: cflag[p][r][c] = k
                for p in range(R):
                    if p != r:
                        while cflag[r][p][c] < k: yield
                for p in range(R):
                    yield
                    assert partial[p][par][c] == (p, k), (protocol, "stale/overwritten", r, c, k, p, partial[p][par][c])
                nb += 1; yield from grid_barrier(r, nb)
            elif protocol == "mode1":
                nb += 1; yield from grid_barrier(r, nb)
                if c == 0:
                    for p in range(R):
                        if p != r: flag[p][r] = k
                for p in range(R):
                    if p != r:
                        while flag[r][p] < k: yield
                # slice-wise pull: CTA c reads range c of every rank
                for p in range(R):
                    yield
                    assert partial[p][par][c] == (p, k), (protocol, r, c, k, p, partial[p][par][c])
                nb += 1; yield from grid_barrier(r, nb)
            elif protocol == "twoshot":   # kernel boundary before (all local partials complete) = barrier
                nb += 1; yield from grid_barrier(r, nb)
                if c == 0:
                    for p in range(R):
                        if p != r: flag[p][r] = k
                for p in range(R):
                    if p != r:
                        while flag[r][p] < k: yield
                # B: rank r reduces its own slice (range index = (r, c)); reads peers' partial for slice r
                for p in range(R):
                    yield
                    assert partial[p][par][c] == (p, k), ("twoshot B", r, c, k, p)
                gather[r][c] = (r, k)
                yield
                done[r] += 1
                if done[r] == C:          # last CTA
                    done[r] = 0
                    for p in range(R):
                        if p != r: flag2[p][r] = k
                for p in range(R):
                    if p != r:
                        while flag2[r][p] < k: yield
                for p in range(R):
                    if p != r:
                        yield
                        assert gather[p][c] == (p, k), ("twoshot C", r, c, k, p, gather[p][c])
                nb += 1; yield from grid_barrier(r, nb)   # kernel end / next kernel boundary
    gens = [cta(r, c) for r in range(R) for c in range(C)]
    live = list(range(len(gens)))
    steps = 0
    while live:
        i = rng.choice(live) if rng.random() < 0.7 else live[rng.randrange(len(live)) // 2]  # skewed scheduling
        try:
            next(gens[i])
        except StopIteration:
            live.remove(i)
        steps += 1
        assert steps < 5_000_000, (protocol, "no progress: deadlock?")
    return steps

def main(n_cfg=40):
    for proto in ("mode1", "mode2", "ll", "twoshot"):
        tot = 0
        for seed in range(n_cfg):
            tot += run(proto, R=random.Random(seed).choice([2, 3, 4]), C=random.Random(seed + 1).choice([1, 2, 5]), K=9, seed=seed)
        print(proto, "ok,", tot, "scheduler steps over", n_cfg, "random configurations")


if __name__ == "__main__":
    main()
