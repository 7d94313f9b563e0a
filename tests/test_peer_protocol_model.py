"""CPU: exhaustive model check of the peer-memory exchange protocol (scaled-mmd-gan_b200/csrc/smmd_peer.cuh).

compute-sanitizer is closed on this pool, so the slot / flag protocol of smmd_mmd2_fwd_bwd_peers is checked here on a
model instead: every interleaving of the per-rank operation sequences that the CUDA stream semantics allow is explored
(depth-first over the global state, memoised), and three properties are asserted in every reachable state:

  * a pull of rank r's rows at step k reads the rows published FOR step k (never a stale or an already-overwritten slot),
  * a combine at step k reads every rank's partial sums OF step k,
  * no deadlock: as long as some rank has work left, some operation is enabled.

Model = what the kernels do: per step k a rank runs, in stream order,
  publish(k): write own data slot k & 1                      signal(k): data_flag[k & 1][self] := k in every peer
  pull(k, r) for every peer r: wait data_flag[k & 1][r] >= k at home, read r's data slot k & 1
  sums(k): write own sums into slot k & 1 of every peer, sums_flag[k & 1][self] := k there
  combine(k): wait sums_flag[k & 1][r] >= k for every r, read the slots
Steps of the same parity are ordered (same stream); consecutive steps may overlap (two streams) -- or not (one stream).
The test also shows that the model has teeth: with a single set of flags shared by both parities (the first version of
the protocol) the two-stream schedule reaches a state where a pull reads rows of the wrong step."""
import itertools

import pytest


def _ops_of_step(k, rank, world):
    ops = [("publish", k), ("signal", k)]
    ops += [("pull", k, r) for r in range(world) if r != rank]
    ops += [("sums", k), ("combine", k)]
    return ops


class Model:
    def __init__(self, world, steps, two_streams, per_parity_flags=True):
        self.world, self.steps, self.two_streams, self.ppf = world, steps, two_streams, per_parity_flags
        # per rank: one op list per stream; with two streams step k runs on stream k & 1
        self.prog = []
        for r in range(world):
            streams = [[], []] if two_streams else [[]]
            for k in range(1, steps + 1):
                streams[(k & 1) if two_streams else 0] += _ops_of_step(k, r, world)
            self.prog.append(streams)

    def initial(self):
        w = self.world
        pcs = tuple(tuple(0 for _ in self.prog[r]) for r in range(w))
        data = tuple((0, 0) for _ in range(w))                       # data[r][slot] = step whose rows are in rank r's slot
        nflag = 2 if self.ppf else 1
        dflag = tuple(tuple(tuple(0 for _ in range(w)) for _ in range(nflag)) for _ in range(w))   # dflag[home][parity][src]
        sflag = dflag
        sums = tuple(tuple(tuple(0 for _ in range(w)) for _ in range(2)) for _ in range(w))        # sums[home][slot][src] = step
        return (pcs, data, dflag, sflag, sums)

    def _fi(self, k):
        return (k & 1) if self.ppf else 0

    def enabled(self, state, r, s):
        pcs, data, dflag, sflag, sums = state
        pc = pcs[r][s]
        if pc >= len(self.prog[r][s]):
            return None
        op = self.prog[r][s][pc]
        if op[0] == "pull":
            _, k, src = op
            if dflag[r][self._fi(k)][src] < k:
                return None
        if op[0] == "combine":
            k = op[1]
            if any(sflag[r][self._fi(k)][q] < k for q in range(self.world) if q != r):
                return None
        return op

    def apply(self, state, r, s, op):
        pcs, data, dflag, sflag, sums = state
        w = self.world
        err = None
        if op[0] == "publish":
            k = op[1]
            row = list(data[r])
            row[k & 1] = k
            data = data[:r] + (tuple(row),) + data[r + 1:]
        elif op[0] == "signal":
            k = op[1]
            dflag = tuple(
                dflag[h] if h == r else tuple(
                    tuple(k if (q == r and p == self._fi(k)) else dflag[h][p][q] for q in range(w)) for p in range(len(dflag[h])))
                for h in range(w))
        elif op[0] == "pull":
            _, k, src = op
            if data[src][k & 1] != k:
                err = "rank %d pulled rank %d's rows at step %d but the slot holds step %d" % (r, src, k, data[src][k & 1])
        elif op[0] == "sums":
            k = op[1]
            sums = tuple(
                tuple(tuple(k if (q == r and sl == (k & 1)) else sums[h][sl][q] for q in range(w)) for sl in range(2))
                for h in range(w))
            sflag = tuple(
                sflag[h] if h == r else tuple(
                    tuple(k if (q == r and p == self._fi(k)) else sflag[h][p][q] for q in range(w)) for p in range(len(sflag[h])))
                for h in range(w))
        elif op[0] == "combine":
            k = op[1]
            for q in range(w):
                if sums[r][k & 1][q] != k:
                    err = "rank %d combined step %d but holds rank %d's sums of step %d" % (r, k, q, sums[r][k & 1][q])
        row = list(pcs[r])
        row[s] += 1
        pcs = pcs[:r] + (tuple(row),) + pcs[r + 1:]
        return (pcs, data, dflag, sflag, sums), err

    def explore(self, limit=2_000_000):
        """Returns (states visited, first violation or None)."""
        start = self.initial()
        seen = {start}
        stack = [start]
        while stack:
            state = stack.pop()
            moves = []
            for r in range(self.world):
                for s in range(len(self.prog[r])):
                    op = self.enabled(state, r, s)
                    if op is not None:
                        moves.append((r, s, op))
            if not moves:
                done = all(state[0][r][s] >= len(self.prog[r][s]) for r in range(self.world) for s in range(len(self.prog[r])))
                if not done:
                    return len(seen), "deadlock at program counters %s" % (state[0],)
                continue
            for (r, s, op) in moves:
                nxt, err = self.apply(state, r, s, op)
                if err:
                    return len(seen), err
                if nxt not in seen:
                    if len(seen) >= limit:
                        return len(seen), "state limit reached"
                    seen.add(nxt)
                    stack.append(nxt)
        return len(seen), None


@pytest.mark.parametrize("world,steps,two_streams", [(2, 5, False), (2, 4, True), (3, 3, False), (3, 2, True)])
def test_peer_protocol_is_safe_and_live(world, steps, two_streams):
    n, violation = Model(world, steps, two_streams).explore()
    assert violation is None, violation
    assert n > 100      # the exploration did cover interleavings


def test_model_detects_the_shared_flag_bug():
    """One set of flags for both parities (the protocol's first version) is unsafe once two consecutive steps of a rank may
    overlap: step k + 1's flag lets a peer's step-k pull through before step k's rows are there."""
    n, violation = Model(2, 4, True, per_parity_flags=False).explore()
    assert violation is not None and ("pulled" in violation or "combined" in violation or "deadlock" in violation), violation
    # ... while on ONE stream that version was fine
    assert Model(2, 5, False, per_parity_flags=False).explore()[1] is None
