"""Event model of the fused kernel's per-CTA pipeline (developer tool, no GPU needed).

Roles: TMA producer (ring of NST Zj stages), one in-order MMA issuer (static order), two epilogue groups
(tile t -> group t%2), NS S buffers, 2 W buffers.  All times in SM cycles.
"""
import sys

def sim(T=200, E=2500, t1=768, t2=512, lam=300, L=1500, NST=5, NS=3, order="static", verbose=False):
    INF = float("inf")
    load_done = [None] * T      # Zj(t) landed
    s_ready = [None] * T        # MMA1(t) complete
    s_free = [None] * T         # epilogue(t) has S in registers
    w_ready = [None] * T        # epilogue(t) wrote W
    mma2_done = [None] * T
    f = [None] * T              # epilogue finish
    # producer: load(t) issued when stage free (mma2_done[t-NST]) and previous load issued
    # MMA issuer: program order list of ops
    ops = []
    if order == "static":       # iteration jj: MMA2(jj-NS), MMA1(jj)
        for jj in range(T + NS):
            if jj - NS >= 0: ops.append(("m2", jj - NS))
            if jj < T: ops.append(("m1", jj))
    elif order == "m1first":    # iteration jj: MMA1(jj), MMA2(jj-NS+... ) issue MMA1 before the MMA2 wait
        for jj in range(T + NS):
            if jj < T: ops.append(("m1", jj))
            if jj - (NS - 1) >= 0 and jj - (NS - 1) < T: ops.append(("m2", jj - (NS - 1)))
    # iterate to fixed point (times only depend on earlier events; simple repeated relaxation)
    grp_free = [0.0, 0.0]
    tensor_free = 0.0
    issue_t = 0.0
    prod_t = 0.0
    # process in a merged loop: we need epilogue times which depend on s_ready, which depend on issue order.
    # Do a simple time-stepped dependency resolution by repeated passes.
    for _ in range(4 * T):
        changed = False
        # producer
        pt = 0.0
        for t in range(T):
            dep = 0.0 if t < NST else mma2_done[t - NST]
            if dep is None: break
            issue = max(pt, dep + (lam if t >= NST else 0))
            pt = issue + 20
            v = issue + L
            if load_done[t] != v: load_done[t] = v; changed = True
        # MMA issuer
        it = 0.0; tf = 0.0
        for kind, t in ops:
            if kind == "m1":
                deps = [load_done[t]]
                if t >= NS: deps.append(s_free[t - NS])
                if any(d is None for d in deps): break
                it = max(it, max(d + lam for d in deps)) + 50
                start = max(it, tf); tf = start + t1
                v = tf + lam
                if s_ready[t] != v: s_ready[t] = v; changed = True
            else:
                if w_ready[t] is None: break
                it = max(it, w_ready[t] + lam) + 30
                start = max(it, tf); tf = start + t2
                if mma2_done[t] != tf: mma2_done[t] = tf; changed = True
        # epilogue groups
        gf = [0.0, 0.0]
        for t in range(T):
            g = t & 1
            deps = [s_ready[t], load_done[t]]
            if t >= 2: deps.append(mma2_done[t - 2])   # W buffer free
            if any(d is None for d in deps): break
            start = max(gf[g], max(deps))
            sf = start + 0.25 * E        # S slice in registers after ~1/4 of the tile (last chunk load) -- optimistic
            wr = start + E
            gf[g] = wr
            if s_free[t] != sf: s_free[t] = sf; changed = True
            if w_ready[t] != wr: w_ready[t] = wr; changed = True
            f[t] = wr
        if not changed: break
    done = [x for x in f if x is not None]
    n = len(done)
    per_tile = (done[-1] - done[n // 2]) / (n - 1 - n // 2)
    return per_tile

if __name__ == "__main__":
    for E in (800, 1500, 2500, 4000):
        for lam in (100, 300, 600):
            row = []
            for order in ("static", "m1first"):
                for NS in (3,):
                    row.append("%s NS=%d: %6.0f" % (order, NS, sim(E=E, lam=lam, order=order, NS=NS)))
            print("E=%5d lam=%4d | " % (E, lam) + " | ".join(row))
