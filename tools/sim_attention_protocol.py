#!/usr/bin/env python
"""Discrete-event model of the mbarrier protocol of attention_tc3_kernel (magicdrive_b200/csrc/attention_tc3.cuh): TMA
producer, MMA issuer and the softmax warps as coroutines over mbarriers with phase-parity waits, an in-order tensor pipe
with commits, asynchronous TMA completion and random latencies.  It checks, over many random schedules,
  * no deadlock,
  * no parity aliasing: every wait that passes does so because the phase it was written for completed,
  * no data hazard: S / P / O_a+O_b / Q / K / V are read in the version the reader expects and never overwritten while a
    reader of the previous version is outstanding.
It is a model of the protocol (transcribed by hand from the kernel), run on the CPU: it validates the design of the global
counters (git / gset / item) that the persistent kernel adds to attention_tc2, not the CUDA code itself.

  python tools/sim_attention_protocol.py [n_schedules]
"""
import heapq
import random
import sys

STAGES = 2


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending, self.completed = 0, self.completed + 1

    def passes(self, parity):
        return (self.completed & 1) != parity


class Sim:
    def __init__(self, seed, n_items, ntiles, n_sets):
        self.rng = random.Random(seed)
        self.now, self.events, self.seq = 0.0, [], 0
        self.n_items, self.ntiles, self.n_sets = n_items, ntiles, n_sets
        self.iters = ntiles * n_sets
        B = Barrier
        self.q_full, self.q_empty = [B("q_full0", 1), B("q_full1", 1)], [B("q_empty0", 1), B("q_empty1", 1)]
        self.kv_full = [B(f"kv_full{i}", 1) for i in range(STAGES)]
        self.kv_empty = [B(f"kv_empty{i}", 1) for i in range(STAGES)]
        self.s_full, self.p_full = B("s_full", 1), B("p_full", 8)
        self.pv_done, self.o_full, self.o_free = B("pv_done", 1), B("o_full", 1), B("o_free", 8)
        # resource versions: what each buffer currently holds
        self.q_ver = [None, None]            # item index
        self.kv_ver = [None] * STAGES        # global iteration whose K/V the stage holds
        self.s_ver = None                    # ("S", g) after QK(g) executes, ("P", g, n_written) while softmax writes
        self.p_writes = 0
        self.o_sets_accum = None             # (gset, tiles accumulated)
        self.o_reads = {}                    # gset -> softmax warps that finished reading that set's O
        self.kv_readers = [0] * STAGES       # outstanding MMA reads per stage
        self.mma_queue, self.mma_busy_until = [], 0.0
        self.procs, self.blocked = [], {}
        self.errors = []

    # ---------------------------------------------------------------- scheduler
    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.events, (t, self.seq, fn))

    def spawn(self, gen, name):
        p = {"gen": gen, "name": name, "done": False}
        self.procs.append(p)
        self.at(self.now, lambda: self.step(p))

    def step(self, p):
        try:
            op = next(p["gen"])
        except StopIteration:
            p["done"] = True
            return
        if op[0] == "delay":
            self.at(self.now + op[1], lambda: self.step(p))
        elif op[0] == "wait":
            _, bar, parity, intent = op
            self.blocked[id(p)] = (p, bar, parity, intent)
            self.poll()

    def poll(self):
        for key, (p, bar, parity, intent) in list(self.blocked.items()):
            if bar.passes(parity):
                if intent is not None and bar.completed != intent + 1:
                    self.errors.append(f"{p['name']}: wait on {bar.name} for phase {intent} passed with {bar.completed} completed")
                del self.blocked[key]
                self.at(self.now + self.rng.uniform(0.01, 0.3), lambda p=p: self.step(p))

    def arrive(self, bar):
        bar.arrive()
        self.poll()

    # ---------------------------------------------------------------- tensor pipe: in order, commits fire when all before are done
    def mma(self, dur, on_start, on_done):
        start = max(self.now, self.mma_busy_until)
        self.mma_busy_until = start + dur
        self.at(start, on_start)
        self.at(start + dur, on_done)

    def commit(self, bar):
        self.at(max(self.now, self.mma_busy_until) + 0.01, lambda: self.arrive(bar))

    def err(self, msg):
        self.errors.append(f"t={self.now:.2f} {msg}")

    # ---------------------------------------------------------------- the three roles
    def tma(self):
        stage, phase, g = 0, 0, 0
        for k in range(self.n_items):
            qb = k & 1
            yield ("wait", self.q_empty[qb], ((k >> 1) & 1) ^ 1, (k >> 1) - 1 if k >= 2 else None)

            def q_done(qb=qb, k=k):
                self.q_ver[qb] = k
                self.arrive(self.q_full[qb])
            self.at(self.now + self.rng.uniform(0.5, 3.0), q_done)
            for _ in range(self.iters):
                n_use = g // STAGES
                yield ("wait", self.kv_empty[stage], phase ^ 1, n_use - 1 if n_use >= 1 else None)
                if self.kv_readers[stage]:
                    self.err(f"TMA overwrites K/V stage {stage} with {self.kv_readers[stage]} MMA reads outstanding")

                def kv_done(stage=stage, g=g):
                    self.kv_ver[stage] = g
                    self.arrive(self.kv_full[stage])
                self.at(self.now + self.rng.uniform(0.5, 3.0), kv_done)
                g += 1
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1
                yield ("delay", 0.02)

    def issue_qk(self, g):
        k, lit = divmod(g, self.iters)
        qb, st = k & 1, g % STAGES
        if lit == 0:
            yield ("wait", self.q_full[qb], (k >> 1) & 1, k >> 1)
        yield ("wait", self.kv_full[st], (g // STAGES) & 1, g // STAGES)
        self.kv_readers[st] += 1

        def start():
            if self.q_ver[qb] != k:
                self.err(f"QK({g}) reads Q buffer {qb} holding item {self.q_ver[qb]}, wants {k}")
            if self.kv_ver[st] != g:
                self.err(f"QK({g}) reads K stage {st} holding iteration {self.kv_ver[st]}")
            if self.s_ver is not None and self.s_ver[0] == "P" and not self.s_ver[2]:
                self.err(f"QK({g}) overwrites P({self.s_ver[1]}) before PV read it")

        def done():
            self.s_ver = ("S", g)
            self.kv_readers[st] -= 1
        self.mma(self.rng.uniform(0.2, 0.6), start, done)
        self.commit(self.s_full)
        if lit == self.iters - 1:
            self.commit(self.q_empty[qb])

    def mma_warp(self):
        total = self.n_items * self.iters
        if total:
            yield from self.issue_qk(0)
        j = gset = 0
        for g in range(total):
            st = g % STAGES
            yield ("wait", self.p_full, g & 1, g)
            if j == 0 and gset > 0:
                yield ("wait", self.o_free, (gset - 1) & 1, gset - 1)
            self.kv_readers[st] += 1

            def start(g=g, st=st, j=j, gset=gset):
                if self.s_ver != ("P", g, False):
                    self.err(f"PV({g}) expects P({g}) in the S buffer, finds {self.s_ver}")
                if self.kv_ver[st] != g:
                    self.err(f"PV({g}) reads V stage {st} holding iteration {self.kv_ver[st]}")
                if j == 0:
                    if gset > 0 and self.o_reads.get(gset - 1, 0) != 8:
                        self.err(f"PV({g}) overwrites O while only {self.o_reads.get(gset - 1, 0)}/8 warps have read set {gset - 1}")
                    self.o_sets_accum = (gset, 0)
                elif self.o_sets_accum is None or self.o_sets_accum[0] != gset:
                    self.err(f"PV({g}) accumulates into O of set {self.o_sets_accum}, wants {gset}")

            def done(g=g, st=st):
                self.s_ver = ("P", g, True)  # P consumed
                self.o_sets_accum = (self.o_sets_accum[0], self.o_sets_accum[1] + 1)
                self.kv_readers[st] -= 1
            self.mma(self.rng.uniform(0.1, 0.4), start, done)
            self.commit(self.pv_done)
            self.commit(self.kv_empty[st])
            if j == self.ntiles - 1:
                self.commit(self.o_full)
            if g + 1 < total:
                yield from self.issue_qk(g + 1)
            j += 1
            if j == self.ntiles:
                j, gset = 0, gset + 1

    def softmax_warp(self, w):
        g = gset = 0
        for k in range(self.n_items):
            for _ in range(self.n_sets):
                for j in range(self.ntiles):
                    yield ("wait", self.s_full, g & 1, g)
                    if self.s_ver not in (("S", g),) and not (self.s_ver is not None and self.s_ver[0] == "P" and self.s_ver[1] == g):
                        self.err(f"softmax warp {w} reads S({g}) but the buffer holds {self.s_ver}")
                    yield ("delay", self.rng.uniform(0.3, 1.5))
                    if j > 0 and self.rng.random() < 0.15:  # rare rescale path
                        yield ("wait", self.pv_done, (g - 1) & 1, None)
                        if self.pv_done.completed < g:
                            self.err(f"rescale at iteration {g} passed pv_done with only {self.pv_done.completed} PVs complete")
                    # write P over S
                    self.p_writes += 1
                    if self.p_writes == 8:
                        self.p_writes = 0
                        self.s_ver = ("P", g, False)
                    self.arrive(self.p_full)
                    g += 1
                yield ("wait", self.o_full, gset & 1, gset)
                if self.o_sets_accum != (gset, self.ntiles):
                    self.err(f"softmax warp {w} reads O of set {gset}, accumulator state {self.o_sets_accum}")
                yield ("delay", self.rng.uniform(0.2, 0.8))
                if self.o_sets_accum != (gset, self.ntiles):
                    self.err(f"O of set {gset} changed under softmax warp {w}: {self.o_sets_accum}")
                self.o_reads[gset] = self.o_reads.get(gset, 0) + 1
                self.arrive(self.o_free)
                gset += 1

    def run(self):
        self.spawn(self.tma(), "tma")
        self.spawn(self.mma_warp(), "mma")
        for w in range(8):
            self.spawn(self.softmax_warp(w), f"softmax{w}")
        while self.events:
            self.now, _, fn = heapq.heappop(self.events)
            fn()
            if self.errors:
                break
        if not self.errors and not all(p["done"] for p in self.procs):
            stuck = [(p["name"], b.name, par, b.completed) for p, b, par, _ in self.blocked.values()]
            self.errors.append(f"deadlock: {stuck}")
        return self.errors


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    bad = 0
    for seed in range(n):
        rng = random.Random(seed)
        cfg = dict(n_items=rng.choice([1, 2, 3, 4, 5]), ntiles=rng.choice([1, 2, 3, 11]), n_sets=rng.choice([1, 2]))
        errs = Sim(seed, **cfg).run()
        if errs:
            bad += 1
            print(f"seed {seed} {cfg}: {errs[0]}")
            if bad > 5:
                break
    print(f"{n} random schedules, {bad} with violations")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
